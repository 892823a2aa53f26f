// kernel_shim.cpp - replaces the HOST half of the reference's kernel.cu for the sphere render path:
// the global scene, onStart() (kernel.cu:1704-1714) and update() (kernel.cu:1762-1792), on top of the
// C ABI in include/ore_render.h.  window.cpp links against this exactly as it links against kernel.cu.
//
// What changes relative to the reference's update(): no per-frame cudaMallocManaged/cudaMalloc/cudaFree
// (kernel.cu:1775-1790) - the context owns persistent device buffers - and the frame lands in a pinned
// host buffer by one async copy instead of managed-memory page migration; setPixelBuff still receives a
// host-readable pointer and copies from it (window.cpp:130-132).
#include "ore_host.h"

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/ore_render.h"
#include "ore_memmanager.h"
#include "ore_mesh.h"
#include "ore_sprite.h"
#include "ore_window_callbacks.h"

#define checkOre(expr) check_ore((expr), #expr, g_ctx ? ore_last_error(g_ctx) : "", __FILE__, __LINE__)

namespace {
ore_context* g_ctx = nullptr;
// kernel.cu:1692-1702 - the reference's globals
int light_size = 3;
ore_camera cam = {{4, 3, 10}, {0, 0, 1}, 0.f, 180.f, -20.f};          // camera cam({4,3,10},{0,0,1},0) + :261
float aspect = (float)tan((90 * 0.5 * 3.1415) / 180);                  // kernel.cu:1701
int g_sphere_count = 64;                                               // the reference ships 0 (kernel.cu:1231)
unsigned g_seed = 1;                                                   // unseeded rand()
std::string g_texture = "proc:smooth:512:512:102";
std::string g_sky = "proc:smooth:1024:512:203";
std::string g_obj;                                                     // loadMesh(file, ...), kernel.cu:1706
sprite* texture = nullptr;
sprite* skyTex = nullptr;
unsigned int* g_pixels = nullptr;  // pinned host frame(s) handed to setPixelBuff: [2][cap] when pipelined
size_t g_pixels_cap = 0;
// pipelined presentation (oreConfigurePresentation(1)): update() enqueues frame f (render + async copy into one of
// two pinned frames) and hands setPixelBuff frame f-1, whose copy overlapped this frame's kernels
bool g_pipelined = false;
volatile uint32_t* g_done = nullptr;   // pinned: number of frames whose copy has landed (written by the copy stream)
uint32_t g_submitted = 0, g_presented = 0;

unsigned msvc_rand(unsigned& s) {
    s = s * 214013u + 2531011u;
    return (s >> 16) & 0x7fff;
}
}  // namespace

void oreConfigureScene(int sphere_count, unsigned seed, const char* tex, const char* sky) {
    g_sphere_count = sphere_count;
    g_seed = seed;
    if (tex) g_texture = tex;
    if (sky) g_sky = sky;
}

void oreSetMeshFile(const char* obj_path) { g_obj = obj_path ? obj_path : ""; }

void oreConfigurePresentation(int pipelined) { g_pipelined = pipelined != 0; }

void oreSetCamera(float x, float y, float z, float yaw_deg, float pitch_deg) {
    cam.org[0] = x;
    cam.org[1] = y;
    cam.org[2] = z;
    cam.yaw = yaw_deg;
    cam.pitch = pitch_deg;
}

void onStart() {
    checkOre(ore_create(&g_ctx, 0));
    // object::loadMesh sphere generator, kernel.cu:1189-1191: centre = (rand()%100)/10, r = (rand()%100)/100;
    // the sphere ctor stores r*r (kernel.cu:287)
    std::vector<float> s((size_t)g_sphere_count * 4);
    unsigned st = g_seed;
    for (int i = 0; i < g_sphere_count; i++) {
        s[4 * i + 0] = (float)(msvc_rand(st) % 100) / 10;
        s[4 * i + 1] = (float)(msvc_rand(st) % 100) / 10;
        s[4 * i + 2] = (float)(msvc_rand(st) % 100) / 10;
        const float r = (float)(msvc_rand(st) % 100) / 100;
        s[4 * i + 3] = r * r;
    }
    checkOre(ore_set_spheres(g_ctx, s.data(), g_sphere_count));
    if (!g_obj.empty()) {
        // objs->loadMesh(file, ...): mesh(file) parses the OBJ and builds the flat BVH on the host (kernel.cu:1183,
        // 577-936); ore_set_mesh replaces mesh::allocMem (kernel.cu:1184,999-1017)
        OreMesh m;
        if (ore_load_obj(g_obj, m) && m.n_tris() > 0)
            checkOre(ore_set_mesh(g_ctx, m.tris.data(), m.n_tris(), m.has_normals ? 1 : 0, m.box_bounds.data(),
                                  m.box_offsets.data(), m.box_indices.data(), m.n_boxes()));
    }
    texture = new sprite(g_texture);   // objs->texture = new sprite(tex), kernel.cu:1201
    skyTex = new sprite(g_sky);        // new skybox(img, 10000), kernel.cu:1700
    checkOre(ore_set_texture(g_ctx, texture->rBuff->data, texture->gBuff->data, texture->bBuff->data, texture->width,
                             texture->height));
    checkOre(ore_set_sky(g_ctx, skyTex->rBuff->data, skyTex->gBuff->data, skyTex->bBuff->data, skyTex->width,
                         skyTex->height, 10000.f));
    // kernel.cu:1708-1712
    const float lights[3][7] = {{20, 20, 20, 20, 1, 0, 0}, {0, 20, -20, 20, 0, 0, 1}, {0, 20, 0, 20, 0, 1, 0}};
    checkOre(ore_set_lights(g_ctx, &lights[0][0], light_size));
}

static void present_next() {
    // frame g_presented is complete once the copy stream has stored g_presented + 1 into the pinned counter
    while ((int32_t)(*g_done - (g_presented + 1)) < 0) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    setPixelBuff(g_pixels + (size_t)(g_presented & 1u) * g_pixels_cap);
    g_presented++;
}

void update() {
    // checkKey() (kernel.cu:1716-1759) is Win32 key polling; headless callers use oreSetCamera()
    const int width = getScreenWidth(), height = getScreenHeight();
    const size_t n = (size_t)width * height;
    cam.aspect = (float)height / width;  // kernel.cu:1773 (never read by the kernel)
    if (n > g_pixels_cap) {
        if (g_pipelined) oreFlush();
        if (g_pixels) checkCudaErrors(cudaFreeHost(g_pixels));
        checkCudaErrors(cudaMallocHost((void**)&g_pixels, 2 * n * sizeof(unsigned int)));
        g_pixels_cap = n;
    }
    if (!g_done) {
        checkCudaErrors(cudaMallocHost((void**)&g_done, 64));
        *g_done = 0;
    }
    ore_frame fr = {width, height, 0, height, 1, aspect, ORE_FLAG_NO_KERNEL_TIMING, 0, 0};
    if (!g_pipelined) {
        checkOre(ore_render(g_ctx, &cam, &fr, g_pixels));  // launch + sync + device->host, like :1783-1788
        setPixelBuff(g_pixels);
        return;
    }
    // throughput mode: one frame of latency, no host stall on this frame's kernels
    checkOre(ore_render_async_signal(g_ctx, &cam, &fr, g_pixels + (size_t)(g_submitted & 1u) * g_pixels_cap,
                                     (uint32_t*)g_done, g_submitted + 1));
    g_submitted++;
    if (g_submitted - g_presented >= 2) present_next();
}

void oreFlush() {
    while (g_presented < g_submitted) present_next();
}

void oreShutdown() {
    if (g_ctx) ore_wait(g_ctx);
    g_presented = g_submitted;
    if (g_ctx) ore_destroy(g_ctx);
    g_ctx = nullptr;
    delete texture;
    delete skyTex;
    texture = skyTex = nullptr;
    if (g_pixels) cudaFreeHost(g_pixels);
    g_pixels = nullptr;
    g_pixels_cap = 0;
    if (g_done) cudaFreeHost((void*)g_done);
    g_done = nullptr;
    g_submitted = g_presented = 0;
}
