// ref_gpu_driver.cu - runs the REFERENCE's own rayTrace kernel and update() on a real GPU, headless.
//
// TEST / MEASUREMENT INFRASTRUCTURE ONLY ("reference kernel on B200", BASELINE.json north_star:
// "If the reference kernel also builds headless for sm_100, it is reported alongside").
// Textually includes oracle/_ref/kernel_patched.inc (the reference's kernel.cu with the mechanical
// signature patch of make_ref.py) and is compiled by nvcc for sm_100 against the REAL CUDA runtime.
// Adds only: scene construction from flat arrays, the window.h callbacks, timing.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

#include "oracle.h"
#include "sprite_raw.h"

#define protected public
#include "kernel_patched.inc"
#undef protected

static int g_w = 0, g_h = 0;
static std::vector<unsigned int> g_frame;
int getScreenWidth() { return g_w; }
int getScreenHeight() { return g_h; }
void setPixelBuff(unsigned int* pixels) { memcpy(g_frame.data(), pixels, sizeof(unsigned int) * g_w * g_h); }  // window.cpp:130-132
void drawPixel(int, int, int) {}
void Set_Background() {}
void Clear_Screen(unsigned int) {}
int make_inbound(int lo, int hi, int v) { return v > hi ? hi : (v < lo ? lo : v); }
int getBuffSize() { return 0; }
void setScreen(int*) {}

// Renders `frames` full frames with the reference's update() (per-frame managed alloc, launch, sync, copy) and,
// separately, times the bare rayTrace kernel with CUDA events.  cams: frames x {x,y,z,yaw,pitch}.
extern "C" int refgpu_render(const oracle_frame* f, const float* cams, int frames, uint32_t* pixels_last,
                             float* ms_update_avg, float* ms_kernel_avg) {
    if (!f || frames <= 0) return 1;
    g_w = f->width;
    g_h = f->height;
    g_frame.assign((size_t)g_w * g_h, 0u);
    sprite_raw_register("ore:tex", raw_image{f->tex_w, f->tex_h, f->tex_r, f->tex_g, f->tex_b});
    sprite_raw_register("ore:sky", raw_image{f->sky_w, f->sky_h, f->sky_r, f->sky_g, f->sky_b});

    // the reference's globals (kernel.cu:1692-1702)
    objs = new object();
    objs->sphere_count = f->n_spheres;
    objs->plane_count = 0;
    objs->cube_count = 0;
    objs->s1 = new sphere[f->n_spheres > 0 ? f->n_spheres : 1];
    for (int i = 0; i < f->n_spheres; i++) {
        const float* s = f->spheres + 4 * (size_t)i;
        objs->s1[i] = sphere({s[0], s[1], s[2]}, 0.f);
        objs->s1[i].radius = s[3];
    }
    objs->sphereAllocMem();
    objs->texture = new sprite("ore:tex");
    objs->mesh1 = new mesh("/nonexistent/ore-none.obj");  // ctor returns early; managed memory is not zeroed:
    objs->mesh1->bvhbox_count = 0;                        // make the triangle loops run zero times
    Skybox = new skybox("ore:sky", f->sky_size);
    light_size = f->n_lights;
    lights = new light[f->n_lights > 0 ? f->n_lights : 1];
    for (int i = 0; i < f->n_lights; i++) {
        const float* l = f->lights + 7 * (size_t)i;
        lights[i] = light({l[0], l[1], l[2]}, l[3], l[4], l[5], l[6]);
    }
    // update() copies lightByteSize = 37 floats x light_size from `lights` (kernel.cu:1697,1778): an over-read
    // of the host array in the reference; give it a large enough source
    {
        light* big = (light*)calloc(1, sizeof(float) * 37 * (f->n_lights > 0 ? f->n_lights : 1) + sizeof(light));
        memcpy(big, lights, sizeof(light) * f->n_lights);
        lights = big;
        lightByteSize = sizeof(float) * 37 * light_size;
    }
    aspect = f->aspect;

    auto set_cam = [&](int i) {
        cam.Org = {cams[5 * i], cams[5 * i + 1], cams[5 * i + 2]};
        cam.Camyaw = cams[5 * i + 3];
        cam.Campitch = cams[5 * i + 4];
    };
    set_cam(0);
    update();  // warm-up
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < frames; i++) {
        set_cam(i);
        update();
    }
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (ms_update_avg) *ms_update_avg = (float)(sec / frames * 1e3);
    if (pixels_last) memcpy(pixels_last, g_frame.data(), sizeof(unsigned int) * g_w * g_h);

    // bare kernel: same launch shape as update() (kernel.cu:1780-1783)
    unsigned int* px = nullptr;
    light* d_lights = nullptr;
    cudaMallocManaged((void**)&px, sizeof(unsigned int) * g_w * g_h);
    cudaMalloc((void**)&d_lights, lightByteSize);
    cudaMemcpy(d_lights, lights, lightByteSize, cudaMemcpyHostToDevice);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    dim3 blocks(g_w / tx + 1, g_h / ty + 1), threads(tx, ty);
    float total = 0.f;
    for (int i = 0; i < frames; i++) {
        set_cam(i);
        cudaEventRecord(a);
        rayTrace<<<blocks, threads>>>(px, g_w, g_h, aspect, *objs, d_lights, light_size, cam, *Skybox);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        total += ms;
    }
    if (ms_kernel_avg) *ms_kernel_avg = total / frames;
    cudaError_t e = cudaGetLastError();
    cudaFree(px);
    cudaFree(d_lights);
    sprite_raw_unregister("ore:tex");
    sprite_raw_unregister("ore:sky");
    return e == cudaSuccess ? 0 : 2;
}
