// ref_driver.cpp - headless host driver around the REFERENCE's own per-pixel code.
//
// TEST INFRASTRUCTURE ONLY.  This TU textually includes oracle/_ref/kernel_patched.inc,
// which make_ref.py produces from /root/reference/kernel.cu (read in place; the 16-line
// mechanical patch touches signatures only, never arithmetic).  With the stub headers in
// stubs/ the reference's __device__/__global__ functions become ordinary host functions:
// rayTrace (kernel.cu:1614-1690) is called once per pixel and writes the pixel itself.
//
// What this file adds (and nothing else): building the scene objects from flat arrays,
// the window.h callbacks, the per-pixel call loop, and hit-id/t extraction by calling the
// reference's castRay (kernel.cu:1287) on a ray built with the reference's own helpers in
// the order rayTrace uses them (kernel.cu:1624-1631).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <limits>
#include <map>
#include <string>
#include <vector>
#include <omp.h>

#include "oracle.h"
#include "sprite_raw.h"

// reach skybox::skyboxTex (protected, kernel.cu:1168-1170) to free it after a frame
#define protected public
#include "kernel_patched.inc"
#undef protected

thread_local uint3 threadIdx = {0, 0, 0};
thread_local uint3 blockIdx = {0, 0, 0};
thread_local dim3 blockDim(1, 1, 1);

// ---- window.h callbacks (window.cpp:86-91,130-132); only update() uses them ----
static int g_w = 0, g_h = 0;
int getScreenWidth() { return g_w; }
int getScreenHeight() { return g_h; }
void setPixelBuff(unsigned int*) {}
void drawPixel(int, int, int) {}
void Set_Background() {}
void Clear_Screen(unsigned int) {}
int make_inbound(int lo, int hi, int v) { return v > hi ? hi : (v < lo ? lo : v); }
int getBuffSize() { return 0; }
void setScreen(int*) {}

extern "C" const char* oracle_kind(void) { return "reference"; }

extern "C" uint32_t oracle_rgb_to_int(int r, int g, int b) { return (uint32_t)rgbToInt(r, g, b); }

extern "C" int oracle_sphere_intersect(const float org[3], const float dir[3], const float centre[3],
                                       float radius_member, float* t) {
    sphere s({centre[0], centre[1], centre[2]}, 0.f);
    s.radius = radius_member;
    ray r({org[0], org[1], org[2]}, {dir[0], dir[1], dir[2]});
    float tt = 0.f;
    bool hit = s.intersect(r, tt);
    if (t) *t = tt;
    return hit ? 1 : 0;
}

// The reference's OWN OBJ loader and BVH builder (mesh::mesh -> createBvhMesh, kernel.cu:577-936), run on a file,
// with the result copied out as flat arrays (the loader prints every leaf box to stdout: silenced here).
extern "C" int oracle_ref_build_mesh(const char* obj_path, float* tris, int32_t cap_tris, int32_t* n_tris,
                                     int32_t* has_normals, float* box_bounds, int32_t* box_offsets, int32_t cap_boxes,
                                     int32_t* n_boxes, int32_t* box_indices, int32_t cap_indices) {
    {
        std::ifstream probe(obj_path);
        if (!probe.is_open()) return 1;
    }
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    mesh* m = new mesh(std::string(obj_path));
    std::cout.rdbuf(old);
    int rc = 0;
    if (m->poly_count > cap_tris || m->bvhbox_count > cap_boxes) rc = 1;
    if (!rc) {
        *n_tris = m->poly_count;
        *has_normals = m->has_normals ? 1 : 0;
        *n_boxes = m->bvhbox_count;
        memcpy(tris, m->h_tri_arr, sizeof(triangle) * (size_t)m->poly_count);
        int off = 0;
        for (int j = 0; j < m->bvhbox_count && !rc; j++) {
            const Bvhbox& b = m->h_box[j];
            if (off + b.length > cap_indices) {
                rc = 1;
                break;
            }
            const vec3d lo = b.bvhbox->bounds[0], hi = b.bvhbox->bounds[1];
            float* bb = box_bounds + 6 * (size_t)j;
            bb[0] = lo.x; bb[1] = lo.y; bb[2] = lo.z; bb[3] = hi.x; bb[4] = hi.y; bb[5] = hi.z;
            box_offsets[j] = off;
            memcpy(box_indices + off, b.indexes, sizeof(int) * (size_t)b.length);
            off += b.length;
        }
        if (!rc) box_offsets[m->bvhbox_count] = off;
    }
    return rc;  // the mesh object is leaked on purpose: the reference never frees one either
}

extern "C" int oracle_render(const oracle_frame* f, uint32_t* pixels, int32_t* hit_id, float* hit_t,
                             uint64_t* counts, int n_threads) {
    if (!f || f->width <= 0 || f->height <= 0 || f->y_step <= 0) return 1;
    if (counts) counts[0] = counts[1] = counts[2] = counts[3] = 0;
    const int W = f->width, H = f->height;

    sprite_raw_register("ore:tex", raw_image{f->tex_w, f->tex_h, f->tex_r, f->tex_g, f->tex_b});
    sprite_raw_register("ore:sky", raw_image{f->sky_w, f->sky_h, f->sky_r, f->sky_g, f->sky_b});

    // scene container exactly as the reference holds it (kernel.cu:1176-1244)
    object* o = new object();
    o->sphere_count = f->n_spheres;
    o->s1 = new sphere[f->n_spheres > 0 ? f->n_spheres : 1];
    for (int i = 0; i < f->n_spheres; i++) {
        const float* s = f->spheres + 4 * (size_t)i;
        o->s1[i] = sphere({s[0], s[1], s[2]}, 0.f);
        o->s1[i].radius = s[3];               // stored member (ctor would have squared r)
    }
    o->sphereAllocMem();                       // kernel.cu:1208-1212 (32-byte AoS copy)
    // cubes (kernel.cu:1186,1193-1199,1224-1228) and the single plane the reference supports (:1187,1213-1217)
    if (f->n_planes < 0 || f->n_planes > 1) return 3;
    o->cube_count = f->n_cubes;
    o->c1 = new cube[f->n_cubes > 0 ? f->n_cubes : 1];
    for (int i = 0; i < f->n_cubes; i++) {
        const float* c = f->cubes + 6 * (size_t)i;
        o->c1[i] = cube({c[0], c[1], c[2]}, {c[3], c[4], c[5]});
    }
    o->cubeAllocMem();
    if (f->n_planes == 1)
        o->planes = new plane({f->planes[0], f->planes[1], f->planes[2]}, {f->planes[3], f->planes[4], f->planes[5]});
    else
        o->planes = new plane({0, -4, 0}, normalise(vec3d({0, 1, 0})));   // kernel.cu:1187
    o->plane_count = f->n_planes;
    o->planeAllocMem();
    o->texture = new sprite("ore:tex");
    // a mesh whose file does not exist: the ctor returns early (kernel.cu:583-585) and the
    // calloc-backed managed allocation leaves bvhbox_count == 0, so the triangle loops
    // (kernel.cu:1293-1328,1475-1497) run zero times - the sphere-only scene.
    o->mesh1 = new mesh("/nonexistent/ore-none.obj");
    if (o->mesh1->bvhbox_count != 0) return 2;
    std::vector<cube*> box_cubes;
    if (f->n_tris > 0) {
        // the reference's mesh object filled from the flat arrays exactly as its loader + createBvhMesh + allocMem
        // leave it (kernel.cu:559-1017): triangle array, has_normals, leaf boxes with their index lists
        mesh* m = o->mesh1;
        static_assert(sizeof(triangle) == 27 * sizeof(float), "triangle layout");
        m->poly_count = f->n_tris;
        m->h_tri_arr = new triangle[f->n_tris];
        memcpy(m->h_tri_arr, f->tris, sizeof(triangle) * (size_t)f->n_tris);
        m->d_tri_arr = m->h_tri_arr;
        m->has_normals = f->mesh_has_normals != 0;
        m->bvhbox_count = f->n_boxes;
        m->h_box = new Bvhbox[f->n_boxes > 0 ? f->n_boxes : 1];
        for (int j = 0; j < f->n_boxes; j++) {
            const float* b = f->box_bounds + 6 * (size_t)j;
            const int len = f->box_offsets[j + 1] - f->box_offsets[j];
            int* idx = new int[len > 0 ? len : 1];
            memcpy(idx, f->box_indices + f->box_offsets[j], sizeof(int) * (size_t)len);
            m->h_box[j] = Bvhbox({b[0], b[1], b[2]}, {b[3], b[4], b[5]}, idx, len);
            // the loader builds the leaf cube with cube(low, high) whose ctor also derives `orgin`; the bounds are
            // what the intersection reads.  AllocMem's device copies alias the host objects here.
            m->h_box[j].d_bvhbox = m->h_box[j].bvhbox;
            m->h_box[j].d_indexes = idx;
            box_cubes.push_back(m->h_box[j].bvhbox);
        }
        m->d_box = m->h_box;
    }

    skybox* sky = new skybox("ore:sky", f->sky_size);

    std::vector<light> L(f->n_lights > 0 ? f->n_lights : 1);
    for (int i = 0; i < f->n_lights; i++) {
        const float* l = f->lights + 7 * (size_t)i;
        L[i] = light({l[0], l[1], l[2]}, l[3], l[4], l[5], l[6]);
    }

    camera c({f->cam_org[0], f->cam_org[1], f->cam_org[2]}, {0, 0, 1}, 0.f);
    c.Camyaw = f->cam_yaw;
    c.Campitch = f->cam_pitch;
    c.aspect = (float)H / W;                   // kernel.cu:1773 (never read by the kernel)

    const float asp = f->aspect;
    const int n_rows = (f->y1 - f->y0 + f->y_step - 1) / f->y_step;
    if (n_threads > 0) omp_set_num_threads(n_threads);

#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < n_rows; k++) {
        const int y = f->y0 + k * f->y_step;
        std::vector<unsigned int> row(W);
        blockDim = dim3(1, 1, 1);
        threadIdx = {0, 0, 0};
        for (int x = 0; x < W; x++) {
            blockIdx = {(unsigned)x, (unsigned)y, 0};
            // rayTrace writes pixels[y*width+x]; bias the base so that lands in row[x]
            unsigned int* base = row.data() - (ptrdiff_t)y * W;
            rayTrace(base, W, H, asp, *o, L.data(), f->n_lights, c, *sky);
            if (hit_id || hit_t) {
                // same calls, same order as kernel.cu:1624-1631, then castRay (kernel.cu:1640)
                float dx = asp * (2 * (x + 0.5) / (float)W) - 1;
                float dy = asp * (2 * (y + 0.5) / (float)H) * ((float)H / W) - 1;
                vec3d eyePos({0, 0, (-1 / asp)});
                vec3d dir = vec3d({dx, dy, 0});
                ray cam_ray(add(eyePos, c.Org), c.rotateDir(normalise(sub(dir, eyePos)), c.Camyaw, c.Campitch));
                int ht = -1, hi = -1;
                float nt, nu, nv, tx, ty;
                vec3d no, nn;
                bool hit = castRay(*o, cam_ray, ht, hi, nt, nu, nv, no, nn, tx, ty);
                if (hit && ht == 3) hi += f->n_spheres;                      // same encoding as oracle.h
                if (hit && ht == 2) hi += f->n_spheres + f->n_cubes;
                if (hit && ht == 0) hi += f->n_spheres + f->n_cubes + f->n_planes;
                if (hit_id) hit_id[(size_t)k * W + x] = hit ? hi : -1;
                if (hit_t) hit_t[(size_t)k * W + x] = nt;
            }
        }
        if (pixels) memcpy(pixels + (size_t)k * W, row.data(), sizeof(unsigned int) * W);
    }

    sprite_raw_free(o->texture);
    sprite_raw_free(sky->skyboxTex);
    delete sky->box;
    delete sky;
    cudaFree(o->d_spheres);
    cudaFree(o->d_cubes);
    cudaFree(o->d_planes);
    delete[] o->c1;
    delete o->planes;
    delete[] o->s1;
    if (f->n_tris > 0) {
        for (int j = 0; j < f->n_boxes; j++) delete[] o->mesh1->h_box[j].indexes;
        for (cube* c : box_cubes) delete c;
        delete[] o->mesh1->h_box;
        delete[] o->mesh1->h_tri_arr;
    }
    delete o->mesh1;
    delete o;
    sprite_raw_unregister("ore:tex");
    sprite_raw_unregister("ore:sky");
    return 0;
}
