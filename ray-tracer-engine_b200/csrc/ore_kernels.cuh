// ore_kernels.cuh - the three kernels of the render hot path (sm_100a, --fmad=false).
//
//   prep_frame_kernel : per-frame tables (dx per column, dy per row, in the reference's
//                       double arithmetic) and per-sphere PRIMARY filter coefficients
//                       (all primary rays share one origin, so L = O - c and C are per-
//                       sphere constants of the frame).
//   primary_kernel<P> : nearest hit for P pixels per thread (one warp = one 32*P pixel
//                       strip of a row); sphere tiles staged in shared memory by TMA bulk
//                       copies; sky lookup + packed store for miss pixels; hit pixels are
//                       compacted into a list.                      (kernel.cu:1614-1640,
//                                                                    1330-1342, 1146-1166)
//   shadow_kernel<NL> : one thread per HIT pixel: shading set-up, 10*NL soft-shadow rays
//                       held in registers, any-hit over the staged sphere tiles, light
//                       accumulation and the packed pixel store.     (kernel.cu:1396-1405,
//                                                                    1643-1684, 1432-1544)
#pragma once
#include "ore_device.cuh"

namespace ore {

constexpr int MAX_LIGHTS = 16;
constexpr int CTA_THREADS = 256;
constexpr int CTA_WARPS = CTA_THREADS / 32;
constexpr int MAX_STAGES = 4;
// shadow kernel launch shape: tunable at build time (see DESIGN.md "Shadow kernel tuning")
#ifndef ORE_SHADOW_THREADS
#define ORE_SHADOW_THREADS 256
#endif
#ifndef ORE_SHADOW_MIN_CTAS
#define ORE_SHADOW_MIN_CTAS 2
#endif
// The shadow-pass kernels are warp-independent persistent kernels: a CTA's registers and shared memory stay
// allocated until its LAST warp has run out of work, so small CTAs hand an SM back to the next kernel (the next
// frame's, or the other stage's) almost warp by warp at the tail of a launch.
#ifndef ORE_BEAM_THREADS
#define ORE_BEAM_THREADS 256
#endif
#ifndef ORE_BEAM_MIN_CTAS
#define ORE_BEAM_MIN_CTAS (768 / ORE_BEAM_THREADS)  // shadow_beam_kernel: 80 registers, 768 threads per SM
#endif
#ifndef ORE_STAGE_A_THREADS
#define ORE_STAGE_A_THREADS 256
#endif
#ifndef ORE_SHADOW_SG
#define ORE_SHADOW_SG 4
#endif
constexpr int SHADOW_THREADS = ORE_SHADOW_THREADS;
constexpr int BEAM_THREADS = ORE_BEAM_THREADS;
constexpr int STAGE_A_THREADS = ORE_STAGE_A_THREADS;
// primary tile kernel: rows per 32-wide pixel tile and CTAs/SM it is bounded for
#ifndef ORE_TILE_P
#define ORE_TILE_P 8
#endif
#ifndef ORE_TILE_MIN_CTAS
#define ORE_TILE_MIN_CTAS 2
#endif
constexpr int TILE_P = ORE_TILE_P;
constexpr int SPHERE_PAD = 16;  // device sphere arrays are padded to a multiple of this

// conservative filter margins (see DESIGN.md "Filter soundness")
#define ORE_KAPPA_SHADOW 3.814697265625e-06f /* 2^-18 */
#define ORE_KAPPA_PRIMARY 7.62939453125e-06  /* 2^-17 */
#define ORE_BIG 3.0e38f

enum CounterSlot {
    CNT_HITS = 0,          // hit-list length
    CNT_SHADOW_CURSOR = 1, // dynamic batch cursor of the shadow kernel
    CNT_EXACT_PRIMARY = 2,
    CNT_EXACT_SHADOW = 3,
    CNT_SHADOW_TESTS_REF = 4,
    CNT_COUNT_CURSOR = 5,
    CNT_BEAM_L1 = 6,   // spheres passing the warp-level beam test (summed over warps and light passes)
    CNT_BEAM_L2 = 7,   // (pixel, sphere) pairs passing the per-pixel cone test
    CNT_STAGE_A0 = 8,                          // per-chunk block cursors of shade_setup_kernel
    CNT_STAGE_B0 = CNT_STAGE_A0 + 32,          // per-chunk block cursors of the staged shadow_beam_kernel
    CNT_SLOTS = CNT_STAGE_B0 + 32
};
constexpr int MAX_STAGE_CHUNKS = 32;

// Staging buffer between shade_setup_kernel and the staged shadow_beam_kernel (DESIGN.md "Two-stage shadow pass"):
// blocks of 32 hit-list items, value-major inside a block (value v of lane i at (block * nv + v) * 32 + i, so every
// access is one coalesced 128-byte line).  Values: 0-2 start, 3-5 texel r,g,b, then per light 30 direction
// components + a = dot(normal, toL).
struct StageArgs {
    float* buf;
    uint32_t first_block;  // hit-list block (32 items) held by stage block 0 of this chunk
    uint32_t cap_blocks;   // stage capacity in blocks
    int chunk;             // cursor slot
    int nv;                // values per item = 6 + 31 * n_lights
};

struct LightP {
    float px, py, pz, size, r, g, b;
};

struct FrameParams {
    int W, H, y0, y_step, n_rows;
    int y_block;     // rows come in blocks of y_block consecutive image rows, block starts y_step apart
    int pitch;       // output row pitch in pixels (pixels[] only; hit records stay packed)
    int out_global;  // 1: pixel rows are stored at their image position relative to y0, 0: packed
    int n_spheres, n_spheres_pad, n_lights;
    uint32_t flags;
    float aspect, ez, fz;  // ez = -1/aspect (kernel.cu:1629), fz = 0 - ez
    float Ox, Oy, Oz;      // add(eyePos, cam.Org), kernel.cu:1631
    float cp, sp, cy, sy;  // cosf/sinf of pitchRad / yawRad (kernel.cu:249-255), host libm
    int chunk, stages, n_chunks, resident;  // sphere tile pipeline
    const float* dx_tab;
    const float* dy_tab;
    const float4* sph_exact;  // cx,cy,cz,radius member
    float4* sph_prim;         // primary filter coefficients a',b',c',0 (per frame)
    float4* sph_cone;         // primary tile-cone record Mx,My,Mz,W in the camera frame (per frame)
    float tile_ca, tile_sa;   // cos/sin of the largest pixel-tile half-angle (+ margins), host-computed
    float px_delta;           // image-plane pixel pitch 2*aspect/width
    const float4* sph_shad;   // cx,cy,cz,R' = effective radius rounded up: R'^2 >= (1+k)*radius^2 (shadow filter)
    // the same records in Morton order of the centres, in clusters of 32 with one bounding sphere per cluster
    // (shadow sweep of the beam kernel: any-hit is order-free); sph_xsort = the exact records in that order
    const float4* sph_sort;
    const float4* sph_xsort;
    const float4* clu_sph;    // cx,cy,cz,radius of cluster j = spheres [32j, 32j+32) of sph_sort; radius >= 1e18: always a candidate
    int n_clusters, beam_resident;
    const float *tex_r, *tex_g, *tex_b;
    int tex_w, tex_h;
    const float *sky_r, *sky_g, *sky_b;
    int sky_w, sky_h;
    float sky_radius;  // skybox sphere member = size*size (kernel.cu:287,1122)
    int32_t* hit_id;
    float* hit_t;
    uint32_t* hit_list;
    unsigned long long* counters;
    uint32_t* pixels;
    // "next" primitives (kernel.cu:360-509): cube i = 3 float4 {bounds[0], bounds[1], orgin}, plane i = 2 float4
    // {orgin, normal}; hit ids continue after the spheres: cube i -> n_spheres + i, plane i -> n_spheres + n_cubes + i
    const float4* cubes;
    const float4* planes;
    int n_cubes, n_planes;
    // triangle mesh + flat BVH as the reference's `mesh` holds it (kernel.cu:559-1017): triangle i = 27 floats;
    // leaf box j = 2 float4 {bounds[0], bounds[1]} with triangles box_indices[box_offsets[j] .. box_offsets[j+1]);
    // hit id of triangle i = n_spheres + n_cubes + n_planes + i
    const float* tris;
    const float4* boxes;
    const float4* box_sph;    // bounding sphere of each leaf box (centre, radius incl. margin) for the cone filters
    float4* box_cone;         // per-frame tile-cone record of each leaf box (camera frame), like sph_cone
    const int* box_offsets;
    const int* box_indices;
    int n_tris, n_boxes, mesh_has_normals;
    // tools only (env ORE_DEBUG_BLOCK_CYCLES at ore_create): SM clocks spent on every hit-list block, [2][dbg_cap]
    // (0: shade_setup_kernel, 1: staged shadow_beam_kernel); null on every product path
    uint32_t* dbg_cycles;
    uint32_t dbg_cap;
    LightP lights[MAX_LIGHTS];
};

// cosf/ORE_SINF((float)j/10*2.f*3.1415f), kernel.cu:1454,1462-1463 - ten frame-independent
// values, evaluated once on the host (same libm as the oracle) at context creation
__constant__ float c_cos_phi[10];
__constant__ float c_sin_phi[10];
// b after k unshadowed samples: `b += 0.1` is float += double (kernel.cu:1537-1539), so b depends only on
// how many of the 10 samples were unshadowed; the 11 values are produced on the host with that arithmetic
__constant__ float c_b_of_k[11];

// ------------------------------------------------------------------------------------
// primary ray of pixel (x, row k): kernel.cu:1624-1631 with dx/dy from the tables
// ------------------------------------------------------------------------------------
__device__ __forceinline__ v3 primary_dir(const FrameParams& prm, float dx, float dy) {
    v3 v = mk(dx - 0.f, dy - 0.f, 0.f - prm.ez);  // sub(dir, eyePos)
    v3 n = ref_normalise(v);
    // camera::rotateDir, kernel.cu:252-255
    float y = n.y * prm.cp - n.z * prm.sp;
    float z = n.y * prm.sp + n.z * prm.cp;
    float x = n.x * prm.cy + z * prm.sy;
    z = -n.x * prm.sy + z * prm.cy;
    return mk(x, y, z);
}

__device__ __forceinline__ int clamp_index(int idx, int n) { return idx < 0 ? 0 : (idx >= n ? n - 1 : idx); }

// image row (relative to y0) of rendered row k, and where its pixels go
__device__ __forceinline__ int image_row_rel(const FrameParams& prm, int k) {
    return (k / prm.y_block) * prm.y_step + (k % prm.y_block);
}
__device__ __forceinline__ size_t out_index(const FrameParams& prm, int k, int x) {
    return (size_t)(prm.out_global ? image_row_rel(prm, k) : k) * prm.pitch + x;
}

// ---- out-of-line helpers: ONE copy of each cold or bulky sequence keeps the kernels' code small enough
// ---- for the instruction caches (measured: -27 % shadow-kernel time when the light set-up stopped being
// ---- inlined three times)
struct SkyArgs {
    const float *r, *g, *b;
    int w, h;
    float radius;
};
// skybox::getFColor + rgbToInt, kernel.cu:1146-1166,1688
__device__ __noinline__ uint32_t sky_pixel(const SkyArgs sk, float Ox, float Oy, float Oz, float Dx, float Dy, float Dz) {
    const v3 O = mk(Ox, Oy, Oz), D = mk(Dx, Dy, Dz);
    float t;
    ref_intersect(O, D, 0.f, 0.f, 0.f, sk.radius, t);
    v3 hp = ref_add(O, ref_scale(D, t));
    v3 n = ref_sub(hp, mk(0.f, 0.f, 0.f));
    ref_normalise(n);
    int sx = (int)((1.f + ORE_ATAN2F(n.z, n.x) / 3.1415f) * 0.5f * (float)sk.w);
    int sy = (int)(ORE_ACOSF(n.y) / 3.1415f * (float)sk.h);
    int index = clamp_index(sy * sk.w + sx, sk.w * sk.h);
    float r = __ldg(&sk.r[index]), g = __ldg(&sk.g[index]), b = __ldg(&sk.b[index]);
    return ref_rgb_to_int((int)(r * 254.f), (int)(g * 254.f), (int)(b * 254.f));
}
// exact sphere test out of line: returns the hit flag, t through the pointer
struct DirArgs {
    float ez, cp, sp, cy, sy;
};
__device__ __noinline__ v3 primary_dir_call(const DirArgs a, float dx, float dy) {
    v3 v = mk(dx - 0.f, dy - 0.f, 0.f - a.ez);
    v3 n = ref_normalise(v);
    float y = n.y * a.cp - n.z * a.sp;
    float z = n.y * a.sp + n.z * a.cp;
    float x = n.x * a.cy + z * a.sy;
    z = -n.x * a.sy + z * a.cy;
    return mk(x, y, z);
}

// ---- cubes and planes: small counts, exact tests, one out-of-line copy ----------------
// castRay's cube loop then plane loop (kernel.cu:1344-1372), continuing the strict '<' search after the spheres
__device__ __noinline__ void nearest_cube_plane(const float4* __restrict__ cubes, int nc, const float4* __restrict__ planes,
                                                int np, int id_base, float Ox, float Oy, float Oz, float Dx, float Dy,
                                                float Dz, float* best_t, int* best_id) {
    const v3 O = mk(Ox, Oy, Oz), D = mk(Dx, Dy, Dz);
    float nt = *best_t;
    int id = *best_id;
    for (int i = 0; i < nc; i++) {
        const float4 b0 = __ldg(&cubes[3 * i]), b1 = __ldg(&cubes[3 * i + 1]);
        float t;
        if (ref_cube_intersect(O, D, mk(b0.x, b0.y, b0.z), mk(b1.x, b1.y, b1.z), t)) {
            if (t < nt) {
                nt = t;
                id = id_base + i;
            }
        }
    }
    for (int i = 0; i < np; i++) {
        const float4 po = __ldg(&planes[2 * i]), no = __ldg(&planes[2 * i + 1]);
        float t;
        if (ref_plane_intersect(O, D, mk(po.x, po.y, po.z), mk(no.x, no.y, no.z), t)) {
            if (t < nt) {
                nt = t;
                id = id_base + nc + i;
            }
        }
    }
    *best_t = nt;
    *best_id = id;
}
// castLightRay's plane loop then cube loop (kernel.cu:1512-1536) for the live rays `live` (bits 0..9) of ONE light:
// any hit blocks.  A cube whose bounding sphere (cubes[3*i+2].xyz = orgin, .w = radius incl. margin) the light's cone
// cannot touch is skipped.  Returns the rays found blocked.
__device__ __noinline__ uint32_t cubes_planes_block_light(const float4* __restrict__ cubes, int nc,
                                                          const float4* __restrict__ planes, int np, float Ox, float Oy,
                                                          float Oz, float Ax, float Ay, float Az, float ca, float sa,
                                                          bool use_cone, const float* __restrict__ dirs, uint32_t live) {
    const v3 O = mk(Ox, Oy, Oz);
    uint32_t hit = 0;
    float t;
    for (int i = 0; i < np && live; i++) {
        const float4 po = __ldg(&planes[2 * i]), no = __ldg(&planes[2 * i + 1]);
        uint32_t todo = live;
        while (todo) {
            const int r = __ffs(todo) - 1;
            todo &= todo - 1;
            if (ref_plane_intersect(O, mk(dirs[r * 3], dirs[r * 3 + 1], dirs[r * 3 + 2]), mk(po.x, po.y, po.z),
                                    mk(no.x, no.y, no.z), t)) {
                hit |= 1u << r;
                live &= ~(1u << r);
            }
        }
    }
    for (int i = 0; i < nc && live; i++) {
        const float4 q = __ldg(&cubes[3 * i + 2]);
        if (use_cone) {
            const float lx = Ox - q.x, ly = Oy - q.y, lz = Oz - q.z;
            const float LL = fmaf(lz, lz, fmaf(ly, ly, lx * lx));
            const float Cm = fmaf(LL, 1.0f - ORE_KAPPA_SHADOW, -(q.w * q.w));
            if (Cm > 1e-20f) {
                const float sv = Cm * rsqrt_approx(Cm);
                const float T = fmaf(ca, sv, -(sa * q.w));
                if (!(fmaf(Ax, lx, fmaf(Ay, ly, fmaf(Az, lz, T))) < 0.f)) continue;  // cone misses the cube
            }
        }
        const float4 b0 = __ldg(&cubes[3 * i]), b1 = __ldg(&cubes[3 * i + 1]);
        uint32_t todo = live;
        while (todo) {
            const int r = __ffs(todo) - 1;
            todo &= todo - 1;
            if (ref_cube_intersect(O, mk(dirs[r * 3], dirs[r * 3 + 1], dirs[r * 3 + 2]), mk(b0.x, b0.y, b0.z),
                                   mk(b1.x, b1.y, b1.z), t)) {
                hit |= 1u << r;
                live &= ~(1u << r);
            }
        }
    }
    return hit;
}

// ---- triangle mesh: linear scan over the leaf boxes, exact tests, one out-of-line copy ----
struct MeshArgs {
    const float* tris;
    const float4* boxes;
    const int* offsets;
    const int* indices;
    int n_boxes;
};
// castRay's triangle loop (kernel.cu:1293-1328) for ONE leaf box: exact slab test, then exact triangle tests in leaf
// order, continuing the strict '<' search.  The caller visits the leaves in ascending order (all of them in the
// reference; here those whose bounding sphere the tile cone can touch - a leaf the ray misses contributes nothing).
__device__ __noinline__ void nearest_in_leaf(const MeshArgs m, int j, int id_base, float Ox, float Oy, float Oz, float Dx,
                                             float Dy, float Dz, float* best_t, int* best_id) {
    const v3 O = mk(Ox, Oy, Oz), D = mk(Dx, Dy, Dz);
    const float4 b0 = __ldg(&m.boxes[2 * j]), b1 = __ldg(&m.boxes[2 * j + 1]);
    float temp;
    if (!ref_cube_intersect(O, D, mk(b0.x, b0.y, b0.z), mk(b1.x, b1.y, b1.z), temp)) return;
    float nt = *best_t;
    int id = *best_id;
    const int k1 = __ldg(&m.offsets[j + 1]);
    for (int k = __ldg(&m.offsets[j]); k < k1; k++) {
        const int ti = __ldg(&m.indices[k]);
        float t, u, v;
        if (ref_tri_intersect(O, D, m.tris + 27 * (size_t)ti, t, u, v)) {
            if (t < nt) {
                nt = t;
                id = id_base + ti;
            }
        }
    }
    *best_t = nt;
    *best_id = id;
}
// castLightRay's triangle part (kernel.cu:1475-1497) for the live rays `live` (bits 0..9) of ONE light: a leaf is
// skipped when the light's cone (A, ca, sa; use_cone) cannot touch its bounding sphere; otherwise every live ray
// runs the exact slab test and the exact triangle tests.  Returns the rays found blocked.
__device__ __noinline__ uint32_t mesh_blocks_light(const MeshArgs m, const float4* __restrict__ box_sph, float Ox, float Oy,
                                                   float Oz, float Ax, float Ay, float Az, float ca, float sa, bool use_cone,
                                                   const float* __restrict__ dirs /* [10][3] */, uint32_t live) {
    const v3 O = mk(Ox, Oy, Oz);
    uint32_t hit = 0;
    for (int j = 0; j < m.n_boxes && live; j++) {
        if (use_cone) {
            const float4 q = __ldg(&box_sph[j]);
            const float lx = Ox - q.x, ly = Oy - q.y, lz = Oz - q.z;
            const float LL = fmaf(lz, lz, fmaf(ly, ly, lx * lx));
            const float Cm = fmaf(LL, 1.0f - ORE_KAPPA_SHADOW, -(q.w * q.w));
            if (Cm > 1e-20f) {
                const float sv = Cm * rsqrt_approx(Cm);
                const float T = fmaf(ca, sv, -(sa * q.w));
                if (!(fmaf(Ax, lx, fmaf(Ay, ly, fmaf(Az, lz, T))) < 0.f)) continue;  // cone misses the leaf
            }
        }
        const float4 b0 = __ldg(&m.boxes[2 * j]), b1 = __ldg(&m.boxes[2 * j + 1]);
        const int k0 = __ldg(&m.offsets[j]), k1 = __ldg(&m.offsets[j + 1]);
        uint32_t todo = live;
        while (todo) {
            const int r = __ffs(todo) - 1;
            todo &= todo - 1;
            const v3 D = mk(dirs[r * 3], dirs[r * 3 + 1], dirs[r * 3 + 2]);
            float temp;
            if (!ref_cube_intersect(O, D, mk(b0.x, b0.y, b0.z), mk(b1.x, b1.y, b1.z), temp)) continue;
            for (int k = k0; k < k1; k++) {
                float t, u, v;
                if (ref_tri_intersect(O, D, m.tris + 27 * (size_t)__ldg(&m.indices[k]), t, u, v)) {
                    hit |= 1u << r;
                    live &= ~(1u << r);
                    break;
                }
            }
        }
    }
    return hit;
}
// triangle hit attributes (kernel.cu:1380-1394): u,v come from re-running the winning test (deterministic)
__device__ __noinline__ void triangle_attributes(const float* __restrict__ tri, int has_normals, float Ox, float Oy, float Oz,
                                                 float Dx, float Dy, float Dz, float nt, v3* normal, v3* new_org, float* tx,
                                                 float* ty) {
    const v3 O = mk(Ox, Oy, Oz), D = mk(Dx, Dy, Dz);
    float t, nu = 0.f, nv = 0.f;
    ref_tri_intersect(O, D, tri, t, nu, nv);
    v3 n;
    if (has_normals) {
        const v3 vn0 = mk(tri[12], tri[13], tri[14]), vn1 = mk(tri[15], tri[16], tri[17]), vn2 = mk(tri[18], tri[19], tri[20]);
        n = ref_add(ref_add(ref_scale(vn0, (1 - nu - nv)), ref_scale(vn1, nu)), ref_scale(vn2, nv));
        ref_normalise(n);
    } else {
        n = mk(tri[9], tri[10], tri[11]);
    }
    *tx = ((1 - nu - nv) * tri[21]) + (nu * tri[23]) + (nv * tri[25]);
    *ty = ((1 - nu - nv) * tri[22]) + (nu * tri[24]) + (nv * tri[26]);
    *normal = n;
    *new_org = ref_add(n, ref_add(O, ref_scale(D, nt)));  // the unit normal is ADDED to the hit point (kernel.cu:1393)
}

// ------------------------------------------------------------------------------------
// prep_frame_kernel
// ------------------------------------------------------------------------------------
__global__ void prep_frame_kernel(const FrameParams prm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < CNT_SLOTS) prm.counters[i] = 0ull;
    if (i < prm.W) {
        // kernel.cu:1624  float dx = aspect * (2 * (x + 0.5) / (float)width) - 1;   (double)
        double v = (double)prm.aspect * (2 * (i + 0.5) / (double)(float)prm.W) - 1;
        const_cast<float*>(prm.dx_tab)[i] = (float)v;
    }
    if (i < prm.n_rows) {
        // kernel.cu:1625  float dy = aspect * (2 * (y + 0.5) / (float)height)*((float)height/width) - 1;
        const int y = prm.y0 + image_row_rel(prm, i);
        float hw = (float)prm.H / (float)prm.W;
        double v = (double)prm.aspect * (2 * (y + 0.5) / (double)(float)prm.H) * (double)hw - 1;
        const_cast<float*>(prm.dy_tab)[i] = (float)v;
    }
    if (i < prm.n_spheres_pad) {
        float4 out = make_float4(0.f, 0.f, ORE_BIG, 0.f);  // padding: never a candidate
        float4 cone = make_float4(0.f, 0.f, 0.f, ORE_BIG);
        if (i < prm.n_spheres) {
            const float4 s = prm.sph_exact[i];
            // L exactly as the reference forms it (float), then the filter works in double
            const double Lx = (double)(prm.Ox - s.x), Ly = (double)(prm.Oy - s.y), Lz = (double)(prm.Oz - s.z);
            const double LL = Lx * Lx + Ly * Ly + Lz * Lz;
            const double r4 = (double)(s.w * s.w);
            const double Cm = LL * (1.0 - ORE_KAPPA_PRIMARY) - r4 * (1.0 + ORE_KAPPA_PRIMARY);
            if (!(Cm > 1e-9 * LL) || !(Cm > 1e-30)) {
                out = make_float4(0.f, 0.f, -ORE_BIG, 0.f);  // origin in/near the sphere: always exact
                cone = make_float4(0.f, 0.f, 0.f, -ORE_BIG);
            } else {
                const double sv = sqrt(Cm);
                const double cp = prm.cp, sp = prm.sp, cy = prm.cy, sy = prm.sy;
                const double Mx = cy * Lx - sy * Lz;
                const double My = sp * sy * Lx + cp * Ly + sp * cy * Lz;
                const double Mz = cp * sy * Lx - sp * Ly + cp * cy * Lz;
                out = make_float4((float)(Mx / sv), (float)(My / sv), (float)((double)prm.fz * Mz / sv), 0.f);
                // tile cone (camera frame): a pixel tile whose directions lie within `a` of its axis A can
                // only contain a hit if  A.M + cos(a) sv - sin(a) sqrt(LL - sv^2) <= 0   (DESIGN.md "Cone filter")
                const double Rpp = sqrt(LL - Cm);
                const double Wd = (double)prm.tile_ca * sv - (double)prm.tile_sa * Rpp;
                cone = make_float4((float)Mx, (float)My, (float)Mz, (float)(Wd - 4e-6 * sqrt(LL) - 1e-30));
            }
        }
        prm.sph_prim[i] = out;
        prm.sph_cone[i] = cone;
    }
    if (i < prm.n_boxes) {
        // tile-cone record of leaf box i from its bounding sphere (same formula as for the spheres)
        const float4 q = prm.box_sph[i];
        const double Lx = (double)prm.Ox - q.x, Ly = (double)prm.Oy - q.y, Lz = (double)prm.Oz - q.z;
        const double LL = Lx * Lx + Ly * Ly + Lz * Lz;
        const double Cm = LL * (1.0 - ORE_KAPPA_PRIMARY) - (double)q.w * q.w * (1.0 + ORE_KAPPA_PRIMARY);
        float4 rec = make_float4(0.f, 0.f, 0.f, -ORE_BIG);  // eye in/near the leaf's sphere: always a candidate
        if (Cm > 1e-9 * LL && Cm > 1e-30) {
            const double sv = sqrt(Cm), Rpp = sqrt(LL - Cm);
            const double cp = prm.cp, sp = prm.sp, cy = prm.cy, sy = prm.sy;
            const double Mx = cy * Lx - sy * Lz;
            const double My = sp * sy * Lx + cp * Ly + sp * cy * Lz;
            const double Mz = cp * sy * Lx - sp * Ly + cp * cy * Lz;
            const double Wd = (double)prm.tile_ca * sv - (double)prm.tile_sa * Rpp;
            rec = make_float4((float)Mx, (float)My, (float)Mz, (float)(Wd - 4e-6 * sqrt(LL) - 1e-30));
        }
        prm.box_cone[i] = rec;
    }
}

// ------------------------------------------------------------------------------------
// sphere tile pipeline: chunks of `chunk` float4 records, `stages` shared-memory slots,
// filled by TMA bulk copies that complete on an mbarrier.  Resident mode (whole array in
// one slot) loads once per CTA; streaming mode re-streams the chunks for every batch.
// ------------------------------------------------------------------------------------
struct TilePipe {
    float4* slots;
    uint64_t* bars;
    const float4* src;
    int chunk, stages, n_chunks, n_pad;
    uint32_t phase_bits;
    bool resident, loaded;

    __device__ __forceinline__ int count(int c) const {
        int rem = n_pad - c * chunk;
        return rem < chunk ? rem : chunk;
    }
    __device__ __forceinline__ void issue(int c) {  // one thread
        const int st = c % stages;
        const uint32_t bytes = (uint32_t)count(c) * 16u;
        mbar_expect_tx(&bars[st], bytes);
        tma_bulk_g2s(slots + (size_t)st * chunk, src + (size_t)c * chunk, bytes, &bars[st]);
    }
    // start of a round (one pass over all chunks)
    __device__ __forceinline__ void begin_round() {
        if (resident && loaded) return;
        if (threadIdx.x == 0) {
            const int n0 = n_chunks < stages ? n_chunks : stages;
            for (int c = 0; c < n0; c++) issue(c);
        }
    }
    __device__ __forceinline__ const float4* acquire(int c) {
        const int st = c % stages;
        if (!(resident && loaded)) {
            mbar_wait(&bars[st], (phase_bits >> st) & 1u);
            phase_bits ^= (1u << st);
        }
        return slots + (size_t)st * chunk;
    }
    // all threads of the CTA must call this after consuming chunk c
    __device__ __forceinline__ void release(int c) {
        if (resident) {
            loaded = true;
            return;
        }
        __syncthreads();
        if (threadIdx.x == 0 && c + stages < n_chunks) issue(c + stages);
    }
};

__device__ __forceinline__ void pipe_init(TilePipe& tp, const FrameParams& prm, const float4* src, float4* slots,
                                          uint64_t* bars) {
    tp.slots = slots;
    tp.bars = bars;
    tp.src = src;
    tp.chunk = prm.chunk;
    tp.stages = prm.stages;
    tp.n_chunks = prm.n_chunks;
    tp.n_pad = prm.n_spheres_pad;
    tp.phase_bits = 0;
    tp.resident = prm.resident != 0;
    tp.loaded = false;
    if (threadIdx.x == 0) {
        for (int s = 0; s < prm.stages; s++) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------
// primary_kernel
// ------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(CTA_THREADS) primary_kernel(const FrameParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[MAX_STAGES];
    __shared__ uint32_t warp_tot[CTA_WARPS];
    __shared__ uint32_t cta_base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    TilePipe tp;
    pipe_init(tp, prm, prm.sph_prim, reinterpret_cast<float4*>(smem_raw), bars);

    const int strip_px = 32 * P;
    const int strips_per_row = (prm.W + strip_px - 1) / strip_px;
    const int total_strips = prm.n_rows * strips_per_row;
    const int n_batches = (total_strips + CTA_WARPS - 1) / CTA_WARPS;
    const v3 O = mk(prm.Ox, prm.Oy, prm.Oz);
    const bool exhaustive = (prm.flags & 1u) != 0;
    unsigned long long n_exact = 0;

    for (int batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
        const int strip = batch * CTA_WARPS + warp;
        const bool strip_ok = strip < total_strips;
        const int k = strip_ok ? strip / strips_per_row : 0;
        const int sx = strip_ok ? strip % strips_per_row : 0;
        const float dy = prm.dy_tab[k];

        float dxp[P], negn[P], best_t[P];
        int best_id[P];
        v3 D[P];
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int x = sx * strip_px + p * 32 + lane;
            const bool ok = strip_ok && x < prm.W;
            dxp[p] = ok ? prm.dx_tab[x] : 0.f;
            D[p] = primary_dir(prm, dxp[p], dy);
            // filter threshold: candidate iff g' <= -|v| (shrunk a little: more candidates)
            const float nv = sqrtf(fmaf(dxp[p], dxp[p], fmaf(dy, dy, prm.fz * prm.fz)));
            negn[p] = ok ? (exhaustive ? INFINITY : -nv * 0.99999905f) : -INFINITY;
            best_t[p] = INFINITY;
            best_id[p] = -1;
        }

        tp.begin_round();
        for (int c = 0; c < tp.n_chunks; c++) {
            const float4* tile = tp.acquire(c);
            const int cnt = tp.count(c);
            const int base = c * tp.chunk;
#pragma unroll 1
            for (int s = 0; s < cnt; s += 4) {
                const float4 q0 = tile[s], q1 = tile[s + 1], q2 = tile[s + 2], q3 = tile[s + 3];
                const float e0 = fmaf(dy, q0.y, q0.z), e1 = fmaf(dy, q1.y, q1.z);
                const float e2 = fmaf(dy, q2.y, q2.z), e3 = fmaf(dy, q3.y, q3.z);
                bool any = false;
#pragma unroll
                for (int p = 0; p < P; p++) {
                    any |= (fmaf(dxp[p], q0.x, e0) <= negn[p]);
                    any |= (fmaf(dxp[p], q1.x, e1) <= negn[p]);
                    any |= (fmaf(dxp[p], q2.x, e2) <= negn[p]);
                    any |= (fmaf(dxp[p], q3.x, e3) <= negn[p]);
                }
                if (any) {
                    // exact re-adjudication in ascending sphere index (strict '<' keeps the
                    // lowest index on ties, kernel.cu:1335)
#pragma unroll 1
                    for (int u = 0; u < 4; u++) {
                        const int idx = base + s + u;
                        if (idx >= prm.n_spheres) break;
                        const float4 q = tile[s + u];
                        const float e = fmaf(dy, q.y, q.z);
                        float4 ex = make_float4(0.f, 0.f, 0.f, 0.f);
                        bool have = false;
#pragma unroll
                        for (int p = 0; p < P; p++) {
                            if (fmaf(dxp[p], q.x, e) <= negn[p]) {
                                if (!have) {
                                    ex = __ldg(&prm.sph_exact[idx]);
                                    have = true;
                                }
                                float t;
                                n_exact++;
                                if (ref_intersect(O, D[p], ex.x, ex.y, ex.z, ex.w, t)) {
                                    if (t < best_t[p]) {
                                        best_t[p] = t;
                                        best_id[p] = idx;
                                    }
                                }
                            }
                        }
                    }
                }
            }
            tp.release(c);
        }

        // ---- epilogue: records, sky for misses, hit-list compaction ----
        uint32_t warp_hits = 0;   // warp-uniform
        uint32_t my_off[P];
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int x = sx * strip_px + p * 32 + lane;
            const bool ok = strip_ok && x < prm.W;
            const bool hit = ok && best_id[p] >= 0;
            if (ok) {
                const size_t o = (size_t)k * prm.W + x;
                prm.hit_id[o] = best_id[p];
                prm.hit_t[o] = best_t[p];
                if (!hit) {
                    // skybox::getFColor, kernel.cu:1146-1166
                    float t;
                    ref_intersect(O, D[p], 0.f, 0.f, 0.f, prm.sky_radius, t);
                    v3 hp = ref_add(O, ref_scale(D[p], t));
                    v3 n = ref_sub(hp, mk(0.f, 0.f, 0.f));
                    ref_normalise(n);
                    int tx = (int)((1.f + ORE_ATAN2F(n.z, n.x) / 3.1415f) * 0.5f * (float)prm.sky_w);
                    int ty = (int)(ORE_ACOSF(n.y) / 3.1415f * (float)prm.sky_h);
                    int index = clamp_index(ty * prm.sky_w + tx, prm.sky_w * prm.sky_h);
                    float r = __ldg(&prm.sky_r[index]), g = __ldg(&prm.sky_g[index]), b = __ldg(&prm.sky_b[index]);
                    prm.pixels[out_index(prm, k, x)] = ref_rgb_to_int((int)(r * 254.f), (int)(g * 254.f), (int)(b * 254.f));
                }
            }
            // hit-list order inside a warp: p-major, then lane => 32 neighbouring pixels stay together
            const uint32_t bal = __ballot_sync(0xffffffffu, hit);
            my_off[p] = warp_hits + __popc(bal & ((1u << lane) - 1u));
            warp_hits += __popc(bal);
        }
        // CTA-level compaction: one global atomic per batch
        if (lane == 0) warp_tot[warp] = warp_hits;
        __syncthreads();
        if (tid == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < CTA_WARPS; w++) {
                uint32_t v = warp_tot[w];
                warp_tot[w] = tot;
                tot += v;
            }
            cta_base = tot ? (uint32_t)atomicAdd(&prm.counters[CNT_HITS], (unsigned long long)tot) : 0u;
        }
        __syncthreads();
        const uint32_t wbase = cta_base + warp_tot[warp];
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int x = sx * strip_px + p * 32 + lane;
            if (strip_ok && x < prm.W && best_id[p] >= 0) prm.hit_list[wbase + my_off[p]] = (uint32_t)((size_t)k * prm.W + x);
        }
        __syncthreads();  // warp_tot / cta_base reuse
    }
    if (n_exact) atomicAdd(&prm.counters[CNT_EXACT_PRIMARY], n_exact);
}

// ------------------------------------------------------------------------------------
// primary_tile_kernel (default primary path)
//
// One warp = one tile of 32 x P pixels.  The 32 lanes first test 32 DIFFERENT spheres
// against the tile's bounding cone (one 3-FMA test per lane, shared-memory records staged
// by TMA), ballot, and only the surviving spheres get the per-pixel filter + exact
// sequence.  Every (tile, sphere) pair is visited in index order; candidates keep
// ascending order, so the strict '<' tie rule (kernel.cu:1335) is preserved.
// ------------------------------------------------------------------------------------
template <int P, bool EXH>
__global__ void __launch_bounds__(CTA_THREADS, ORE_TILE_MIN_CTAS) primary_tile_kernel(const FrameParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[MAX_STAGES];
    __shared__ uint32_t warp_tot[CTA_WARPS];
    __shared__ uint32_t cta_base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    TilePipe tp;
    pipe_init(tp, prm, prm.sph_cone, reinterpret_cast<float4*>(smem_raw), bars);

    const int tiles_x = (prm.W + 31) / 32;
    const int tiles_y = (prm.n_rows + P - 1) / P;
    const int total_tiles = tiles_x * tiles_y;
    const int n_batches = (total_tiles + CTA_WARPS - 1) / CTA_WARPS;
    const v3 O = mk(prm.Ox, prm.Oy, prm.Oz);
    const DirArgs da = {prm.ez, prm.cp, prm.sp, prm.cy, prm.sy};
    const SkyArgs sk = {prm.sky_r, prm.sky_g, prm.sky_b, prm.sky_w, prm.sky_h, prm.sky_radius};
    unsigned long long n_exact = 0;

    for (int batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
        const int tile_id = batch * CTA_WARPS + warp;
        const bool tile_ok = tile_id < total_tiles;
        const int ty = tile_ok ? tile_id / tiles_x : 0;
        const int tx = tile_ok ? tile_id % tiles_x : 0;
        const int x = tx * 32 + lane;
        const bool x_ok = tile_ok && x < prm.W;
        const float dx = x_ok ? prm.dx_tab[x] : 0.f;

        float dyp[P], negn[P], best_t[P];
        int best_id[P];
        v3 D[P];
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int k = ty * P + p;
            const bool ok = x_ok && k < prm.n_rows;
            dyp[p] = prm.dy_tab[k < prm.n_rows ? k : prm.n_rows - 1];
            D[p] = primary_dir_call(da, dx, dyp[p]);
            const float nv = sqrtf(fmaf(dx, dx, fmaf(dyp[p], dyp[p], prm.fz * prm.fz)));
            negn[p] = ok ? (EXH ? INFINITY : -nv * 0.99999905f) : -INFINITY;
            best_t[p] = INFINITY;
            best_id[p] = -1;
        }
        // tile axis in the camera frame: nominal tile centre on the image plane (warp-uniform)
        float ax, ay, az;
        {
            const int k0 = ty * P;
            const float cx = prm.dx_tab[min(tx * 32, prm.W - 1)] + 15.5f * prm.px_delta;
            const float cy = 0.5f * (prm.dy_tab[min(k0, prm.n_rows - 1)] + prm.dy_tab[min(k0 + P - 1, prm.n_rows - 1)]);
            const float inv = rsqrtf(fmaf(cx, cx, fmaf(cy, cy, prm.fz * prm.fz)));
            ax = cx * inv;
            ay = cy * inv;
            az = prm.fz * inv;
        }

        // ---- triangles first (kernel.cu:1293-1328): they seed the strict '<' search the spheres continue.
        //      Lane i tests leaf box s0+i (its bounding sphere) against the tile cone; surviving leaves, in
        //      ascending order, get the exact slab + triangle tests per pixel. ----
        if (prm.n_boxes) {
            const MeshArgs ma = {prm.tris, prm.boxes, prm.box_offsets, prm.box_indices, prm.n_boxes};
            const int id_base = prm.n_spheres + prm.n_cubes + prm.n_planes;
#pragma unroll 1
            for (int s0 = 0; s0 < prm.n_boxes; s0 += 32) {
                bool cand = false;
                if (s0 + lane < prm.n_boxes) {
                    const float4 rec = __ldg(&prm.box_cone[s0 + lane]);
                    cand = EXH || fmaf(ax, rec.x, fmaf(ay, rec.y, fmaf(az, rec.z, rec.w))) <= 0.f;
                }
                uint32_t mask = __ballot_sync(0xffffffffu, cand && tile_ok);
                while (mask) {
                    const int i = __ffs(mask) - 1;
                    mask &= mask - 1;
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        if (x_ok && ty * P + p < prm.n_rows)
                            nearest_in_leaf(ma, s0 + i, id_base, O.x, O.y, O.z, D[p].x, D[p].y, D[p].z, &best_t[p], &best_id[p]);
                    }
                }
            }
        }

        tp.begin_round();
        for (int c = 0; c < tp.n_chunks; c++) {
            const float4* tile = tp.acquire(c);
            const int cnt = tp.count(c);
            const int base = c * tp.chunk;
#pragma unroll 1
            for (int s = 0; s < cnt; s += 32) {
                // n_spheres_pad is a multiple of 16, chunks are multiples of 32 except possibly the last
                const bool in = s + lane < cnt;
                const float4 rec = in ? tile[s + lane] : make_float4(0.f, 0.f, 0.f, ORE_BIG);
                const float hA = fmaf(ax, rec.x, fmaf(ay, rec.y, fmaf(az, rec.z, rec.w)));
                const bool cand = in && (base + s + lane) < prm.n_spheres && (EXH || hA <= 0.f);
                uint32_t mask = __ballot_sync(0xffffffffu, cand && tile_ok);
                while (mask) {
                    const int i = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int idx = base + s + i;
                    const float4 q = __ldg(&prm.sph_prim[idx]);
                    const float e = fmaf(dx, q.x, q.z);
                    bool any = false;
#pragma unroll
                    for (int p = 0; p < P; p++) any |= (fmaf(dyp[p], q.y, e) <= negn[p]);
                    if (any) {
                        const float4 ex = __ldg(&prm.sph_exact[idx]);
#pragma unroll
                        for (int p = 0; p < P; p++) {
                            if (fmaf(dyp[p], q.y, e) <= negn[p]) {
                                float t;
                                n_exact++;
                                if (ref_intersect(O, D[p], ex.x, ex.y, ex.z, ex.w, t)) {
                                    if (t < best_t[p]) {
                                        best_t[p] = t;
                                        best_id[p] = idx;
                                    }
                                }
                            }
                        }
                    }
                }
            }
            tp.release(c);
        }

        // ---- cubes, then planes (kernel.cu:1344-1372): exact tests continuing the same strict '<' search ----
        if (prm.n_cubes | prm.n_planes) {
#pragma unroll
            for (int p = 0; p < P; p++) {
                if (x_ok && ty * P + p < prm.n_rows)
                    nearest_cube_plane(prm.cubes, prm.n_cubes, prm.planes, prm.n_planes, prm.n_spheres, O.x, O.y, O.z, D[p].x,
                                       D[p].y, D[p].z, &best_t[p], &best_id[p]);
            }
        }

        // ---- epilogue: records, sky for misses, hit-list compaction (row-major inside the tile) ----
        uint32_t warp_hits = 0;
        uint32_t my_off[P];
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int k = ty * P + p;
            const bool ok = x_ok && k < prm.n_rows;
            const bool hit = ok && best_id[p] >= 0;
            if (ok) {
                const size_t o = (size_t)k * prm.W + x;
                prm.hit_id[o] = best_id[p];
                prm.hit_t[o] = best_t[p];
                if (!hit) {
                    prm.pixels[out_index(prm, k, x)] = sky_pixel(sk, O.x, O.y, O.z, D[p].x, D[p].y, D[p].z);
                }
            }
        }
        // hit-list order inside the tile: grouped by hit sphere (ascending id), row-major inside a group, so the
        // 32 consecutive entries a shadow warp takes mostly lie on ONE sphere (tight beams)
        {
            uint32_t rem = 0;
#pragma unroll
            for (int p = 0; p < P; p++) {
                const int k = ty * P + p;
                my_off[p] = 0;
                if (x_ok && k < prm.n_rows && best_id[p] >= 0) rem |= 1u << p;
            }
            while (__any_sync(0xffffffffu, rem != 0)) {
                int cur = 0x7fffffff;
#pragma unroll
                for (int p = 0; p < P; p++)
                    if ((rem >> p) & 1u) cur = min(cur, best_id[p]);
                cur = __reduce_min_sync(0xffffffffu, cur);
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const bool m = ((rem >> p) & 1u) && best_id[p] == cur;
                    const uint32_t bal = __ballot_sync(0xffffffffu, m);
                    if (m) {
                        my_off[p] = warp_hits + __popc(bal & ((1u << lane) - 1u));
                        rem &= ~(1u << p);
                    }
                    warp_hits += __popc(bal);
                }
            }
        }
        if (lane == 0) warp_tot[warp] = warp_hits;
        __syncthreads();
        if (tid == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < CTA_WARPS; w++) {
                uint32_t v = warp_tot[w];
                warp_tot[w] = tot;
                tot += v;
            }
            cta_base = tot ? (uint32_t)atomicAdd(&prm.counters[CNT_HITS], (unsigned long long)tot) : 0u;
        }
        __syncthreads();
        const uint32_t wbase = cta_base + warp_tot[warp];
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int k = ty * P + p;
            if (x_ok && k < prm.n_rows && best_id[p] >= 0) prm.hit_list[wbase + my_off[p]] = (uint32_t)((size_t)k * prm.W + x);
        }
        // (the barriers also keep the CTA's warps in step, which keeps its code resident in the instruction cache:
        // per-warp reservations without them measured 14 % slower)
        __syncthreads();  // warp_tot / cta_base reuse
    }
    if (n_exact) atomicAdd(&prm.counters[CNT_EXACT_PRIMARY], n_exact);
}

// ------------------------------------------------------------------------------------
// shadow ray directions of one light: castLightRay set-up, kernel.cu:1438-1468.
// Writes 10 directions to dir[] and returns a = dot(normal, toL) with the toL left
// after the loop (kernel.cu:1541).  toL is re-normalised in place twice per sample.
// ------------------------------------------------------------------------------------
__device__ __noinline__ float light_directions(const LightP L, const v3 start, const v3 normal,
                                               float* __restrict__ dir /* [10][3] */) {
    const v3 lpos = mk(L.px, L.py, L.pz);
    v3 tmp = ref_sub(lpos, start);
    v3 toL = ref_normalise(tmp);
    const v3 up = mk(0.f, 1.f, 0.f), fwd = mk(0.f, 0.f, 1.f);
#pragma unroll 1
    for (int j = 0; j < 10; j++) {
        v3 P = ref_cross(toL, up);
        v3 e = ref_sub(ref_add(lpos, ref_scale(P, L.size)), start);
        v3 toEdge = ref_normalise(e);
        float angle = ORE_COSF((ref_dot(toL, toEdge)) * 2);
        float _z = (float)j / 10 * (1.0f - angle) + angle;
        float sq = sqrtf(1.f - _z * _z);
        float x = sq * c_cos_phi[j];
        float y = sq * c_sin_phi[j];
        v3 n1 = ref_normalise(toL);
        v3 ax = ref_cross(fwd, n1);
        v3 axis = ref_normalise(ax);
        v3 n2 = ref_normalise(toL);
        float nAngle = ORE_ACOSF(ref_dot(n2, fwd));
        v3 nd = ref_sub(lpos, ref_rotate_apply(nAngle, axis, mk(x, y, _z)));
        v3 nn = ref_normalise(nd);
        dir[j * 3 + 0] = nn.x;
        dir[j * 3 + 1] = nn.y;
        dir[j * 3 + 2] = nn.z;
    }
    return ref_dot(normal, toL);
}

// ------------------------------------------------------------------------------------
// light_directions_n: the same set-up for up to NL lights in lockstep (independent chains
// overlap), with exact reuse: every per-sample quantity except _z/x/y is a pure function of
// the toL seen at the top of the iteration, and that toL stops changing once the in-place
// re-normalisation reaches a fixed point (usually after one or two samples).  While toL
// repeats bit for bit, angle and the rotate() matrix are reused instead of recomputed.
// ------------------------------------------------------------------------------------
template <int NL>
__device__ __forceinline__ void light_directions_n(const FrameParams& prm, int l0, int n_act, const v3 start,
                                                   const v3 normal, float* __restrict__ dirs /* [NL][10][3] */,
                                                   float (&a_out)[NL]) {
    const v3 up = mk(0.f, 1.f, 0.f), fwd = mk(0.f, 0.f, 1.f);
    v3 lpos[NL], toL[NL], prev[NL];
    float lsize[NL], angle[NL];
    RotM M[NL];
#pragma unroll
    for (int l = 0; l < NL; l++) {
        if (l < n_act) {
            const LightP L = prm.lights[l0 + l];
            lpos[l] = mk(L.px, L.py, L.pz);
            lsize[l] = L.size;
            v3 tmp = ref_sub(lpos[l], start);
            toL[l] = ref_normalise(tmp);
        } else {
            lpos[l] = toL[l] = mk(0.f, 0.f, 0.f);
            lsize[l] = 0.f;
        }
        prev[l] = mk(__int_as_float(0x7fc00001), 0.f, 0.f);  // matches nothing
        angle[l] = 0.f;
        M[l] = RotM{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    }
#pragma unroll 1
    for (int j = 0; j < 10; j++) {
#pragma unroll
        for (int l = 0; l < NL; l++) {
            if (l < n_act) {
                if (!same_bits(toL[l], prev[l])) {
                    prev[l] = toL[l];
                    v3 P = ref_cross(toL[l], up);
                    v3 e = ref_sub(ref_add(lpos[l], ref_scale(P, lsize[l])), start);
                    v3 toEdge = ref_normalise(e);
                    angle[l] = ORE_COSF((ref_dot(toL[l], toEdge)) * 2);
                    v3 n1 = ref_normalise(toL[l]);
                    v3 ax = ref_cross(fwd, n1);
                    v3 axis = ref_normalise(ax);
                    v3 n2 = ref_normalise(toL[l]);
                    float nAngle = ORE_ACOSF(ref_dot(n2, fwd));
                    M[l] = ref_rotate_matrix(nAngle, axis);
                }
                const float _z = (float)j / 10 * (1.0f - angle[l]) + angle[l];
                const float sq = sqrtf(1.f - _z * _z);
                const float x = sq * c_cos_phi[j];
                const float y = sq * c_sin_phi[j];
                v3 nd = ref_sub(lpos[l], ref_matrix_apply(M[l], mk(x, y, _z)));
                v3 nn = ref_normalise(nd);
                float* d = dirs + (l * 10 + j) * 3;
                d[0] = nn.x;
                d[1] = nn.y;
                d[2] = nn.z;
            }
        }
    }
#pragma unroll
    for (int l = 0; l < NL; l++) a_out[l] = (l < n_act) ? ref_dot(normal, toL[l]) : 0.f;
}

// ------------------------------------------------------------------------------------
// light_directions_reuse: one light, one copy of the code (instruction-cache friendly), same exact reuse
// of angle / rotate() matrix while toL repeats bit for bit.  Returns a = dot(normal, toL_final).
// ------------------------------------------------------------------------------------
__device__ __noinline__ float light_directions_reuse(const LightP L, const v3 start, const v3 normal,
                                                     float* __restrict__ dir /* [10][3] */) {
    const v3 up = mk(0.f, 1.f, 0.f), fwd = mk(0.f, 0.f, 1.f);
    const v3 lpos = mk(L.px, L.py, L.pz);
    v3 tmp = ref_sub(lpos, start);
    v3 toL = ref_normalise(tmp);
    v3 prev = mk(__int_as_float(0x7fc00001), 0.f, 0.f);
    float angle = 0.f;
    RotM M = RotM{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int j = 0; j < 10; j++) {
        if (!same_bits(toL, prev)) {
            prev = toL;
            v3 P = ref_cross(toL, up);
            v3 e = ref_sub(ref_add(lpos, ref_scale(P, L.size)), start);
            v3 toEdge = ref_normalise(e);
            angle = ORE_COSF((ref_dot(toL, toEdge)) * 2);
            v3 n1 = ref_normalise(toL);
            v3 ax = ref_cross(fwd, n1);
            v3 axis = ref_normalise(ax);
            v3 n2 = ref_normalise(toL);
            float nAngle = ORE_ACOSF(ref_dot(n2, fwd));
            M = ref_rotate_matrix(nAngle, axis);
        }
        const float _z = (float)j / 10 * (1.0f - angle) + angle;
        const float sq = sqrtf(1.f - _z * _z);
        const float x = sq * c_cos_phi[j];
        const float y = sq * c_sin_phi[j];
        v3 nd = ref_sub(lpos, ref_matrix_apply(M, mk(x, y, _z)));
        v3 nn = ref_normalise(nd);
        dir[j * 3 + 0] = nn.x;
        dir[j * 3 + 1] = nn.y;
        dir[j * 3 + 2] = nn.z;
    }
    return ref_dot(normal, toL);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// direction of sample ray j (0..29: light j / 10, sample j % 10) in a per-thread bundle with `ls` floats per light
__device__ __forceinline__ v3 load_dir(const float* __restrict__ dirs, int ls, int j) {
    const int l = (j >= 20) ? 2 : (j >= 10 ? 1 : 0);
    const float* __restrict__ d = dirs + ls * l + 3 * (j - 10 * l);
    return mk(d[0], d[1], d[2]);
}

// cone_of10: cone of one light's 10 sample directions: axis = normalised sum, w = min_j axis.D_j, from a
// 16-byte aligned bundle loaded once as vectors and reduced from registers; w = -1 for a degenerate bundle (zero-length
// or non-unit direction).  One out-of-line copy for all lights.
__device__ __noinline__ float4 cone_of10(const float* __restrict__ d /* 32 floats, 30 used */) {
    const float4* __restrict__ d4 = reinterpret_cast<const float4*>(d);
    float v[32];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const float4 w = d4[q];
        v[4 * q] = w.x;
        v[4 * q + 1] = w.y;
        v[4 * q + 2] = w.z;
        v[4 * q + 3] = w.w;
    }
    float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
    for (int j = 0; j < 10; j++) {
        sx += v[3 * j];
        sy += v[3 * j + 1];
        sz += v[3 * j + 2];
    }
    const float inv = rsqrtf(fmaf(sx, sx, fmaf(sy, sy, sz * sz)));
    float cmin = 1.f;
    bool ok = isfinite(inv);
    sx *= inv;
    sy *= inv;
    sz *= inv;
#pragma unroll
    for (int j = 0; j < 10; j++) {
        const float dd = fmaf(v[3 * j], v[3 * j], fmaf(v[3 * j + 1], v[3 * j + 1], v[3 * j + 2] * v[3 * j + 2]));
        // the filters and the sure-hit shortcut assume |D| = 1: the reference's normalise() leaves | |D|^2 - 1 |
        // below 5e-7 unless it degenerated (zero or denormal input), which is what this check catches
        ok = ok && fabsf(dd - 1.f) < 4e-6f;
        cmin = fminf(cmin, fmaf(sx, v[3 * j], fmaf(sy, v[3 * j + 1], sz * v[3 * j + 2])));
    }
    return make_float4(sx, sy, sz, ok ? cmin : -1.f);
}

// ------------------------------------------------------------------------------------
// shadow_kernel
// ------------------------------------------------------------------------------------
template <int NL, bool EXH>
__global__ void __launch_bounds__(SHADOW_THREADS, ORE_SHADOW_MIN_CTAS) shadow_kernel(const FrameParams prm) {
    constexpr int NR = 10 * NL;
    constexpr int SG = ORE_SHADOW_SG;  // spheres per inner step
    constexpr uint32_t ALL = (NR == 32) ? 0xffffffffu : ((1u << NR) - 1u);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[MAX_STAGES];
    __shared__ int s_batch;

    const int tid = threadIdx.x;
    TilePipe tp;
    pipe_init(tp, prm, prm.sph_shad, reinterpret_cast<float4*>(smem_raw), bars);

    const uint32_t n_items = (uint32_t)prm.counters[CNT_HITS];
    const v3 O0 = mk(prm.Ox, prm.Oy, prm.Oz);
    unsigned long long n_exact = 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_batch = (int)atomicAdd(&prm.counters[CNT_SHADOW_CURSOR], 1ull);
        __syncthreads();
        const uint32_t batch = (uint32_t)s_batch;
        if ((unsigned long long)batch * SHADOW_THREADS >= n_items) break;
        const uint32_t item = batch * SHADOW_THREADS + tid;
        const bool valid = item < n_items;

        // ---- shading set-up (kernel.cu:1396-1405, 1643-1655) ----
        uint32_t o = 0;
        size_t o_out = 0;
        v3 start = mk(1e9f, 1e9f, 1e9f), normal = mk(0.f, 0.f, 0.f);
        float tr = 0.f, tg = 0.f, tb = 0.f;
        if (valid) {
            o = prm.hit_list[item];
            const int k = (int)(o / (uint32_t)prm.W), x = (int)(o % (uint32_t)prm.W);
            o_out = out_index(prm, k, x);
            const v3 D = primary_dir(prm, prm.dx_tab[x], prm.dy_tab[k]);
            const float nt = prm.hit_t[o];
            const float4 sc = __ldg(&prm.sph_exact[prm.hit_id[o]]);
            const v3 new_org = ref_add(O0, ref_scale(D, nt));
            normal = ref_sub(new_org, mk(sc.x, sc.y, sc.z));
            ref_normalise(normal);
            const float txf = (float)((1 + (double)ORE_ATAN2F(normal.z, normal.x) / 3.1415) * 0.5);
            const float tyf = (float)((double)ORE_ACOSF(normal.y) / 3.1415);
            const int maxX = prm.tex_w, maxY = prm.tex_h;
            start = ref_add(ref_scale(normal, 0.00001f), new_org);
            int c_index = (int)(tyf * (float)maxY) * maxX + (int)(txf * (float)maxX);
            c_index = clamp_index(c_index, maxX * maxY);
            tr = __ldg(&prm.tex_r[c_index]);
            tg = __ldg(&prm.tex_g[c_index]);
            tb = __ldg(&prm.tex_b[c_index]);
        }
        float fr = 0.f, fg = 0.f, fb = 0.f;

        for (int l0 = 0; l0 < prm.n_lights; l0 += NL) {
            float dx[NR], dy[NR], dz[NR], a_l[NL];
            float dirs[NR * 3];  // local-memory copy, indexed dynamically by the rare exact path
            uint32_t blocked = ALL;
#pragma unroll
            for (int l = 0; l < NL; l++) {
                const bool lit = valid && (l0 + l) < prm.n_lights;
                a_l[l] = 0.f;
                if (lit) {
                    a_l[l] = light_directions(prm.lights[l0 + l], start, normal, dirs + 30 * l);
                    blocked &= ~(0x3ffu << (10 * l));
                }
#pragma unroll
                for (int j = 0; j < 10; j++) {
                    dx[l * 10 + j] = lit ? dirs[(l * 10 + j) * 3 + 0] : 0.f;
                    dy[l * 10 + j] = lit ? dirs[(l * 10 + j) * 3 + 1] : 0.f;
                    dz[l * 10 + j] = lit ? dirs[(l * 10 + j) * 3 + 2] : 0.f;
                }
            }

            // ---- any-hit over all spheres (kernel.cu:1501-1510), filter + exact ----
            tp.begin_round();
            bool warp_done = false;
            for (int c = 0; c < tp.n_chunks; c++) {
                const float4* tile = tp.acquire(c);
                const int cnt = tp.count(c);
                const int base = c * tp.chunk;
                if (!warp_done) {
#pragma unroll 1
                    for (int s = 0; s < cnt; s += SG) {
                        // SG spheres at a time: SG independent 3-FMA chains per ray keep the FMA pipe
                        // busy without waiting on its 4-cycle latency
                        float Lx[SG], Ly[SG], Lz[SG], sv[SG];
#pragma unroll
                        for (int u = 0; u < SG; u++) {
                            const float4 q = tile[s + u];
                            Lx[u] = start.x - q.x;
                            Ly[u] = start.y - q.y;
                            Lz[u] = start.z - q.z;
                            const float LL = fmaf(Lz[u], Lz[u], fmaf(Ly[u], Ly[u], Lx[u] * Lx[u]));
                            const float Cm = fmaf(LL, 1.0f - ORE_KAPPA_SHADOW, -(q.w * q.w));
                            const float sq = Cm * rsqrt_approx(Cm);
                            sv[u] = EXH ? -ORE_BIG : ((Cm > 1e-20f) ? sq : -ORE_BIG);
                        }
                        int acc = 0;
#pragma unroll
                        for (int j = 0; j < NR; j++) {
                            float h[SG];
#pragma unroll
                            for (int u = 0; u < SG; u++) h[u] = fmaf(dz[j], Lz[u], sv[u]);
#pragma unroll
                            for (int u = 0; u < SG; u++) h[u] = fmaf(dy[j], Ly[u], h[u]);
#pragma unroll
                            for (int u = 0; u < SG; u++) h[u] = fmaf(dx[j], Lx[u], h[u]);
#pragma unroll
                            for (int u = 0; u < SG; u += 2) acc |= __float_as_int(h[u]) | __float_as_int(h[u + 1]);
                        }
                        if (acc < 0 && blocked != ALL) {
                            // rare: some ray of this thread may hit one of the SG spheres -> exact sequence
#pragma unroll 1
                            for (int u = 0; u < SG; u++) {
                                const int idx = base + s + u;
                                if (idx >= prm.n_spheres) break;
                                const float4 q = tile[s + u];
                                const float lx = start.x - q.x, ly = start.y - q.y, lz = start.z - q.z;
                                const float LL = fmaf(lz, lz, fmaf(ly, ly, lx * lx));
                                const float Cm = fmaf(LL, 1.0f - ORE_KAPPA_SHADOW, -(q.w * q.w));
                                const float sq = Cm * rsqrt_approx(Cm);
                                const float svu = EXH ? -ORE_BIG : ((Cm > 1e-20f) ? sq : -ORE_BIG);
                                // candidate rays of this sphere (filter), then the exact sequence in a
                                // rolled loop that reads the directions from their local-memory copy
                                uint32_t cand = 0;
#pragma unroll
                                for (int j = 0; j < NR; j++) {
                                    const float h = fmaf(dx[j], lx, fmaf(dy[j], ly, fmaf(dz[j], lz, svu)));
                                    cand |= (h < 0.f) ? (1u << j) : 0u;
                                }
                                cand &= ~blocked;
                                if (cand) {
                                    const float4 ex = __ldg(&prm.sph_exact[idx]);
                                    uint32_t newly = 0;
                                    while (cand) {
                                        const int j = __ffs(cand) - 1;
                                        cand &= cand - 1;
                                        float t;
                                        n_exact++;
                                        if (ref_intersect(start, mk(dirs[j * 3], dirs[j * 3 + 1], dirs[j * 3 + 2]), ex.x, ex.y,
                                                          ex.z, ex.w, t))
                                            newly |= 1u << j;
                                    }
                                    if (newly) {
                                        blocked |= newly;
#pragma unroll
                                        for (int j = 0; j < NR; j++) {
                                            const bool nb = (newly >> j) & 1u;
                                            dx[j] = nb ? 0.f : dx[j];
                                            dy[j] = nb ? 0.f : dy[j];
                                            dz[j] = nb ? 0.f : dz[j];
                                        }
                                    }
                                }
                            }
                        }
                        if ((s & (2 * SG - 1)) == SG) {
                            if (__all_sync(0xffffffffu, blocked == ALL)) {
                                warp_done = true;
                                break;
                            }
                        }
                    }
                }
                tp.release(c);
            }

            // ---- light accumulation (kernel.cu:1537-1543, 1673-1675) ----
            if (valid) {
#pragma unroll
                for (int l = 0; l < NL; l++) {
                    if (l0 + l < prm.n_lights) {
                        float b = c_b_of_k[10 - __popc((blocked >> (l * 10)) & 0x3ffu)];
                        const float a = a_l[l];
                        b *= a > 0 ? a : 0;
                        const LightP L = prm.lights[l0 + l];
                        fr += b * L.r * tr;
                        fg += b * L.g * tg;
                        fb += b * L.b * tb;
                    }
                }
            }
        }
        if (valid) prm.pixels[o_out] = ref_rgb_to_int((int)(fr * 254.f), (int)(fg * 254.f), (int)(fb * 254.f));
    }
    if (n_exact) atomicAdd(&prm.counters[CNT_EXACT_SHADOW], n_exact);
}

// ------------------------------------------------------------------------------------
// shadow_cone_kernel (default shadow path)
//
// One thread per HIT pixel.  The 10 sample rays of a light leave the same origin inside a
// narrow cone (new_dir = normalise(l.pos - M*(x,y,_z)), kernel.cu:1468: the jitter vector
// has length <= ~1.7 against |l.pos| of tens of units).  For every sphere the thread first
// runs ONE conservative cone-vs-sphere test per light (3 FMA + threshold) and only when a
// cone can touch the sphere does it test that light's individual rays (filter, then the
// exact reference sequence).  Every (pixel, sphere) pair is still visited in index order;
// no spatial structure is built.  Results are identical to the per-ray kernel.
//
// Cone test (DESIGN.md "Cone filter"): with L = O - c, s' = sqrt(Cm) the per-ray filter
// says a ray D can only hit if D.L + s' <= 0, i.e. its angle to the centre direction is at
// most b' = acos(s'/|L|).  Rays lie within a of the axis A, so a hit needs
// angle(A, centre) <= a + b'  <=>  A.L + cos(a) s' - sin(a) sqrt(|L|^2 - s'^2) <= 0,
// and sqrt(|L|^2 - s'^2) <= 1.002 R' + 0.00196 s'.
// ------------------------------------------------------------------------------------
template <bool EXH>
__global__ void __launch_bounds__(CTA_THREADS, 3) shadow_cone_kernel(const FrameParams prm) {
    constexpr int NL = 3;        // lights per pass
    constexpr int NR = 10 * NL;
    constexpr uint32_t ALL = (1u << NR) - 1u;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[MAX_STAGES];
    __shared__ int s_batch;

    const int tid = threadIdx.x;
    TilePipe tp;
    pipe_init(tp, prm, prm.sph_shad, reinterpret_cast<float4*>(smem_raw), bars);

    const uint32_t n_items = (uint32_t)prm.counters[CNT_HITS];
    const v3 O0 = mk(prm.Ox, prm.Oy, prm.Oz);
    unsigned long long n_exact = 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_batch = (int)atomicAdd(&prm.counters[CNT_SHADOW_CURSOR], 1ull);
        __syncthreads();
        const uint32_t batch = (uint32_t)s_batch;
        if ((unsigned long long)batch * CTA_THREADS >= n_items) break;
        const uint32_t item = batch * CTA_THREADS + tid;
        const bool valid = item < n_items;

        // ---- shading set-up (kernel.cu:1396-1405, 1643-1655) ----
        size_t o_out = 0;
        v3 start = mk(1e9f, 1e9f, 1e9f), normal = mk(0.f, 0.f, 0.f);
        float tr = 0.f, tg = 0.f, tb = 0.f;
        if (valid) {
            const uint32_t o = prm.hit_list[item];
            const int k = (int)(o / (uint32_t)prm.W), x = (int)(o % (uint32_t)prm.W);
            o_out = out_index(prm, k, x);
            const v3 D = primary_dir(prm, prm.dx_tab[x], prm.dy_tab[k]);
            const float nt = prm.hit_t[o];
            const float4 sc = __ldg(&prm.sph_exact[prm.hit_id[o]]);
            const v3 new_org = ref_add(O0, ref_scale(D, nt));
            normal = ref_sub(new_org, mk(sc.x, sc.y, sc.z));
            ref_normalise(normal);
            const float txf = (float)((1 + (double)ORE_ATAN2F(normal.z, normal.x) / 3.1415) * 0.5);
            const float tyf = (float)((double)ORE_ACOSF(normal.y) / 3.1415);
            const int maxX = prm.tex_w, maxY = prm.tex_h;
            start = ref_add(ref_scale(normal, 0.00001f), new_org);
            int c_index = (int)(tyf * (float)maxY) * maxX + (int)(txf * (float)maxX);
            c_index = clamp_index(c_index, maxX * maxY);
            tr = __ldg(&prm.tex_r[c_index]);
            tg = __ldg(&prm.tex_g[c_index]);
            tb = __ldg(&prm.tex_b[c_index]);
        }
        float fr = 0.f, fg = 0.f, fb = 0.f;

        for (int l0 = 0; l0 < prm.n_lights; l0 += NL) {
            float dirs[NR * 3];  // local memory (L1): only the rare per-ray path reads it
            float Ax[NL], Ay[NL], Az[NL], ca[NL], sa[NL], a_l[NL];
            uint32_t blocked = ALL;
            bool force = false;  // some light cannot use the cone test: every sphere is a candidate
            {
                const int n_act = valid ? min(NL, prm.n_lights - l0) : 0;
                light_directions_n<NL>(prm, l0, n_act, start, normal, dirs, a_l);
            }
#pragma unroll
            for (int l = 0; l < NL; l++) {
                const bool lit = valid && (l0 + l) < prm.n_lights;
                Ax[l] = Ay[l] = Az[l] = ca[l] = sa[l] = 0.f;
                if (lit) {
                    float* d = dirs + 30 * l;
                    blocked &= ~(0x3ffu << (10 * l));
                    // cone of the 10 rays: axis = normalised sum, cos(a) = min_j axis.D_j
                    float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll 1
                    for (int j = 0; j < 10; j++) {
                        sx += d[j * 3];
                        sy += d[j * 3 + 1];
                        sz += d[j * 3 + 2];
                    }
                    const float inv = rsqrtf(fmaf(sx, sx, fmaf(sy, sy, sz * sz)));
                    float cmin = 1.f;
                    bool ok = isfinite(inv);
                    if (ok) {
                        sx *= inv;
                        sy *= inv;
                        sz *= inv;
#pragma unroll 1
                        for (int j = 0; j < 10; j++) {
                            const float dd = fmaf(d[j * 3], d[j * 3], fmaf(d[j * 3 + 1], d[j * 3 + 1], d[j * 3 + 2] * d[j * 3 + 2]));
                            ok = ok && fabsf(dd - 1.f) < 1e-4f;  // the filters assume |D| = 1 (normalised by the reference)
                            cmin = fminf(cmin, fmaf(sx, d[j * 3], fmaf(sy, d[j * 3 + 1], sz * d[j * 3 + 2])));
                        }
                    }
                    if (ok && cmin > 0.5f && !EXH) {
                        const float cosa = cmin - 4e-6f;
                        const float sina = sqrtf(fmaxf(0.f, fmaf(-cosa, cosa, 1.f))) * 1.0001f + 1e-6f;
                        Ax[l] = sx;
                        Ay[l] = sy;
                        Az[l] = sz;
                        ca[l] = cosa - 0.00196f * sina;
                        sa[l] = 1.002f * sina;
                    } else {
                        // degenerate bundle (zero direction, very wide cone) or exhaustive mode
                        force = true;
                    }
                }
            }

            tp.begin_round();
            bool warp_done = false;
            for (int c = 0; c < tp.n_chunks; c++) {
                const float4* tile = tp.acquire(c);
                const int cnt = tp.count(c);
                const int base = c * tp.chunk;
                if (!warp_done) {
                    const float4* __restrict__ tp_ptr = tile;
#pragma unroll 1
                    for (int s = 0; s < cnt; s += 2, tp_ptr += 2) {
                        float m[2];
#pragma unroll
                        for (int u = 0; u < 2; u++) {
                            const float4 q = tp_ptr[u];
                            const float Lx = start.x - q.x, Ly = start.y - q.y, Lz = start.z - q.z;
                            const float LL = fmaf(Lz, Lz, fmaf(Ly, Ly, Lx * Lx));
                            const float Cm = fmaf(LL, 1.0f - ORE_KAPPA_SHADOW, -(q.w * q.w));
                            const float sq = Cm * rsqrt_approx(Cm);
                            const float sv = (Cm > 1e-20f) ? sq : -ORE_BIG;
                            float mm = INFINITY;
#pragma unroll
                            for (int l = 0; l < NL; l++) {
                                const float T = fmaf(ca[l], sv, -(sa[l] * q.w));
                                mm = fminf(mm, fmaf(Ax[l], Lx, fmaf(Ay[l], Ly, fmaf(Az[l], Lz, T))));
                            }
                            m[u] = mm;
                        }
                        if ((force || fminf(m[0], m[1]) < 0.f) && blocked != ALL) {
#pragma unroll 1
                            for (int u = 0; u < 2; u++) {
                                const int idx = base + s + u;
                                if (idx >= prm.n_spheres) break;
                                // (rare) recompute this sphere's filter terms, then test the live rays
                                const float4 q = tp_ptr[u];
                                const float lx = start.x - q.x, ly = start.y - q.y, lz = start.z - q.z;
                                const float LL = fmaf(lz, lz, fmaf(ly, ly, lx * lx));
                                const float Cm = fmaf(LL, 1.0f - ORE_KAPPA_SHADOW, -(q.w * q.w));
                                const float sq = Cm * rsqrt_approx(Cm);
                                const float svu = (EXH || !(Cm > 1e-20f)) ? -ORE_BIG : sq;
                                uint32_t live = ~blocked & ALL;
                                if (!force) {
                                    // only the lights whose cone touches this sphere
                                    uint32_t lm = 0;
#pragma unroll
                                    for (int l = 0; l < NL; l++) {
                                        const float T = fmaf(ca[l], svu, -(sa[l] * q.w));
                                        if (fmaf(Ax[l], lx, fmaf(Ay[l], ly, fmaf(Az[l], lz, T))) < 0.f) lm |= 0x3ffu << (10 * l);
                                    }
                                    live &= lm;
                                }
                                if (!live) continue;
                                const float4 ex = __ldg(&prm.sph_exact[idx]);
                                while (live) {
                                    const int j = __ffs(live) - 1;
                                    live &= live - 1;
                                    const v3 D = mk(dirs[j * 3], dirs[j * 3 + 1], dirs[j * 3 + 2]);
                                    const float h = fmaf(D.x, lx, fmaf(D.y, ly, fmaf(D.z, lz, svu)));
                                    if (h < 0.f) {
                                        float t;
                                        n_exact++;
                                        if (ref_intersect(start, D, ex.x, ex.y, ex.z, ex.w, t)) blocked |= 1u << j;
                                    }
                                }
                            }
                            // a light whose 10 rays are all blocked drops out of the cone test
#pragma unroll
                            for (int l = 0; l < NL; l++) {
                                if (((blocked >> (10 * l)) & 0x3ffu) == 0x3ffu) {
                                    Ax[l] = Ay[l] = Az[l] = 0.f;
                                    ca[l] = 0.f;
                                    sa[l] = 0.f;
                                }
                            }
                        }
                        if ((s & 14) == 14) {
                            if (__all_sync(0xffffffffu, blocked == ALL)) {
                                warp_done = true;
                                break;
                            }
                        }
                    }
                }
                tp.release(c);
            }

            // ---- light accumulation (kernel.cu:1537-1543, 1673-1675) ----
            if (valid) {
#pragma unroll
                for (int l = 0; l < NL; l++) {
                    if (l0 + l < prm.n_lights) {
                        float b = c_b_of_k[10 - __popc((blocked >> (l * 10)) & 0x3ffu)];
                        const float a = a_l[l];
                        b *= a > 0 ? a : 0;
                        const LightP L = prm.lights[l0 + l];
                        fr += b * L.r * tr;
                        fg += b * L.g * tg;
                        fb += b * L.b * tb;
                    }
                }
            }
        }
        if (valid) prm.pixels[o_out] = ref_rgb_to_int((int)(fr * 254.f), (int)(fg * 254.f), (int)(fb * 254.f));
    }
    if (n_exact) atomicAdd(&prm.counters[CNT_EXACT_SHADOW], n_exact);
}

// ------------------------------------------------------------------------------------
// shade_point: the shading set-up of one hit pixel (kernel.cu:1380-1425, 1643-1655): hit point, normal, shadow-ray
// origin `start`, texel colour.  `item` indexes the hit list.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void shade_point(const FrameParams& prm, uint32_t item, const v3 O0, size_t& o_out, int& my_id,
                                            v3& start, v3& normal, float& tr, float& tg, float& tb) {
    const uint32_t o = prm.hit_list[item];
    const int k = (int)(o / (uint32_t)prm.W), x = (int)(o % (uint32_t)prm.W);
    o_out = out_index(prm, k, x);
    const v3 D = primary_dir(prm, prm.dx_tab[x], prm.dy_tab[k]);
    const float nt = prm.hit_t[o];
    my_id = prm.hit_id[o];
    v3 new_org = ref_add(O0, ref_scale(D, nt));
    float txf, tyf;
    if (my_id >= prm.n_spheres + prm.n_cubes + prm.n_planes) {
        // triangle hit, kernel.cu:1380-1394
        triangle_attributes(prm.tris + 27 * (size_t)(my_id - prm.n_spheres - prm.n_cubes - prm.n_planes),
                            prm.mesh_has_normals, O0.x, O0.y, O0.z, D.x, D.y, D.z, nt, &normal, &new_org, &txf, &tyf);
    } else if (my_id >= prm.n_spheres + prm.n_cubes) {
        // plane hit, kernel.cu:1407-1416
        const float4 no = __ldg(&prm.planes[2 * (my_id - prm.n_spheres - prm.n_cubes) + 1]);
        normal = mk(no.x, no.y, no.z);
        txf = 0.5f;
        tyf = 0.5f;
    } else {
        // sphere hit kernel.cu:1396-1405 / cube hit :1417-1425: normal from the primitive's `orgin`
        const float4 sc = (my_id < prm.n_spheres) ? __ldg(&prm.sph_exact[my_id])
                                                  : __ldg(&prm.cubes[3 * (my_id - prm.n_spheres) + 2]);
        normal = ref_sub(new_org, mk(sc.x, sc.y, sc.z));
        ref_normalise(normal);
        txf = (float)((1 + (double)ORE_ATAN2F(normal.z, normal.x) / 3.1415) * 0.5);
        tyf = (float)((double)ORE_ACOSF(normal.y) / 3.1415);
    }
    const int maxX = prm.tex_w, maxY = prm.tex_h;
    start = ref_add(ref_scale(normal, 0.00001f), new_org);
    int c_index = (int)(tyf * (float)maxY) * maxX + (int)(txf * (float)maxX);
    c_index = clamp_index(c_index, maxX * maxY);
    tr = __ldg(&prm.tex_r[c_index]);
    tg = __ldg(&prm.tex_g[c_index]);
    tb = __ldg(&prm.tex_b[c_index]);
}

// ------------------------------------------------------------------------------------
// shade_setup_kernel (stage A of the default shadow pass)
//
// Per hit pixel: shading set-up + the 10 shadow-ray directions of every light, written to the staging buffer.
// This is all the per-pixel transcendental code (atan2f/acosf/cosf/sinf, ~30 KB of instructions); keeping it out
// of the sweep kernel lets each kernel's hot loop stay resident in the SM instruction cache.
// Warps run independently: each fetches blocks of 32 consecutive hit-list items.
// ------------------------------------------------------------------------------------
#ifndef ORE_STAGE_A_MIN_CTAS
#define ORE_STAGE_A_MIN_CTAS (1024 / ORE_STAGE_A_THREADS)
#endif
__global__ void __launch_bounds__(STAGE_A_THREADS, ORE_STAGE_A_MIN_CTAS) shade_setup_kernel(const FrameParams prm, const StageArgs st) {
    const int lane = threadIdx.x & 31;
    const uint32_t n_items = (uint32_t)prm.counters[CNT_HITS];
    const v3 O0 = mk(prm.Ox, prm.Oy, prm.Oz);
    for (;;) {
        uint32_t wb = 0;
        if (lane == 0) wb = (uint32_t)atomicAdd(&prm.counters[CNT_STAGE_A0 + st.chunk], 1ull);
        wb = __shfl_sync(0xffffffffu, wb, 0);
        if (wb >= st.cap_blocks) break;
        const uint32_t blk = st.first_block + wb;
        if ((unsigned long long)blk * 32ull >= n_items) break;
        const uint32_t item = blk * 32u + lane;
        const bool valid = item < n_items;
        const long long dbg_t0 = prm.dbg_cycles ? clock64() : 0;
        size_t o_out = 0;
        int my_id = -1;
        v3 start = mk(1e9f, 1e9f, 1e9f), normal = mk(0.f, 0.f, 0.f);
        float tr = 0.f, tg = 0.f, tb = 0.f;
        if (valid) shade_point(prm, item, O0, o_out, my_id, start, normal, tr, tg, tb);
        float* __restrict__ sp = st.buf + ((size_t)wb * (size_t)st.nv) * 32u + lane;
        sp[0] = start.x;
        sp[32] = start.y;
        sp[64] = start.z;
        sp[96] = tr;
        sp[128] = tg;
        sp[160] = tb;
#pragma unroll 1
        for (int li = 0; li < prm.n_lights; li++) {
            __align__(16) float d[32];
            float a = 0.f;
            if (valid) {
                LightP L;
                const LightP* __restrict__ src = &prm.lights[0];
                L.px = src[li].px; L.py = src[li].py; L.pz = src[li].pz; L.size = src[li].size;
                L.r = src[li].r; L.g = src[li].g; L.b = src[li].b;
                a = light_directions_reuse(L, start, normal, d);
            }
            float* __restrict__ q = sp + (size_t)(6 + 31 * li) * 32u;
            if (valid) {
                const float4* __restrict__ d4 = reinterpret_cast<const float4*>(d);
#pragma unroll
                for (int v = 0; v < 8; v++) {
                    const float4 w = d4[v];
                    q[(4 * v) * 32] = w.x;
                    q[(4 * v + 1) * 32] = w.y;
                    if (4 * v + 2 < 30) q[(4 * v + 2) * 32] = w.z;
                    if (4 * v + 3 < 30) q[(4 * v + 3) * 32] = w.w;
                }
            }
            q[30 * 32] = a;
        }
        if (prm.dbg_cycles && lane == 0 && blk < prm.dbg_cap) prm.dbg_cycles[blk] = (uint32_t)(clock64() - dbg_t0);
    }
}

// beam_may_touch: can any ray of the warp's (up to three) beams touch the ball q = (centre, radius)?  (DESIGN.md 2.4;
// used for single spheres and for the bounding spheres of 32-sphere clusters.)  bx,by,bz = centroid of the group's
// origins; per light: axis wA*, tan of the beam half-angle, k1 = -a_min, k2 = rho_perp.  A radius >= 1e18 (or NaN)
// means "always".
__device__ __forceinline__ bool beam_may_touch(const float4 q, float bx, float by, float bz, const float (&wAx)[3],
                                               const float (&wAy)[3], const float (&wAz)[3], const float (&wtan)[3],
                                               const float (&wk1)[3], const float (&wk2)[3], bool wforce) {
    const float Lx = bx - q.x, Ly = by - q.y, Lz = bz - q.z;
    const float LL = fmaf(Lz, Lz, fmaf(Ly, Ly, Lx * Lx));
    const float Rq = fmaf(q.w, 1.0001f, fmaf(LL, 1e-12f, 1e-6f));  // radius + rounding slack
    const float slack = LL * 2e-6f;                                 // cancellation in LL - sc^2
    bool wc = wforce || !(q.w < 1e18f);
#pragma unroll
    for (int l = 0; l < 3; l++) {
        const float sc = -fmaf(wAx[l], Lx, fmaf(wAy[l], Ly, wAz[l] * Lz));  // centre's axial coordinate
        const float u = sc + Rq + wk1[l];                                     // >= 0 unless wholly behind
        const float thr = fmaf(u, wtan[l], Rq + wk2[l]);
        const float d2 = fmaf(-sc, sc, LL);
        wc = wc || (u >= 0.f && d2 <= fmaf(thr, thr, slack));
    }
    return wc;
}

// ------------------------------------------------------------------------------------
// shadow_beam_kernel (default shadow path)
//
// shadow_cone_kernel with one more level in front: a warp's 32 hit pixels are neighbours
// (hit-list order = row-major inside a 32 x P tile), so their ray origins fit in a small
// region and their per-light cones in one slightly wider cone around a common axis line.
// Level 1: lane i tests sphere s0+i against the three warp beams (axial / lateral distance
// to the axis line through the origins' centroid) and the warp ballots.  Level 2: each lane
// runs its own per-pixel cone test, then the per-ray filter and the exact sequence, on the
// survivors.  A hit at X = S_i + tD (t >= 0) has axial coordinate <= centre + R and lateral
// offset <= rho_perp + t sin(a), which is exactly what level 1 bounds.
// ------------------------------------------------------------------------------------
template <bool EXH, bool STAGED>
__global__ void __launch_bounds__(BEAM_THREADS, ORE_BEAM_MIN_CTAS) shadow_beam_kernel(const FrameParams prm, const StageArgs st) {
    constexpr int NL = 3;        // lights per pass
    constexpr int NR = 10 * NL;
    constexpr uint32_t ALL = (1u << NR) - 1u;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[MAX_STAGES];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    // Sphere records (cx,cy,cz,R'): the whole array is staged once per CTA into shared memory by one
    // TMA bulk copy when it fits; larger scenes are read through L1/L2 (256 KiB for 16384 spheres).
    // Either way warps run independently: each fetches 32 consecutive hit pixels at a time.
    const float4* __restrict__ spheres = prm.sph_sort;
    const float4* __restrict__ clusters = prm.clu_sph;
    const int n_clu = prm.n_clusters;
    if (prm.beam_resident) {
        float4* slot = reinterpret_cast<float4*>(smem_raw);
        if (tid == 0) {
            mbar_init(&bars[0], 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t sbytes = (uint32_t)n_clu * 32u * 16u, cbytes = (uint32_t)((n_clu + 3) & ~3) * 16u;
            mbar_expect_tx(&bars[0], sbytes + cbytes);
            tma_bulk_g2s(slot, prm.sph_sort, sbytes, &bars[0]);
            tma_bulk_g2s(slot + (size_t)n_clu * 32, prm.clu_sph, cbytes, &bars[0]);
        }
        mbar_wait(&bars[0], 0);
        spheres = slot;
        clusters = slot + (size_t)n_clu * 32;
    }
    const int n_sph = prm.n_spheres;

    const uint32_t n_items = (uint32_t)prm.counters[CNT_HITS];
    const v3 O0 = mk(prm.Ox, prm.Oy, prm.Oz);
    unsigned long long n_exact = 0;
    unsigned int n_l1 = 0, n_l2 = 0;

    // the block cursor is fetched one block ahead, so that the next block's staged values can be pulled into L2
    // while this one is processed
    unsigned long long* const cursor = &prm.counters[STAGED ? CNT_STAGE_B0 + st.chunk : CNT_SHADOW_CURSOR];
    uint32_t wb = 0;
    if (lane == 0) wb = (uint32_t)atomicAdd(cursor, 1ull);
    wb = __shfl_sync(0xffffffffu, wb, 0);
    for (;; ) {
        if (STAGED && wb >= st.cap_blocks) break;
        const uint32_t blk = st.first_block + wb;  // fused: 0, or the first block past the staged chunks (catch-all)
        if ((unsigned long long)blk * 32ull >= n_items) break;
        const uint32_t item = blk * 32u + lane;
        const bool valid = item < n_items;
        const long long dbg_t0 = prm.dbg_cycles ? clock64() : 0;
        // not near the end of the list / chunk, where a reserved block would wait behind this one while other
        // warps run dry
        const bool ahead = (unsigned long long)(blk + 8192u) * 32ull < n_items && (!STAGED || wb + 8192u < st.cap_blocks);
        uint32_t wb_next = 0;
        if (ahead) {
            if (lane == 0) wb_next = (uint32_t)atomicAdd(cursor, 1ull);
            wb_next = __shfl_sync(0xffffffffu, wb_next, 0);
            if (STAGED && wb_next < st.cap_blocks) {
                const float* nb = st.buf + ((size_t)wb_next * (size_t)st.nv) * 32u;
                for (int v = lane; v < st.nv; v += 32) prefetch_l2(nb + (size_t)v * 32u);
            }
        }

        // ---- shading set-up (kernel.cu:1396-1405, 1643-1655), or its staged result ----
        size_t o_out = 0;
        int my_id = -1;
        v3 start = mk(1e9f, 1e9f, 1e9f), normal = mk(0.f, 0.f, 0.f);
        float tr = 0.f, tg = 0.f, tb = 0.f;
        const float* __restrict__ sp = STAGED ? st.buf + ((size_t)wb * (size_t)st.nv) * 32u + lane : nullptr;
        if (STAGED) {
            if (valid) {
                const uint32_t o = prm.hit_list[item];
                const int k = (int)(o / (uint32_t)prm.W), x = (int)(o % (uint32_t)prm.W);
                o_out = out_index(prm, k, x);
                my_id = prm.hit_id[o];
                start = mk(sp[0], sp[32], sp[64]);
                tr = sp[96];
                tg = sp[128];
                tb = sp[160];
            }
        } else if (valid) {
            shade_point(prm, item, O0, o_out, my_id, start, normal, tr, tg, tb);
        }
        float fr = 0.f, fg = 0.f, fb = 0.f;

        for (int l0 = 0; l0 < prm.n_lights; l0 += NL) {
            // local memory (L1), 32 floats per light so that a light's 10 directions load as 16-byte vectors
            constexpr int LS = 32;
            __align__(16) float dirs[NL * LS];
            float Ax[NL], Ay[NL], Az[NL], ca[NL], sa[NL], a_l[NL];
            uint32_t blocked = ALL;
            bool force = false;  // some light cannot use the cone test: every sphere is a candidate
#pragma unroll 1
            for (int l = 0; l < NL; l++) {
                float a = 0.f;
                if (valid && (l0 + l) < prm.n_lights) {
                    if (STAGED) {
                        const float* __restrict__ q = sp + (size_t)(6 + 31 * (l0 + l)) * 32u;
                        float* __restrict__ d = dirs + LS * l;
#pragma unroll
                        for (int j = 0; j < 30; j++) d[j] = q[j * 32];
                        a = q[30 * 32];
                    } else {
                        LightP L;
                        const LightP* __restrict__ src = &prm.lights[0];
                        // copy out of the parameter bank without taking its address into local memory
                        const int li = l0 + l;
                        L.px = src[li].px; L.py = src[li].py; L.pz = src[li].pz; L.size = src[li].size;
                        L.r = src[li].r; L.g = src[li].g; L.b = src[li].b;
                        a = light_directions_reuse(L, start, normal, dirs + LS * l);
                    }
                }
                if (l == 0) a_l[0] = a;
                if (l == 1) a_l[1] = a;
                if (l == 2) a_l[2] = a;
            }
#pragma unroll
            for (int l = 0; l < NL; l++) {
                const bool lit = valid && (l0 + l) < prm.n_lights;
                Ax[l] = Ay[l] = Az[l] = ca[l] = sa[l] = 0.f;
                if (lit) {
                    blocked &= ~(0x3ffu << (10 * l));
                    const float4 cn = cone_of10(dirs + LS * l);
                    const float sx = cn.x, sy = cn.y, sz = cn.z, cmin = cn.w;
                    const bool ok = cmin > 0.f;
                    if (ok && cmin > 0.5f && !EXH) {
                        const float cosa = cmin - 4e-6f;
                        const float sina = sqrtf(fmaxf(0.f, fmaf(-cosa, cosa, 1.f))) * 1.0001f + 1e-6f;
                        Ax[l] = sx;
                        Ay[l] = sy;
                        Az[l] = sz;
                        ca[l] = cosa - 0.00196f * sina;
                        sa[l] = 1.002f * sina;
                    } else {
                        // degenerate bundle (zero direction, very wide cone) or exhaustive mode
                        force = true;
                    }
                }
            }

            // Lanes are processed in groups that hit the SAME sphere (the hit list is grouped that way, so a warp
            // normally is one group; a warp straddling a silhouette is two or three): origins on one sphere give
            // a narrow beam.  At most 4 passes; the last pass takes every lane that is left.
            uint32_t pending = __ballot_sync(0xffffffffu, valid);
#ifdef ORE_EXPERIMENT_A_ONLY
            pending = 0;
#endif
#pragma unroll 1
            for (int pass = 0; pending; pass++) {
                const int leader = __ffs(pending) - 1;
                const int gid = __shfl_sync(0xffffffffu, my_id, leader);
                const bool ing = valid && ((pending >> (tid & 31)) & 1u) && (pass >= 3 || my_id == gid);
                const uint32_t gmask = __ballot_sync(0xffffffffu, ing);
                pending &= ~gmask;
                // ---- warp beams: per light one axis line through the origins' centroid; the warp's rays of
                //      that light stay within rho_perp + (axial distance) * tan(a) of it (DESIGN.md "Beam test")
                const float nvalid = (float)__popc(gmask);
                float bx = ing ? start.x : 0.f, by = ing ? start.y : 0.f, bz = ing ? start.z : 0.f;
    #pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    bx += __shfl_xor_sync(0xffffffffu, bx, d);
                    by += __shfl_xor_sync(0xffffffffu, by, d);
                    bz += __shfl_xor_sync(0xffffffffu, bz, d);
                }
                const float inv_n = 1.f / fmaxf(nvalid, 1.f);
                bx *= inv_n;
                by *= inv_n;
                bz *= inv_n;
                const float ex = ing ? start.x - bx : 0.f, ey = ing ? start.y - by : 0.f, ez = ing ? start.z - bz : 0.f;
                const float escale = 1e-5f * (fabsf(bx) + fabsf(by) + fabsf(bz) + 1.f);  // rounding slack on offsets
                bool wforce = __any_sync(0xffffffffu, ing && force);
                // One rolled copy of the beam code (instruction-cache footprint): light l's beam is computed by the
                // whole warp, parked in lane l, and handed back to every lane by shuffles afterwards.
                float pAx = 0.f, pAy = 0.f, pAz = 0.f, ptan = 0.f, pk1 = -3e38f, pk2 = 0.f;
#pragma unroll 1
                for (int l = 0; l < NL; l++) {
                    const float Al_x = l == 0 ? Ax[0] : (l == 1 ? Ax[1] : Ax[2]);
                    const float Al_y = l == 0 ? Ay[0] : (l == 1 ? Ay[1] : Ay[2]);
                    const float Al_z = l == 0 ? Az[0] : (l == 1 ? Az[1] : Az[2]);
                    const float ca_l = l == 0 ? ca[0] : (l == 1 ? ca[1] : ca[2]);
                    const float sa_l = l == 0 ? sa[0] : (l == 1 ? sa[1] : sa[2]);
                    const bool part = ing && ca_l > 0.f;  // lanes of the group whose light l takes part (lit, not degenerate)
                    float sx = part ? Al_x : 0.f, sy = part ? Al_y : 0.f, sz = part ? Al_z : 0.f;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) {
                        sx += __shfl_xor_sync(0xffffffffu, sx, d);
                        sy += __shfl_xor_sync(0xffffffffu, sy, d);
                        sz += __shfl_xor_sync(0xffffffffu, sz, d);
                    }
                    const float n2 = fmaf(sx, sx, fmaf(sy, sy, sz * sz));
                    const float inv = rsqrtf(fmaxf(n2, 1e-30f));
                    sx *= inv;
                    sy *= inv;
                    sz *= inv;
                    // widest angle between the warp axis and any participating ray: theta_lane + a_lane
                    float cw = 1.f, amin = 3e38f, rp = 0.f;
                    if (part) {
                        const float sina = sa_l * (1.f / 1.002f);
                        const float cosa = ca_l + 0.00196f * sina;
                        const float c1 = fminf(1.f, fmaf(sx, Al_x, fmaf(sy, Al_y, sz * Al_z)));
                        const float s1 = sqrtf(fmaxf(0.f, fmaf(-c1, c1, 1.f))) + 1e-6f;
                        cw = fmaf(c1, cosa, -(s1 * sina)) - 2e-6f;
                        const float ai = fmaf(ex, sx, fmaf(ey, sy, ez * sz));  // axial offset of this origin
                        const float px = ex - ai * sx, py = ey - ai * sy, pz = ez - ai * sz;
                        amin = ai;
                        rp = sqrtf(fmaf(px, px, fmaf(py, py, pz * pz)));
                    }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) {
                        cw = fminf(cw, __shfl_xor_sync(0xffffffffu, cw, d));
                        amin = fminf(amin, __shfl_xor_sync(0xffffffffu, amin, d));
                        rp = fmaxf(rp, __shfl_xor_sync(0xffffffffu, rp, d));
                    }
                    if (__any_sync(0xffffffffu, part)) {
                        if (!(n2 > 1e-12f) || !(cw > 0.3f)) {
                            wforce = true;  // bundle axes disagree wildly: no warp-level culling
                        } else if (lane == l) {
                            const float sinw = sqrtf(fmaxf(0.f, fmaf(-cw, cw, 1.f))) * 1.0001f + 1e-6f;
                            pAx = sx;
                            pAy = sy;
                            pAz = sz;
                            ptan = sinw / cw * 1.0001f;
                            pk1 = -(amin - escale);   // u = sc + R' + k1 < 0: never a candidate
                            pk2 = rp * 1.0001f + escale;
                        }
                    }
                }
                float wAx[NL], wAy[NL], wAz[NL], wtan[NL], wk1[NL], wk2[NL];  // k1 = -a_min, k2 = rho_perp
#pragma unroll
                for (int l = 0; l < NL; l++) {
                    wAx[l] = __shfl_sync(0xffffffffu, pAx, l);
                    wAy[l] = __shfl_sync(0xffffffffu, pAy, l);
                    wAz[l] = __shfl_sync(0xffffffffu, pAz, l);
                    wtan[l] = __shfl_sync(0xffffffffu, ptan, l);
                    wk1[l] = __shfl_sync(0xffffffffu, pk1, l);
                    wk2[l] = __shfl_sync(0xffffffffu, pk2, l);
                }
                if (EXH) wforce = true;

                bool warp_done = false;
    #pragma unroll 1
                for (int c0 = 0; c0 < n_clu && !warp_done; c0 += 32) {
                    // ---- level 0: lane i tests the bounding sphere of cluster c0+i (32 spheres) against the beams ----
                    uint32_t cmask = __ballot_sync(0xffffffffu, c0 + lane < n_clu &&
                                                   beam_may_touch(clusters[min(c0 + lane, n_clu - 1)], bx, by, bz, wAx, wAy, wAz, wtan, wk1, wk2, wforce));
                  while (cmask && !warp_done) {
                    const int s0 = (c0 + __ffs(cmask) - 1) * 32;
                    cmask &= cmask - 1;
                    // ---- level 1: lane i tests sphere s0+i of that cluster ----
                    uint32_t wmask = __ballot_sync(0xffffffffu, s0 + lane < n_sph &&
                                                   beam_may_touch(spheres[s0 + lane], bx, by, bz, wAx, wAy, wAz, wtan, wk1, wk2, wforce));
                    n_l1 += __popc(wmask);
                    // ---- level 2: every lane runs its own cone test on the surviving spheres ----
                    while (wmask) {
                        const int i = __ffs(wmask) - 1;
                        wmask &= wmask - 1;
                        const int s = s0 + i;
                        const float4 q = spheres[s];
                        const float lx = start.x - q.x, ly = start.y - q.y, lz = start.z - q.z;
                        const float LL = fmaf(lz, lz, fmaf(ly, ly, lx * lx));
                        const float Cm = fmaf(LL, 1.0f - ORE_KAPPA_SHADOW, -(q.w * q.w));
                        const float sq = Cm * rsqrt_approx(Cm);
                        const float svu = (EXH || !(Cm > 1e-20f)) ? -ORE_BIG : sq;
                        uint32_t live = ing ? (~blocked & ALL) : 0u;
                        if (!force) {
                            uint32_t lm = 0;
    #pragma unroll
                            for (int l = 0; l < NL; l++) {
                                const float T = fmaf(ca[l], svu, -(sa[l] * q.w));
                                if (fmaf(Ax[l], lx, fmaf(Ay[l], ly, fmaf(Az[l], lz, T))) < 0.f) lm |= 0x3ffu << (10 * l);
                            }
                            live &= lm;
                        }
                        if (live) {
                            n_l2++;
                            const float4 ex4 = __ldg(&prm.sph_xsort[s]);
                            // per-ray filter h = D.L + s' < 0 on all 10 rays of each live light at once
                            // (independent loads and FMAs), then the exact sequence on the few that pass
                            // "Sure hit" (DESIGN.md 2.5): sphere::intersect returns true whenever its discriminant is
                            // positive and the far root passes the 0.0001 gate.  With b = D.L < 0 and | |D|^2 - 1 | <
                            // 5e-6 (cone_of10 admits nothing else), b^2 > (|L|^2 - r^2) + 2e-5 |L|^2 keeps the
                            // float discriminant positive (its rounding error is below 7e-6 |L|^2) and b^2 >
                            // 1e-6 |L|^2 + 1e-6 keeps the far root above 9e-4: the ray is blocked without running
                            // the exact sequence.  Everything in between is re-adjudicated exactly as before.
                            // (disabled = +inf, not a large finite number: b*b is +inf for a sphere at infinity)
                            const float sure_thr = (EXH || force)
                                                       ? INFINITY
                                                       : fmaxf(fmaf(LL, 1e-6f, 1e-6f), fmaf(LL, 2e-5f, fmaf(-ex4.w, ex4.w, LL)));
                            uint32_t cand = 0;
#pragma unroll 1
                            for (int l = 0; l < NL; l++) {
                                const uint32_t lv = (live >> (10 * l)) & 0x3ffu;
                                if (lv) {
                                    const float4* __restrict__ d4 = reinterpret_cast<const float4*>(dirs + LS * l);
                                    float dd[32];
#pragma unroll
                                    for (int v = 0; v < 8; v++) {
                                        const float4 w = d4[v];
                                        dd[4 * v] = w.x;
                                        dd[4 * v + 1] = w.y;
                                        dd[4 * v + 2] = w.z;
                                        dd[4 * v + 3] = w.w;
                                    }
                                    uint32_t m = 0, sure = 0;
#pragma unroll
                                    for (int j = 0; j < 10; j++) {
                                        const float b = fmaf(dd[3 * j], lx, fmaf(dd[3 * j + 1], ly, dd[3 * j + 2] * lz));
                                        if (b + svu < 0.f) m |= 1u << j;
                                        if (b < 0.f && b * b > sure_thr) sure |= 1u << j;
                                    }
                                    sure &= m & lv;  // only rays the filter lets through, that are still unblocked
                                    blocked |= sure << (10 * l);
                                    cand |= (m & lv & ~sure) << (10 * l);
                                }
                            }
                            if (cand) {
                                // exact re-adjudication; the next candidate's direction is loaded (local memory)
                                // before the current one's exact sequence runs
                                int j = __ffs(cand) - 1;
                                cand &= cand - 1;
                                v3 D = load_dir(dirs, LS, j);
                                for (;;) {
                                    int jn = -1;
                                    v3 Dn = D;
                                    if (cand) {
                                        jn = __ffs(cand) - 1;
                                        cand &= cand - 1;
                                        Dn = load_dir(dirs, LS, jn);
                                    }
                                    float t;
                                    n_exact++;
                                    if (ref_intersect(start, D, ex4.x, ex4.y, ex4.z, ex4.w, t)) blocked |= 1u << j;
                                    if (jn < 0) break;
                                    j = jn;
                                    D = Dn;
                                }
                            }
    #pragma unroll
                            for (int l = 0; l < NL; l++) {
                                if (((blocked >> (10 * l)) & 0x3ffu) == 0x3ffu) {
                                    Ax[l] = Ay[l] = Az[l] = 0.f;
                                    ca[l] = 0.f;
                                    sa[l] = 0.f;
                                }
                            }
                        }
                    }
                    if (__all_sync(0xffffffffu, !ing || blocked == ALL)) warp_done = true;
                  }
                }

            }

            // ---- triangles (kernel.cu:1475-1497; tested first in the reference - the result is an OR, order-free):
            //      per light, leaves outside the light's cone are skipped ----
            if (prm.n_boxes && valid) {
                const MeshArgs ma = {prm.tris, prm.boxes, prm.box_offsets, prm.box_indices, prm.n_boxes};
#pragma unroll
                for (int l = 0; l < NL; l++) {
                    const uint32_t live = (~blocked >> (10 * l)) & 0x3ffu;
                    if (live) {
                        const bool use_cone = !EXH && !force && ca[l] > 0.f;
                        const uint32_t hit = mesh_blocks_light(ma, prm.box_sph, start.x, start.y, start.z, Ax[l], Ay[l], Az[l],
                                                               ca[l], sa[l], use_cone, dirs + LS * l, live);
                        blocked |= hit << (10 * l);
                    }
                }
            }
            // ---- planes, then cubes (kernel.cu:1512-1536) for the rays nothing blocked yet; cubes outside a light's
            //      cone are skipped ----
            if ((prm.n_cubes | prm.n_planes) && valid) {
#pragma unroll
                for (int l = 0; l < NL; l++) {
                    const uint32_t live = (~blocked >> (10 * l)) & 0x3ffu;
                    if (live) {
                        const bool use_cone = !EXH && !force && ca[l] > 0.f;
                        const uint32_t hit = cubes_planes_block_light(prm.cubes, prm.n_cubes, prm.planes, prm.n_planes, start.x,
                                                                      start.y, start.z, Ax[l], Ay[l], Az[l], ca[l], sa[l],
                                                                      use_cone, dirs + LS * l, live);
                        blocked |= hit << (10 * l);
                    }
                }
            }

            // ---- light accumulation (kernel.cu:1537-1543, 1673-1675) ----
            if (valid) {
#pragma unroll
                for (int l = 0; l < NL; l++) {
                    if (l0 + l < prm.n_lights) {
                        float b = c_b_of_k[10 - __popc((blocked >> (l * 10)) & 0x3ffu)];
                        const float a = a_l[l];
                        b *= a > 0 ? a : 0;
                        const LightP L = prm.lights[l0 + l];
                        fr += b * L.r * tr;
                        fg += b * L.g * tg;
                        fb += b * L.b * tb;
                    }
                }
            }
        }
        if (valid) prm.pixels[o_out] = ref_rgb_to_int((int)(fr * 254.f), (int)(fg * 254.f), (int)(fb * 254.f));
        if (STAGED && prm.dbg_cycles && lane == 0 && blk < prm.dbg_cap)
            prm.dbg_cycles[prm.dbg_cap + blk] = (uint32_t)(clock64() - dbg_t0);
        if (!ahead) {
            if (lane == 0) wb_next = (uint32_t)atomicAdd(cursor, 1ull);
            wb_next = __shfl_sync(0xffffffffu, wb_next, 0);
        }
        wb = wb_next;
    }
    if (n_exact) atomicAdd(&prm.counters[CNT_EXACT_SHADOW], n_exact);
    if (lane == 0 && n_l1) atomicAdd(&prm.counters[CNT_BEAM_L1], (unsigned long long)n_l1);
    if (n_l2) atomicAdd(&prm.counters[CNT_BEAM_L2], (unsigned long long)n_l2);
}

// ------------------------------------------------------------------------------------
// count_reference_tests_kernel (ORE_FLAG_COUNT_REFERENCE_TESTS, never on the timed path)
// The reference's own any-hit loop, literally: one ray at a time, spheres in index order,
// exact sequence, break at the first hit (kernel.cu:1501-1510).  Sums the number of
// sphere::intersect calls that loop order makes - the "tests" of the FP32 roofline
// (SURVEY.md 8d).  One thread per (hit pixel, light).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CTA_THREADS) count_reference_tests_kernel(const FrameParams prm) {
    const uint32_t n_items = (uint32_t)prm.counters[CNT_HITS];
    const unsigned long long total = (unsigned long long)n_items * (unsigned long long)prm.n_lights;
    const v3 O0 = mk(prm.Ox, prm.Oy, prm.Oz);
    unsigned long long mine = 0;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < total;
         w += (unsigned long long)gridDim.x * blockDim.x) {
        // consecutive threads = consecutive hit pixels of one light (coherent loops)
        const uint32_t light = (uint32_t)(w / n_items), item = (uint32_t)(w % n_items);
        const uint32_t o = prm.hit_list[item];
        const int k = (int)(o / (uint32_t)prm.W), x = (int)(o % (uint32_t)prm.W);
        const v3 D = primary_dir(prm, prm.dx_tab[x], prm.dy_tab[k]);
        const float nt = prm.hit_t[o];
        const float4 sc = __ldg(&prm.sph_exact[prm.hit_id[o]]);
        const v3 new_org = ref_add(O0, ref_scale(D, nt));
        v3 normal = ref_sub(new_org, mk(sc.x, sc.y, sc.z));
        ref_normalise(normal);
        const v3 start = ref_add(ref_scale(normal, 0.00001f), new_org);
        float dir[30];
        light_directions(prm.lights[light], start, normal, dir);
#pragma unroll 1
        for (int j = 0; j < 10; j++) {
            const v3 d = mk(dir[j * 3 + 0], dir[j * 3 + 1], dir[j * 3 + 2]);
            int i = 0;
            bool shadow = false;
            for (; i < prm.n_spheres; i++) {
                const float4 s = __ldg(&prm.sph_exact[i]);
                float t;
                if (ref_intersect(start, d, s.x, s.y, s.z, s.w, t)) {
                    shadow = true;
                    break;
                }
            }
            mine += (unsigned long long)(shadow ? i + 1 : prm.n_spheres);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, d);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&prm.counters[CNT_SHADOW_TESTS_REF], mine);
}

// ------------------------------------------------------------------------------------
// fp32_burn_kernel: dependent-chain-free FFMA burn used to MEASURE the FP32 roofline
// denominator on the box (MEASURED_PEAKS.json carries HBM and bf16 peaks only).
// 16 independent accumulators per thread, 2 FLOP per FFMA.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CTA_THREADS) fp32_burn_kernel(float* out, int iters, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = (float)(threadIdx.x + i);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = fmaf(acc[i], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    if (s == 12345.678f) out[0] = s;  // never true; keeps the chain alive
}

// ------------------------------------------------------------------------------------
// libm_probe_kernel: evaluates the device libm used by the path on caller-supplied inputs (tests only)
// op: 0 cosf, 1 sinf, 2 acosf, 3 atan2f(y = a, x = b)
// ------------------------------------------------------------------------------------
__global__ void libm_probe_kernel(int op, int n, const float* a, const float* b, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r;
    if (op == 0)
        r = ORE_COSF(a[i]);
    else if (op == 1)
        r = ORE_SINF(a[i]);
    else if (op == 2)
        r = ORE_ACOSF(a[i]);
    else if (op == 3)
        r = ORE_ATAN2F(a[i], b[i]);
    else {
        // ops 4-6: component (op - 4) of ref_normalise((a[i], b[i], a[(i + 1) % n]))
        v3 v = mk(a[i], b[i], a[(i + 1) % n]);
        ref_normalise(v);
        r = op == 4 ? v.x : (op == 5 ? v.y : v.z);
    }
    out[i] = r;
}

}  // namespace ore
