// ore_capi.cu - C ABI (include/ore_render.h) over the sm_100a render kernels.
//
// Replaces the host half of the reference's kernel.cu for this path: the global scene
// set-up of onStart() (kernel.cu:1704-1714), object::sphereAllocMem (:1208-1212) and the
// per-frame body of update() (:1762-1792).  Where the reference allocates and frees
// managed memory every frame, the context keeps persistent SoA device buffers, a pinned
// host staging buffer for uploads / read-back and its own stream.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -shared
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <algorithm>
#include <utility>
#include <vector>

#include "../../include/ore_render.h"
#include "ore_clusters.h"
#include "ore_kernels.cuh"

using namespace ore;

// second instantiation of the default-path kernels with CUDA's libm (ore_fast.cu)
extern "C" int ore_fast_set_tables(const float* cphi, const float* sphi, const float* bk);
extern "C" int ore_fast_primary_tile(const void* prm, int sm_count, size_t smem, long long n_batches, int exh,
                                     cudaStream_t stream);
// stage: a StageArgs of identical layout; staged = 0 selects the fused form (only first_block is used then)
extern "C" int ore_fast_shadow_sweep(const void* prm, const void* stage, int staged, int sm_count, size_t smem, int exh,
                                     cudaStream_t stream);
extern "C" int ore_fast_shade_setup(const void* prm, const void* stage, int sm_count, cudaStream_t stream);

struct ore_context {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_rendered[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    // device framebuffers of the pipelined presentation: two sets (one being copied out while the other is rendered)
    // of `async_k` frames of `async_px` pixels each, one allocation
    uint32_t* async_pool = nullptr;
    size_t async_pool_cap = 0;   // pixels
    size_t async_px = 0;
    int async_k = 0;
    unsigned long long async_batches = 0;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ev_valid = false;
    bool ran_count = false;
    std::string err;

    // pinned staging (uploads and read-back)
    void* pinned = nullptr;
    size_t pinned_cap = 0;

    // scene (structure of arrays on the device)
    int n_spheres = 0;
    float4* sph_exact = nullptr;   // cx,cy,cz,radius member in ORIGINAL order (hit attributes by hit id)
    // the spheres in Morton order with two levels of bounding balls over them (ore_clusters.h)
    float4* sph_xsort = nullptr;   // exact records, sorted
    float4* sph_sort = nullptr;    // shadow records cx,cy,cz,R', sorted
    int* sort_index = nullptr;     // sorted position -> original index
    float4* leaf_sph = nullptr;    // ball of spheres [8j, 8j+8)
    float4* super_sph = nullptr;   // ball of leaves [32k, 32k+32)
    // per-frame (camera-space) records of the primary kernel
    float4* prim_sorted = nullptr;
    float4* cone_sorted = nullptr;
    float4* leaf_cone = nullptr;
    float4* super_cone = nullptr;
    int n_sort = 0, n_leaves = 0, n_leaves_pad = 0, n_supers = 0, n_supers_pad = 0;
    size_t sph_cap = 0, sort_cap = 0, leaf_cap = 0, super_cap = 0;
    int n_lights = 0;
    LightP lights[MAX_LIGHTS];
    float* tex[3] = {nullptr, nullptr, nullptr};
    int tex_w = 0, tex_h = 0;
    bool tex_finite = false;   // every texel of the object texture is finite (checked at upload; FrameParams::skip_dark)
    float* sky[3] = {nullptr, nullptr, nullptr};
    int sky_w = 0, sky_h = 0;
    float sky_radius = 0.f;
    float4* cubes = nullptr;   // 3 float4 per cube: bounds[0], bounds[1], orgin
    float4* planes = nullptr;  // 2 float4 per plane: orgin, normal
    int n_cubes = 0, n_planes = 0;
    float* tris = nullptr;     // 27 floats per triangle
    float4* boxes = nullptr;   // 2 float4 per leaf box
    float4* box_sph = nullptr; // bounding sphere per leaf box (cone filters)
    float4* box_cone = nullptr;  // per-frame tile-cone record per leaf box
    int* box_offsets = nullptr;
    int* box_indices = nullptr;
    int n_tris = 0, n_boxes = 0, mesh_has_normals = 0;

    // per-frame buffers
    float* dx_tab = nullptr;
    size_t dx_cap = 0;
    float* dy_tab = nullptr;
    size_t dy_cap = 0;
    uint32_t* hit_list = nullptr;  // compact hit records (pixel, id, t), one per hit pixel
    int32_t* hit_ids = nullptr;
    float* hit_ts = nullptr;
    uint32_t* pixels = nullptr;    // the context's own framebuffer (ore_render)
    size_t px_cap = 0, pixels_cap = 0;
    int32_t* hit_id_map = nullptr;  // per-pixel id / t maps: only built by ore_get_hits
    float* hit_t_map = nullptr;
    size_t map_cap = 0;
    unsigned long long* counters = nullptr;
    // band DMA (ORE_FLAG_BAND_DMA): packed local rows of the primary kernel + the stream / events of their copy
    uint32_t* band_buf = nullptr;
    size_t band_cap = 0;   // pixels
    cudaStream_t band_stream = nullptr;
    cudaEvent_t ev_band_go = nullptr, ev_band_done = nullptr;
    float* stage = nullptr;  // staging buffer between the two kernels of the default shadow pass
    size_t stage_cap = 0;    // floats
    size_t stage_blocks_override = 0;  // test hook (env ORE_STAGE_BLOCKS at ore_create): staging capacity in 32-item blocks
    size_t stage_max_items = (size_t)8 << 20;  // env ORE_STAGE_MAX_ITEMS: upper bound of the staging buffer in hit pixels
    bool no_memops = false;            // test hook (env ORE_NO_STREAM_MEMOPS=1): flags through the one-thread kernels
    std::string dbg_cycles_path;       // tools hook (env ORE_DEBUG_BLOCK_CYCLES=file): per-block SM clocks of the shadow pass
    uint32_t* dbg_cycles = nullptr;
    size_t dbg_cap = 0;
    // hit count of this context's previous frame, copied to pinned memory at the end of every frame and read
    // WITHOUT synchronisation by the next one: only a hint for how many staged chunk pairs to launch - whatever
    // lies beyond them is swept by one catch-all fused launch, so any value (stale, zero, mid-copy) is safe
    unsigned long long* hits_hint = nullptr;
    bool hits_hint_set = false;
    FrameParams last_prm;              // the last frame's parameters (ore_get_hits expands its hit records)
    bool last_prm_valid = false;

    // occupancy-sized grids, computed once per (kernel, dynamic shared memory, block size)
    struct GridEntry {
        const void* fn;
        size_t smem;
        int threads, grid;
    };
    std::vector<GridEntry> grid_cache;
    // end of the last render on whichever stream it ran (scene setters wait for it before touching scene buffers)
    cudaEvent_t ev_done = nullptr;
    bool ev_done_set = false;
    // ordered flag writes (ore_flag_write_after)
    cudaStream_t signal_stream = nullptr;
    cudaEvent_t ev_signal[8] = {};
    unsigned n_signals = 0;

    // last frame
    int last_frames = 1;
    size_t last_px = 0;
    int last_n_spheres = 0;
    uint64_t last_launches = 0;
};

#define ORE_CUDA(ctx, expr)                                                                         \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            char _b[512];                                                                           \
            snprintf(_b, sizeof _b, "CUDA error = %u at %s:%d '%s' (%s)", (unsigned)_e, __FILE__,   \
                     __LINE__, #expr, cudaGetErrorString(_e));                                      \
            (ctx)->err = _b;                                                                        \
            return ORE_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

static int fail(ore_context* ctx, int code, const char* msg) {
    if (ctx) ctx->err = msg;
    return code;
}

static int ensure_pinned(ore_context* ctx, size_t bytes) {
    if (bytes <= ctx->pinned_cap) return ORE_OK;
    if (ctx->pinned) ORE_CUDA(ctx, cudaFreeHost(ctx->pinned));
    ctx->pinned = nullptr;
    ctx->pinned_cap = 0;
    size_t cap = bytes + bytes / 4 + 4096;
    ORE_CUDA(ctx, cudaMallocHost(&ctx->pinned, cap));
    ctx->pinned_cap = cap;
    return ORE_OK;
}

template <typename T>
static int ensure_dev(ore_context* ctx, T** p, size_t* cap, size_t n) {
    if (n <= *cap && *p) return ORE_OK;
    if (*p) ORE_CUDA(ctx, cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    size_t want = n + n / 8 + 64;
    ORE_CUDA(ctx, cudaMalloc((void**)p, want * sizeof(T)));
    *cap = want;
    return ORE_OK;
}

// scene setters: the previous render may still be running on a caller-supplied stream
static int wait_last_render(ore_context* ctx) {
    if (ctx->ev_done_set) ORE_CUDA(ctx, cudaEventSynchronize(ctx->ev_done));
    return ORE_OK;
}

extern "C" int ore_abi_version(void) { return ORE_ABI_VERSION; }

extern "C" const char* ore_last_error(const ore_context* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int ore_create(ore_context** out, int device) {
    if (!out) return ORE_ERR_INVALID;
    *out = nullptr;
    ore_context* ctx = new (std::nothrow) ore_context();
    if (!ctx) return ORE_ERR_NOMEM;
    *out = ctx;  // returned even on failure so the caller can read ore_last_error
    ctx->device = device;
    ORE_CUDA(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    ORE_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        char b[160];
        snprintf(b, sizeof b, "device %d is sm_%d%d; this library only carries an sm_100a image (no fallback)", device,
                 prop.major, prop.minor);
        return fail(ctx, ORE_ERR_CUDA, b);
    }
    ctx->sm_count = prop.multiProcessorCount;
    ORE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ORE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        ORE_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_rendered[i], cudaEventDisableTiming));
        ORE_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 5; i++) ORE_CUDA(ctx, cudaEventCreate(&ctx->ev[i]));
    ORE_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming));
    ORE_CUDA(ctx, cudaMallocHost((void**)&ctx->hits_hint, sizeof(unsigned long long)));
    *ctx->hits_hint = 0;
    if (const char* e = getenv("ORE_NO_STREAM_MEMOPS")) ctx->no_memops = atoi(e) != 0;
    if (const char* e = getenv("ORE_STAGE_MAX_ITEMS")) {
        const long long v = atoll(e);
        if (v >= 32) ctx->stage_max_items = (size_t)v;
    }
    if (const char* e = getenv("ORE_DEBUG_BLOCK_CYCLES")) ctx->dbg_cycles_path = e;
    if (const char* e = getenv("ORE_STAGE_BLOCKS")) {
        const long v = atol(e);
        if (v > 0) ctx->stage_blocks_override = (size_t)v;
    }
    ORE_CUDA(ctx, cudaMalloc((void**)&ctx->counters, CNT_SLOTS * sizeof(unsigned long long)));
    ORE_CUDA(ctx, cudaMemset(ctx->counters, 0, CNT_SLOTS * sizeof(unsigned long long)));
    // kernel.cu:1454,1462-1463: phi = (float)j/10 * 2.f * 3.1415f; cosf(phi), sinf(phi)
    float cphi[10], sphi[10];
    for (int j = 0; j < 10; j++) {
        float phi = (float)j / 10 * 2.f * 3.1415f;
        cphi[j] = cosf(phi);
        sphi[j] = sinf(phi);
    }
    ORE_CUDA(ctx, cudaMemcpyToSymbol(c_cos_phi, cphi, sizeof cphi));
    ORE_CUDA(ctx, cudaMemcpyToSymbol(c_sin_phi, sphi, sizeof sphi));
    float bk[11];
    {
        float b = 0;
        bk[0] = b;
        for (int k = 1; k <= 10; k++) {
            b += 0.1;  // float += double, kernel.cu:1538
            bk[k] = b;
        }
    }
    ORE_CUDA(ctx, cudaMemcpyToSymbol(c_b_of_k, bk, sizeof bk));
    if (ore_fast_set_tables(cphi, sphi, bk)) return fail(ctx, ORE_ERR_CUDA, "constant tables of the fast-libm kernels");
    return ORE_OK;
}

extern "C" int ore_destroy(ore_context* ctx) {
    if (!ctx) return ORE_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->dbg_cycles) {
        // tools hook: [2][dbg_cap] uint32 of the LAST frame, preceded by dbg_cap as one uint64
        cudaDeviceSynchronize();
        std::vector<uint32_t> h(2 * ctx->dbg_cap);
        if (cudaMemcpy(h.data(), ctx->dbg_cycles, h.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess) {
            if (FILE* f = fopen(ctx->dbg_cycles_path.c_str(), "wb")) {
                const unsigned long long cap = ctx->dbg_cap;
                fwrite(&cap, 8, 1, f);
                fwrite(h.data(), 4, h.size(), f);
                fclose(f);
            }
        }
        cudaFree(ctx->dbg_cycles);
    }
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_rendered[i]) cudaEventDestroy(ctx->ev_rendered[i]);
        if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
    if (ctx->signal_stream) {
        cudaStreamSynchronize(ctx->signal_stream);
        for (auto& e : ctx->ev_signal)
            if (e) cudaEventDestroy(e);
        cudaStreamDestroy(ctx->signal_stream);
    }
    if (ctx->band_stream) {
        cudaStreamSynchronize(ctx->band_stream);
        cudaEventDestroy(ctx->ev_band_go);
        cudaEventDestroy(ctx->ev_band_done);
        cudaStreamDestroy(ctx->band_stream);
    }
    if (ctx->async_pool) cudaFree(ctx->async_pool);
    void* dev[] = {ctx->band_buf, ctx->stage, ctx->sph_exact, ctx->sph_xsort, ctx->sph_sort, ctx->sort_index, ctx->leaf_sph, ctx->super_sph,
                   ctx->prim_sorted, ctx->cone_sorted, ctx->leaf_cone, ctx->super_cone, ctx->tex[0], ctx->tex[1], ctx->tex[2],
                   ctx->sky[0], ctx->sky[1], ctx->sky[2], ctx->dx_tab, ctx->dy_tab, ctx->hit_list, ctx->hit_ids, ctx->hit_ts,
                   ctx->hit_id_map, ctx->hit_t_map, ctx->pixels, ctx->counters, ctx->cubes, ctx->planes, ctx->tris, ctx->boxes,
                   ctx->box_offsets, ctx->box_indices, ctx->box_sph, ctx->box_cone};
    for (void* p : dev)
        if (p) cudaFree(p);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->hits_hint) cudaFreeHost(ctx->hits_hint);
    for (int i = 0; i < 5; i++)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ORE_OK;
}

// ---- scene upload -----------------------------------------------------------------------

// Replaces object::sphereAllocMem: the n records go to the device twice - in their original order (hit attributes are
// looked up by hit id) and in the Morton order of their centres, with a bounding ball per leaf of 8 and per
// super-cluster of 32 leaves (csrc/ore_clusters.h, unit-tested on the CPU by tests/test_clusters_cpu.py).  Both
// the primary kernel (tile cones) and the shadow sweep (beams) descend that hierarchy; neither result depends on
// the order spheres are visited in (nearest hit: candidates are adjudicated by (t, original index)).
static int upload_spheres(ore_context* ctx, const float* src, size_t stride_floats, size_t first, int32_t n) {
    if (!ctx || n < 0 || (n > 0 && !src)) return fail(ctx, ORE_ERR_INVALID, "ore_set_spheres: bad arguments");
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int rcw = wait_last_render(ctx)) return rcw;
    std::vector<ore_host::Rec4> ex((size_t)std::max(n, 1)), sh((size_t)std::max(n, 1));
    for (int i = 0; i < n; i++) {
        const float* s = src + (size_t)i * stride_floats + first;
        ex[i] = ore_host::Rec4{s[0], s[1], s[2], s[3]};
        const float r4 = s[3] * s[3];  // the test squares the stored member, kernel.cu:334
        // R' = effective radius, rounded up so that R'^2 >= (1+k) r4 in float
        float rp = (float)sqrt((double)r4 * (1.0 + (double)ORE_KAPPA_SHADOW));
        rp = nextafterf(rp, INFINITY);
        while (rp * rp < (float)((double)r4 * (1.0 + (double)ORE_KAPPA_SHADOW))) rp = nextafterf(rp, INFINITY);
        sh[i] = ore_host::Rec4{s[0], s[1], s[2], rp};
    }
    std::vector<ore_host::Rec4> ss, xs, cl32, leaves, supers;
    std::vector<int> order;
    ore_host::build_clusters(ex.data(), sh.data(), n, ss, xs, cl32, &order);
    ore_host::build_hierarchy(ss, n, leaves, supers);
    static_assert(sizeof(ore_host::Rec4) == sizeof(float4), "Rec4 must have float4's layout");
    if (ss.empty()) {   // empty scene: one block of padding records (never read: every loop is bounded by n = 0)
        ss.assign(32, ore_host::Rec4{3e18f, -1e18f, 2e17f, 0.f});
        xs = ss;
        order.assign(32, 0);
    }
    const size_t n_sort = ss.size(), n_leaf_pad = leaves.size(), n_sup_pad = supers.size();
    int rc;
    size_t c0 = ctx->sph_cap;
    if ((rc = ensure_dev(ctx, &ctx->sph_exact, &c0, (size_t)std::max(n, 1)))) return rc;
    ctx->sph_cap = c0;
    size_t c1 = ctx->sort_cap, c2 = c1, c3 = c1, c4 = c1, c5 = c1;
    if ((rc = ensure_dev(ctx, &ctx->sph_xsort, &c1, n_sort))) return rc;
    if ((rc = ensure_dev(ctx, &ctx->sph_sort, &c2, n_sort))) return rc;
    if ((rc = ensure_dev(ctx, &ctx->sort_index, &c3, n_sort))) return rc;
    // camera-space records: one set per frame of a batch
    c4 *= MAX_BATCH;
    c5 *= MAX_BATCH;
    if ((rc = ensure_dev(ctx, &ctx->prim_sorted, &c4, n_sort * MAX_BATCH))) return rc;
    if ((rc = ensure_dev(ctx, &ctx->cone_sorted, &c5, n_sort * MAX_BATCH))) return rc;
    ctx->sort_cap = std::min(std::min(c1, c2), std::min(std::min(c3, c4 / MAX_BATCH), c5 / MAX_BATCH));
    size_t l1 = ctx->leaf_cap, l2 = l1;
    if ((rc = ensure_dev(ctx, &ctx->leaf_sph, &l1, n_leaf_pad))) return rc;
    l2 *= MAX_BATCH;
    if ((rc = ensure_dev(ctx, &ctx->leaf_cone, &l2, n_leaf_pad * MAX_BATCH))) return rc;
    ctx->leaf_cap = std::min(l1, l2 / MAX_BATCH);
    size_t s1 = ctx->super_cap, s2 = s1;
    if ((rc = ensure_dev(ctx, &ctx->super_sph, &s1, n_sup_pad))) return rc;
    s2 *= MAX_BATCH;
    if ((rc = ensure_dev(ctx, &ctx->super_cone, &s2, n_sup_pad * MAX_BATCH))) return rc;
    ctx->super_cap = std::min(s1, s2 / MAX_BATCH);
    // pinned staging: the five arrays back to back (16-byte records first), one H2D copy each
    const size_t b_ex = (size_t)std::max(n, 1) * 16, b_sort = n_sort * 16, b_leaf = n_leaf_pad * 16, b_sup = n_sup_pad * 16;
    if ((rc = ensure_pinned(ctx, b_ex + 2 * b_sort + b_leaf + b_sup + n_sort * sizeof(int)))) return rc;
    char* h = (char*)ctx->pinned;
    char* h_ex = h;
    char* h_xs = h_ex + b_ex;
    char* h_ss = h_xs + b_sort;
    char* h_lv = h_ss + b_sort;
    char* h_sp = h_lv + b_leaf;
    char* h_ix = h_sp + b_sup;
    memcpy(h_ex, ex.data(), b_ex);
    memcpy(h_xs, xs.data(), b_sort);
    memcpy(h_ss, ss.data(), b_sort);
    memcpy(h_lv, leaves.data(), b_leaf);
    memcpy(h_sp, supers.data(), b_sup);
    memcpy(h_ix, order.data(), n_sort * sizeof(int));
    ORE_CUDA(ctx, cudaMemcpyAsync(ctx->sph_exact, h_ex, b_ex, cudaMemcpyHostToDevice, ctx->stream));
    ORE_CUDA(ctx, cudaMemcpyAsync(ctx->sph_xsort, h_xs, b_sort, cudaMemcpyHostToDevice, ctx->stream));
    ORE_CUDA(ctx, cudaMemcpyAsync(ctx->sph_sort, h_ss, b_sort, cudaMemcpyHostToDevice, ctx->stream));
    ORE_CUDA(ctx, cudaMemcpyAsync(ctx->leaf_sph, h_lv, b_leaf, cudaMemcpyHostToDevice, ctx->stream));
    ORE_CUDA(ctx, cudaMemcpyAsync(ctx->super_sph, h_sp, b_sup, cudaMemcpyHostToDevice, ctx->stream));
    ORE_CUDA(ctx, cudaMemcpyAsync(ctx->sort_index, h_ix, n_sort * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    ORE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n_spheres = n;
    ctx->n_sort = (int)n_sort;
    ctx->n_leaves = (n + ore_host::LEAF_SPHERES - 1) / ore_host::LEAF_SPHERES;
    ctx->n_leaves_pad = (int)n_leaf_pad;
    ctx->n_supers = (ctx->n_leaves + ore_host::SUPER_LEAVES - 1) / ore_host::SUPER_LEAVES;
    ctx->n_supers_pad = (int)n_sup_pad;
    return ORE_OK;
}

extern "C" int ore_set_spheres(ore_context* ctx, const float* xyz_radius, int32_t n) {
    return upload_spheres(ctx, xyz_radius, 4, 0, n);
}

extern "C" int ore_set_spheres_aos32(ore_context* ctx, const void* records, int32_t n) {
    // reference record: vptr@0, orgin@8 (3 floats), reflective@20, radius@24, pad@28 (kernel.cu:265-358)
    // => floats 2,3,4 = centre; float 6 = radius.  Repack to x,y,z,radius through a temporary.
    if (!ctx || n < 0 || (n > 0 && !records)) return fail(ctx, ORE_ERR_INVALID, "ore_set_spheres_aos32: bad arguments");
    std::vector<float> tmp((size_t)n * 4);
    const float* f = (const float*)records;
    for (int i = 0; i < n; i++) {
        tmp[4 * (size_t)i + 0] = f[8 * (size_t)i + 2];
        tmp[4 * (size_t)i + 1] = f[8 * (size_t)i + 3];
        tmp[4 * (size_t)i + 2] = f[8 * (size_t)i + 4];
        tmp[4 * (size_t)i + 3] = f[8 * (size_t)i + 6];
    }
    return upload_spheres(ctx, tmp.data(), 4, 0, n);
}

extern "C" int ore_set_cubes(ore_context* ctx, const float* c1_c2, int32_t n) {
    if (!ctx || n < 0 || (n > 0 && !c1_c2)) return fail(ctx, ORE_ERR_INVALID, "ore_set_cubes: bad arguments");
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int rcw = wait_last_render(ctx)) return rcw;
    if (ctx->cubes) ORE_CUDA(ctx, cudaFree(ctx->cubes));
    ctx->cubes = nullptr;
    ctx->n_cubes = 0;
    if (n == 0) return ORE_OK;
    int rc;
    if ((rc = ensure_pinned(ctx, (size_t)n * 3 * sizeof(float4)))) return rc;
    float4* h = (float4*)ctx->pinned;
    for (int i = 0; i < n; i++) {
        const float* c = c1_c2 + 6 * (size_t)i;
        h[3 * i + 0] = make_float4(c[0], c[1], c[2], 0.f);  // bounds[0] = c1, kernel.cu:393
        h[3 * i + 1] = make_float4(c[3], c[4], c[5], 0.f);  // bounds[1] = c2
        // orgin = divide(add(c1, c2), 2), kernel.cu:395 (float add, float divide)
        // .w = radius of a bounding sphere about orgin (half diagonal + margins) for the light-cone filter
        const double dx = (double)c[3] - c[0], dy = (double)c[4] - c[1], dz = (double)c[5] - c[2];
        const double rad = 0.5 * sqrt(dx * dx + dy * dy + dz * dz);
        const double mag = fabs((double)c[0]) + fabs((double)c[1]) + fabs((double)c[2]) + fabs((double)c[3]) +
                           fabs((double)c[4]) + fabs((double)c[5]);
        h[3 * i + 2] = make_float4((c[0] + c[3]) / 2, (c[1] + c[4]) / 2, (c[2] + c[5]) / 2,
                                   (float)(rad * 1.001 + 1e-5 * mag + 1e-6));
    }
    ORE_CUDA(ctx, cudaMalloc((void**)&ctx->cubes, (size_t)n * 3 * sizeof(float4)));
    ORE_CUDA(ctx, cudaMemcpyAsync(ctx->cubes, h, (size_t)n * 3 * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    ORE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n_cubes = n;
    return ORE_OK;
}

extern "C" int ore_set_planes(ore_context* ctx, const float* pos_normal, int32_t n) {
    if (!ctx || n < 0 || (n > 0 && !pos_normal)) return fail(ctx, ORE_ERR_INVALID, "ore_set_planes: bad arguments");
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int rcw = wait_last_render(ctx)) return rcw;
    if (ctx->planes) ORE_CUDA(ctx, cudaFree(ctx->planes));
    ctx->planes = nullptr;
    ctx->n_planes = 0;
    if (n == 0) return ORE_OK;
    int rc;
    if ((rc = ensure_pinned(ctx, (size_t)n * 2 * sizeof(float4)))) return rc;
    float4* h = (float4*)ctx->pinned;
    for (int i = 0; i < n; i++) {
        const float* c = pos_normal + 6 * (size_t)i;
        h[2 * i + 0] = make_float4(c[0], c[1], c[2], 0.f);
        h[2 * i + 1] = make_float4(c[3], c[4], c[5], 0.f);
    }
    ORE_CUDA(ctx, cudaMalloc((void**)&ctx->planes, (size_t)n * 2 * sizeof(float4)));
    ORE_CUDA(ctx, cudaMemcpyAsync(ctx->planes, h, (size_t)n * 2 * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
    ORE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n_planes = n;
    return ORE_OK;
}

extern "C" int ore_set_mesh(ore_context* ctx, const float* tris27, int32_t n_tris, int32_t has_normals,
                            const float* box_bounds6, const int32_t* box_offsets, const int32_t* box_indices,
                            int32_t n_boxes) {
    if (!ctx || n_tris < 0 || n_boxes < 0) return fail(ctx, ORE_ERR_INVALID, "ore_set_mesh: bad arguments");
    if (n_tris > 0 && n_boxes > 0 && (!tris27 || !box_bounds6 || !box_offsets || !box_indices))
        return fail(ctx, ORE_ERR_INVALID, "ore_set_mesh: null array");
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int rcw = wait_last_render(ctx)) return rcw;
    void* old[] = {ctx->tris, ctx->boxes, ctx->box_offsets, ctx->box_indices, ctx->box_sph, ctx->box_cone};
    for (void* p : old)
        if (p) ORE_CUDA(ctx, cudaFree(p));
    ctx->tris = nullptr;
    ctx->boxes = nullptr;
    ctx->box_offsets = nullptr;
    ctx->box_indices = nullptr;
    ctx->box_sph = nullptr;
    ctx->box_cone = nullptr;
    ctx->n_tris = ctx->n_boxes = ctx->mesh_has_normals = 0;
    if (n_tris == 0 || n_boxes == 0) return ORE_OK;
    const int n_idx = box_offsets[n_boxes];
    if (box_offsets[0] != 0 || n_idx < 0) return fail(ctx, ORE_ERR_INVALID, "ore_set_mesh: bad offsets");
    for (int j = 0; j < n_boxes; j++)
        if (box_offsets[j + 1] < box_offsets[j]) return fail(ctx, ORE_ERR_INVALID, "ore_set_mesh: offsets must ascend");
    for (int k = 0; k < n_idx; k++)
        if (box_indices[k] < 0 || box_indices[k] >= n_tris) return fail(ctx, ORE_ERR_INVALID, "ore_set_mesh: triangle index out of range");
    const size_t tb = (size_t)n_tris * 27 * sizeof(float), bb = (size_t)n_boxes * 2 * sizeof(float4);
    const size_t ob = (size_t)(n_boxes + 1) * sizeof(int), ib = (size_t)(n_idx > 0 ? n_idx : 1) * sizeof(int);
    int rc;
    const size_t sb = (size_t)n_boxes * sizeof(float4);
    // pinned staging layout: the two float4 sections first (16-byte aligned: bb and sb are multiples of 16 and the
    // pinned base is page aligned), then the 4-byte-aligned sections
    const size_t off_b = 0, off_s = bb, off_t = bb + sb, off_o = off_t + tb, off_i = off_o + ob;
    if ((rc = ensure_pinned(ctx, off_i + ib))) return rc;
    char* h = (char*)ctx->pinned;
    float4* hb = (float4*)(h + off_b);
    float4* hs = (float4*)(h + off_s);
    memcpy(h + off_t, tris27, tb);
    for (int j = 0; j < n_boxes; j++) {
        const float* b = box_bounds6 + 6 * (size_t)j;
        hb[2 * j] = make_float4(b[0], b[1], b[2], 0.f);
        hb[2 * j + 1] = make_float4(b[3], b[4], b[5], 0.f);
        // bounding sphere for the conservative cone filters: centre, half diagonal + 0.1 % + a little absolute slack
        // (the float slab test can accept a ray that misses the true box by rounding)
        const double cx = 0.5 * ((double)b[0] + b[3]), cy = 0.5 * ((double)b[1] + b[4]), cz = 0.5 * ((double)b[2] + b[5]);
        const double dx = (double)b[3] - b[0], dy = (double)b[4] - b[1], dz = (double)b[5] - b[2];
        const double rad = 0.5 * sqrt(dx * dx + dy * dy + dz * dz);
        const double mag = fabs(cx) + fabs(cy) + fabs(cz) + rad;
        hs[j] = make_float4((float)cx, (float)cy, (float)cz, (float)(rad * 1.001 + 1e-5 * mag + 1e-6));
    }
    memcpy(h + off_o, box_offsets, ob);
    memcpy(h + off_i, box_indices, (size_t)n_idx * sizeof(int));
    ORE_CUDA(ctx, cudaMalloc((void**)&ctx->tris, tb));
    ORE_CUDA(ctx, cudaMalloc((void**)&ctx->boxes, bb));
    ORE_CUDA(ctx, cudaMalloc((void**)&ctx->box_offsets, ob));
    ORE_CUDA(ctx, cudaMalloc((void**)&ctx->box_indices, ib));
    ORE_CUDA(ctx, cudaMalloc((void**)&ctx->box_sph, sb));
    ORE_CUDA(ctx, cudaMalloc((void**)&ctx->box_cone, sb * MAX_BATCH));
    ORE_CUDA(ctx, cudaMemcpyAsync(ctx->box_sph, hs, sb, cudaMemcpyHostToDevice, ctx->stream));
    ORE_CUDA(ctx, cudaMemcpyAsync(ctx->tris, h + off_t, tb, cudaMemcpyHostToDevice, ctx->stream));
    ORE_CUDA(ctx, cudaMemcpyAsync(ctx->boxes, hb, bb, cudaMemcpyHostToDevice, ctx->stream));
    ORE_CUDA(ctx, cudaMemcpyAsync(ctx->box_offsets, h + off_o, ob, cudaMemcpyHostToDevice, ctx->stream));
    ORE_CUDA(ctx, cudaMemcpyAsync(ctx->box_indices, h + off_i, ib, cudaMemcpyHostToDevice, ctx->stream));
    ORE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n_tris = n_tris;
    ctx->n_boxes = n_boxes;
    ctx->mesh_has_normals = has_normals ? 1 : 0;
    return ORE_OK;
}

extern "C" int ore_set_lights(ore_context* ctx, const float* lights7, int32_t n) {
    if (!ctx || n < 0 || n > MAX_LIGHTS || (n > 0 && !lights7))
        return fail(ctx, ORE_ERR_INVALID, "ore_set_lights: 0..16 lights of 7 floats");
    // lights travel as kernel arguments: nothing on the device to wait for
    for (int i = 0; i < n; i++) {
        const float* l = lights7 + 7 * (size_t)i;
        ctx->lights[i] = LightP{l[0], l[1], l[2], l[3], l[4], l[5], l[6]};
    }
    ctx->n_lights = n;
    return ORE_OK;
}

static int upload_planes(ore_context* ctx, float* dst[3], const float* r, const float* g, const float* b, int32_t w,
                         int32_t h) {
    if (!ctx || w <= 0 || h <= 0 || !r || !g || !b) return fail(ctx, ORE_ERR_INVALID, "sprite planes: bad arguments");
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int rcw = wait_last_render(ctx)) return rcw;
    const size_t n = (size_t)w * h;
    int rc;
    if ((rc = ensure_pinned(ctx, n * sizeof(float)))) return rc;
    const float* src[3] = {r, g, b};
    for (int c = 0; c < 3; c++) {
        if (dst[c]) ORE_CUDA(ctx, cudaFree(dst[c]));
        dst[c] = nullptr;
        ORE_CUDA(ctx, cudaMalloc((void**)&dst[c], n * sizeof(float)));
        memcpy(ctx->pinned, src[c], n * sizeof(float));
        ORE_CUDA(ctx, cudaMemcpyAsync(dst[c], ctx->pinned, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        ORE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return ORE_OK;
}

extern "C" int ore_set_texture(ore_context* ctx, const float* r, const float* g, const float* b, int32_t width,
                               int32_t height) {
    int rc = upload_planes(ctx, ctx ? ctx->tex : nullptr, r, g, b, width, height);
    if (rc) return rc;
    ctx->tex_w = width;
    ctx->tex_h = height;
    // (0 x texel must be 0 for lights that do not face a pixel to be skippable: see FrameParams::skip_dark)
    bool finite = true;
    const size_t n = (size_t)width * height;
    const float* planes[3] = {r, g, b};
    for (int c = 0; c < 3 && finite; c++)
        for (size_t i = 0; i < n; i++)
            if (!std::isfinite(planes[c][i])) {
                finite = false;
                break;
            }
    ctx->tex_finite = finite;
    return ORE_OK;
}

extern "C" int ore_set_sky(ore_context* ctx, const float* r, const float* g, const float* b, int32_t width,
                           int32_t height, float size) {
    int rc = upload_planes(ctx, ctx ? ctx->sky : nullptr, r, g, b, width, height);
    if (rc) return rc;
    ctx->sky_w = width;
    ctx->sky_h = height;
    ctx->sky_radius = size * size;  // sphere ctor stores r*r (kernel.cu:287)
    return ORE_OK;
}

// ---- render -------------------------------------------------------------------------------

static constexpr int TILE_ROWS = TILE_P;   // rows per tile of the primary kernel
// shared memory of the primary kernel: the leaf / super-cluster cone records always, the sphere-level records when
// everything stays within this budget (6 CTAs of 128 threads per SM)
static constexpr size_t PRIMARY_RESIDENT_BYTES = 30 * 1024;

template <typename K>
static int grid_for(ore_context* ctx, K kernel, size_t smem, int* grid, int threads = CTA_THREADS) {
    const void* fn = reinterpret_cast<const void*>(kernel);
    for (const auto& e : ctx->grid_cache)
        if (e.fn == fn && e.smem == smem && e.threads == threads) {
            *grid = e.grid;
            return ORE_OK;
        }
    int occ = 0;
    // the opt-in limit only ever grows, so that earlier (cached) configurations of this kernel stay launchable
    size_t limit = smem;
    for (const auto& e : ctx->grid_cache)
        if (e.fn == fn && e.smem > limit) limit = e.smem;
    ORE_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
    ORE_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
    if (occ < 1) return fail(ctx, ORE_ERR_CUDA, "kernel does not fit on an SM");
    *grid = occ * ctx->sm_count;
    ctx->grid_cache.push_back({fn, smem, threads, *grid});
    return ORE_OK;
}

// rows y0 + m*y_step + j (j < y_block) below y1
static int frame_rows(const ore_frame* fr) {
    const int yb = fr->y_block > 0 ? fr->y_block : 1;
    const int span = fr->y1 - fr->y0;
    if (span <= 0 || fr->y_step <= 0) return 0;
    if (yb >= fr->y_step) return span;  // contiguous
    const int full = span / fr->y_step, rem = span % fr->y_step;
    return full * yb + (rem < yb ? rem : yb);
}

// one launch of the sweep kernel (stage B / fused / catch-all)
static int launch_sweep(ore_context* ctx, const FrameParams& prm, const StageArgs& st, bool staged, bool exh, bool fast_libm,
                        cudaStream_t stream) {
    int rc, grid = 0;
    const size_t smem = SWEEP_SMEM_BYTES;
    if (fast_libm) {
        ORE_CUDA(ctx, (cudaError_t)ore_fast_shadow_sweep(&prm, &st, staged ? 1 : 0, ctx->sm_count, smem, exh ? 1 : 0, stream));
        return ORE_OK;
    }
#define ORE_SWEEP(E, S)                                                                                  \
    do {                                                                                                 \
        if ((rc = grid_for(ctx, shadow_sweep_kernel<E, S>, smem, &grid, SWEEP_THREADS))) return rc;      \
        shadow_sweep_kernel<E, S><<<grid, SWEEP_THREADS, smem, stream>>>(prm, st);                       \
    } while (0)
    if (staged) {
        if (exh) ORE_SWEEP(true, true); else ORE_SWEEP(false, true);
    } else {
        if (exh) ORE_SWEEP(true, false); else ORE_SWEEP(false, false);
    }
#undef ORE_SWEEP
    ORE_CUDA(ctx, cudaGetLastError());
    return ORE_OK;
}

// One launch set for a batch of n_frames cameras.  outs: n_frames device framebuffers, or null (n_frames == 1 only): the
// context's own framebuffer.
static int render_impl(ore_context* ctx, const ore_camera* cams, int n_frames, const ore_frame* fr, uint32_t* const* outs,
                       cudaStream_t stream) {
    if (!ctx) return ORE_ERR_INVALID;
    if (!cams || !fr) return fail(ctx, ORE_ERR_INVALID, "ore_render: null camera/frame");
    if (n_frames < 1 || n_frames > MAX_BATCH || (!outs && n_frames != 1))
        return fail(ctx, ORE_ERR_INVALID, "ore_render: a batch holds 1..8 frames");
    // a band may be empty (y0 >= y1: a rank with no rows of a short frame) but never reaches past the image
    if (fr->width <= 0 || fr->height <= 0 || fr->y_step <= 0 || fr->y0 < 0 || fr->y1 < 0 || fr->y1 > fr->height ||
        (fr->out_pitch != 0 && fr->out_pitch < fr->width))
        return fail(ctx, ORE_ERR_INVALID, "ore_render: bad frame geometry");
    if (fr->flags & ~(uint32_t)(ORE_FLAG_EXHAUSTIVE | ORE_FLAG_COUNT_REFERENCE_TESTS | ORE_FLAG_FAST_LIBM | ORE_FLAG_FUSED_SHADOW |
                                ORE_FLAG_NO_KERNEL_TIMING | ORE_FLAG_BAND_DMA))
        return fail(ctx, ORE_ERR_INVALID, "ore_render: unknown flag");
    if (!ctx->tex[0] || !ctx->sky[0]) return fail(ctx, ORE_ERR_INVALID, "ore_render: texture and sky must be set first");
    if (!ctx->sph_exact) {
        int rc = upload_spheres(ctx, nullptr, 4, 0, 0);
        if (rc) return rc;
    }
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    const int W = fr->width;
    const int yb = fr->y_block > 0 ? fr->y_block : 1;
    if (yb > fr->y_step && fr->y_step != 1) return fail(ctx, ORE_ERR_INVALID, "ore_render: y_block must not exceed y_step");
    const int n_rows = frame_rows(fr);
    const size_t n_px_frame = (size_t)n_rows * W;
    const size_t n_px = n_px_frame * (size_t)n_frames;   // pixels of the whole batch
    if (n_px_frame >= (size_t)1 << 31 || n_px >= ((size_t)1 << 32) - 64) return fail(ctx, ORE_ERR_INVALID, "ore_render: band / batch too large");
    ctx->last_px = n_px_frame;
    ctx->last_frames = n_frames;
    ctx->last_n_spheres = ctx->n_spheres;
    ctx->last_launches = 0;
    ctx->ev_valid = false;
    ctx->ran_count = false;
    ctx->last_prm_valid = false;
    if (n_px == 0) return ORE_OK;

    int rc;
    const int W_pad = (W + 31) & ~31;
    if ((rc = ensure_dev(ctx, &ctx->dx_tab, &ctx->dx_cap, (size_t)W_pad))) return rc;
    if ((rc = ensure_dev(ctx, &ctx->dy_tab, &ctx->dy_cap, (size_t)n_rows))) return rc;
    if (n_px > ctx->px_cap || !ctx->hit_list) {
        if ((rc = wait_last_render(ctx))) return rc;   // (an earlier frame may still be running on another stream)
        size_t c1 = ctx->hit_list ? ctx->px_cap : 0, c2 = c1, c3 = c1;
        if ((rc = ensure_dev(ctx, &ctx->hit_list, &c1, n_px))) return rc;
        if ((rc = ensure_dev(ctx, &ctx->hit_ids, &c2, n_px))) return rc;
        if ((rc = ensure_dev(ctx, &ctx->hit_ts, &c3, n_px))) return rc;
        ctx->px_cap = std::min(std::min(c1, c2), c3);
    }
    if (!outs && (n_px_frame > ctx->pixels_cap || !ctx->pixels)) {
        // the context's own framebuffer: only ore_render (one frame, host output) renders into it
        if ((rc = wait_last_render(ctx))) return rc;
        if ((rc = ensure_dev(ctx, &ctx->pixels, &ctx->pixels_cap, n_px_frame))) return rc;
    }

    FrameParams prm;
    memset(&prm, 0, sizeof prm);
    prm.W = W;
    prm.W_pad = W_pad;
    prm.H = fr->height;
    prm.y0 = fr->y0;
    prm.y_step = (yb >= fr->y_step) ? 1 : fr->y_step;
    prm.n_rows = n_rows;
    prm.n_frames = n_frames;
    prm.n_px_frame = (uint32_t)n_px_frame;
    prm.y_block = (yb >= fr->y_step) ? 1 : yb;
    prm.out_global = (outs && fr->out_pitch > 0) ? 1 : 0;
    prm.pitch = prm.out_global ? fr->out_pitch : W;
    prm.n_spheres = ctx->n_spheres;
    prm.n_lights = ctx->n_lights;
    prm.flags = fr->flags;
    prm.aspect = fr->aspect;
    prm.ez = (-1 / fr->aspect);  // kernel.cu:1629
    prm.fz = 0.f - prm.ez;
    for (int f = 0; f < n_frames; f++) {
        const ore_camera* cam = &cams[f];
        CamP& c = prm.cam[f];
        c.Ox = 0.f + cam->org[0];  // add(eyePos, cam.Org), kernel.cu:1631
        c.Oy = 0.f + cam->org[1];
        c.Oz = prm.ez + cam->org[2];
        // camera::rotateDir, kernel.cu:249-255: frame-uniform, evaluated once on the host
        float yawRad = cam->yaw * (3.1415 / 180);
        float pitchRad = cam->pitch * (3.1415 / 180);
        c.cp = cosf(pitchRad);
        c.sp = sinf(pitchRad);
        c.cy = cosf(yawRad);
        c.sy = sinf(yawRad);
        prm.pixels[f] = outs ? outs[f] : ctx->pixels;
        if (!prm.pixels[f]) return fail(ctx, ORE_ERR_INVALID, "ore_render: null framebuffer in the batch");
        prm.sky_pixels[f] = prm.pixels[f];
    }
    prm.sky_pitch = prm.pitch;
    prm.sky_global = prm.out_global;
    // Band DMA (include/ore_render.h; opt-in): rows whose 8-row blocks are contiguous at the destination (pitch == width)
    const bool band_dma = prm.out_global && fr->out_pitch == W && (fr->flags & ORE_FLAG_BAND_DMA) != 0;
    if (band_dma) {
        if (n_px > ctx->band_cap || !ctx->band_buf) {
            if ((rc = wait_last_render(ctx))) return rc;
            if ((rc = ensure_dev(ctx, &ctx->band_buf, &ctx->band_cap, n_px))) return rc;
        }
        if (!ctx->band_stream) {
            ORE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->band_stream, cudaStreamNonBlocking));
            ORE_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_band_go, cudaEventDisableTiming));
            ORE_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_band_done, cudaEventDisableTiming));
        }
        for (int f = 0; f < n_frames; f++) prm.sky_pixels[f] = ctx->band_buf + (size_t)f * n_px_frame;
        prm.sky_pitch = W;
        prm.sky_global = 0;
    }
    {
        // largest half-angle of a 32 x TILE_ROWS pixel tile seen from the eye (tile cone of the primary
        // kernel): |v_pixel - v_centre| <= halfdiag on the image plane and |v| >= fz, so sin(a) <= halfdiag/fz
        const double delta = 2.0 * (double)fr->aspect / (double)W;
        // tallest image-row span of TILE_ROWS consecutive RENDERED rows (rows come in blocks of y_block)
        int max_span = 0;
        {
            const int B = prm.y_block, S = prm.y_step;
            for (int k0 = 0; k0 < B * TILE_ROWS; k0 += TILE_ROWS) {
                const int k1 = k0 + TILE_ROWS - 1;
                const int sp = ((k1 / B) * S + k1 % B) - ((k0 / B) * S + k0 % B);
                if (sp > max_span) max_span = sp;
            }
        }
        const double hx = 16.0 * delta, hy = (0.5 * max_span + 0.5) * delta;
        const double xr = sqrt(hx * hx + hy * hy) / fabs((double)prm.fz);
        prm.px_delta = (float)delta;
        if (!(xr < 0.9)) {
            prm.tile_ca = -1e20f;  // tiles too wide for a cone: every sphere is a candidate
            prm.tile_sa = 0.f;
        } else {
            const double a = asin(xr) * 1.001 + 1e-6;
            prm.tile_ca = (float)(cos(a) * (1.0 - 1e-6) - 1e-6);
            prm.tile_sa = (float)(sin(a) * (1.0 + 1e-6) + 1e-6);
        }
    }
    prm.dx_tab = ctx->dx_tab;
    prm.dy_tab = ctx->dy_tab;
    prm.sph_exact = ctx->sph_exact;
    prm.sph_xsort = ctx->sph_xsort;
    prm.sph_sort = ctx->sph_sort;
    prm.sort_index = ctx->sort_index;
    prm.leaf_sph = ctx->leaf_sph;
    prm.super_sph = ctx->super_sph;
    prm.n_sort = ctx->n_sort;
    prm.n_leaves = ctx->n_leaves;
    prm.n_leaves_pad = ctx->n_leaves_pad;
    prm.n_supers = ctx->n_supers;
    prm.n_supers_pad = ctx->n_supers_pad;
    prm.prim_sorted = ctx->prim_sorted;
    prm.cone_sorted = ctx->cone_sorted;
    prm.leaf_cone = ctx->leaf_cone;
    prm.super_cone = ctx->super_cone;
    const size_t tree_bytes = ((size_t)ctx->n_supers_pad + (size_t)ctx->n_leaves_pad) * sizeof(float4);
    const size_t all_bytes = tree_bytes + (size_t)ctx->n_sort * sizeof(float4);
    prm.cone_resident = all_bytes <= PRIMARY_RESIDENT_BYTES ? 1 : 0;
    const size_t psmem = prm.cone_resident ? all_bytes : tree_bytes;
    if (psmem > 200 * 1024) return fail(ctx, ORE_ERR_INVALID, "ore_render: more than ~100 000 spheres are not supported");
    prm.tex_r = ctx->tex[0];
    prm.tex_g = ctx->tex[1];
    prm.tex_b = ctx->tex[2];
    prm.tex_w = ctx->tex_w;
    prm.tex_h = ctx->tex_h;
    prm.sky_r = ctx->sky[0];
    prm.sky_g = ctx->sky[1];
    prm.sky_b = ctx->sky[2];
    prm.sky_w = ctx->sky_w;
    prm.sky_h = ctx->sky_h;
    prm.sky_radius = ctx->sky_radius;
    prm.cubes = ctx->cubes;
    prm.planes = ctx->planes;
    prm.n_cubes = ctx->n_cubes;
    prm.n_planes = ctx->n_planes;
    prm.tris = ctx->tris;
    prm.boxes = ctx->boxes;
    prm.box_offsets = ctx->box_offsets;
    prm.box_indices = ctx->box_indices;
    prm.box_sph = ctx->box_sph;
    prm.box_cone = ctx->box_cone;
    prm.n_tris = ctx->n_tris;
    prm.n_boxes = ctx->n_boxes;
    prm.mesh_has_normals = ctx->mesh_has_normals;
    {
        bool skip = ctx->tex_finite && !(fr->flags & ORE_FLAG_EXHAUSTIVE);
        for (int i = 0; i < ctx->n_lights && skip; i++)
            skip = std::isfinite(ctx->lights[i].r) && std::isfinite(ctx->lights[i].g) && std::isfinite(ctx->lights[i].b);
        prm.skip_dark = skip ? 1 : 0;
    }
    prm.hit_list = ctx->hit_list;
    prm.hit_ids = ctx->hit_ids;
    prm.hit_ts = ctx->hit_ts;
    prm.counters = ctx->counters;
    if (!ctx->dbg_cycles_path.empty()) {
        const size_t nb = (n_px + 31) / 32;
        if (nb > ctx->dbg_cap) {
            if (ctx->dbg_cycles) ORE_CUDA(ctx, cudaFree(ctx->dbg_cycles));
            ORE_CUDA(ctx, cudaMalloc((void**)&ctx->dbg_cycles, 2 * nb * sizeof(uint32_t)));
            ctx->dbg_cap = nb;
        }
        ORE_CUDA(ctx, cudaMemsetAsync(ctx->dbg_cycles, 0, 2 * ctx->dbg_cap * sizeof(uint32_t), stream));
        prm.dbg_cycles = ctx->dbg_cycles;
        prm.dbg_cap = (uint32_t)ctx->dbg_cap;
    }
    for (int i = 0; i < ctx->n_lights; i++) prm.lights[i] = ctx->lights[i];

    const bool timing = !(fr->flags & ORE_FLAG_NO_KERNEL_TIMING);
    if (timing) ORE_CUDA(ctx, cudaEventRecord(ctx->ev[0], stream));
    {
        int m = W_pad > n_rows ? W_pad : n_rows;
        m = std::max(m, std::max(ctx->n_sort, std::max(ctx->n_leaves_pad, ctx->n_supers_pad)));
        m = std::max(m, std::max(ctx->n_boxes, (int)CNT_SLOTS));
        prep_frame_kernel<<<(m + 255) / 256, 256, 0, stream>>>(prm);
        ORE_CUDA(ctx, cudaGetLastError());
        ctx->last_launches++;
    }
    if (timing) ORE_CUDA(ctx, cudaEventRecord(ctx->ev[1], stream));
    const bool exh = (fr->flags & ORE_FLAG_EXHAUSTIVE) != 0;
    const bool fast_libm = (fr->flags & ORE_FLAG_FAST_LIBM) != 0;
    {
        int grid = 0;
        const long long tiles = (long long)((W + 31) / 32) * ((n_rows + TILE_ROWS - 1) / TILE_ROWS);
        const long long n_batches = ((tiles + PRIMARY_WARPS - 1) / PRIMARY_WARPS) * n_frames;
        if (fast_libm) {
            ORE_CUDA(ctx, (cudaError_t)ore_fast_primary_tile(&prm, ctx->sm_count, psmem, n_batches, exh ? 1 : 0, stream));
        } else if (exh) {
            if ((rc = grid_for(ctx, primary_tile_kernel<TILE_ROWS, true>, psmem, &grid, PRIMARY_THREADS))) return rc;
            if (grid > n_batches) grid = (int)n_batches;
            primary_tile_kernel<TILE_ROWS, true><<<grid, PRIMARY_THREADS, psmem, stream>>>(prm);
        } else {
            if ((rc = grid_for(ctx, primary_tile_kernel<TILE_ROWS, false>, psmem, &grid, PRIMARY_THREADS))) return rc;
            if (grid > n_batches) grid = (int)n_batches;
            primary_tile_kernel<TILE_ROWS, false><<<grid, PRIMARY_THREADS, psmem, stream>>>(prm);
        }
        ORE_CUDA(ctx, cudaGetLastError());
        ctx->last_launches++;
    }
    if (timing) ORE_CUDA(ctx, cudaEventRecord(ctx->ev[2], stream));
    bool band_pending = false;
    if (band_dma) {
        // the rows the primary kernel has just written (sky pixels, 0 for hit pixels) travel to their place in the
        // destination frames on a copy engine, behind the primary kernel and concurrently with stage A of the shadow
        // pass; the sweep, which stores the hit pixels into the same rows, waits for them
        ORE_CUDA(ctx, cudaEventRecord(ctx->ev_band_go, stream));
        ORE_CUDA(ctx, cudaStreamWaitEvent(ctx->band_stream, ctx->ev_band_go, 0));
        const int B = prm.y_block, S = prm.y_step;           // (1, 1 for a contiguous band)
        const int full = n_rows / B, rem = n_rows % B;
        for (int f = 0; f < n_frames; f++) {
            const uint32_t* src = ctx->band_buf + (size_t)f * n_px_frame;
            uint32_t* dst = outs[f];
            if (full > 0)
                ORE_CUDA(ctx, cudaMemcpy2DAsync(dst, (size_t)S * W * sizeof(uint32_t), src, (size_t)B * W * sizeof(uint32_t),
                                                (size_t)B * W * sizeof(uint32_t), (size_t)full, cudaMemcpyDefault, ctx->band_stream));
            if (rem > 0)
                ORE_CUDA(ctx, cudaMemcpyAsync(dst + (size_t)full * S * W, src + (size_t)full * B * W,
                                              (size_t)rem * W * sizeof(uint32_t), cudaMemcpyDefault, ctx->band_stream));
        }
        ORE_CUDA(ctx, cudaEventRecord(ctx->ev_band_done, ctx->band_stream));
        band_pending = true;
    }
    auto join_band = [&]() -> int {
        if (band_pending) {
            ORE_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->ev_band_done, 0));
            band_pending = false;
        }
        return ORE_OK;
    };
    {
        // Two-stage pass (default): shade_setup_kernel -> staging buffer -> staged shadow_sweep_kernel.  The host does
        // not know the hit count (no sync): the first frame of a context sizes the staging buffer from the pixel count,
        // later frames from the previous frame's hit count (a hint read without synchronisation) - normally ONE chunk
        // pair per frame -, and one catch-all launch of the fused sweep takes whatever lies beyond the staged chunks.
        // With no lights there is nothing to sweep: the primary kernel has already written 0 to every hit pixel.
        StageArgs st{};
        st.nv = STAGE_HEADER + STAGE_PER_LIGHT * ctx->n_lights;
        const size_t n_blocks_px = (n_px + 31) / 32;
        bool staged = !(fr->flags & ORE_FLAG_FUSED_SHADOW);
        size_t cap_blocks = 0, est_blocks = n_blocks_px;
        if (ctx->n_lights == 0) {
            staged = false;
        } else if (staged) {
            size_t est_items;
            if (ctx->hits_hint_set) {
                // previous frame's hit count + 12.5 % + 32 K items; the catch-all below covers a wrong guess
                const unsigned long long h = *(volatile unsigned long long*)ctx->hits_hint;
                est_items = (size_t)(h + h / 8 + 32768ull);
            } else {
                // no frame of this context has finished yet: half the pixels (growing the buffer later costs a
                // device-wide synchronisation; the cap below bounds the memory)
                est_items = n_px / 2 > ((size_t)1 << 20) ? n_px / 2 : ((size_t)1 << 20);
            }
            if (est_items > n_px) est_items = n_px;
            est_blocks = (est_items + 31) / 32;
            // the staging buffer never grows past stage_max_items (448 bytes each with three lights: 3.6 GB); a longer hit
            // list - a batch of several 8K frames - is walked in chunk pairs of that size, each a full-size launch
            const size_t max_blocks = (ctx->stage_max_items + 31) / 32;
            const size_t have_blocks = ctx->stage ? ctx->stage_cap / (32 * (size_t)st.nv) : 0;
            cap_blocks = have_blocks;
            if (ctx->stage_blocks_override) {
                cap_blocks = ctx->stage_blocks_override;
                while ((n_blocks_px + cap_blocks - 1) / cap_blocks > (size_t)MAX_STAGE_CHUNKS) cap_blocks *= 2;
            } else if (est_blocks > have_blocks && have_blocks < max_blocks) {
                cap_blocks = est_blocks + est_blocks / 4;   // grow with some slack: reallocations stay rare
                if (cap_blocks > n_blocks_px) cap_blocks = n_blocks_px;
                if (cap_blocks > max_blocks) cap_blocks = max_blocks;
            }
            while ((est_blocks + cap_blocks - 1) / cap_blocks > (size_t)MAX_STAGE_CHUNKS) cap_blocks *= 2;
            const size_t need = cap_blocks * 32 * (size_t)st.nv;
            if (need > ctx->stage_cap || !ctx->stage) {
                // (an earlier frame of this context may still be reading the old buffer on another stream)
                if ((rc = wait_last_render(ctx))) return rc;
                if (ctx->stage) ORE_CUDA(ctx, cudaFree(ctx->stage));
                ctx->stage = nullptr;
                ctx->stage_cap = 0;
                if (cudaMalloc((void**)&ctx->stage, need * sizeof(float)) == cudaSuccess) {
                    ctx->stage_cap = need;
                } else {
                    (void)cudaGetLastError();  // no room for the staging buffer: the fused sweep needs none
                    ctx->stage = nullptr;
                    staged = false;
                }
            }
        }
        if (ctx->n_lights == 0) {
            // nothing to launch
        } else if (!staged) {
            if ((rc = join_band())) return rc;
            if ((rc = launch_sweep(ctx, prm, st, false, exh, fast_libm, stream))) return rc;
            ctx->last_launches++;
        } else {
            st.buf = ctx->stage;
            int n_chunks = (int)((est_blocks + cap_blocks - 1) / cap_blocks);
            if (n_chunks < 1) n_chunks = 1;
            // equal chunks: the expected hit list is split evenly instead of into full buffers plus a small remainder
            if (!ctx->stage_blocks_override) cap_blocks = (est_blocks + n_chunks - 1) / n_chunks;
            st.cap_blocks = (uint32_t)cap_blocks;
            const int n_chunks_max = (int)((n_blocks_px + cap_blocks - 1) / cap_blocks);
            if (n_chunks > n_chunks_max) n_chunks = n_chunks_max;
            int grid_a = 0;
            if (!fast_libm && (rc = grid_for(ctx, shade_setup_kernel, 0, &grid_a, STAGE_A_THREADS))) return rc;
            for (int c = 0; c < n_chunks; c++) {
                st.chunk = c;
                st.first_block = (uint32_t)((size_t)c * cap_blocks);
                if (fast_libm) {
                    ORE_CUDA(ctx, (cudaError_t)ore_fast_shade_setup(&prm, &st, ctx->sm_count, stream));
                } else {
                    shade_setup_kernel<<<grid_a, STAGE_A_THREADS, 0, stream>>>(prm, st);
                    ORE_CUDA(ctx, cudaGetLastError());
                }
                if ((rc = join_band())) return rc;   // (band DMA: the first stage A ran beside the copy)
                if ((rc = launch_sweep(ctx, prm, st, true, exh, fast_libm, stream))) return rc;
                ctx->last_launches += 2;
            }
            if (n_chunks < n_chunks_max) {
                // catch-all: hit-list blocks beyond the staged chunks (normally none: exits at once), one fused launch
                StageArgs rest{};
                rest.first_block = (uint32_t)((size_t)n_chunks * cap_blocks);
                rest.nv = st.nv;
                if ((rc = launch_sweep(ctx, prm, rest, false, exh, fast_libm, stream))) return rc;
                ctx->last_launches++;
            }
        }
    }
    if ((rc = join_band())) return rc;   // (no lights: nothing but the copy writes the frames)
    if (timing) ORE_CUDA(ctx, cudaEventRecord(ctx->ev[3], stream));
    // hint for the next frame of this context (see hits_hint): 8 bytes, no synchronisation
    ORE_CUDA(ctx, cudaMemcpyAsync(ctx->hits_hint, ctx->counters + CNT_HITS, sizeof(unsigned long long),
                                  cudaMemcpyDeviceToHost, stream));
    ctx->hits_hint_set = true;
    if (fr->flags & ORE_FLAG_COUNT_REFERENCE_TESTS) {
        count_reference_tests_kernel<<<ctx->sm_count * 8, CTA_THREADS, 0, stream>>>(prm);
        ORE_CUDA(ctx, cudaGetLastError());
        ctx->last_launches++;
        ctx->ran_count = true;
    }
    if (timing) ORE_CUDA(ctx, cudaEventRecord(ctx->ev[4], stream));
    ctx->ev_valid = timing;
    ORE_CUDA(ctx, cudaEventRecord(ctx->ev_done, stream));
    ctx->ev_done_set = true;
    ctx->last_prm = prm;
    ctx->last_prm_valid = true;
    return ORE_OK;
}

extern "C" int ore_render_device(ore_context* ctx, const ore_camera* cam, const ore_frame* frame,
                                 uint32_t* out_device, void* stream) {
    if (!ctx) return ORE_ERR_INVALID;
    if (!out_device) return fail(ctx, ORE_ERR_INVALID, "ore_render_device: null output");
    uint32_t* outs[1] = {out_device};
    return render_impl(ctx, cam, 1, frame, outs, stream ? (cudaStream_t)stream : ctx->stream);
}

extern "C" int ore_render_batch_device(ore_context* ctx, const ore_camera* cams, int32_t n_frames, const ore_frame* frame,
                                       uint32_t* const* out_device, void* stream) {
    if (!ctx) return ORE_ERR_INVALID;
    if (!out_device) return fail(ctx, ORE_ERR_INVALID, "ore_render_batch_device: null output");
    return render_impl(ctx, cams, n_frames, frame, out_device, stream ? (cudaStream_t)stream : ctx->stream);
}

extern "C" int ore_render(ore_context* ctx, const ore_camera* cam, const ore_frame* frame, uint32_t* out_host) {
    if (!ctx) return ORE_ERR_INVALID;
    if (!out_host) return fail(ctx, ORE_ERR_INVALID, "ore_render: null output");
    int rc = render_impl(ctx, cam, 1, frame, nullptr, ctx->stream);
    if (rc) return rc;
    if (ctx->last_px) {
        // device -> host of the band; CUDA stages pageable destinations through its own pinned pool
        ORE_CUDA(ctx, cudaMemcpyAsync(out_host, ctx->pixels, ctx->last_px * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                      ctx->stream));
    }
    ORE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ORE_OK;
}

// ---- stream-ordered 32-bit flags (cuStreamWriteValue32 / cuStreamWaitValue32, no SM involved) ----------------
// The driver entry points are resolved at run time (the library links against cudart only).
typedef int (*ore_cu_memop32)(cudaStream_t, unsigned long long /*CUdeviceptr*/, unsigned int, unsigned int);
static ore_cu_memop32 cu_write32 = nullptr, cu_wait32 = nullptr;
static bool memops_resolved = false;
static void resolve_memops() {
    if (memops_resolved) return;
    memops_resolved = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
        cu_write32 = (ore_cu_memop32)f;
    f = nullptr;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
        cu_wait32 = (ore_cu_memop32)f;
    (void)cudaGetLastError();
}

// fallbacks on an SM: one thread.  The wait kernel is only ever used when the driver offers no stream wait; it
// polls a flag that ANOTHER GPU (or the host) writes, never another launch on this GPU.
__global__ void flag_write_kernel(uint32_t* flag, uint32_t value) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
__global__ void flag_wait_kernel(const uint32_t* flag, uint32_t value) {
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int32_t)(v - value) >= 0) break;
        __nanosleep(200);
    }
}

static cudaStream_t pick_stream(ore_context* ctx, void* stream) { return stream ? (cudaStream_t)stream : ctx->stream; }

extern "C" void* ore_get_stream(ore_context* ctx, int which) {
    if (!ctx) return nullptr;
    if (which == 2) return (void*)ctx->signal_stream;  // null until the first ore_flag_write_after
    return which == 1 ? (void*)ctx->copy_stream : (void*)ctx->stream;
}

extern "C" int ore_flag_write(ore_context* ctx, void* stream, uint32_t* flag, uint32_t value) {
    if (!ctx || !flag) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    resolve_memops();
    cudaStream_t st = pick_stream(ctx, stream);
    // peer-mapped device memory (another GPU's flag, imported with ore_ipc_import) is written by a one-thread
    // kernel (st.release.sys over NVLink); local device memory and registered host memory by the stream itself
    cudaPointerAttributes at;
    bool local = false;
    if (cudaPointerGetAttributes(&at, flag) == cudaSuccess)
        local = (at.type == cudaMemoryTypeHost) || (at.type == cudaMemoryTypeDevice && at.device == ctx->device);
    (void)cudaGetLastError();
    if (local && cu_write32 && !ctx->no_memops) {
        void* dptr = flag;
        if (at.type == cudaMemoryTypeHost) ORE_CUDA(ctx, cudaHostGetDevicePointer(&dptr, flag, 0));
        if (cu_write32(st, (unsigned long long)(uintptr_t)dptr, value, 0) == 0) return ORE_OK;
    }
    flag_write_kernel<<<1, 1, 0, st>>>(flag, value);
    ORE_CUDA(ctx, cudaGetLastError());
    return ORE_OK;
}

// Ordered variant for frames rendered on SEVERAL streams (frames in flight): the flag is written on the context's
// signal stream once everything enqueued on `stream` so far has completed.  The signal stream is in-order, so flag
// values written through one context appear in call order even when the frames finish out of order.
extern "C" int ore_flag_write_after(ore_context* ctx, void* stream, uint32_t* flag, uint32_t value) {
    if (!ctx || !flag) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->signal_stream) {
        ORE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->signal_stream, cudaStreamNonBlocking));
        for (auto& e : ctx->ev_signal) ORE_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    cudaEvent_t e = ctx->ev_signal[ctx->n_signals++ % 8];
    ORE_CUDA(ctx, cudaEventRecord(e, pick_stream(ctx, stream)));
    ORE_CUDA(ctx, cudaStreamWaitEvent(ctx->signal_stream, e, 0));
    return ore_flag_write(ctx, ctx->signal_stream, flag, value);
}

extern "C" int ore_flag_wait_geq(ore_context* ctx, void* stream, const uint32_t* flag, uint32_t value) {
    if (!ctx || !flag) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    resolve_memops();
    cudaStream_t st = pick_stream(ctx, stream);
    cudaPointerAttributes at;
    void* dptr = (void*)flag;
    if (cudaPointerGetAttributes(&at, flag) == cudaSuccess && at.type == cudaMemoryTypeHost)
        ORE_CUDA(ctx, cudaHostGetDevicePointer(&dptr, (void*)flag, 0));
    (void)cudaGetLastError();
    if (cu_wait32 && !ctx->no_memops) {
        if (cu_wait32(st, (unsigned long long)(uintptr_t)dptr, value, 0 /* CU_STREAM_WAIT_VALUE_GEQ */) == 0) return ORE_OK;
    }
    flag_wait_kernel<<<1, 1, 0, st>>>((const uint32_t*)dptr, value);
    ORE_CUDA(ctx, cudaGetLastError());
    return ORE_OK;
}

extern "C" int ore_host_register(ore_context* ctx, void* host_ptr, size_t bytes) {
    if (!ctx || !host_ptr || bytes == 0) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    ORE_CUDA(ctx, cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return ORE_OK;
}
extern "C" int ore_host_unregister(ore_context* ctx, void* host_ptr) {
    if (!ctx || !host_ptr) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    ORE_CUDA(ctx, cudaHostUnregister(host_ptr));
    return ORE_OK;
}

// Pipelined presentation.  A batch of frames is rendered into one of two sets of device framebuffers and copied to
// the host on the copy stream while the next batch renders.  frame->out_pitch == 0: rows packed at out_host[f].
// out_pitch == width: out_host[f] is image row y0 of a FULL host frame and every rendered row lands at its image
// position (a rank of a multi-GPU job copies its own row blocks into the shared host frame over its own PCIe link).
// done_flag (optional): first_value + f is stored there, in stream order, once frame f's copy has landed.
static int render_async_impl(ore_context* ctx, const ore_camera* cams, int n_frames, const ore_frame* frame,
                             uint32_t* const* out_host, uint32_t* done_flag, uint32_t first_value) {
    if (!ctx) return ORE_ERR_INVALID;
    if (!out_host || !frame || !cams) return fail(ctx, ORE_ERR_INVALID, "ore_render_async: null output/frame");
    if (n_frames < 1 || n_frames > MAX_BATCH) return fail(ctx, ORE_ERR_INVALID, "ore_render_async: a batch holds 1..8 frames");
    for (int f = 0; f < n_frames; f++)
        if (!out_host[f]) return fail(ctx, ORE_ERR_INVALID, "ore_render_async: null host frame in the batch");
    if (frame->out_pitch != 0 && frame->out_pitch != frame->width)
        return fail(ctx, ORE_ERR_INVALID, "ore_render_async: out_pitch must be 0 (packed) or the frame width (rows in place)");
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n_rows = frame_rows(frame);
    const size_t W = (size_t)(frame->width > 0 ? frame->width : 0);
    const size_t n_px = (size_t)(n_rows > 0 ? n_rows : 0) * W;
    const int set = (int)(ctx->async_batches & 1ull);
    int rc;
    if (n_px != ctx->async_px || n_frames != ctx->async_k || !ctx->async_pool) {
        // new geometry: the two sets are laid out afresh, so nothing of the old layout may still be in flight
        ORE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ORE_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
        if ((rc = ensure_dev(ctx, &ctx->async_pool, &ctx->async_pool_cap, 2 * (size_t)n_frames * (n_px ? n_px : 1)))) return rc;
        ctx->async_px = n_px;
        ctx->async_k = n_frames;
    }
    uint32_t* targets[MAX_BATCH];
    for (int f = 0; f < n_frames; f++) targets[f] = ctx->async_pool + ((size_t)set * n_frames + f) * n_px;
    // do not overwrite framebuffers whose previous copy is still in flight
    ORE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[set], 0));
    ore_frame fr = *frame;
    fr.out_pitch = 0;
    if ((rc = render_impl(ctx, cams, n_frames, &fr, targets, ctx->stream))) return rc;
    ORE_CUDA(ctx, cudaEventRecord(ctx->ev_rendered[set], ctx->stream));
    ORE_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_rendered[set], 0));
    for (int f = 0; f < n_frames; f++) {
        if (n_px) {
            const int yb = frame->y_block > 0 ? frame->y_block : 1;
            if (frame->out_pitch == 0 || yb >= frame->y_step) {
                // packed, or a contiguous band: one linear copy
                ORE_CUDA(ctx, cudaMemcpyAsync(out_host[f], targets[f], n_px * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
            } else {
                // block-interleaved rows: blocks of yb rows, y_step image rows apart -> one strided copy (+ a ragged tail)
                const size_t blk_bytes = (size_t)yb * W * sizeof(uint32_t);
                const size_t full = (size_t)n_rows / yb, rem = (size_t)n_rows % yb;
                if (full)
                    ORE_CUDA(ctx, cudaMemcpy2DAsync(out_host[f], (size_t)frame->y_step * W * sizeof(uint32_t), targets[f], blk_bytes,
                                                    blk_bytes, full, cudaMemcpyDeviceToHost, ctx->copy_stream));
                if (rem)
                    ORE_CUDA(ctx, cudaMemcpyAsync(out_host[f] + full * (size_t)frame->y_step * W, targets[f] + full * (size_t)yb * W,
                                                  rem * W * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
            }
        }
        if (done_flag && (rc = ore_flag_write(ctx, ctx->copy_stream, done_flag, first_value + (uint32_t)f))) return rc;
    }
    ORE_CUDA(ctx, cudaEventRecord(ctx->ev_copied[set], ctx->copy_stream));
    ctx->async_batches++;
    return ORE_OK;
}

extern "C" int ore_render_async(ore_context* ctx, const ore_camera* cam, const ore_frame* frame, uint32_t* out_host) {
    uint32_t* outs[1] = {out_host};
    return render_async_impl(ctx, cam, 1, frame, outs, nullptr, 0);
}
extern "C" int ore_render_async_signal(ore_context* ctx, const ore_camera* cam, const ore_frame* frame, uint32_t* out_host,
                                       uint32_t* done_flag, uint32_t done_value) {
    if (!done_flag) return fail(ctx, ORE_ERR_INVALID, "ore_render_async_signal: null flag");
    uint32_t* outs[1] = {out_host};
    return render_async_impl(ctx, cam, 1, frame, outs, done_flag, done_value);
}
extern "C" int ore_render_batch_async(ore_context* ctx, const ore_camera* cams, int32_t n_frames, const ore_frame* frame,
                                      uint32_t* const* out_host, uint32_t* done_flag, uint32_t first_done_value) {
    return render_async_impl(ctx, cams, n_frames, frame, out_host, done_flag, first_done_value);
}

extern "C" int ore_wait(ore_context* ctx) {
    if (!ctx) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    ORE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ORE_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    return ORE_OK;
}

extern "C" int ore_host_alloc(ore_context* ctx, size_t bytes, void** host_ptr) {
    if (!ctx || !host_ptr || bytes == 0) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    ORE_CUDA(ctx, cudaMallocHost(host_ptr, bytes));
    return ORE_OK;
}
extern "C" int ore_host_free(ore_context* ctx, void* host_ptr) {
    if (!ctx) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaFreeHost(host_ptr));
    return ORE_OK;
}

extern "C" int ore_synchronize(ore_context* ctx) {
    if (!ctx) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    ORE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ORE_CUDA(ctx, cudaDeviceSynchronize());
    return ORE_OK;
}

extern "C" int ore_dev_alloc(ore_context* ctx, size_t bytes, void** dev_ptr) {
    if (!ctx || !dev_ptr || bytes == 0) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    ORE_CUDA(ctx, cudaMalloc(dev_ptr, bytes));
    ORE_CUDA(ctx, cudaMemset(*dev_ptr, 0, bytes));
    return ORE_OK;
}
extern "C" int ore_dev_free(ore_context* ctx, void* dev_ptr) {
    if (!ctx) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    ORE_CUDA(ctx, cudaFree(dev_ptr));
    return ORE_OK;
}
extern "C" int ore_ipc_export(ore_context* ctx, void* dev_ptr, unsigned char handle[64]) {
    if (!ctx || !dev_ptr || !handle) return ORE_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    ORE_CUDA(ctx, cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle, &h, 64);
    return ORE_OK;
}
extern "C" int ore_ipc_import(ore_context* ctx, const unsigned char handle[64], void** dev_ptr) {
    if (!ctx || !dev_ptr || !handle) return ORE_ERR_INVALID;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    ORE_CUDA(ctx, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return ORE_OK;
}
extern "C" int ore_ipc_close(ore_context* ctx, void* dev_ptr) {
    if (!ctx) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    ORE_CUDA(ctx, cudaIpcCloseMemHandle(dev_ptr));
    return ORE_OK;
}
extern "C" int ore_copy_to_host(ore_context* ctx, void* host_dst, const void* dev_src, size_t bytes) {
    if (!ctx || !host_dst || !dev_src) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    ORE_CUDA(ctx, cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ORE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ORE_OK;
}

extern "C" int ore_get_hits(ore_context* ctx, int32_t* hit_id_host, float* hit_t_host) {
    if (!ctx) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    ORE_CUDA(ctx, cudaDeviceSynchronize());
    if (ctx->last_px == 0) return ORE_OK;
    if (!ctx->last_prm_valid) return fail(ctx, ORE_ERR_INVALID, "ore_get_hits: no frame rendered");
    if (ctx->last_frames != 1) return fail(ctx, ORE_ERR_INVALID, "ore_get_hits: the last render was a batch (render the frame alone)");
    // the render path keeps compact hit records only; the per-pixel maps are built here, on demand
    int rc;
    size_t c1 = ctx->map_cap, c2 = c1;
    if ((rc = ensure_dev(ctx, &ctx->hit_id_map, &c1, ctx->last_px))) return rc;
    if ((rc = ensure_dev(ctx, &ctx->hit_t_map, &c2, ctx->last_px))) return rc;
    ctx->map_cap = std::min(c1, c2);
    fill_hits_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->hit_id_map, ctx->hit_t_map, ctx->last_px);
    expand_hits_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->last_prm, ctx->hit_id_map, ctx->hit_t_map);
    ORE_CUDA(ctx, cudaGetLastError());
    ORE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (hit_id_host)
        ORE_CUDA(ctx, cudaMemcpy(hit_id_host, ctx->hit_id_map, ctx->last_px * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (hit_t_host)
        ORE_CUDA(ctx, cudaMemcpy(hit_t_host, ctx->hit_t_map, ctx->last_px * sizeof(float), cudaMemcpyDeviceToHost));
    return ORE_OK;
}

extern "C" int ore_get_counters(ore_context* ctx, ore_counters* out) {
    if (!ctx || !out) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    ORE_CUDA(ctx, cudaDeviceSynchronize());
    unsigned long long c[CNT_SLOTS];
    ORE_CUDA(ctx, cudaMemcpy(c, ctx->counters, sizeof c, cudaMemcpyDeviceToHost));
    memset(out, 0, sizeof *out);
    const uint64_t batch_px = (uint64_t)ctx->last_px * (uint64_t)ctx->last_frames;
    out->pixels = batch_px;
    out->hit_pixels = batch_px ? c[CNT_HITS] : 0;
    out->primary_tests = batch_px * (uint64_t)ctx->last_n_spheres;
    out->shadow_tests_ref = c[CNT_SHADOW_TESTS_REF];
    out->sky_tests = batch_px ? batch_px - c[CNT_HITS] : 0;
    out->exact_primary = c[CNT_EXACT_PRIMARY];
    out->exact_shadow = c[CNT_EXACT_SHADOW];
    out->kernel_launches = ctx->last_launches;
    out->beam_l1 = c[CNT_BEAM_L1];
    out->beam_l2 = c[CNT_BEAM_L2];
    out->primary_steps = c[CNT_PRIMARY_STEPS];
    out->sweep_steps = c[CNT_SWEEP_STEPS];
    out->sky_exact = c[CNT_SKY_EXACT];
    return ORE_OK;
}

extern "C" int ore_get_kernel_ms(ore_context* ctx, float ms[4]) {
    if (!ctx || !ms) return ORE_ERR_INVALID;
    ms[0] = ms[1] = ms[2] = ms[3] = 0.f;
    if (!ctx->ev_valid) return ORE_OK;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    ORE_CUDA(ctx, cudaEventSynchronize(ctx->ev[4]));
    for (int i = 0; i < 4; i++) ORE_CUDA(ctx, cudaEventElapsedTime(&ms[i], ctx->ev[i], ctx->ev[i + 1]));
    return ORE_OK;
}

extern "C" int ore_debug_libm(ore_context* ctx, int op, int n, const float* a_host, const float* b_host, float* out_host) {
    if (!ctx || n <= 0 || !a_host || !out_host || op < 0 || op > 6) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    float *a = nullptr, *b = nullptr, *o = nullptr;
    const size_t bytes = (size_t)n * sizeof(float);
    ORE_CUDA(ctx, cudaMalloc((void**)&a, bytes));
    ORE_CUDA(ctx, cudaMalloc((void**)&b, bytes));
    ORE_CUDA(ctx, cudaMalloc((void**)&o, bytes));
    ORE_CUDA(ctx, cudaMemcpy(a, a_host, bytes, cudaMemcpyHostToDevice));
    ORE_CUDA(ctx, cudaMemcpy(b, b_host ? b_host : a_host, bytes, cudaMemcpyHostToDevice));
    libm_probe_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(op, n, a, b, o);
    ORE_CUDA(ctx, cudaGetLastError());
    ORE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ORE_CUDA(ctx, cudaMemcpy(out_host, o, bytes, cudaMemcpyDeviceToHost));
    cudaFree(a);
    cudaFree(b);
    cudaFree(o);
    return ORE_OK;
}

extern "C" int ore_measure_fp32_peak(ore_context* ctx, double* tflops, double* sm_clock_mhz_nominal) {
    if (!ctx || !tflops) return ORE_ERR_INVALID;
    ORE_CUDA(ctx, cudaSetDevice(ctx->device));
    float* d = nullptr;
    ORE_CUDA(ctx, cudaMalloc((void**)&d, sizeof(float)));
    const int iters = 4096, grid = ctx->sm_count * 8;
    cudaEvent_t a, b;
    ORE_CUDA(ctx, cudaEventCreate(&a));
    ORE_CUDA(ctx, cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        ORE_CUDA(ctx, cudaEventRecord(a, ctx->stream));
        fp32_burn_kernel<<<grid, CTA_THREADS, 0, ctx->stream>>>(d, iters, 1.0000001f, 1e-9f);
        ORE_CUDA(ctx, cudaGetLastError());
        ORE_CUDA(ctx, cudaEventRecord(b, ctx->stream));
        ORE_CUDA(ctx, cudaEventSynchronize(b));
        float ms = 0.f;
        ORE_CUDA(ctx, cudaEventElapsedTime(&ms, a, b));
        const double flop = 2.0 * 16 * 8 * (double)iters * CTA_THREADS * (double)grid;
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (rep >= 1 && tf > best) best = tf;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    *tflops = best;
    if (sm_clock_mhz_nominal) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
        *sm_clock_mhz_nominal = khz / 1000.0;
    }
    return ORE_OK;
}
