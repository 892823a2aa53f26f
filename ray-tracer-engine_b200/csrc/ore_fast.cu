// ore_fast.cu - second instantiation of the default-path kernels with CUDA's own cosf/sinf/acosf/atan2f
// (ORE_FLAG_FAST_LIBM).  Same source (ore_kernels.cuh), different libm macros; the `ore` namespace is renamed
// for this translation unit so both sets of kernels and __constant__ tables coexist in one library.
//
// Results: nearest-hit ids and t are unaffected (no libm on that path); pixels stay within 1 LSB on >= 99.9 %
// (measured 99.999 %) of the bit-exact default, which uses the glibc-compatible functions of ore_libm.cuh.
#define ore ore_fast
#define ORE_CUDA_LIBM 1
#include "ore_kernels.cuh"
#undef ore

#define ORE_HIDDEN __attribute__((visibility("hidden")))

extern "C" ORE_HIDDEN int ore_fast_set_tables(const float* cphi, const float* sphi, const float* bk) {
    if (cudaMemcpyToSymbol(ore_fast::c_cos_phi, cphi, 10 * sizeof(float)) != cudaSuccess) return 1;
    if (cudaMemcpyToSymbol(ore_fast::c_sin_phi, sphi, 10 * sizeof(float)) != cudaSuccess) return 1;
    if (cudaMemcpyToSymbol(ore_fast::c_b_of_k, bk, 11 * sizeof(float)) != cudaSuccess) return 1;
    return 0;
}

template <typename K, typename... A>
static cudaError_t launch(K kernel, int sm_count, size_t smem, long long max_grid, cudaStream_t stream, A... args) {
    int occ = 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, ore_fast::CTA_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    long long grid = (long long)occ * sm_count;
    if (max_grid > 0 && grid > max_grid) grid = max_grid;
    kernel<<<(int)grid, ore_fast::CTA_THREADS, smem, stream>>>(args...);
    return cudaGetLastError();
}

// prm / stage point at a FrameParams / StageArgs of identical layout (same header)
extern "C" ORE_HIDDEN int ore_fast_primary_tile(const void* prm, int sm_count, size_t smem, long long n_batches, int exh,
                                               cudaStream_t stream) {
    const ore_fast::FrameParams& p = *static_cast<const ore_fast::FrameParams*>(prm);
    return (int)(exh ? launch(ore_fast::primary_tile_kernel<ore_fast::TILE_P, true>, sm_count, smem, n_batches, stream, p)
                     : launch(ore_fast::primary_tile_kernel<ore_fast::TILE_P, false>, sm_count, smem, n_batches, stream, p));
}
extern "C" ORE_HIDDEN int ore_fast_shadow_beam(const void* prm, const void* stage, int staged, int sm_count, size_t smem,
                                              int exh, cudaStream_t stream) {
    const ore_fast::FrameParams& p = *static_cast<const ore_fast::FrameParams*>(prm);
    const ore_fast::StageArgs& st = *static_cast<const ore_fast::StageArgs*>(stage);
    if (!staged)
        return (int)(exh ? launch(ore_fast::shadow_beam_kernel<true, false>, sm_count, smem, 0, stream, p, st)
                         : launch(ore_fast::shadow_beam_kernel<false, false>, sm_count, smem, 0, stream, p, st));
    return (int)(exh ? launch(ore_fast::shadow_beam_kernel<true, true>, sm_count, smem, 0, stream, p, st)
                     : launch(ore_fast::shadow_beam_kernel<false, true>, sm_count, smem, 0, stream, p, st));
}
extern "C" ORE_HIDDEN int ore_fast_shade_setup(const void* prm, const void* stage, int sm_count, cudaStream_t stream) {
    const ore_fast::FrameParams& p = *static_cast<const ore_fast::FrameParams*>(prm);
    const ore_fast::StageArgs& st = *static_cast<const ore_fast::StageArgs*>(stage);
    return (int)launch(ore_fast::shade_setup_kernel, sm_count, 0, 0, stream, p, st);
}
