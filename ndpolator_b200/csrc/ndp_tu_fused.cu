// k_fused instantiations for NDP_FUSED_N axes (compiled once per axis count, -DNDP_FUSED_N=2..6)
#define NDP_TU_FUSED
#include <algorithm>
#include "ndp_launch.h"

#ifndef NDP_FUSED_N
#error "compile with -DNDP_FUSED_N=<axes>"
#endif
// software pipeline inside a warp (plan of the next tile while this tile's cells are in flight): see k_fused
#ifndef NDP_FUSED_PIPE
// 2..4 axes: PIPE 2 (profiles/r02c_variants.jsonl).  5 and 6 axes: PIPE 3 -- one query buffer; for 6 axes its smaller
// footprint lets a third CTA fit on an SM (profiles/r02n_variants_pipe3.jsonl: C5 linear +9.7 %, C5 nearest +27 %,
// C4 +2.1 %, C2 -1.4 % -- hence not for the small cells)
#define NDP_FUSED_PIPE (NDP_FUSED_N >= 5 ? 3 : 2)
#endif
// resident CTAs per SM the kernel is compiled for (register budget 65536 / (128 * OCC); shared memory per
// CTA: 4 warps x ndpf_warp_bytes)
#ifndef NDP_FUSED_OCC
#if NDP_FUSED_N <= 4
#define NDP_FUSED_OCC (NDP_FUSED_PIPE ? 5 : 6)
#elif NDP_FUSED_N == 5
#define NDP_FUSED_OCC (NDP_FUSED_PIPE == 1 ? 3 : (NDP_FUSED_PIPE == 0 ? 5 : 4))
#else
#define NDP_FUSED_OCC (NDP_FUSED_PIPE == 3 ? 3 : (NDP_FUSED_PIPE ? 2 : 3))
#endif
#endif

template <int METHOD, int MODE, int VS, int CACHED = 0>
static cudaError_t go(const FusedArgs &A, int sm_count, size_t smem_fixed, cudaStream_t st)
{
    constexpr int N = NDP_FUSED_N;
    auto kern = k_fused<N, METHOD, MODE, NDP_FUSED_OCC, VS, NDP_FUSED_PIPE, CACHED>;
    const size_t smem = smem_fixed + (size_t) NDP_FUSED_WARPS * ndpf_warp_bytes(N, CACHED ? 0 : NDP_FUSED_PIPE);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NDP_FUSED_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long tiles = (A.n + 31) / 32;
    const long long want = (tiles + NDP_FUSED_WARPS - 1) / NDP_FUSED_WARPS;
    const int blocks = (int) std::max<long long>(1, std::min<long long>(want, (long long) sm_count * per_sm));
    kern<<<blocks, NDP_FUSED_WARPS * 32, smem, st>>>(A);
    return cudaGetLastError();
}

template <int METHOD, int MODE, int VS>
static cudaError_t go_vec(const FusedArgs &A, int sm_count, size_t smem_fixed, cudaStream_t st)
{
    constexpr int N = NDP_FUSED_N;
    auto kern = k_fused_vec<N, METHOD, MODE, VS>;
    int per_sm = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NDP_FUSED_WARPS * 32, smem_fixed);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long tiles = (A.n + 31) / 32;
    const long long want = (tiles + NDP_FUSED_WARPS - 1) / NDP_FUSED_WARPS;
    const int blocks = (int) std::max<long long>(1, std::min<long long>(want, (long long) sm_count * per_sm));
    kern<<<blocks, NDP_FUSED_WARPS * 32, smem_fixed, st>>>(A);
    return cudaGetLastError();
}

template <int VS>
static cudaError_t by_mode_vec(const FusedArgs &A, int method, int mode, int sm_count, size_t smem_fixed, cudaStream_t st)
{
    switch (mode) {
    case NDP_MODE_KD_NEAREST: return go_vec<NDP_M_NEAREST, NDP_MODE_KD_NEAREST, VS>(A, sm_count, smem_fixed, st);
    case NDP_MODE_KD_LINEAR: return go_vec<NDP_M_LINEAR, NDP_MODE_KD_LINEAR, VS>(A, sm_count, smem_fixed, st);
    case NDP_MODE_LS_NEAREST: return go_vec<NDP_M_NEAREST, NDP_MODE_LS_NEAREST, VS>(A, sm_count, smem_fixed, st);
    default: return go_vec<NDP_M_LINEAR, NDP_MODE_LS_LINEAR, VS>(A, sm_count, smem_fixed, st);
    }
}

template <int VS, int CACHED = 0>
static cudaError_t by_mode(const FusedArgs &A, int method, int mode, int sm_count, size_t smem_fixed, cudaStream_t st)
{
    switch (mode) {
    case NDP_MODE_KD_NEAREST: return go<NDP_M_NEAREST, NDP_MODE_KD_NEAREST, VS, CACHED>(A, sm_count, smem_fixed, st);
    case NDP_MODE_KD_LINEAR: return go<NDP_M_LINEAR, NDP_MODE_KD_LINEAR, VS, CACHED>(A, sm_count, smem_fixed, st);
    case NDP_MODE_LS_NEAREST: return go<NDP_M_NEAREST, NDP_MODE_LS_NEAREST, VS, CACHED>(A, sm_count, smem_fixed, st);
    default: return go<NDP_M_LINEAR, NDP_MODE_LS_LINEAR, VS, CACHED>(A, sm_count, smem_fixed, st);
    }
}

#define NDP_CAT2(a, b) a##b
#define NDP_CAT(a, b) NDP_CAT2(a, b)
cudaError_t NDP_CAT(ndp_launch_fused_, NDP_FUSED_N)(const FusedArgs &A, int method, int mode, int sm_count, size_t smem_fixed, cudaStream_t st)
{
    if (A.g.vdim > 1) {          // vector-valued table: gather from the original grid, G lanes per query
        if (method == NDP_M_NONE) return go_vec<NDP_M_NONE, 0, 0>(A, sm_count, smem_fixed, st);
        return nsm_leading_pair(A.M) ? by_mode_vec<1>(A, method, mode, sm_count, smem_fixed, st)
                                     : by_mode_vec<0>(A, method, mode, sm_count, smem_fixed, st);
    }
    if (A.c_idx) {               // cached import products (ndpb_ndpolate_cached): the general candidate loop only
        if (method == NDP_M_NONE) return go<NDP_M_NONE, 0, 0, 1>(A, sm_count, smem_fixed, st);
        return by_mode<0, 1>(A, method, mode, sm_count, smem_fixed, st);
    }
    if (method == NDP_M_NONE) return go<NDP_M_NONE, 0, 0>(A, sm_count, smem_fixed, st);
    // tables whose voids depend on exactly their two leading axes take the specialised candidate loop
    return nsm_leading_pair(A.M) ? by_mode<1>(A, method, mode, sm_count, smem_fixed, st)
                                 : by_mode<0>(A, method, mode, sm_count, smem_fixed, st);
}
