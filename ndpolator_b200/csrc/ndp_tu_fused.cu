// k_fused instantiations for NDP_FUSED_N axes (compiled once per axis count, -DNDP_FUSED_N=2..6)
#define NDP_TU_FUSED
#include <algorithm>
#include "ndp_launch.h"

#ifndef NDP_FUSED_N
#error "compile with -DNDP_FUSED_N=<axes>"
#endif
// resident CTAs per SM the kernel is compiled for (register budget 65536 / (128 * OCC); shared memory:
// 4 warps x 32 slots x ((8 << N) + 16) bytes per CTA)
#ifndef NDP_FUSED_OCC
#if NDP_FUSED_N <= 4
#define NDP_FUSED_OCC 6
#elif NDP_FUSED_N == 5
#define NDP_FUSED_OCC 5
#else
#define NDP_FUSED_OCC 3
#endif
#endif

template <int METHOD, int MODE>
static cudaError_t go(const FusedArgs &A, int sm_count, size_t smem_fixed, cudaStream_t st)
{
    constexpr int N = NDP_FUSED_N;
    auto kern = k_fused<N, METHOD, MODE, NDP_FUSED_OCC>;
    const size_t smem = smem_fixed + (size_t) NDP_FUSED_WARPS * 32 * (((size_t) 8 << N) + 16);
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

#define NDP_CAT2(a, b) a##b
#define NDP_CAT(a, b) NDP_CAT2(a, b)
cudaError_t NDP_CAT(ndp_launch_fused_, NDP_FUSED_N)(const FusedArgs &A, int method, int mode, int sm_count, size_t smem_fixed, cudaStream_t st)
{
    if (method == NDP_M_NONE) return go<NDP_M_NONE, 0>(A, sm_count, smem_fixed, st);
    switch (mode) {
    case NDP_MODE_KD_NEAREST: return go<NDP_M_NEAREST, NDP_MODE_KD_NEAREST>(A, sm_count, smem_fixed, st);
    case NDP_MODE_KD_LINEAR: return go<NDP_M_LINEAR, NDP_MODE_KD_LINEAR>(A, sm_count, smem_fixed, st);
    case NDP_MODE_LS_NEAREST: return go<NDP_M_NEAREST, NDP_MODE_LS_NEAREST>(A, sm_count, smem_fixed, st);
    default: return go<NDP_M_LINEAR, NDP_MODE_LS_LINEAR>(A, sm_count, smem_fixed, st);
    }
}
