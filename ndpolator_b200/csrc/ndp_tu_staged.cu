// k_gather_staged instantiations (cp.async staged gather)
#define NDP_TU_STAGED
#include <algorithm>
#include "ndp_launch.h"

template <int N, int MODE, int STAGES, int OCC = 5>
static cudaError_t go(const StagedArgs &A, int sm_count, long long items, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(k_gather_staged<N, MODE, STAGES, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gather_staged<N, MODE, STAGES, OCC>, NDP_STAGED_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long tiles = (items + 31) / 32;
    const long long want = (tiles + NDP_STAGED_WARPS - 1) / NDP_STAGED_WARPS;
    const int blocks = (int) std::max<long long>(1, std::min<long long>(want, (long long) sm_count * per_sm));
    k_gather_staged<N, MODE, STAGES, OCC><<<blocks, NDP_STAGED_WARPS * 32, smem, st>>>(A);
    return cudaGetLastError();
}

template <int MODE, int STAGES, int OCC = 5>
static cudaError_t by_axes(const StagedArgs &A, int naxes, int sm_count, long long items, size_t smem, cudaStream_t st)
{
    switch (naxes) {
    case 2: return go<2, MODE, STAGES, OCC>(A, sm_count, items, smem, st);
    case 3: return go<3, MODE, STAGES, OCC>(A, sm_count, items, smem, st);
    case 4: return go<4, MODE, STAGES, OCC>(A, sm_count, items, smem, st);
    case 5: return go<5, MODE, STAGES, OCC>(A, sm_count, items, smem, st);
    default: return go<6, MODE, STAGES, OCC>(A, sm_count, items, smem, st);
    }
}

cudaError_t ndp_launch_staged(const StagedArgs &A, int naxes, int mode, int stages, int sm_count, long long items, size_t smem, cudaStream_t st)
{
    if (stages == 3)       // experiment: one stage, 4 CTAs per SM (more registers, fewer warps)
        return mode == 0 ? by_axes<0, 1, 4>(A, naxes, sm_count, items, smem, st) : by_axes<1, 1, 4>(A, naxes, sm_count, items, smem, st);
    if (stages == 1)
        return mode == 0 ? by_axes<0, 1>(A, naxes, sm_count, items, smem, st) : by_axes<1, 1>(A, naxes, sm_count, items, smem, st);
    return mode == 0 ? by_axes<0, 2>(A, naxes, sm_count, items, smem, st) : by_axes<1, 2>(A, naxes, sm_count, items, smem, st);
}
