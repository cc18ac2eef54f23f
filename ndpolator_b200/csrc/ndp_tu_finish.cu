// k_finish instantiations
#define NDP_TU_FINISH
#include "ndp_launch.h"

template <int N, bool CM>
static cudaError_t by_group(const FinishArgs &A, int G, int blocks, size_t smem, cudaStream_t st)
{
    switch (G) {
    case 1: k_finish<N, 1, CM><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 2: k_finish<N, 2, CM><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 4: k_finish<N, 4, CM><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 8: k_finish<N, 8, CM><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 16: k_finish<N, 16, CM><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    default: k_finish<N, 32, CM><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    }
    return cudaGetLastError();
}

template <bool CM>
static cudaError_t go(const FinishArgs &A, int naxes, int G, int blocks, size_t smem, cudaStream_t st)
{
    switch (naxes) {
    case 1: return by_group<1, CM>(A, G, blocks, smem, st);
    case 2: return by_group<2, CM>(A, G, blocks, smem, st);
    case 3: return by_group<3, CM>(A, G, blocks, smem, st);
    case 4: return by_group<4, CM>(A, G, blocks, smem, st);
    case 5: return by_group<5, CM>(A, G, blocks, smem, st);
    case 6: return by_group<6, CM>(A, G, blocks, smem, st);
    default: return by_group<0, false>(A, G, blocks, smem, st);
    }
}

cudaError_t ndp_launch_finish(const FinishArgs &A, int naxes, int G, bool cellmajor, int blocks, size_t smem, cudaStream_t st)
{
    return (cellmajor && naxes <= 6) ? go<true>(A, naxes, G, blocks, smem, st) : go<false>(A, naxes, G, blocks, smem, st);
}
