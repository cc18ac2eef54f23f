// k_finish instantiations
#define NDP_TU_FINISH
#include "ndp_launch.h"

template <bool CM>
static cudaError_t go(const FinishArgs &A, int naxes, int blocks, size_t smem, cudaStream_t st)
{
    switch (naxes) {
    case 1: k_finish<1, CM><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 2: k_finish<2, CM><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 3: k_finish<3, CM><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 4: k_finish<4, CM><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 5: k_finish<5, CM><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 6: k_finish<6, CM><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    default: k_finish<0, false><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    }
    return cudaGetLastError();
}

cudaError_t ndp_launch_finish(const FinishArgs &A, int naxes, bool cellmajor, int blocks, size_t smem, cudaStream_t st)
{
    return (cellmajor && naxes <= 6) ? go<true>(A, naxes, blocks, smem, st) : go<false>(A, naxes, blocks, smem, st);
}
