// k_main instantiations for the original (strided) grid layout
#define NDP_TU_MAIN
#include "ndp_launch.h"

template <int N>
static cudaError_t go(const MainArgs &A, int G, int blocks, size_t smem, cudaStream_t st)
{
    switch (G) {
    case 1: k_main<N, 1, false><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 2: k_main<N, 2, false><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 4: k_main<N, 4, false><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 8: k_main<N, 8, false><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 16: k_main<N, 16, false><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    default: k_main<N, 32, false><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    }
    return cudaGetLastError();
}

cudaError_t ndp_launch_main_strided(const MainArgs &A, int naxes, int G, int blocks, size_t smem, cudaStream_t st)
{
    switch (naxes) {
    case 1: return go<1>(A, G, blocks, smem, st);
    case 2: return go<2>(A, G, blocks, smem, st);
    case 3: return go<3>(A, G, blocks, smem, st);
    case 4: return go<4>(A, G, blocks, smem, st);
    case 5: return go<5>(A, G, blocks, smem, st);
    case 6: return go<6>(A, G, blocks, smem, st);
    default: return go<0>(A, G, blocks, smem, st);
    }
}
