// nearest-search kernels and the find / hypercubes taps
#define NDP_TU_SEARCH
#define NDP_TU_TAPS
#include "ndp_launch.h"

cudaError_t ndp_launch_import(const ImportArgs &A, int blocks, size_t smem, cudaStream_t st)
{
    k_import<<<blocks, NDP_BLOCK, smem, st>>>(A);
    return cudaGetLastError();
}

cudaError_t ndp_launch_hypercubes(const HypercubeArgs &A, int blocks, cudaStream_t st)
{
    k_hypercubes<<<blocks, NDP_BLOCK, 0, st>>>(A);
    return cudaGetLastError();
}

template <int MODE>
static cudaError_t fast_by_axes(const SearchArgs &A, int naxes, int blocks, size_t smem, cudaStream_t st)
{
    switch (naxes) {
    case 1: k_search_fast<1, MODE><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 2: k_search_fast<2, MODE><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 3: k_search_fast<3, MODE><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 4: k_search_fast<4, MODE><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 5: k_search_fast<5, MODE><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case 6: k_search_fast<6, MODE><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    default: k_search_fast<0, MODE><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    }
    return cudaGetLastError();
}

cudaError_t ndp_launch_search_fast(const SearchArgs &A, int naxes, int blocks, size_t smem, cudaStream_t st)
{
    switch (ndp_mode_of(A.method, A.search)) {
    case NDP_MODE_KD_NEAREST: return fast_by_axes<NDP_MODE_KD_NEAREST>(A, naxes, blocks, smem, st);
    case NDP_MODE_KD_LINEAR: return fast_by_axes<NDP_MODE_KD_LINEAR>(A, naxes, blocks, smem, st);
    case NDP_MODE_LS_NEAREST: return fast_by_axes<NDP_MODE_LS_NEAREST>(A, naxes, blocks, smem, st);
    default: return fast_by_axes<NDP_MODE_LS_LINEAR>(A, naxes, blocks, smem, st);
    }
}

cudaError_t ndp_launch_search_hard(const SearchArgs &A, int blocks, size_t smem, cudaStream_t st)
{
    switch (ndp_mode_of(A.method, A.search)) {
    case NDP_MODE_KD_NEAREST: k_search_hard<NDP_MODE_KD_NEAREST><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case NDP_MODE_KD_LINEAR: k_search_hard<NDP_MODE_KD_LINEAR><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case NDP_MODE_LS_NEAREST: k_search_hard<NDP_MODE_LS_NEAREST><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    default: k_search_hard<NDP_MODE_LS_LINEAR><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    }
    return cudaGetLastError();
}

cudaError_t ndp_launch_search_brute(const SearchArgs &A, int blocks, size_t smem, cudaStream_t st)
{
    k_search_brute<<<blocks, NDP_BLOCK, smem, st>>>(A);
    return cudaGetLastError();
}

cudaError_t ndp_launch_search_sparse(const SearchArgs &A, int blocks, size_t smem, cudaStream_t st)
{
    switch (ndp_mode_of(A.method, A.search)) {
    case NDP_MODE_KD_NEAREST: k_search_sparse<NDP_MODE_KD_NEAREST><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case NDP_MODE_KD_LINEAR: k_search_sparse<NDP_MODE_KD_LINEAR><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case NDP_MODE_LS_NEAREST: k_search_sparse<NDP_MODE_LS_NEAREST><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    default: k_search_sparse<NDP_MODE_LS_LINEAR><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    }
    return cudaGetLastError();
}

cudaError_t ndp_launch_slow(const SlowArgs &A, int blocks, size_t smem, cudaStream_t st)
{
    switch (ndp_mode_of(A.method, A.search)) {
    case NDP_MODE_KD_NEAREST: k_slow<NDP_MODE_KD_NEAREST><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case NDP_MODE_KD_LINEAR: k_slow<NDP_MODE_KD_LINEAR><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    case NDP_MODE_LS_NEAREST: k_slow<NDP_MODE_LS_NEAREST><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    default: k_slow<NDP_MODE_LS_LINEAR><<<blocks, NDP_BLOCK, smem, st>>>(A); break;
    }
    return cudaGetLastError();
}

cudaError_t ndp_launch_nsm_count(const NdpSiteMap &M, int *counts, cudaStream_t st)
{
    k_nsm_count<<<(M.nregions + 127) / 128, 128, 0, st>>>(M, counts);
    return cudaGetLastError();
}

cudaError_t ndp_launch_nsm_fill(const NdpSiteMap &M, uint32_t *ents, uint32_t *head, cudaStream_t st)
{
    k_nsm_fill<<<(M.nregions + 127) / 128, 128, 0, st>>>(M, ents, head);
    return cudaGetLastError();
}

cudaError_t ndp_launch_assoc_consistent(const double *grid, int nverts, long long per_vertex, int vdim, int *flag, int blocks, cudaStream_t st)
{
    k_assoc_consistent<<<blocks, NDP_BLOCK, 0, st>>>(grid, nverts, per_vertex, vdim, flag);
    return cudaGetLastError();
}
