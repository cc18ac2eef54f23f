// ndp_launch.h -- host-callable launchers, one translation unit per kernel family so that
// the sm_100a build compiles in parallel (ndpolator_b200/build.py).
#pragma once
#include <cuda_runtime.h>
#include "ndp_kernels.cuh"
#include "ndp_fused_kernels.cuh"

cudaError_t ndp_launch_import(const ImportArgs &A, int blocks, size_t smem, cudaStream_t st);
cudaError_t ndp_launch_hypercubes(const HypercubeArgs &A, int blocks, cudaStream_t st);
cudaError_t ndp_launch_main_strided(const MainArgs &A, int naxes, int G, int blocks, size_t smem, cudaStream_t st);
cudaError_t ndp_launch_main_cellmajor(const MainArgs &A, int naxes, int G, int blocks, size_t smem, cudaStream_t st);
cudaError_t ndp_launch_finish(const FinishArgs &A, int naxes, int G, bool cellmajor, int blocks, size_t smem, cudaStream_t st);
cudaError_t ndp_launch_search_fast(const SearchArgs &A, int naxes, int blocks, size_t smem, cudaStream_t st);
cudaError_t ndp_launch_search_hard(const SearchArgs &A, int blocks, size_t smem, cudaStream_t st);
cudaError_t ndp_launch_search_brute(const SearchArgs &A, int blocks, size_t smem, cudaStream_t st);
cudaError_t ndp_launch_search_sparse(const SearchArgs &A, int blocks, size_t smem, cudaStream_t st);
// mode 0: k_main work, mode 1: linear-extrapolation finish; picks the grid from the occupancy
cudaError_t ndp_launch_staged(const StagedArgs &A, int naxes, int mode, int stages, int sm_count, long long items, size_t smem, cudaStream_t st);
// fused path (ndp_fused_kernels.cuh): one translation unit per axis count
#define NDP_FUSED_DECL(N) cudaError_t ndp_launch_fused_##N(const FusedArgs &A, int method, int mode, int sm_count, size_t smem_fixed, cudaStream_t st);
NDP_FUSED_DECL(2) NDP_FUSED_DECL(3) NDP_FUSED_DECL(4) NDP_FUSED_DECL(5) NDP_FUSED_DECL(6)
#undef NDP_FUSED_DECL
cudaError_t ndp_launch_slow(const SlowArgs &A, int blocks, size_t smem, cudaStream_t st);
cudaError_t ndp_launch_nsm_count(const NdpSiteMap &M, int *counts, cudaStream_t st);
cudaError_t ndp_launch_nsm_fill(const NdpSiteMap &M, uint32_t *ents, uint32_t *head, cudaStream_t st);
cudaError_t ndp_launch_assoc_consistent(const double *grid, int nverts, long long per_vertex, int vdim, int *flag, int blocks, cudaStream_t st);
