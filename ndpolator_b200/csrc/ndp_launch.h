// ndp_launch.h -- host-callable launchers, one translation unit per kernel family so that
// the sm_100a build compiles in parallel (ndpolator_b200/build.py).
#pragma once
#include <cuda_runtime.h>
#include "ndp_kernels.cuh"

cudaError_t ndp_launch_import(const ImportArgs &A, int blocks, size_t smem, cudaStream_t st);
cudaError_t ndp_launch_hypercubes(const HypercubeArgs &A, int blocks, cudaStream_t st);
cudaError_t ndp_launch_main_strided(const MainArgs &A, int naxes, int G, int blocks, size_t smem, cudaStream_t st);
cudaError_t ndp_launch_main_cellmajor(const MainArgs &A, int naxes, int G, int blocks, size_t smem, cudaStream_t st);
cudaError_t ndp_launch_finish(const FinishArgs &A, int naxes, int G, bool cellmajor, int blocks, size_t smem, cudaStream_t st);
cudaError_t ndp_launch_search_fast(const SearchArgs &A, int naxes, int blocks, size_t smem, cudaStream_t st);
cudaError_t ndp_launch_search_hard(const SearchArgs &A, int blocks, size_t smem, cudaStream_t st);
cudaError_t ndp_launch_search_brute(const SearchArgs &A, int blocks, size_t smem, cudaStream_t st);
// mode 0: k_main work, mode 1: linear-extrapolation finish; picks the grid from the occupancy
cudaError_t ndp_launch_staged(const StagedArgs &A, int naxes, int mode, int stages, int sm_count, long long items, size_t smem, cudaStream_t st);
