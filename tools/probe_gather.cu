// probe_gather.cu -- microbenchmark (development tool, not part of the product):
// how fast can a B200 gather random, aligned 256-byte blocks out of a table much larger
// than L2?  Bounds what k_main / k_finish can reach with the cell-major layout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/probe tools/probe_gather.cu && /tmp/probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void k_fill_idx(int *idx, long long nq, long long nblocks)
{
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (long long) gridDim.x * blockDim.x)
        idx[i] = (int) (mix(i) % nblocks);
}

// (a) one thread per block: 16 x LDG.128
__global__ void k_thread(const double2 *__restrict__ tab, const int *__restrict__ idx, long long nq, double *out)
{
    double acc = 0;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (long long) gridDim.x * blockDim.x) {
        const double2 *b = tab + (long long) idx[i] * 16;
        double2 v[16];
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = b[j];
#pragma unroll
        for (int j = 0; j < 16; j++) acc += v[j].x + v[j].y;
    }
    if (acc == 1.2345) out[0] = acc;
}

// (b) half-warp per block: 16 lanes x 16 B = one coalesced 256-byte request; DEPTH blocks in flight per lane
template <int DEPTH>
__global__ void k_halfwarp(const double2 *__restrict__ tab, const int *__restrict__ idx, long long nq, double *out)
{
    double acc = 0;
    const int sub = threadIdx.x & 15;
    const long long hw = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const long long nhw = ((long long) gridDim.x * blockDim.x) >> 4;
    for (long long i = hw * DEPTH; i + DEPTH <= nq; i += nhw * DEPTH) {
        int b[DEPTH];
#pragma unroll
        for (int k = 0; k < DEPTH; k++) b[k] = idx[i + k];
        double2 v[DEPTH];
#pragma unroll
        for (int k = 0; k < DEPTH; k++) v[k] = tab[(long long) b[k] * 16 + sub];
#pragma unroll
        for (int k = 0; k < DEPTH; k++) acc += v[k].x + v[k].y;
    }
    if (acc == 1.2345) out[0] = acc;
}

// (c) sequential 256-byte blocks (streaming reference)
__global__ void k_stream(const double2 *__restrict__ tab, long long nblocks, long long nq, double *out)
{
    double acc = 0;
    const int sub = threadIdx.x & 15;
    const long long hw = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const long long nhw = ((long long) gridDim.x * blockDim.x) >> 4;
    for (long long i = hw; i < nq; i += nhw) { double2 v = tab[(i % nblocks) * 16 + sub]; acc += v.x + v.y; }
    if (acc == 1.2345) out[0] = acc;
}

int main()
{
    const long long maxblocks = 90000000;          // 23 GB like the C4 cell-major table
    const long long nq = 1ll << 26;
    double2 *tab; double *out; int *idx;
    CK(cudaMalloc(&tab, maxblocks * 256)); CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&idx, nq * 4));
    CK(cudaMemset(tab, 0, maxblocks * 256));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char *name, long long nblocks, auto launch) {
        launch(); cudaDeviceSynchronize();
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-26s table %8.1f MB  %8.3f ms  %7.1f GB/s  %6.2f Gblocks/s\n", name, nblocks * 256.0 / 1e6, ms, nq * 256.0 / ms / 1e6, nq / ms / 1e6);
    };
    for (long long nblocks : {200000ll, 3200000ll, 90000000ll}) {
        k_fill_idx<<<148 * 8, 256>>>(idx, nq, nblocks);
        CK(cudaDeviceSynchronize());
        run("thread 2368x256", nblocks, [&] { k_thread<<<2368, 256>>>(tab, idx, nq, out); });
        run("thread 1184x128", nblocks, [&] { k_thread<<<1184, 128>>>(tab, idx, nq, out); });
        run("halfwarp d1 2368x256", nblocks, [&] { k_halfwarp<1><<<2368, 256>>>(tab, idx, nq, out); });
        run("halfwarp d2 2368x256", nblocks, [&] { k_halfwarp<2><<<2368, 256>>>(tab, idx, nq, out); });
        run("halfwarp d4 2368x256", nblocks, [&] { k_halfwarp<4><<<2368, 256>>>(tab, idx, nq, out); });
        run("halfwarp d8 2368x256", nblocks, [&] { k_halfwarp<8><<<2368, 256>>>(tab, idx, nq, out); });
        run("halfwarp d16 2368x256", nblocks, [&] { k_halfwarp<16><<<2368, 256>>>(tab, idx, nq, out); });
        run("halfwarp d8 1184x256", nblocks, [&] { k_halfwarp<8><<<1184, 256>>>(tab, idx, nq, out); });
    }
    run("stream 256B blocks", maxblocks, [&] { k_stream<<<148 * 16, 256>>>(tab, maxblocks, nq, out); });
    return 0;
}
