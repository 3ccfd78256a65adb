// Microbenchmark (not product code): how fast can a persistent kernel write the [B, N] node-mass slab on B200,
// as a function of the store shape?  Mirrors tile_kernel's emit: item = (tile of `chunk` consecutive nodes, group of
// R rows); values come from registers so only the store path is measured.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_probe tools/store_probe.cu && ./store_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

enum Policy { P_DEFAULT = 0, P_CS = 1, P_WT = 2 };

template <int POL> __device__ __forceinline__ void st1(float* p, float v) {
    if (POL == P_CS) __stcs(p, v); else if (POL == P_WT) __stwt(p, v); else *p = v;
}
template <int POL> __device__ __forceinline__ void st2(float* p, float2 v) {
    if (POL == P_CS) __stcs((float2*)p, v); else if (POL == P_WT) __stwt((float2*)p, v); else *(float2*)p = v;
}
template <int POL> __device__ __forceinline__ void st4(float* p, float4 v) {
    if (POL == P_CS) __stcs((float4*)p, v); else if (POL == P_WT) __stwt((float4*)p, v); else *(float4*)p = v;
}

// VEC = floats per lane per store (1, 2, 4); R rows per item
template <int VEC, int R, int POL>
__global__ void __launch_bounds__(512) emit_probe(float* __restrict__ out, long ld, int n_rows, int N, int chunk, int first) {
    const int NT = (N - first + chunk - 1) / chunk;
    const int RG = n_rows / R;
    const int items = NT * RG;
    const int i0 = (int)((long)blockIdx.x * items / gridDim.x), i1 = (int)((long)(blockIdx.x + 1) * items / gridDim.x);
    const float v = (float)threadIdx.x;
    for (int it = i0; it < i1; ++it) {
        const int t = it / RG, g = it - t * RG;
        const int n0 = first + t * chunk, n1 = min(N, n0 + chunk);
        float* row[R];
#pragma unroll
        for (int r = 0; r < R; ++r) row[r] = out + (long)(g * R + r) * ld;
        if (VEC == 1) {
            for (int n = n0 + threadIdx.x; n < n1; n += 512)
#pragma unroll
                for (int r = 0; r < R; ++r) st1<POL>(row[r] + n, v);
        } else if (VEC == 2) {
            for (int n = n0 + 2 * threadIdx.x; n + 1 < n1; n += 1024)
#pragma unroll
                for (int r = 0; r < R; ++r) st2<POL>(row[r] + n, make_float2(v, v));
        } else {
            for (int n = n0 + 4 * threadIdx.x; n + 3 < n1; n += 2048)
#pragma unroll
                for (int r = 0; r < R; ++r) st4<POL>(row[r] + n, make_float4(v, v, v, v));
        }
    }
}

template <int VEC, int R, int POL>
static void run(const char* name, float** bufs, int nbuf, long ld, int n_rows, int N, int chunk, int first, int ctas_per_sm) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * ctas_per_sm;
    for (int i = 0; i < 3; ++i) emit_probe<VEC, R, POL><<<grid, 512>>>(bufs[i % nbuf], ld, n_rows, N, chunk, first);
    CK(cudaDeviceSynchronize());
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 40;
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) emit_probe<VEC, R, POL><<<grid, 512>>>(bufs[i % nbuf], ld, n_rows, N, chunk, first);
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double us = ms * 1e3 / iters;
    const double bytes = (double)n_rows * (N - first) * 4;
    printf("%-46s vec%d R%d ld=%ld chunk=%d first=%d cta/sm=%d : %7.1f us  %7.0f GB/s\n", name, VEC, R, ld, chunk, first,
           ctas_per_sm, us, bytes / us * 1e-3);
}

int main() {
    const int n_rows = 64, N = 345180;
    const long ld_pad = 345216;  // multiple of 32 floats (128 B)
    const int nbuf = 4;
    float* bufs[nbuf];
    for (int i = 0; i < nbuf; ++i) CK(cudaMalloc(&bufs[i], (size_t)n_rows * ld_pad * 4 + 4096));
    // baseline: cudaMemset of the same bytes
    {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int i = 0; i < 3; ++i) cudaMemsetAsync(bufs[i % nbuf], 0, (size_t)n_rows * N * 4);
        cudaEventRecord(a);
        for (int i = 0; i < 40; ++i) cudaMemsetAsync(bufs[i % nbuf], 0, (size_t)n_rows * N * 4);
        cudaEventRecord(b); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("cudaMemset 88 MB: %.1f us %.0f GB/s\n", ms * 1e3 / 40, (double)n_rows * N * 4 / (ms * 1e3 / 40) * 1e-3);
    }
    for (int cps = 1; cps <= 4; cps *= 2) {
        printf("---- %d CTAs/SM\n", cps);
        run<1, 4, P_CS>("ours: lane/node, 4 rows, stcs, unaligned", bufs, nbuf, N, n_rows, N, 2740, 0, cps);
        run<1, 4, P_CS>("lane/node, 4 rows, stcs, chunk%32==0", bufs, nbuf, N, n_rows, N, 2752, 0, cps);
        run<1, 4, P_CS>("lane/node, 4 rows, stcs, ld%32==0 aligned", bufs, nbuf, ld_pad, n_rows, N, 2752, 0, cps);
        run<1, 4, P_DEFAULT>("lane/node, 4 rows, default, aligned", bufs, nbuf, ld_pad, n_rows, N, 2752, 0, cps);
        run<1, 4, P_DEFAULT>("lane/node, 4 rows, default, unaligned", bufs, nbuf, N, n_rows, N, 2740, 0, cps);
        run<1, 1, P_CS>("lane/node, 1 row, stcs, unaligned", bufs, nbuf, N, n_rows, N, 2740, 0, cps);
        run<1, 1, P_CS>("lane/node, 1 row, stcs, aligned", bufs, nbuf, ld_pad, n_rows, N, 2752, 0, cps);
        run<1, 2, P_CS>("lane/node, 2 rows, stcs, unaligned", bufs, nbuf, N, n_rows, N, 2740, 0, cps);
        run<4, 4, P_CS>("lane/4 nodes (STG.128), 4 rows, stcs, ld%4", bufs, nbuf, N, n_rows, N, 2740, 0, cps);
        run<4, 4, P_CS>("lane/4 nodes (STG.128), 4 rows, stcs, aligned", bufs, nbuf, ld_pad, n_rows, N, 2752, 0, cps);
        run<4, 4, P_DEFAULT>("lane/4 nodes (STG.128), 4 rows, default, ld%4", bufs, nbuf, N, n_rows, N, 2740, 0, cps);
        run<4, 1, P_CS>("lane/4 nodes (STG.128), 1 row, stcs, ld%4", bufs, nbuf, N, n_rows, N, 2740, 0, cps);
        run<4, 1, P_CS>("lane/4 nodes (STG.128), 1 row, big chunk", bufs, nbuf, N, n_rows, N, 16384, 0, cps);
        run<2, 4, P_CS>("lane/2 nodes (STG.64), 4 rows, stcs", bufs, nbuf, N, n_rows, N, 2740, 0, cps);
        run<4, 4, P_WT>("lane/4 nodes (STG.128), 4 rows, wt, ld%4", bufs, nbuf, N, n_rows, N, 2740, 0, cps);
    }
    return 0;
}
