// How fast can an SM issue scattered stores that hit L2?  Each lane writes 16 (or 4 / 8) bytes to a pseudo-random,
// naturally aligned place of a 33 MB buffer (the size of the staging buffer z); `adj` consecutive lanes write adjacent
// places (adj = 1: every lane its own 128-byte line; 8: a warp store covers 4 whole lines).  Reports cycles per
// warp-level store request ("line touched") per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_tools/scatter_store_probe tools/scatter_store_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

template <int BYTES>
__global__ void __launch_bounds__(512) scatter(unsigned char* buf, unsigned n_slots, int iters, int adj, int active_warps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= active_warps) return;
    unsigned h = (blockIdx.x * 16 + warp) * 2654435761u + 12345u;
    for (int i = 0; i < iters; ++i) {
        // one random group per `adj` lanes
        const unsigned grp = lane / adj;
        unsigned x = h + grp * 0x9E3779B9u + (unsigned)i * 0x85EBCA6Bu;
        x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
        const unsigned slot = ((x % (n_slots / adj)) * adj + (lane % adj));
        unsigned char* p = buf + (size_t)slot * BYTES;
        if (BYTES == 16) *reinterpret_cast<float4*>(p) = make_float4(1.f, 2.f, 3.f, (float)i);
        else if (BYTES == 8) *reinterpret_cast<float2*>(p) = make_float2(1.f, (float)i);
        else *reinterpret_cast<float*>(p) = (float)i;
    }
}

int main() {
    const size_t bytes = 33u << 20;
    unsigned char* buf;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 256;
    printf("bytes/lane adj warps/SM   us    lane-stores/SM  cycles per lane-store per SM   cycles per line per SM   GB/s\n");
    for (int sz : {16, 8, 4})
        for (int adj : {1, 2, 4, 8})
            for (int aw : {2, 4, 16, 32}) {
                const int ctas = 2 * sms, warps_per_cta = aw / 2 > 0 ? aw / 2 : 1;
                auto run = [&]() {
                    const unsigned n_slots = (unsigned)(bytes / sz);
                    if (sz == 16) scatter<16><<<ctas, 512>>>(buf, n_slots, iters, adj, warps_per_cta);
                    else if (sz == 8) scatter<8><<<ctas, 512>>>(buf, n_slots, iters, adj, warps_per_cta);
                    else scatter<4><<<ctas, 512>>>(buf, n_slots, iters, adj, warps_per_cta);
                };
                run(); run();
                cudaEventRecord(a);
                for (int r = 0; r < 5; ++r) run();
                cudaEventRecord(b);
                cudaEventSynchronize(b);
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                const double us = ms * 1e3 / 5;
                const double lane_stores = 2.0 * warps_per_cta * 32 * iters;  // per SM
                const double cycles = us * 1e-6 * clk * 1e3;
                const double lines = lane_stores * sz / (double)(adj * sz >= 128 ? 128 : adj * sz);  // distinct <=128B pieces
                printf("%6d %6d %6d  %8.2f  %10.0f  %10.2f  %24.2f  %8.1f\n", sz, adj, 2 * warps_per_cta, us, lane_stores,
                       cycles / lane_stores, cycles / lines, lane_stores * sms * sz / us * 1e-3);
            }
    return 0;
}
