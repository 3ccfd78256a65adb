// Can the TMA engine scatter 16-byte pieces faster than the load/store path?  Every thread issues `iters` bulk stores
// (cp.async.bulk.global.shared::cta, 16 / 32 / 64 bytes each) from shared memory to pseudo-random, naturally aligned
// places of a 33 MB buffer (the size of the staging buffer z).  Reports cycles per piece per SM next to the plain-store
// figure of tools/scatter_store_probe.cu (1.57 cycles per 16-byte lane-store).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_tools/bulk_scatter_probe tools/bulk_scatter_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int BYTES>
__global__ void __launch_bounds__(512) bulk_scatter(unsigned char* buf, unsigned n_slots, int iters, int active_warps) {
    __shared__ __align__(128) unsigned char src[512 * 64];
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 512 * 64 / 4; i += blockDim.x) reinterpret_cast<float*>(src)[i] = (float)i;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp >= active_warps) return;
    unsigned h = (blockIdx.x * 512 + threadIdx.x) * 2654435761u + 12345u;
    const unsigned s = (unsigned)__cvta_generic_to_shared(src + threadIdx.x * 64);
    for (int i = 0; i < iters; ++i) {
        unsigned x = h + (unsigned)i * 0x85EBCA6Bu;
        x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
        unsigned char* p = buf + (size_t)(x % n_slots) * BYTES;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p), "r"(s), "n"(BYTES) : "memory");
        if ((i & 7) == 7) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    const size_t bytes = 33u << 20;
    unsigned char* buf;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    int sms = 148, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 64;
    printf("bytes/piece warps/SM      us   pieces/SM   cycles per piece per SM     GB/s\n");
    for (int sz : {16, 32, 64})
        for (int aw : {2, 8, 32}) {
            const int ctas = 2 * sms, wpc = aw / 2;
            auto run = [&]() {
                const unsigned n_slots = (unsigned)(bytes / sz);
                if (sz == 16) bulk_scatter<16><<<ctas, 512>>>(buf, n_slots, iters, wpc);
                else if (sz == 32) bulk_scatter<32><<<ctas, 512>>>(buf, n_slots, iters, wpc);
                else bulk_scatter<64><<<ctas, 512>>>(buf, n_slots, iters, wpc);
            };
            run(); run();
            cudaEventRecord(a);
            for (int r = 0; r < 5; ++r) run();
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            const double us = ms * 1e3 / 5, pieces = 2.0 * wpc * 32 * iters, cycles = us * 1e-6 * clk * 1e3;
            printf("%6d %8d  %8.2f  %10.0f  %10.2f  %24.1f\n", sz, aw, us, pieces, cycles / pieces, pieces * sms * sz / us * 1e-3);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
