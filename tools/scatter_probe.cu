// Microbenchmark (not product code): cost of the slot-interleaved staging write considered for permute_kernel --
// every thread stores 16 bytes (one leaf slot, 4 rows) at a pseudo-random slot of a 2 MB row-group block that lives in
// L2 -- against the current layout (runs of ~16 contiguous 16-byte stores).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scatter_probe tools/scatter_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void __launch_bounds__(512) scatter16(float4* __restrict__ z, const int* __restrict__ dst, int n_slots, int n_groups) {
    // item = (segment of 2048 positions, row group); persistent, segment-major
    const int NS = (n_slots + 2047) / 2048;
    const int items = NS * n_groups;
    const int i0 = (int)((long)blockIdx.x * items / gridDim.x), i1 = (int)((long)(blockIdx.x + 1) * items / gridDim.x);
    for (int it = i0; it < i1; ++it) {
        const int s = it / n_groups, g = it - s * n_groups;
        float4* base = z + (size_t)g * n_slots;
        for (int p = s * 2048 + threadIdx.x; p < min(n_slots, (s + 1) * 2048); p += 512) {
            const float v = (float)p;
            base[__ldg(dst + p)] = make_float4(v, v, v, v);
        }
    }
}

int main() {
    const int n_slots = 126 * 1024, n_groups = 16;
    std::vector<int> ident(n_slots), rnd(n_slots), runs(n_slots);
    for (int i = 0; i < n_slots; ++i) ident[i] = i;
    rnd = ident;
    std::mt19937 rng(1);
    std::shuffle(rnd.begin(), rnd.end(), rng);
    // runs of 16 consecutive slots, run starts shuffled (roughly the current layout's locality)
    {
        std::vector<int> starts(n_slots / 16);
        for (size_t i = 0; i < starts.size(); ++i) starts[i] = (int)i * 16;
        std::shuffle(starts.begin(), starts.end(), rng);
        for (int i = 0; i < n_slots; ++i) runs[i] = starts[i / 16] + i % 16;
    }
    float4* z; int* d;
    CK(cudaMalloc(&z, (size_t)n_groups * n_slots * 16));
    CK(cudaMalloc(&d, n_slots * 4));
    const char* names[3] = {"identity (fully coalesced)", "runs of 16 slots", "random slot per thread"};
    std::vector<int>* tabs[3] = {&ident, &runs, &rnd};
    for (int cps = 2; cps <= 4; ++cps)
        for (int k = 0; k < 3; ++k) {
            CK(cudaMemcpy(d, tabs[k]->data(), n_slots * 4, cudaMemcpyHostToDevice));
            for (int i = 0; i < 3; ++i) scatter16<<<148 * cps, 512>>>(z, d, n_slots, n_groups);
            CK(cudaDeviceSynchronize());
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            for (int i = 0; i < 50; ++i) scatter16<<<148 * cps, 512>>>(z, d, n_slots, n_groups);
            cudaEventRecord(b); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, a, b);
            const double us = ms * 1e3 / 50, bytes = (double)n_groups * n_slots * 16;
            printf("%d CTAs/SM  %-28s %6.1f us  %6.0f GB/s (%.0f MB, L2-resident target)\n", cps, names[k], us, bytes / us * 1e-3, bytes / 1e6);
        }
    return 0;
}
