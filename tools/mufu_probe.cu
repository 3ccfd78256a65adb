// How many ex2.approx.f32 (MUFU.EX2) per clock does an SM retire, next to FFMA?  The sampler executes one per element.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_tools/mufu_probe tools/mufu_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE> __global__ void __launch_bounds__(512) k(float* out, int iters) {
    float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f, e = a + .4f, f = a + .5f, g = a + .6f, h = a + .7f;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
            asm volatile("ex2.approx.f32 %0, %0;" : "+f"(a)); asm volatile("ex2.approx.f32 %0, %0;" : "+f"(b));
            asm volatile("ex2.approx.f32 %0, %0;" : "+f"(c)); asm volatile("ex2.approx.f32 %0, %0;" : "+f"(d));
            asm volatile("ex2.approx.f32 %0, %0;" : "+f"(e)); asm volatile("ex2.approx.f32 %0, %0;" : "+f"(f));
            asm volatile("ex2.approx.f32 %0, %0;" : "+f"(g)); asm volatile("ex2.approx.f32 %0, %0;" : "+f"(h));
        } else if (MODE == 1) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(g)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(h));
        } else {
            asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a)); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(b));
            asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(c)); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(d));
            asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(e)); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f));
            asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(g)); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(h));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + e + f + g + h;
}

int main() {
    int sms = 148, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, 4 * sms * 4 * 512);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 4096;
    const char* names[3] = {"ex2.approx.f32", "ex2.approx.ftz.f32", "fma.rn.f32"};
    for (int mode = 0; mode < 3; ++mode) {
        auto run = [&]() { if (mode == 0) k<0><<<sms * 4, 512>>>(out, iters); else if (mode == 1) k<1><<<sms * 4, 512>>>(out, iters); else k<2><<<sms * 4, 512>>>(out, iters); };
        run();
        cudaEventRecord(a); run(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double ops = 8.0 * iters * 512 * 4;  // per SM
        printf("%-20s %8.3f ms  %6.2f lane-ops / clk / SM\n", names[mode], ms, ops / (ms * 1e-3 * clk * 1e3));
    }
    return 0;
}
