// Microbenchmark (not product code): how fast can persistent CTAs pull the staged rows (z, L2-resident, 35 MB) into
// shared memory, as a function of the copy mechanism and prefetch depth?  Item = (tile, group of R rows): R chunks of
// `zlen` floats at z + row*Zrow + t*zlen.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o load_probe tools/load_probe.cu && ./load_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void cp_async16(void* s, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* s, const void* g, unsigned bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(s)),
                 "l"(g), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b))
                 : "memory");
}

// MODE 0: LDG.128 -> STS.128 (no prefetch)   1: cp.async, DEPTH stages   2: cp.async.bulk + mbarrier, DEPTH stages
template <int MODE, int R, int DEPTH>
__global__ void __launch_bounds__(512) load_probe(const float* __restrict__ z, long Zrow, int n_rows, int NT, int zlen, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* stage = reinterpret_cast<float*>(smem);  // [DEPTH][R][zlen]
    __shared__ uint64_t bar[DEPTH > 0 ? DEPTH : 1];
    const int tid = threadIdx.x;
    const int RG = n_rows / R, items = NT * RG;
    const int i0 = (int)((long)blockIdx.x * items / gridDim.x), i1 = (int)((long)(blockIdx.x + 1) * items / gridDim.x);
    const int nch = zlen / 4;  // 16-byte chunks per row
    float acc = 0.f;
    auto src_of = [&](int it, int r) { const int t = it / RG, g = it - t * RG; return z + (long)(g * R + r) * Zrow + (long)t * zlen; };
    if (MODE == 2) {
        if (tid == 0) for (int d = 0; d < DEPTH; ++d) mbar_init(&bar[d], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
    auto issue = [&](int it) {
        float* dst = stage + (size_t)((it - i0) % DEPTH) * R * zlen;
        if (MODE == 1) {
            for (int i = tid; i < R * nch; i += 512) { const int r = i / nch, c = i - r * nch; cp_async16(dst + r * zlen + 4 * c, src_of(it, r) + 4 * c); }
        } else if (MODE == 2) {
            if (tid == 0) {
                uint64_t* b = &bar[(it - i0) % DEPTH];
                mbar_expect_tx(b, (unsigned)(R * zlen * 4));
                for (int r = 0; r < R; ++r) bulk_g2s(dst + r * zlen, src_of(it, r), (unsigned)(zlen * 4), b);
            }
        }
    };
    if (MODE != 0) {
        for (int d = 0; d < DEPTH - 1; ++d) { if (i0 + d < i1) issue(i0 + d); if (MODE == 1) cp_commit(); }
    }
    for (int it = i0; it < i1; ++it) {
        float* cur = stage + (size_t)((it - i0) % DEPTH) * R * zlen;
        if (MODE == 0) {
            for (int i = tid; i < R * nch; i += 512) {
                const int r = i / nch, c = i - r * nch;
                const float4 v = __ldcg(reinterpret_cast<const float4*>(src_of(it, r)) + c);
                *reinterpret_cast<float4*>(stage + r * zlen + 4 * c) = v;
            }
            __syncthreads();
            cur = stage;
        } else {
            if (it + DEPTH - 1 < i1) issue(it + DEPTH - 1);
            if (MODE == 1) { cp_commit(); cp_wait<DEPTH - 1>(); __syncthreads(); }
            else mbar_wait(&bar[(it - i0) % DEPTH], (unsigned)(((it - i0) / DEPTH) & 1));
        }
        // token consumption: every thread reads one 16-byte chunk per row
#pragma unroll
        for (int r = 0; r < R; ++r) acc += cur[r * zlen + (tid * 4) % zlen];
        __syncthreads();  // buffer free for re-use
    }
    if (acc == 123.456f) *sink = acc;
}

template <int MODE, int R, int DEPTH>
static void run(const char* name, const float* z, long Zrow, int n_rows, int NT, int zlen, int cps, float* sink) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = (size_t)(DEPTH > 0 ? DEPTH : 1) * R * zlen * 4;
    CK(cudaFuncSetAttribute(load_probe<MODE, R, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = sms * cps;
    for (int i = 0; i < 3; ++i) load_probe<MODE, R, DEPTH><<<grid, 512, smem>>>(z, Zrow, n_rows, NT, zlen, sink);
    CK(cudaDeviceSynchronize());
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 40;
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) load_probe<MODE, R, DEPTH><<<grid, 512, smem>>>(z, Zrow, n_rows, NT, zlen, sink);
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double us = ms * 1e3 / iters, bytes = (double)n_rows * NT * zlen * 4;
    printf("%-28s R%d depth%d zlen=%d cta/sm=%d smem=%zuKB : %7.1f us %7.0f GB/s\n", name, R, DEPTH, zlen, cps, smem / 1024, us, bytes / us * 1e-3);
}

int main() {
    const int n_rows = 64;
    float *z, *sink;
    const long Zrow = 126L * 1088;
    CK(cudaMalloc(&z, (size_t)n_rows * Zrow * 4));
    CK(cudaMemset(z, 0, (size_t)n_rows * Zrow * 4));
    CK(cudaMalloc(&sink, 4));
    for (int cps = 1; cps <= 4; cps *= 2) {
        printf("---- %d CTAs/SM\n", cps);
        run<0, 4, 1>("ldg+sts", z, Zrow, n_rows, 126, 1088, cps, sink);
        run<1, 4, 2>("cp.async", z, Zrow, n_rows, 126, 1088, cps, sink);
        run<1, 4, 3>("cp.async", z, Zrow, n_rows, 126, 1088, cps, sink);
        run<2, 4, 2>("bulk", z, Zrow, n_rows, 126, 1088, cps, sink);
        run<2, 4, 3>("bulk", z, Zrow, n_rows, 126, 1088, cps, sink);
        run<2, 4, 4>("bulk", z, Zrow, n_rows, 126, 1088, cps, sink);
        run<2, 2, 4>("bulk", z, Zrow, n_rows, 126, 1088, cps, sink);
        run<2, 4, 2>("bulk T=2048", z, Zrow, n_rows, 63, 2176, cps, sink);
        run<1, 4, 2>("cp.async T=2048", z, Zrow, n_rows, 63, 2176, cps, sink);
    }
    return 0;
}
