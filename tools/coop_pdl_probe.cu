// Can a kernel be launched both cooperatively (grid-wide sync) and with programmatic dependent launch?  And what does a
// grid.sync() of a full-occupancy persistent grid cost next to a kernel boundary?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_tools/coop_pdl_probe tools/coop_pdl_probe.cu && build_tools/coop_pdl_probe
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(512, 2) body(float* x, int n, int syncs) {
    extern __shared__ unsigned char smem[];
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    cg::grid_group g = cg::this_grid();
    for (int s = 0; s < syncs; ++s) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] += 1.f;
        g.sync();
    }
    if (syncs == 0)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] += 1.f;
}

static cudaError_t launch(bool coop, bool pdl, int grid, size_t smem, cudaStream_t st, float* x, int n, int syncs) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (coop) { attr[na].id = cudaLaunchAttributeCooperative; attr[na].val.cooperative = 1; ++na; }
    if (pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
    cfg.attrs = attr; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, body, x, n, syncs);
}

int main() {
    const size_t smem = 110 * 1024;
    cudaFuncSetAttribute(body, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = 2 * sms, n = 1 << 20;
    float* x; cudaMalloc(&x, n * sizeof(float)); cudaMemset(x, 0, n * sizeof(float));
    cudaStream_t st; cudaStreamCreate(&st);
    for (int coop = 0; coop < 2; ++coop)
        for (int pdl = 0; pdl < 2; ++pdl) {
            cudaError_t e = launch(coop, pdl, grid, smem, st, x, n, coop ? 1 : 0);
            cudaError_t e2 = cudaStreamSynchronize(st);
            printf("cooperative=%d pdl=%d: launch %s, sync %s\n", coop, pdl, cudaGetErrorString(e), cudaGetErrorString(e2));
            (void)cudaGetLastError();
        }
    // cost: chains of 64 launches, events around them
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    struct { const char* name; bool coop, pdl; int syncs; } cases[] = {
        {"plain launches, 0 syncs", false, false, 0}, {"pdl launches, 0 syncs", false, true, 0},
        {"cooperative, 1 sync", true, false, 1}, {"cooperative, 2 syncs", true, false, 2},
        {"cooperative + pdl, 1 sync", true, true, 1}, {"cooperative + pdl, 2 syncs", true, true, 2}};
    for (auto& c : cases) {
        for (int i = 0; i < 8; ++i) launch(c.coop, c.pdl, grid, smem, st, x, n, c.syncs);
        cudaStreamSynchronize(st);
        if (cudaGetLastError() != cudaSuccess) { printf("%s: not launchable\n", c.name); continue; }
        cudaEventRecord(a, st);
        for (int i = 0; i < 64; ++i) launch(c.coop, c.pdl, grid, smem, st, x, n, c.syncs);
        cudaEventRecord(b, st);
        cudaStreamSynchronize(st);
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        printf("%-28s %7.2f us per launch (%s)\n", c.name, ms / 64 * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    // the same inside a CUDA graph
    cudaGraph_t gr; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    cudaError_t ec = cudaSuccess;
    for (int i = 0; i < 16 && ec == cudaSuccess; ++i) ec = launch(true, true, grid, smem, st, x, n, 1);
    cudaError_t ee = cudaStreamEndCapture(st, &gr);
    printf("graph capture of cooperative + pdl launches: launch %s, end %s\n", cudaGetErrorString(ec), cudaGetErrorString(ee));
    if (ee == cudaSuccess && cudaGraphInstantiate(&ge, gr, 0) == cudaSuccess) {
        cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
        cudaEventRecord(a, st);
        for (int i = 0; i < 8; ++i) cudaGraphLaunch(ge, st);
        cudaEventRecord(b, st); cudaStreamSynchronize(st);
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        printf("graph of 16 cooperative + pdl launches: %7.2f us per launch (%s)\n", ms / 128 * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
