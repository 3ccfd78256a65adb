// sm_100a kernels for the trie mass path (weight_sum / weight_max over a batch of rows).
//
// Replaces genlm/backend/trie/parallel.py:92-145 (sparse.mm / scatter_reduce amax) and
// genlm/backend/trie/base.py:346-393 (numba loops).  HBM-bound integer/float streaming work:
// no tensor cores; the design rules are coalesced 128-bit global access, shared-memory staging of
// every irregular access, and enough CTAs in flight to keep HBM busy.
//
//   phase 1  permute_kernel : row segment (vocabulary order, coalesced 128-bit loads, exp/cast fused)
//                             -> shared memory -> tile-major staging rows z (coalesced 128-bit stores,
//                             L2-resident scratch).  All scattered accesses hit shared memory only.
//   phase 2  tile_kernel    : staged tile -> DFS-ordered leaf values in shared memory -> aligned-block
//                             pyramid (warp shuffles) -> multi-term nodes -> coalesced emit of the
//                             tile's node-id interval.
//   phase 3  span_kernel    : the few nodes whose leaf range crosses a tile boundary, reduced from the
//                             already written values of their maximal in-tile descendants (fp64 for sums).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <type_traits>
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

#include "trie_internal.h"

namespace gt {

struct PlanView {
    int32_t T, logT, Q, NT, NS, SV;  // SV = per-row value-array length in shared memory
    int64_t V, N, Zrow;
    const int32_t* p1_chunk_ptr; const int32_t* p1_zoff; const uint16_t* p1_src;
    const int32_t* z_tile_off; const uint16_t* p2_slot;
    const int32_t* br_ptr; const int32_t* br_child_ptr; const uint16_t* br_child;
    const int32_t* tile_node_lo; const uint16_t* node_slot;
    int32_t n_span; const int32_t* span_node; const int32_t* span_ptr; const int32_t* span_term;
};

struct DevicePlan {
    int device = -1;
    void* blob = nullptr;  // one allocation holding all metadata
    size_t blob_bytes = 0;
    PlanView view{};
    bool attrs_set = false;
};

void free_device_plan(DevicePlan* d) {
    if (!d) return;
    if (d->blob) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(d->device);
        cudaFree(d->blob);
        cudaSetDevice(cur);
    }
    delete d;
}

#define GT_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            gt::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            (void)cudaGetLastError(); /* do not leave the error for the next, unrelated call */        \
            return GT_ERR_CUDA;                                                           \
        }                                                                                 \
    } while (0)

constexpr int kThreads = 512;
constexpr int OP_SUM = 1, OP_MAX = 2;
constexpr size_t kMaxSmem = 200 * 1024;  // dynamic shared memory we are willing to ask for per CTA
constexpr int kSegPad = 8;  // slack so a row segment can be stored at its global 16-byte phase

// ---- small device helpers -----------------------------------------------------------------------

template <int OP, typename VT> __device__ __forceinline__ VT op_apply(VT a, VT b) {
    if constexpr (OP == OP_SUM) return a + b;
    else return fmax(a, b);  // fmaxf/fmax overloads; NaN operands are ignored like numba's max()
}
template <int OP, typename VT> __device__ __forceinline__ VT op_ident() {
    if constexpr (OP == OP_SUM) return VT(0);
    else return -std::numeric_limits<VT>::infinity();
}

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <typename VT> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

template <typename VT> __device__ __forceinline__ void store4(VT* p, VT a, VT b, VT c, VT d);
template <> __device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void store4<double>(double* p, double a, double b, double c, double d) {
    reinterpret_cast<double2*>(p)[0] = make_double2(a, b);
    reinterpret_cast<double2*>(p)[1] = make_double2(c, d);
}
template <typename VT> __device__ __forceinline__ void load4(const VT* p, VT& a, VT& b, VT& c, VT& d);
template <> __device__ __forceinline__ void load4<float>(const float* p, float& a, float& b, float& c, float& d) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    a = v.x; b = v.y; c = v.z; d = v.w;
}
template <> __device__ __forceinline__ void load4<double>(const double* p, double& a, double& b, double& c, double& d) {
    const double2 u = reinterpret_cast<const double2*>(p)[0], w = reinterpret_cast<const double2*>(p)[1];
    a = u.x; b = u.y; c = w.x; d = w.y;
}

template <typename IN_T> __device__ __forceinline__ float in_to_float(IN_T x);
template <> __device__ __forceinline__ float in_to_float<float>(float x) { return x; }
template <> __device__ __forceinline__ float in_to_float<__half>(__half x) { return __half2float(x); }
template <> __device__ __forceinline__ float in_to_float<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }

template <typename VT, typename IN_T> __device__ __forceinline__ VT convert_in(IN_T x, bool log_input) {
    if constexpr (sizeof(IN_T) == 8) {
        const double d = log_input ? exp((double)x) : (double)x;
        return (VT)d;
    } else {
        const float f = in_to_float<IN_T>(x);
        if constexpr (sizeof(VT) == 8) return log_input ? exp((double)f) : (double)f;
        else return log_input ? expf(f) : f;
    }
}

// ---- phase 1: permute a row segment into the tile-major staging layout ------------------------------

// Loads n elements starting at `row` into dst[phase + i], where phase = element offset of `row` inside
// its 16-byte line, so that the vector body is aligned on both sides.  Returns nothing; caller syncs.
template <typename VT, typename IN_T>
__device__ __forceinline__ void load_segment(const IN_T* __restrict__ row, int n, VT* dst, bool log_input) {
    constexpr int EPV = 16 / (int)sizeof(IN_T);
    const int tid = threadIdx.x;
    const int phase = (int)((reinterpret_cast<uintptr_t>(row) & 15) / sizeof(IN_T));
    int head = (EPV - phase) & (EPV - 1);
    if (head > n) head = n;
    VT* d = dst + phase;
    for (int i = tid; i < head; i += kThreads) d[i] = convert_in<VT, IN_T>(row[i], log_input);
    const int nvec = (n - head) / EPV;
    const uint4* v = reinterpret_cast<const uint4*>(row + head);
    for (int i = tid; i < nvec; i += kThreads) {
        const uint4 raw = ldg_stream(v + i);
        const IN_T* e = reinterpret_cast<const IN_T*>(&raw);
        VT* o = d + head + i * EPV;  // (phase + head) % EPV == 0 -> aligned vector stores
#pragma unroll
        for (int k = 0; k < EPV; k += 4)
            store4<VT>(o + k, convert_in<VT, IN_T>(e[k], log_input), convert_in<VT, IN_T>(e[k + 1], log_input),
                       convert_in<VT, IN_T>(e[k + 2], log_input), convert_in<VT, IN_T>(e[k + 3], log_input));
    }
    for (int i = head + nvec * EPV + tid; i < n; i += kThreads) d[i] = convert_in<VT, IN_T>(row[i], log_input);
}
template <typename VT>
__device__ __forceinline__ void load_segment_f64(const double* __restrict__ row, int n, VT* dst, bool log_input) {
    // fp64 rows: 2 elements per 16 bytes; plain coalesced 8-byte loads are already full-width per warp.
    const int phase = (int)((reinterpret_cast<uintptr_t>(row) & 15) / sizeof(double));
    VT* d = dst + phase;
    for (int i = threadIdx.x; i < n; i += kThreads) d[i] = convert_in<VT, double>(row[i], log_input);
}

template <typename VT, typename IN_T, int R>
__global__ void __launch_bounds__(kThreads) permute_kernel(PlanView P, const IN_T* __restrict__ ws, int64_t ld_ws,
                                                           VT* __restrict__ z, int n_rows, int log_input) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    VT* seg = reinterpret_cast<VT*>(smem_raw);  // [R][Q + kSegPad]
    const int pitch = P.Q + kSegPad;
    const int s = blockIdx.x;
    const int b0 = blockIdx.y * R;
    const int nrows = min(R, n_rows - b0);
    const int seg_lo = s * P.Q;
    const int seg_n = (int)min((int64_t)P.Q, P.V - seg_lo);

    int phase[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        phase[r] = 0;
        if (r < nrows) {
            const IN_T* row = ws + (size_t)(b0 + r) * ld_ws + seg_lo;
            phase[r] = (int)((reinterpret_cast<uintptr_t>(row) & 15) / sizeof(IN_T));
            if constexpr (sizeof(IN_T) == 8) load_segment_f64<VT>(row, seg_n, seg + r * pitch, log_input != 0);
            else load_segment<VT, IN_T>(row, seg_n, seg + r * pitch, log_input != 0);
        }
    }
    __syncthreads();

    const int c0 = P.p1_chunk_ptr[s], c1 = P.p1_chunk_ptr[s + 1];
    const ushort4* src4 = reinterpret_cast<const ushort4*>(P.p1_src);
    for (int c = c0 + threadIdx.x; c < c1; c += kThreads) {
        const int zo = P.p1_zoff[c];
        const ushort4 src = src4[c];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r < nrows) {
                const VT* sr = seg + r * pitch + phase[r];
                const VT a = src.x != 0xFFFF ? sr[src.x] : VT(0);
                const VT b = src.y != 0xFFFF ? sr[src.y] : VT(0);
                const VT cc = src.z != 0xFFFF ? sr[src.z] : VT(0);
                const VT d = src.w != 0xFFFF ? sr[src.w] : VT(0);
                store4<VT>(z + (size_t)(b0 + r) * P.Zrow + zo, a, b, cc, d);
            }
        }
    }
}

// ---- phase 2: per-tile pyramid, multi-term nodes, emit -----------------------------------------------

template <typename VT, int R, int OP>
__global__ void __launch_bounds__(kThreads) tile_kernel(PlanView P, const VT* __restrict__ z, VT* __restrict__ out,
                                                        int64_t ld_out, int n_rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    VT* vals = reinterpret_cast<VT*>(smem_raw);  // [R][SV]
    const int T = P.T, SV = P.SV;
    const int t = blockIdx.x;
    const int b0 = blockIdx.y * R;
    const int nrows = min(R, n_rows - b0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kThreads / 32;

    // 1. staged tile -> DFS-ordered leaf slots
    {
        const int zlo = P.z_tile_off[t];
        const int zn4 = (P.z_tile_off[t + 1] - zlo) >> 2;
        const ushort4* slot4 = reinterpret_cast<const ushort4*>(P.p2_slot + zlo);
        for (int i = tid; i < zn4; i += kThreads) {
            const ushort4 sl = slot4[i];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r < nrows) {
                    VT a, b, c, d;
                    load4<VT>(z + (size_t)(b0 + r) * P.Zrow + zlo + 4 * i, a, b, c, d);
                    VT* lv = vals + r * SV;
                    if (sl.x != 0xFFFF) lv[sl.x] = a;
                    if (sl.y != 0xFFFF) lv[sl.y] = b;
                    if (sl.z != 0xFFFF) lv[sl.z] = c;
                    if (sl.w != 0xFFFF) lv[sl.w] = d;
                }
            }
        }
        const int nleaf = (int)min((int64_t)T, P.V - (int64_t)t * T);
        for (int i = nleaf + tid; i < T; i += kThreads)
#pragma unroll
            for (int r = 0; r < R; ++r) vals[r * SV + i] = op_ident<OP, VT>();
    }
    __syncthreads();

    // 2. pyramid of aligned blocks: level k block i at 2T - (T >> (k-1)) + i.
    //    Levels 1..3 inside a thread (8 consecutive leaves), 4..8 by warp shuffles (256 leaves per warp).
    for (int ub = warp * 32; ub < (T >> 3); ub += kWarps * 32) {
        const int u = ub + lane;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r < nrows) {
                VT* lv = vals + r * SV;
                VT x0, x1, x2, x3, x4, x5, x6, x7;
                load4<VT>(lv + 8 * u, x0, x1, x2, x3);
                load4<VT>(lv + 8 * u + 4, x4, x5, x6, x7);
                const VT a0 = op_apply<OP>(x0, x1), a1 = op_apply<OP>(x2, x3);
                const VT a2 = op_apply<OP>(x4, x5), a3 = op_apply<OP>(x6, x7);
                store4<VT>(lv + T + 4 * u, a0, a1, a2, a3);
                const VT c0 = op_apply<OP>(a0, a1), c1 = op_apply<OP>(a2, a3);
                lv[2 * T - (T >> 1) + 2 * u] = c0;
                lv[2 * T - (T >> 1) + 2 * u + 1] = c1;
                VT x = op_apply<OP>(c0, c1);
                lv[2 * T - (T >> 2) + u] = x;
#pragma unroll
                for (int j = 1; j <= 5; ++j) {
                    const VT y = __shfl_down_sync(0xffffffffu, x, 1 << (j - 1));
                    x = op_apply<OP>(x, y);
                    if ((lane & ((1 << j) - 1)) == 0) lv[2 * T - (T >> (2 + j)) + (u >> j)] = x;
                }
            }
        }
    }
    __syncthreads();
    //    Levels 9..logT: T/256 <= 32 level-8 blocks, one warp per row.
    if (warp < nrows && P.logT > 8) {
        VT* lv = vals + warp * SV;
        const int n8 = T >> 8;
        VT x = lane < n8 ? lv[2 * T - (T >> 7) + lane] : op_ident<OP, VT>();
        for (int j = 1; j <= P.logT - 8; ++j) {
            const VT y = __shfl_down_sync(0xffffffffu, x, 1 << (j - 1));
            x = op_apply<OP>(x, y);
            if ((lane & ((1 << j) - 1)) == 0 && lane < n8) lv[2 * T - (T >> (7 + j)) + (lane >> j)] = x;
        }
    }
    __syncthreads();

    // 3. nodes that need more than one block: short reductions over leaf / pyramid slots
    {
        const int j0 = P.br_ptr[t];
        const int nb = P.br_ptr[t + 1] - j0;
        for (int j = tid; j < nb; j += kThreads) {
            const int p0 = P.br_child_ptr[j0 + j], p1 = P.br_child_ptr[j0 + j + 1];
            VT acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = op_ident<OP, VT>();
            for (int p = p0; p < p1; ++p) {
                const int sl = P.br_child[p];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = op_apply<OP>(acc[r], vals[r * SV + sl]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) vals[r * SV + 2 * T + j] = acc[r];
        }
    }
    __syncthreads();

    // 4. emit the tile's node-id interval, coalesced
    {
        const int n0 = P.tile_node_lo[t], n1 = P.tile_node_lo[t + 1];
        for (int n = n0 + tid; n < n1; n += kThreads) {
            const int sl = P.node_slot[n];
            if (sl != 0xFFFF) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (r < nrows) __stcs(out + (size_t)(b0 + r) * ld_out + n, vals[r * SV + sl]);
            }
        }
    }
}

// ---- phase 3: nodes spanning tiles -------------------------------------------------------------------

template <typename VT, int OP>
__global__ void __launch_bounds__(256) span_kernel(PlanView P, VT* __restrict__ out, int64_t ld_out, int n_rows) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= P.n_span) return;
    const int node = P.span_node[i];
    const int p0 = P.span_ptr[i], p1 = P.span_ptr[i + 1];
    using AT = typename std::conditional<OP == OP_SUM, double, VT>::type;
    for (int b = blockIdx.y; b < n_rows; b += gridDim.y) {
        const VT* row = out + (size_t)b * ld_out;
        AT acc = op_ident<OP, AT>();
        for (int p = p0 + lane; p < p1; p += 32) acc = op_apply<OP, AT>(acc, (AT)row[P.span_term[p]]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc = op_apply<OP, AT>(acc, __shfl_xor_sync(0xffffffffu, acc, o));
        if (lane == 0) out[(size_t)b * ld_out + node] = p1 > p0 ? (VT)acc : VT(0);
    }
}

// ---- host side ---------------------------------------------------------------------------------------

static DevicePlan* upload_plan(const Layout& L, const Plan& P, int device) {
    struct Part { const void* src; size_t bytes; size_t off; };
    std::vector<Part> parts;
    size_t total = 0;
    auto add = [&](const void* src, size_t bytes) {
        const size_t off = total;
        parts.push_back({src, bytes, off});
        total += (bytes + 255) & ~size_t(255);
        return off;
    };
#define ADDV(v) add((v).data(), (v).size() * sizeof((v)[0]))
    const size_t o_p1_chunk_ptr = ADDV(P.p1_chunk_ptr), o_p1_zoff = ADDV(P.p1_zoff), o_p1_src = ADDV(P.p1_src);
    const size_t o_z_tile_off = ADDV(P.z_tile_off), o_p2_slot = ADDV(P.p2_slot);
    const size_t o_br_ptr = ADDV(P.br_ptr), o_br_child_ptr = ADDV(P.br_child_ptr), o_br_child = ADDV(P.br_child);
    const size_t o_tile_node_lo = ADDV(P.tile_node_lo), o_node_slot = ADDV(P.node_slot);
    const size_t o_span_node = ADDV(P.span_node), o_span_ptr = ADDV(P.span_ptr), o_span_term = ADDV(P.span_term);
#undef ADDV
    total += 256;

    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) {
        set_error("cannot select CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    DevicePlan* d = new DevicePlan();
    d->device = device;
    cudaError_t e = cudaMalloc(&d->blob, total);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu) for trie metadata failed: %s", total, cudaGetErrorString(e));
        d->blob = nullptr;
        cudaSetDevice(cur);
        free_device_plan(d);
        return nullptr;
    }
    d->blob_bytes = total;
    std::vector<unsigned char> host(total, 0);
    for (const Part& p : parts) if (p.bytes) memcpy(host.data() + p.off, p.src, p.bytes);
    e = cudaMemcpy(d->blob, host.data(), total, cudaMemcpyHostToDevice);
    cudaSetDevice(cur);
    if (e != cudaSuccess) {
        set_error("metadata upload failed: %s", cudaGetErrorString(e));
        free_device_plan(d);
        return nullptr;
    }

    unsigned char* base = static_cast<unsigned char*>(d->blob);
    PlanView& v = d->view;
    v.T = P.T; v.logT = 0; while ((1 << (v.logT + 1)) <= P.T) ++v.logT;
    v.Q = P.Q; v.NT = P.NT; v.NS = P.NS;
    v.SV = (P.max_tile_values + 3) & ~3;
    v.V = L.V; v.N = L.N; v.Zrow = P.Zrow;
    v.p1_chunk_ptr = (const int32_t*)(base + o_p1_chunk_ptr); v.p1_zoff = (const int32_t*)(base + o_p1_zoff);
    v.p1_src = (const uint16_t*)(base + o_p1_src);
    v.z_tile_off = (const int32_t*)(base + o_z_tile_off); v.p2_slot = (const uint16_t*)(base + o_p2_slot);
    v.br_ptr = (const int32_t*)(base + o_br_ptr); v.br_child_ptr = (const int32_t*)(base + o_br_child_ptr);
    v.br_child = (const uint16_t*)(base + o_br_child);
    v.tile_node_lo = (const int32_t*)(base + o_tile_node_lo); v.node_slot = (const uint16_t*)(base + o_node_slot);
    v.n_span = (int32_t)P.span_node.size();
    v.span_node = (const int32_t*)(base + o_span_node); v.span_ptr = (const int32_t*)(base + o_span_ptr);
    v.span_term = (const int32_t*)(base + o_span_term);
    return d;
}

constexpr int R_F32 = 2;  // rows per CTA, float pipeline
constexpr int R_F64 = 1;  // rows per CTA, double pipeline

template <typename VT, int R> static size_t permute_smem(const PlanView& v) { return (size_t)R * (v.Q + kSegPad) * sizeof(VT); }
template <typename VT, int R> static size_t tile_smem(const PlanView& v) { return (size_t)R * v.SV * sizeof(VT); }

// Opt in to > 48 KB dynamic shared memory once per (kernel, device, size): the attribute call is kept off the
// steady-state launch path (and out of CUDA graph captures).  Keyed by the kernel's address: instantiations
// with the same signature share a function type.
static cudaError_t allow_smem_impl(const void* kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> granted;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(kernel, dev);
    auto it = granted.find(key);
    if (it != granted.end() && it->second >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) granted[key] = bytes;
    return e;
}
template <typename K> static cudaError_t allow_smem(K kernel, size_t bytes) {
    return allow_smem_impl(reinterpret_cast<const void*>(kernel), bytes);
}

template <typename VT, typename IN_T, int R>
static int launch_permute(const PlanView& v, const void* ws, int64_t ld_ws, VT* z, int rows, bool log_input, cudaStream_t st) {
    const size_t smem = permute_smem<VT, R>(v);
    GT_CUDA(allow_smem(permute_kernel<VT, IN_T, R>, smem));
    dim3 grid((unsigned)v.NS, (unsigned)((rows + R - 1) / R));
    permute_kernel<VT, IN_T, R><<<grid, kThreads, smem, st>>>(v, static_cast<const IN_T*>(ws), ld_ws, z, rows, log_input ? 1 : 0);
    GT_CUDA(cudaGetLastError());
    return GT_OK;
}

template <typename VT, int R, int OP>
static int launch_tile_and_span(const PlanView& v, const VT* z, VT* out, int64_t ld_out, int rows, unsigned phases,
                                cudaStream_t st) {
    if (v.NT > 0 && (phases & GT_FLAG_PHASE_TILE)) {
        const size_t smem = tile_smem<VT, R>(v);
        GT_CUDA(allow_smem(tile_kernel<VT, R, OP>, smem));
        dim3 grid((unsigned)v.NT, (unsigned)((rows + R - 1) / R));
        tile_kernel<VT, R, OP><<<grid, kThreads, smem, st>>>(v, z, out, ld_out, rows);
        GT_CUDA(cudaGetLastError());
    }
    if (v.n_span > 0 && (phases & GT_FLAG_PHASE_SPAN)) {
        dim3 grid((unsigned)((v.n_span + 7) / 8), (unsigned)std::min(rows, 1024));
        span_kernel<VT, OP><<<grid, 256, 0, st>>>(v, out, ld_out, rows);
        GT_CUDA(cudaGetLastError());
    }
    return GT_OK;
}

template <typename VT, int R>
static int reduce_typed(const PlanView& v, const void* ws, int in_type, int64_t n_rows, int64_t ld_ws, void* out_sum,
                        void* out_max, int64_t ld_out, unsigned ops, unsigned flags, void* workspace,
                        size_t workspace_bytes, cudaStream_t st) {
    const size_t row_bytes = (size_t)v.Zrow * sizeof(VT);
    int64_t chunk = row_bytes ? (int64_t)(workspace_bytes / row_bytes) : n_rows;
    if (v.Zrow > 0 && chunk < 1) { set_error("workspace too small: need at least %zu bytes", row_bytes); return GT_ERR_STATE; }
    chunk = std::min<int64_t>(chunk, 32768);
    if (v.Zrow == 0) chunk = 32768;
    const bool log_input = (flags & GT_FLAG_LOG_INPUT) != 0;
    const unsigned phases = (flags & GT_FLAG_PHASE_MASK) ? (flags & GT_FLAG_PHASE_MASK) : GT_FLAG_PHASE_MASK;
    VT* z = static_cast<VT*>(workspace);
    const size_t in_size = in_type == GT_F64 ? 8 : in_type == GT_F32 ? 4 : 2;
    for (int64_t r0 = 0; r0 < n_rows; r0 += chunk) {
        const int rows = (int)std::min<int64_t>(chunk, n_rows - r0);
        const void* wsr = static_cast<const char*>(ws) + (size_t)r0 * ld_ws * in_size;
        int rc = GT_OK;
        if (v.NT > 0 && (phases & GT_FLAG_PHASE_PERMUTE)) {
            // rows per CTA in the permute phase: R unless the segment buffer would not fit in shared memory
            const bool wide = permute_smem<VT, R>(v) <= kMaxSmem;
#define GT_PERMUTE(IN_T) (wide ? launch_permute<VT, IN_T, R>(v, wsr, ld_ws, z, rows, log_input, st) \
                               : launch_permute<VT, IN_T, 1>(v, wsr, ld_ws, z, rows, log_input, st))
            switch (in_type) {
                case GT_F32: rc = GT_PERMUTE(float); break;
                case GT_F64: rc = GT_PERMUTE(double); break;
                case GT_F16: rc = GT_PERMUTE(__half); break;
                case GT_BF16: rc = GT_PERMUTE(__nv_bfloat16); break;
                default: set_error("unknown input type %d", in_type); return GT_ERR_ARG;
            }
#undef GT_PERMUTE
            if (rc != GT_OK) return rc;
        }
        if (ops & GT_OP_SUM) {
            rc = launch_tile_and_span<VT, R, OP_SUM>(v, z, static_cast<VT*>(out_sum) + (size_t)r0 * ld_out, ld_out, rows, phases, st);
            if (rc != GT_OK) return rc;
        }
        if (ops & GT_OP_MAX) {
            rc = launch_tile_and_span<VT, R, OP_MAX>(v, z, static_cast<VT*>(out_max) + (size_t)r0 * ld_out, ld_out, rows, phases, st);
            if (rc != GT_OK) return rc;
        }
    }
    return GT_OK;
}

}  // namespace gt

extern "C" {

int gt_upload(gt_trie* t, int device) {
    if (!t) { gt::set_error("gt_upload: null trie"); return GT_ERR_ARG; }
    if (t->dev.count(device)) return GT_OK;
    if (!t->plan) {
        const int rc = gt_plan(t, 0, 0);
        if (rc != GT_OK) return rc;
    }
    gt::DevicePlan* d = gt::upload_plan(t->layout, *t->plan, device);
    if (!d) return GT_ERR_CUDA;
    t->dev[device] = d;
    return GT_OK;
}

int gt_get_plan_info(const gt_trie* t, gt_plan_info* info) {
    if (!t || !info) { gt::set_error("gt_get_plan_info: bad argument"); return GT_ERR_ARG; }
    if (!t->plan) { gt::set_error("gt_get_plan_info: trie has no plan yet (call gt_upload)"); return GT_ERR_STATE; }
    const gt::Plan& P = *t->plan;
    memset(info, 0, sizeof *info);
    info->n_tokens = t->layout.V; info->n_nodes = t->layout.N;
    info->tile_leaves = P.T; info->seg_positions = P.Q; info->n_tiles = P.NT; info->n_segs = P.NS;
    info->rows_per_item = gt::R_F32;
    info->n_span = (int32_t)P.span_node.size(); info->span_terms = (int64_t)P.span_term.size();
    info->max_levels = P.max_levels; info->max_tile_values = P.max_tile_values;
    info->staged_row_elems = P.Zrow;
    size_t meta = 0;
    for (auto& kv : t->dev) { meta = kv.second->blob_bytes; break; }
    info->meta_bytes = (int64_t)meta;
    return GT_OK;
}

size_t gt_workspace_bytes(const gt_trie* t, int64_t max_rows) {
    if (!t || !t->plan || max_rows <= 0) return 0;
    return (size_t)max_rows * (size_t)t->plan->Zrow * sizeof(double) + 256;
}

int gt_weight_reduce(const gt_trie* t, const void* ws, int in_type, int64_t n_rows, int64_t ld_ws, void* out_sum,
                     void* out_max, int out_type, int64_t ld_out, unsigned ops, unsigned flags, void* workspace,
                     size_t workspace_bytes, gt_stream stream) {
    if (!t) { gt::set_error("gt_weight_reduce: null trie"); return GT_ERR_ARG; }
    if (n_rows < 0 || !(ops & (GT_OP_SUM | GT_OP_MAX)) || (ops & ~(unsigned)(GT_OP_SUM | GT_OP_MAX))) {
        gt::set_error("gt_weight_reduce: bad n_rows / ops"); return GT_ERR_ARG;
    }
    if (n_rows == 0) return GT_OK;
    if ((t->layout.V > 0 && !ws) || ((ops & GT_OP_SUM) && !out_sum) || ((ops & GT_OP_MAX) && !out_max)) {
        gt::set_error("gt_weight_reduce: null data pointer"); return GT_ERR_ARG;
    }
    if (ld_ws < t->layout.V || ld_out < t->layout.N) {
        gt::set_error("gt_weight_reduce: row stride smaller than row length (ld_ws=%lld V=%lld ld_out=%lld N=%lld)",
                      (long long)ld_ws, (long long)t->layout.V, (long long)ld_out, (long long)t->layout.N);
        return GT_ERR_ARG;
    }
    int device = -1;
    GT_CUDA(cudaGetDevice(&device));
    auto it = t->dev.find(device);
    if (it == t->dev.end()) { gt::set_error("trie metadata is not resident on device %d (call gt_upload)", device); return GT_ERR_STATE; }
    const gt::PlanView& v = it->second->view;
    if (v.Zrow > 0 && !workspace) { gt::set_error("gt_weight_reduce: null workspace"); return GT_ERR_ARG; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (out_type == GT_F32)
        return gt::reduce_typed<float, gt::R_F32>(v, ws, in_type, n_rows, ld_ws, out_sum, out_max, ld_out, ops, flags,
                                                  workspace, workspace_bytes, st);
    if (out_type == GT_F64)
        return gt::reduce_typed<double, gt::R_F64>(v, ws, in_type, n_rows, ld_ws, out_sum, out_max, ld_out, ops, flags,
                                                   workspace, workspace_bytes, st);
    gt::set_error("gt_weight_reduce: output type must be GT_F32 or GT_F64");
    return GT_ERR_ARG;
}

}  // extern "C"
