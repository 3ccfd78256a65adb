// sm_100a kernels for the trie mass path (weight_sum / weight_max over a batch of rows).
//
// Replaces genlm/backend/trie/parallel.py:92-145 (sparse.mm / scatter_reduce amax) and
// genlm/backend/trie/base.py:346-393 (numba loops).  HBM-bound streaming work: no tensor cores; the design rules
// are coalesced global access, shared-memory staging of every irregular access, bulk copies (TMA) for everything
// that is contiguous, and one persistent grid that keeps HBM busy from the first row to the last.
//
// A reduction is three launches, chained with programmatic dependent launch:
//   permute_kernel  reads the rows in vocabulary order (coalesced) and writes every weight to its DFS leaf slot of
//                   the staging buffer z (16-byte scattered stores that stay in L2):
//                   z[row group][tile][leaf slot][row] is, tile by tile, the leaf region of the tile's shared-memory
//                   value array, so the tile kernel needs no scatter phase of its own;
//   mass_kernel     persistent, warp-specialised: a producer thread fetches (row group, tile) pairs -- leaf block +
//                   tile metadata -- by bulk copies (TMA); compute warps build the aligned-block pyramid (warp
//                   shuffles) and the multi-term ranges (ELL-packed term lists); emit warps stream the tile's node-id
//                   interval out with coalesced stores and write the spanning-node pieces;
//   span_kernel     the few nodes whose leaf range crosses tiles, reduced from their pieces (fp64 for sums).
// (Variants that hide the permute inside mass_kernel were built and measured in round 2 and are not here: publishing
// permuted row groups to other SMs of the same launch needs a device-scope fence, which takes ~10 us on an SM saturated
// with output stores; staging the *next* launch's rows from the producer warp works but one warp per CTA is 2.5x too
// slow, and rider code in the other warps costs the tile work 9 us of register allocation; profiles/r02_experiments.md.)
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
    int32_t T, logT, NT, SV;  // SV = value slots per tile in shared memory (leaves + pyramid + multi-term ranges)
    int32_t R, max_tile_nodes, max_tile_ell_rows, max_tile_chunks;
    int64_t V, N, ZG;         // ZG = NT * T: value slots per row group of the staging buffer
    const int32_t* leaf_dest;  // [V] item position -> value slot of the row group's staging block
    const int32_t* ell_chunk_ptr; const int2* ell_desc; const uint16_t* ell_terms; const int32_t* ell_row_ptr;
    const int32_t* tile_node_lo; const uint16_t* node_slot;
    const int32_t* piece_ptr; const uint16_t* piece_slot; const int32_t* piece_idx;
    int32_t n_span, n_pieces; const int32_t* span_node; const int32_t* span_pp;
    const int32_t* leaf_rank;  // [V] DFS rank of each item's leaf (inverse of Layout::perm)
    const int32_t* node_lo; const int32_t* node_hi;  // [N] DFS leaf range of every node
    int32_t debug_stop;  // profiling aid (GT_DEBUG_STOP): 0 = normal; 3 = no emit stores; 9 = emit only
    long long* trace;    // profiling aid (GT_TRACE=1): per (CTA, item) SM-clock stamps of the pipeline events, else null
};

struct DevicePlan {
    int device = -1;
    void* blob = nullptr;  // one allocation holding all metadata
    size_t blob_bytes = 0;
    PlanView view{};
    bool attrs_set = false;
    long long* trace = nullptr;  // GT_TRACE=1 only
};

void free_device_plan(DevicePlan* d) {
    if (!d) return;
    if (d->blob) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(d->device);
        cudaFree(d->blob);
        if (d->trace) cudaFree(d->trace);
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

constexpr int OP_SUM = 1, OP_MAX = 2;

// ---- programmatic dependent launch ---------------------------------------------------------------------
// Consecutive kernels of one call (mass -> span -> mass -> span ...) are launched with the programmatic
// stream-serialisation attribute: a kernel's launch set-up and prologue overlap the tail of its predecessor, and
// pdl_wait() -- executed before the first access to memory another kernel of the chain touches -- blocks until the
// predecessor grid has completed and flushed.  pdl_trigger() lets the successor's CTAs start launching.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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
template <typename VT> __device__ __forceinline__ void store4_stream(VT* p, VT a, VT b, VT c, VT d);
template <> __device__ __forceinline__ void store4_stream<float>(float* p, float a, float b, float c, float d) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(a, b, c, d));
}
template <> __device__ __forceinline__ void store4_stream<double>(double* p, double a, double b, double c, double d) {
    __stcs(reinterpret_cast<double2*>(p), make_double2(a, b));
    __stcs(reinterpret_cast<double2*>(p) + 1, make_double2(c, d));
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


// mbarrier + bulk-copy (TMA) primitives.  One elected thread arms the barrier with the byte count and issues the
// copies; the copy engine signals the barrier when the bytes have landed, so no LSU instruction is spent per 16 bytes.
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(b);
    unsigned done;
    for (;;) {  // try_wait suspends the thread in hardware for a bounded time; back off between polls so that
                // waiting warps do not take issue slots from the working ones
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(40);
    }
}
// bytes: multiple of 16; both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b))
                 : "memory");
}

// Device twin of gt::swizzle_slot (trie_internal.h) for slots below 2T; B = bytes per slot.
template <int B> __device__ __forceinline__ int swz(int s) {
    constexpr int cs = B == 4 ? 2 : (B == 8 ? 1 : 0);
    const int c = s >> cs;
    return ((c ^ ((c >> 3) & (B / 2 - 1))) << cs) | (s & ((1 << cs) - 1));
}

// R consecutive rows share a CTA and live interleaved in shared memory: vals[slot * R + r].  One vector
// shared-memory access then serves all R rows of a slot.
template <typename VT, int R> struct RowVec {
    VT v[R];
    __device__ __forceinline__ static RowVec load(const VT* p) {
        RowVec x;
        if constexpr (sizeof(VT) * R == 8) { const float2 t = *reinterpret_cast<const float2*>(p); memcpy(x.v, &t, 8); }
        else if constexpr (sizeof(VT) * R == 16) { const float4 t = *reinterpret_cast<const float4*>(p); memcpy(x.v, &t, 16); }
        else {
#pragma unroll
            for (int r = 0; r < R; ++r) x.v[r] = p[r];
        }
        return x;
    }
    __device__ __forceinline__ void store(VT* p) const {
        if constexpr (sizeof(VT) * R == 8) { float2 t; memcpy(&t, v, 8); *reinterpret_cast<float2*>(p) = t; }
        else if constexpr (sizeof(VT) * R == 16) { float4 t; memcpy(&t, v, 16); *reinterpret_cast<float4*>(p) = t; }
        else {
#pragma unroll
            for (int r = 0; r < R; ++r) p[r] = v[r];
        }
    }
    template <int OP> __device__ __forceinline__ static RowVec combine(const RowVec& a, const RowVec& b) {
        RowVec x;
#pragma unroll
        for (int r = 0; r < R; ++r) x.v[r] = op_apply<OP>(a.v[r], b.v[r]);
        return x;
    }
    template <int OP> __device__ __forceinline__ static RowVec ident() {
        RowVec x;
#pragma unroll
        for (int r = 0; r < R; ++r) x.v[r] = op_ident<OP, VT>();
        return x;
    }
    __device__ __forceinline__ RowVec shfl_down(int delta) const {
        RowVec x;
#pragma unroll
        for (int r = 0; r < R; ++r) x.v[r] = __shfl_down_sync(0xffffffffu, v[r], delta);
        return x;
    }
};

// Streaming stores of one slot's R row values to R row pointers at a compile-time byte offset, all under one
// predicate: address = register + immediate, so a store is one instruction and no pointer is recomputed per store.
template <int OFF, typename VT, int R> struct EmitStore {
    __device__ __forceinline__ static void run(VT* const (&p)[R], const RowVec<VT, R>& x, bool ok) {
        if (ok) {
#pragma unroll
            for (int r = 0; r < R; ++r) __stcs(reinterpret_cast<VT*>(reinterpret_cast<unsigned char*>(p[r]) + OFF), x.v[r]);
        }
    }
};
template <int OFF> struct EmitStore<OFF, float, 4> {
    __device__ __forceinline__ static void run(float* const (&p)[4], const RowVec<float, 4>& x, bool ok) {
        asm volatile(
            "{\n.reg .pred q;\nsetp.ne.u32 q, %8, 0;\n"
            "@q st.global.cs.f32 [%0+%9], %4;\n@q st.global.cs.f32 [%1+%9], %5;\n"
            "@q st.global.cs.f32 [%2+%9], %6;\n@q st.global.cs.f32 [%3+%9], %7;\n}\n" ::"l"(p[0]), "l"(p[1]), "l"(p[2]), "l"(p[3]),
            "f"(x.v[0]), "f"(x.v[1]), "f"(x.v[2]), "f"(x.v[3]), "r"((unsigned)ok), "n"(OFF)
            : "memory");
    }
};
template <int OFF> struct EmitStore<OFF, double, 2> {
    __device__ __forceinline__ static void run(double* const (&p)[2], const RowVec<double, 2>& x, bool ok) {
        asm volatile(
            "{\n.reg .pred q;\nsetp.ne.u32 q, %4, 0;\n"
            "@q st.global.cs.f64 [%0+%5], %2;\n@q st.global.cs.f64 [%1+%5], %3;\n}\n" ::"l"(p[0]), "l"(p[1]), "d"(x.v[0]), "d"(x.v[1]),
            "r"((unsigned)ok), "n"(OFF)
            : "memory");
    }
};


// ---- the tile kernel -----------------------------------------------------------------------------------------------
//
// Work decomposition.  A *pair* is (row group g of R rows, tile t); pairs are numbered g-major (p = g * NT + t) and
// CTA b takes pairs b, b + grid, b + 2 grid, ....  A pair holds one *item* per requested reduction (sum, then max);
// both items read the same leaf block.
//
// Shared memory of a CTA:
//     leaf[2]   leaf blocks of the current and the next pair (T slots of 16 bytes: R rows interleaved per slot)
//     rest[2]   pyramid + multi-term-range slots of the item being built and the item being drained
//     meta[2]   per pair: ELL term rows + chunk descriptors (compute group), node -> slot table (emit group), header
// and eight mbarriers:
//     pairFull[2]   bulk copies of a pair have landed (producer -> both groups)
//     pairEmpty[2]  the emit group has drained the last item of the pair (-> producer)
//     full[2] / empty[2]   hand-over of the rest buffers between compute and emit group, per item
constexpr int kThreads = 512;
#ifndef GT_COMPUTE_WARPS
#define GT_COMPUTE_WARPS 7
#endif
#ifndef GT_EMIT_WARPS
#define GT_EMIT_WARPS 8
#endif
constexpr int kComputeWarps = GT_COMPUTE_WARPS, kEmitWarps = GT_EMIT_WARPS;
constexpr int kProducerWarp = kComputeWarps + kEmitWarps;  // one warp: lane 0 issues every bulk copy
static_assert(kProducerWarp + 1 == kThreads / 32, "compute + emit + producer warps fill the CTA");
constexpr int kComputeThreads = 32 * kComputeWarps, kEmitThreads = 32 * kEmitWarps;
#ifndef GT_PERM_UT
#define GT_PERM_UT 128
#endif
constexpr int kUnitTokens = GT_PERM_UT;  // vocabulary positions per permute unit (half for fp64 rows)

#ifndef GT_ELL_BATCH
#define GT_ELL_BATCH 4
#endif

// trace layout: [kTraceCtas][kTraceItems][kTraceEvents] SM-clock stamps
constexpr int kTraceCtas = 512, kTraceItems = 32, kTraceEvents = 12;
#define GT_TRACE(ev)                                                                                         \
    do {                                                                                                     \
        if (P.trace && gtid == 0 && blockIdx.x < kTraceCtas && k < kTraceItems)                               \
            P.trace[((size_t)blockIdx.x * kTraceItems + k) * kTraceEvents + (ev)] = clock64();               \
    } while (0)
#define GT_PTRACE(cond, kk, ev)                                                                              \
    do {                                                                                                     \
        if (P.trace && (cond) && blockIdx.x < kTraceCtas && (kk) < kTraceItems)                               \
            P.trace[((size_t)blockIdx.x * kTraceItems + (kk)) * kTraceEvents + (ev)] = clock64();            \
    } while (0)

// Shared-memory carve-up of mass_kernel (all sections 128-byte aligned: the bank group of a slot is slot % 8 in
// the leaf blocks and in the rest buffers alike).
struct MassSmem {
    size_t leaf, leaf_bytes, rest, rest_bytes, meta, meta_bytes, terms, desc, slots, hdr, bars, total;
    __host__ __device__ MassSmem(const PlanView& P, int slot_bytes) {
        auto up = [](size_t x) { return (x + 127) & ~size_t(127); };
        size_t o = 0;
        leaf_bytes = up((size_t)P.T * slot_bytes);
        leaf = o; o += 2 * leaf_bytes;
        rest_bytes = up((size_t)(P.SV - P.T) * slot_bytes);
        rest = o; o += 2 * rest_bytes;
        // inside one pair's metadata block
        size_t m = 0;
        terms = m; m += up((size_t)P.max_tile_ell_rows * 64);
        desc = m;  m += up((size_t)(P.max_tile_chunks + 2) * 8);
        slots = m; m += up((size_t)P.max_tile_nodes * 2 + 16);
        hdr = m;   m += 128;
        meta_bytes = m;
        meta = o; o += 2 * meta_bytes;
        bars = o; o += 128;
        total = o;
    }
};

__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
template <int ID, int COUNT> __device__ __forceinline__ void group_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}

// What one launch works on.
template <typename VT> struct MassArgs {
    const void* ws; int in_type; int log_input; int64_t ld_ws;
    int dfs_order;  // the rows are already in DFS leaf order (GT_FLAG_DFS_ORDER): destinations are computed, not looked up
    VT* z;
    VT* out_sum; VT* out_max;
    VT* part_sum; VT* part_max;
    int64_t ld_out;
    int n_rows;
    unsigned ops;
};

// ---- permute ---------------------------------------------------------------------------------------------------
// Unit w = (row group g, positions [u * UT, (u + 1) * UT)), UT = 128 (64 for fp64 rows); one warp per unit, a grid
// sized to the job.  Lane l handles positions u * UT + l + 32 j: coalesced loads of the R rows and of leaf_dest, all
// in flight at once, then one 16-byte store per position holding the R rows' weights (exp / conversion fused).  The
// stores are scattered -- every lane its own line of z -- which an SM retires at ~1.6 cycles each when they hit L2
// (tools/scatter_store_probe.cu); that, not HBM, bounds this kernel.  Rows past the end of the batch alias the last
// valid row; nothing outside a row's V elements is read, whatever the alignment.
__device__ __forceinline__ int perm_unit_tokens_rt(int in_type) { return in_type == GT_F64 ? kUnitTokens / 2 : kUnitTokens; }

template <typename VT, typename IN_T, int R>
__device__ __forceinline__ void permute_unit(const PlanView& P, const MassArgs<VT>& A, unsigned w, int n_units, int lane) {
    using RV = RowVec<VT, R>;
    constexpr int UT = sizeof(IN_T) == 8 ? kUnitTokens / 2 : kUnitTokens, TPL = UT / 32;
    const bool log_input = A.log_input != 0;
    const int g = (int)(w / (unsigned)n_units), v0 = (int)(w - (unsigned)g * (unsigned)n_units) * UT;
    const int ntok = min(UT, (int)P.V - v0);
    const IN_T* __restrict__ ws = static_cast<const IN_T*>(A.ws);
    VT* zg = A.z + (size_t)g * P.ZG * R;
    int dst[TPL];
    IN_T x[TPL][R];
    const bool dfs = A.dfs_order != 0;
#pragma unroll
    for (int j = 0; j < TPL; ++j) {
        const int i = min(lane + 32 * j, ntok - 1);
        // position = DFS rank r: slot of leaf r % T in tile r / T (what leaf_dest holds for the item of rank r); consecutive
        // positions then go to consecutive slots and a warp's stores cover whole lines
        const int r = v0 + i;
        dst[j] = dfs ? ((r & ~(P.T - 1)) | swz<(int)sizeof(VT) * R>(r & (P.T - 1))) : __ldg(P.leaf_dest + r);
#pragma unroll
        for (int r = 0; r < R; ++r) x[j][r] = ws[(size_t)min(g * R + r, A.n_rows - 1) * A.ld_ws + v0 + i];
    }
#pragma unroll
    for (int j = 0; j < TPL; ++j) {
        if (lane + 32 * j < ntok) {
            RV y;
#pragma unroll
            for (int r = 0; r < R; ++r) y.v[r] = convert_in<VT, IN_T>(x[j][r], log_input);
            y.store(zg + (size_t)dst[j] * R);
        }
    }
}

// Input-type dispatch (one switch per warp; the loops inside are type-specific).
#define GT_IN_TYPE_SWITCH(in_type, CALL)                          \
    switch (in_type) {                                            \
        case GT_F32: { using IN_T = float; CALL; } break;         \
        case GT_F64: { using IN_T = double; CALL; } break;        \
        case GT_F16: { using IN_T = __half; CALL; } break;        \
        default: { using IN_T = __nv_bfloat16; CALL; } break;     \
    }

#ifndef GT_PERM_THREADS
#define GT_PERM_THREADS 256
#endif
#ifndef GT_PERM_MINB
#define GT_PERM_MINB 1
#endif
constexpr int kPermThreads = GT_PERM_THREADS;
template <typename VT, int R>
__global__ void __launch_bounds__(kPermThreads, GT_PERM_MINB) permute_kernel(PlanView P, MassArgs<VT> A, unsigned total_units) {
    pdl_wait();  // the rows may come from the previous kernel of the stream; z may still be read by the previous call
    pdl_trigger();
    const unsigned w = blockIdx.x * (kPermThreads / 32) + (threadIdx.x >> 5);
    if (w >= total_units) return;
    const int UT = perm_unit_tokens_rt(A.in_type);
    const int n_units = (int)((P.V + UT - 1) / UT);
    GT_IN_TYPE_SWITCH(A.in_type, (permute_unit<VT, IN_T, R>(P, A, w, n_units, (int)(threadIdx.x & 31))));
}

// ---- compute-group phases (OP is a compile-time parameter; the kernel branches once per phase) ----------------------
// Value slot s lives in the pair's leaf block when s < T and in the item's rest buffer otherwise; `restv` is biased
// by -T slots so that both cases are base + s * R.

// 1. pyramid of aligned blocks: level k block i at (swizzled) slot 2T - (T >> (k-1)) + i, levels 1 .. kPyramidTop.
//    Lane u owns 8 consecutive leaves: three levels in registers, five more by warp shuffles, so a warp covers 256
//    leaves and no level needs a cross-warp step (longer aligned blocks are multi-term ranges of the ELL phase).
//    The swizzle makes the 16-byte chunk loads and the strided level stores bank-conflict free.
template <typename VT, int R, int OP, int NWARPS>
__device__ __forceinline__ void phase_pyramid(const VT* leafv, VT* restv, int T, int warp, int lane) {
    using RV = RowVec<VT, R>;
    constexpr int B = (int)sizeof(VT) * R;
    constexpr int SPC = 16 / B;  // slots per 16-byte chunk
    static_assert(kPyramidTop == 8, "the pyramid is written for 8 leaves per lane");
    auto level_slot = [&](int k, int i) { return 2 * T - (T >> (k - 1)) + i; };  // level k >= 1, block i
    for (int ub = warp * 32; ub < (T >> 3); ub += NWARPS * 32) {
        const int u = ub + lane;
        RV x[8];
#pragma unroll
        for (int ch = 0; ch < 8 / SPC; ++ch) {
            const int c = (8 * u) / SPC + ch;
            const int cc = c ^ ((c >> 3) & (B / 2 - 1));
            const float4 raw = *reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(leafv) + (size_t)cc * 16);
            memcpy(&x[ch * SPC], &raw, 16);
        }
        RV a[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            a[e] = RV::template combine<OP>(x[2 * e], x[2 * e + 1]);
            a[e].store(restv + swz<B>(level_slot(1, 4 * u + e)) * R);
        }
        const RV c0 = RV::template combine<OP>(a[0], a[1]), c1 = RV::template combine<OP>(a[2], a[3]);
        c0.store(restv + swz<B>(level_slot(2, 2 * u)) * R);
        c1.store(restv + swz<B>(level_slot(2, 2 * u + 1)) * R);
        RV y = RV::template combine<OP>(c0, c1);
        y.store(restv + swz<B>(level_slot(3, u)) * R);
#pragma unroll
        for (int j = 1; j <= 5; ++j) {
            y = RV::template combine<OP>(y, y.shfl_down(1 << (j - 1)));
            if ((lane & ((1 << j) - 1)) == 0) y.store(restv + swz<B>(level_slot(3 + j, u >> j)) * R);
        }
    }
}

// 2. ranges that need more than one block: ELL chunks of 32 ranges, one warp per chunk, term rows read from shared
//    memory.  Chunks are sorted by descending term count: rounds alternate direction so that the warps that drew the
//    longest chunks of one round draw the shortest of the next.
template <typename VT, int R, int OP, int NWARPS>
__device__ __forceinline__ void phase_ell(const VT* leafv, VT* restv, const uint16_t* s_terms, const int2* dsc, int er0,
                                          int nchunks, int T, int warp, int lane) {
    using RV = RowVec<VT, R>;
    auto val = [&](int sl) { return RV::load((sl < T ? leafv : (const VT*)restv) + sl * R); };
    for (int base = 0, rr = 0; base < nchunks; base += NWARPS, ++rr) {
        const int c = base + ((rr & 1) ? NWARPS - 1 - warp : warp);
        if (c >= nchunks) continue;
        const int2 d = dsc[c];
        const uint16_t* tp = s_terms + (d.x - er0) * 32 + lane;
        RV acc = RV::template ident<OP>();
        int kb = 0;
        for (; kb + 4 <= d.y; kb += 4) {
            int sl[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) sl[e] = tp[(kb + e) * 32];
#pragma unroll
            for (int e = 0; e < 4; ++e) acc = RV::template combine<OP>(acc, val(sl[e]));
        }
        if (kb + 2 <= d.y) {  // no padding rows: a pair, then a single term
            const int s0 = tp[kb * 32], s1 = tp[(kb + 1) * 32];
            acc = RV::template combine<OP>(acc, RV::template combine<OP>(val(s0), val(s1)));
            kb += 2;
        }
        if (kb < d.y) acc = RV::template combine<OP>(acc, val((int)tp[kb * 32]));
        acc.store(restv + (2 * T + c * 32 + lane) * R);
    }
}

// header of a fetched pair (written by the producer thread before it arms pairFull)
struct PairHdr { int t, g, n0, n1, er0, ec0, nchunks, pc0, pc1; };

template <typename VT, int R>
__global__ void __launch_bounds__(kThreads, 2) mass_kernel(PlanView P, MassArgs<VT> A) {
    using RV = RowVec<VT, R>;
    constexpr int B = (int)sizeof(VT) * R;  // bytes per slot
    static_assert(B == 16, "value slots are 16 bytes");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const MassSmem L(P, B);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bars);
    uint64_t* pairFull = bars;       // [2]
    uint64_t* pairEmpty = bars + 2;  // [2]
    uint64_t* full = bars + 4;       // [2] rest buffer filled by the compute group
    uint64_t* empty = bars + 6;      // [2] rest buffer drained by the emit group

    const int T = P.T;
    const int n_rows = A.n_rows;
    const int RG = (n_rows + R - 1) / R;
    const int nops = (A.ops == (unsigned)(GT_OP_SUM | GT_OP_MAX)) ? 2 : 1;
    const int first_op = (A.ops & GT_OP_SUM) ? OP_SUM : OP_MAX;
    const int G = (int)gridDim.x;
    const int n_pairs = P.NT * RG;
    const int my_pairs = (int)blockIdx.x < n_pairs ? (n_pairs - (int)blockIdx.x + G - 1) / G : 0;
    const int dbg = P.debug_stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(pairFull, 1); mbar_init(pairFull + 1, 1);
        mbar_init(pairEmpty, kEmitThreads); mbar_init(pairEmpty + 1, kEmitThreads);
        mbar_init(full, kComputeThreads); mbar_init(full + 1, kComputeThreads);
        mbar_init(empty, kEmitThreads); mbar_init(empty + 1, kEmitThreads);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    GT_PTRACE(threadIdx.x == 0, kTraceItems - 1, 11);  // CTA start

    if (warp == kProducerWarp) {
        // =========================== producer: one thread fetches pair after pair ====================================
        // Plan metadata may be fetched while the permute kernel is still running (programmatic dependent launch); the
        // leaf blocks may not: griddepcontrol.wait precedes the first one.
        if (lane == 0) {
            for (int q = 0; q < my_pairs; ++q) {
                const int p = (int)blockIdx.x + q * G;
                const int g = p / P.NT, t = p - g * P.NT;
                const int buf = q & 1;
                unsigned char* meta = smem_raw + L.meta + (size_t)buf * L.meta_bytes;
                // tile boundaries (independent loads, one round trip, requested before the wait for the buffers)
                const int n0 = __ldg(P.tile_node_lo + t), n1 = __ldg(P.tile_node_lo + t + 1);
                const int er0 = __ldg(P.ell_row_ptr + t), er1 = __ldg(P.ell_row_ptr + t + 1);
                const int ec0 = __ldg(P.ell_chunk_ptr + t), ec1 = __ldg(P.ell_chunk_ptr + t + 1);
                const int pc0 = __ldg(P.piece_ptr + t), pc1 = __ldg(P.piece_ptr + t + 1);
                mbar_wait(pairEmpty + buf, ((unsigned)(q >> 1) & 1u) ^ 1u);  // the pair two back has been drained
                PairHdr* h = reinterpret_cast<PairHdr*>(meta + L.hdr);
                h->t = t; h->g = g; h->n0 = n0; h->n1 = n1; h->er0 = er0; h->ec0 = ec0; h->nchunks = ec1 - ec0;
                h->pc0 = pc0; h->pc1 = pc1;
                const int ea = ec0 & ~1, na = n0 & ~7;
                const unsigned lb = (unsigned)T * (unsigned)B;
                const unsigned tb = (unsigned)(er1 - er0) * 64u, db = (unsigned)((ec1 - ea + 1) >> 1) * 16u;
                const unsigned sb = (unsigned)((n1 - na + 7) >> 3) * 16u;
                mbar_expect_tx(pairFull + buf, lb + tb + db + sb);
                if (tb) bulk_g2s(meta + L.terms, P.ell_terms + (size_t)er0 * 32, tb, pairFull + buf);
                if (db) bulk_g2s(meta + L.desc, P.ell_desc + ea, db, pairFull + buf);
                bulk_g2s(meta + L.slots, P.node_slot + na, sb, pairFull + buf);
                if (q == 0) { pdl_wait(); pdl_trigger(); }  // z comes from permute_kernel
                bulk_g2s(smem_raw + L.leaf + (size_t)buf * L.leaf_bytes, A.z + ((size_t)g * P.ZG + (size_t)t * T) * R, lb, pairFull + buf);
                if (q == 0) GT_PTRACE(true, kTraceItems - 1, 0);
            }
        }
    } else {
        if (warp < kComputeWarps) {
            // =========================== compute group ===========================================================
            const int gtid = threadIdx.x;
            for (int q = 0; q < my_pairs; ++q) {
                const int buf = q & 1;
                const unsigned char* meta = smem_raw + L.meta + (size_t)buf * L.meta_bytes;
                mbar_wait(pairFull + buf, (unsigned)(q >> 1) & 1u);
                const PairHdr* h = reinterpret_cast<const PairHdr*>(meta + L.hdr);
                const int er0 = h->er0, ec0 = h->ec0, nchunks = h->nchunks;
                const uint16_t* s_terms = reinterpret_cast<const uint16_t*>(meta + L.terms);
                const int2* dsc = reinterpret_cast<const int2*>(meta + L.desc) + (ec0 & 1);
                const VT* leafv = reinterpret_cast<const VT*>(smem_raw + L.leaf + (size_t)buf * L.leaf_bytes);
                for (int j = 0; j < nops; ++j) {
                    const int k = q * nops + j;
                    VT* restv = reinterpret_cast<VT*>(smem_raw + L.rest + (size_t)(k & 1) * L.rest_bytes) - (size_t)T * R;
                    const bool is_sum = (j == 0 ? first_op : OP_MAX) == OP_SUM;
                    GT_TRACE(0);
                    mbar_wait(empty + (k & 1), ((unsigned)(k >> 1) & 1u) ^ 1u);  // the emit group has drained this rest buffer
                    GT_TRACE(1);
                    if (dbg != 9) {
                        if (is_sum) phase_pyramid<VT, R, OP_SUM, kComputeWarps>(leafv, restv, T, warp, lane);
                        else phase_pyramid<VT, R, OP_MAX, kComputeWarps>(leafv, restv, T, warp, lane);
                    }
                    if (gtid == kComputeThreads - 1) {  // identity slot (spanning nodes' placeholder)
                        if (is_sum) RV::template ident<OP_SUM>().store(restv + swz<B>(2 * T - 1) * R);
                        else RV::template ident<OP_MAX>().store(restv + swz<B>(2 * T - 1) * R);
                    }
                    GT_TRACE(2);
                    group_sync<1, kComputeThreads>();
                    GT_TRACE(3);
                    if (dbg != 9) {
                        if (is_sum) phase_ell<VT, R, OP_SUM, kComputeWarps>(leafv, restv, s_terms, dsc, er0, nchunks, T, warp, lane);
                        else phase_ell<VT, R, OP_MAX, kComputeWarps>(leafv, restv, s_terms, dsc, er0, nchunks, T, warp, lane);
                    }
                    GT_TRACE(4);
                    mbar_arrive(full + (k & 1));  // this thread's share of the value array is complete
                }
            }
        } else if (warp < kProducerWarp) {
            // =========================== emit group ==============================================================
            const int gtid = threadIdx.x - kComputeThreads;
            pdl_wait();  // the outputs and piece buffers may still be in use by the previous kernels of the stream
            for (int q = 0; q < my_pairs; ++q) {
                const int buf = q & 1;
                const unsigned char* meta = smem_raw + L.meta + (size_t)buf * L.meta_bytes;
                mbar_wait(pairFull + buf, (unsigned)(q >> 1) & 1u);
                const PairHdr* h = reinterpret_cast<const PairHdr*>(meta + L.hdr);
                const int g = h->g, n0 = h->n0, n1 = h->n1, pc0 = h->pc0, pc1 = h->pc1;
                const uint16_t* s_slots = reinterpret_cast<const uint16_t*>(meta + L.slots);
                const VT* leafv = reinterpret_cast<const VT*>(smem_raw + L.leaf + (size_t)buf * L.leaf_bytes);
                // this thread's spanning-node piece of the tile, requested long before its first use
                int my_pslot = 0, my_pidx = 0;
                if (pc0 + gtid < pc1) { my_pslot = __ldg(P.piece_slot + pc0 + gtid); my_pidx = __ldg(P.piece_idx + pc0 + gtid); }
                const int b0 = g * R;
                for (int j = 0; j < nops; ++j) {
                    const int k = q * nops + j;
                    const VT* restv = reinterpret_cast<const VT*>(smem_raw + L.rest + (size_t)(k & 1) * L.rest_bytes) - (size_t)T * R;
                    const bool is_sum = (j == 0 ? first_op : OP_MAX) == OP_SUM;
                    VT* out = is_sum ? A.out_sum : A.out_max;
                    VT* part = is_sum ? A.part_sum : A.part_max;
                    VT* orow[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        orow[r] = out + (size_t)min(b0 + r, n_rows - 1) * A.ld_out;
                        asm volatile("" : "+l"(orow[r]));  // keep the row pointers in registers (no rematerialisation per store)
                    }
                    GT_TRACE(8);
                    mbar_wait(full + (k & 1), (unsigned)(k >> 1) & 1u);  // the compute group has filled this rest buffer
                    GT_TRACE(9);
                    // emit the tile's node-id interval: lane = consecutive node id, so the slot reads of a warp cluster on
                    // a few neighbouring slots (unary chains broadcast) and every store instruction writes 128 contiguous
                    // bytes per row.  The sweep starts at the 128-byte line of row 0 that holds node n0, so with a row
                    // stride that is a multiple of 32 elements every store instruction covers exactly one line.
                    // Spanning nodes inside the interval carry the identity slot: what is written for them here is
                    // overwritten by span_kernel.
                    if (dbg != 3) {
                        constexpr int U = 4;
                        const int na = n0 & ~7;
                        const int lead = (int)(((reinterpret_cast<uintptr_t>(orow[0]) / sizeof(VT)) + (unsigned)n0) & 31u);
                        const unsigned count = (unsigned)(n1 - n0);
                        const uint16_t* sl_base = s_slots - na;
                        // a warp's U node groups are consecutive and its stores go row by row: U * 128 contiguous bytes
                        // per row and warp trip
                        const int nb0 = n0 - lead + (gtid >> 5) * (U * 32) + (gtid & 31);
                        for (int nb = nb0; nb < n1; nb += U * kEmitThreads) {
                            VT* p[R];
#pragma unroll
                            for (int r = 0; r < R; ++r) p[r] = orow[r] + nb;
                            RV x[U];
                            bool ok[U];
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                const int n = nb + u * 32;
                                ok[u] = (unsigned)(n - n0) < count;
                                if (ok[u]) {
                                    const int sl = sl_base[n];
                                    x[u] = RV::load((sl < T ? leafv : restv) + sl * R);
                                }
                            }
#pragma unroll
                            for (int r = 0; r < R; ++r)
#pragma unroll
                                for (int u = 0; u < U; ++u)
                                    if (ok[u]) __stcs(p[r] + u * 32, x[u].v[r]);
                        }
                        // pieces of spanning nodes that overlap this tile (reduced by span_kernel, which runs next on the stream)
                        for (int i = pc0 + gtid; i < pc1; i += kEmitThreads) {
                            const bool mine = i == pc0 + gtid;
                            const int sl = mine ? my_pslot : (int)__ldg(P.piece_slot + i);
                            const RV x = RV::load((sl < T ? leafv : restv) + sl * R);
                            const int idx = mine ? my_pidx : __ldg(P.piece_idx + i);
#pragma unroll
                            for (int r = 0; r < R; ++r) part[(size_t)min(b0 + r, n_rows - 1) * P.n_pieces + idx] = x.v[r];
                        }
                    }
                    GT_TRACE(10);
                    mbar_arrive(empty + (k & 1));  // this thread no longer reads the rest buffer
                }
                mbar_arrive(pairEmpty + buf);  // ... nor the pair's leaf block and metadata
            }
        }
    }

}

// ---- phase 3: nodes whose leaf range crosses tiles, reduced from their per-tile pieces (fp64 for sums) ------------
// One thread per (spanning node, row); consecutive lanes take consecutive spanning nodes of one row, whose pieces
// are adjacent in `part`.  Most spanning nodes have two or three pieces, which a thread loads all at once; the few
// with many (the root has one per tile) are reduced by the whole warp, 32 pieces per step, so that no thread walks a
// long chain of dependent loads.  Runs after mass_kernel in stream order and overwrites the placeholder it emitted.
// blockIdx.z selects the reduction when both were requested.
constexpr int kSpanInline = 8;  // pieces a thread reduces by itself

template <typename VT, bool SUM> __device__ __forceinline__ VT span_reduce(const VT* __restrict__ pr, int q0, int q1, int lane) {
    using AT = std::conditional_t<SUM, double, VT>;  // sums across tiles accumulate in fp64; max is exact
    const AT ident = SUM ? AT(0) : -std::numeric_limits<AT>::infinity();
    const int cnt = q1 - q0;
    AT acc = ident;
    if (cnt <= kSpanInline) {
        VT v[kSpanInline];
#pragma unroll
        for (int i = 0; i < kSpanInline; ++i) v[i] = i < cnt ? pr[q0 + i] : (VT)ident;  // independent loads
#pragma unroll
        for (int i = 0; i < kSpanInline; ++i) acc = SUM ? acc + (AT)v[i] : (AT)fmax((VT)acc, v[i]);
    }
    // nodes with many pieces, one at a time, by the whole warp (fixed order: results do not depend on the launch)
    unsigned heavy = __ballot_sync(0xffffffffu, cnt > kSpanInline);
    while (heavy) {
        const int src = __ffs(heavy) - 1;
        heavy &= heavy - 1;
        const int h0 = __shfl_sync(0xffffffffu, q0, src), h1 = __shfl_sync(0xffffffffu, q1, src);
        AT a = ident;
        for (int q = h0 + lane; q < h1; q += 32) a = SUM ? a + (AT)pr[q] : (AT)fmax((VT)a, pr[q]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const AT y = __shfl_xor_sync(0xffffffffu, a, o);
            a = SUM ? a + y : (AT)fmax((VT)a, (VT)y);
        }
        if (lane == src) acc = a;
    }
    return cnt > 0 ? (VT)acc : VT(0);  // an empty range (root of an empty vocabulary) has no mass
}

template <typename VT>
__global__ void __launch_bounds__(256) span_kernel(PlanView P, MassArgs<VT> A) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool live = k < P.n_span;
    // plan metadata first: these loads may run ahead of the tile kernel's completion
    const int q0 = live ? __ldg(P.span_pp + k) : 0, q1 = live ? __ldg(P.span_pp + k + 1) : 0;
    const int node = live ? __ldg(P.span_node + k) : 0;
    pdl_wait();  // pieces and placeholders come from mass_kernel
    pdl_trigger();
    const bool is_sum = (A.ops & GT_OP_SUM) && blockIdx.z == 0;
    const VT* part = is_sum ? A.part_sum : A.part_max;
    VT* out = is_sum ? A.out_sum : A.out_max;
    for (int b = blockIdx.y; b < A.n_rows; b += gridDim.y) {  // whole warps stay together: dead lanes have no pieces
        const VT* pr = part + (size_t)b * P.n_pieces;
        const VT res = is_sum ? span_reduce<VT, true>(pr, q0, q1, lane) : span_reduce<VT, false>(pr, q0, q1, lane);
        if (live) out[(size_t)b * A.ld_out + node] = res;
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
    const size_t o_leaf_dest = ADDV(P.leaf_dest);
    const size_t o_ell_chunk_ptr = ADDV(P.ell_chunk_ptr), o_ell_desc = ADDV(P.ell_desc), o_ell_terms = ADDV(P.ell_terms);
    const size_t o_ell_row_ptr = ADDV(P.ell_row_ptr);
    const size_t o_tile_node_lo = ADDV(P.tile_node_lo), o_node_slot = ADDV(P.node_slot);
    const size_t o_piece_ptr = ADDV(P.piece_ptr), o_piece_slot = ADDV(P.piece_slot), o_piece_idx = ADDV(P.piece_idx);
    const size_t o_span_node = ADDV(P.span_node), o_span_pp = ADDV(P.span_pp);
    std::vector<int32_t> leaf_rank((size_t)L.V);
    for (int64_t r = 0; r < L.V; ++r) leaf_rank[(size_t)L.perm[(size_t)r]] = (int32_t)r;
    const size_t o_leaf_rank = ADDV(leaf_rank), o_node_lo = ADDV(L.lo), o_node_hi = ADDV(L.hi);
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
    // The source is pageable: the copy may return once the data is staged, before the DMA has finished, and only the
    // legacy stream is ordered after it.  The kernels run on the caller's (possibly non-blocking) streams, so wait for
    // the device here -- once per trie and device.
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaSetDevice(cur);
    if (e != cudaSuccess) {
        set_error("metadata upload failed: %s", cudaGetErrorString(e));
        free_device_plan(d);
        return nullptr;
    }

    unsigned char* base = static_cast<unsigned char*>(d->blob);
    PlanView& v = d->view;
    v.T = P.T; v.logT = 0; while ((1 << (v.logT + 1)) <= P.T) ++v.logT;
    v.NT = P.NT;
    v.SV = (P.max_tile_values + 3) & ~3;
    v.V = L.V; v.N = L.N; v.ZG = (int64_t)P.NT * P.T;
    v.leaf_dest = (const int32_t*)(base + o_leaf_dest);
    v.ell_chunk_ptr = (const int32_t*)(base + o_ell_chunk_ptr); v.ell_desc = (const int2*)(base + o_ell_desc);
    v.ell_terms = (const uint16_t*)(base + o_ell_terms); v.ell_row_ptr = (const int32_t*)(base + o_ell_row_ptr);
    v.R = P.R; v.max_tile_nodes = P.max_tile_nodes; v.max_tile_ell_rows = P.max_tile_ell_rows;
    v.max_tile_chunks = P.max_tile_chunks;
    v.tile_node_lo = (const int32_t*)(base + o_tile_node_lo); v.node_slot = (const uint16_t*)(base + o_node_slot);
    v.piece_ptr = (const int32_t*)(base + o_piece_ptr); v.piece_slot = (const uint16_t*)(base + o_piece_slot);
    v.piece_idx = (const int32_t*)(base + o_piece_idx);
    v.n_span = (int32_t)P.span_node.size(); v.n_pieces = P.n_pieces;
    v.span_node = (const int32_t*)(base + o_span_node); v.span_pp = (const int32_t*)(base + o_span_pp);
    v.leaf_rank = (const int32_t*)(base + o_leaf_rank);
    v.node_lo = (const int32_t*)(base + o_node_lo); v.node_hi = (const int32_t*)(base + o_node_hi);
    { const char* e = getenv("GT_DEBUG_STOP"); v.debug_stop = e && *e ? atoi(e) : 0; }
    v.trace = nullptr;
    if (const char* e = getenv("GT_TRACE")) {
        if (*e && atoi(e) != 0) {
            const size_t bytes = (size_t)kTraceCtas * kTraceItems * kTraceEvents * sizeof(long long);
            cudaSetDevice(device);
            if (cudaMalloc(&d->trace, bytes) == cudaSuccess) { cudaMemset(d->trace, 0, bytes); v.trace = d->trace; }
            else { d->trace = nullptr; (void)cudaGetLastError(); }
            cudaSetDevice(cur);
        }
    }
    return d;
}

template <typename VT, int R> static size_t mass_smem(const PlanView& v) { return MassSmem(v, (int)sizeof(VT) * R).total; }

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

// Scratch layout: a batch is reduced in *chunks* of at most kMaxChunkRows rows (what the staging buffer holds), and the
// pieces of spanning nodes are collected over a *span group* of up to kMaxSpanRows rows, so that span_kernel runs once
// per span group instead of once per chunk (it costs ~6 us in the chain for 2.5 us of work: it cannot become resident
// before the tile kernel's CTAs leave).
//   z [chunk row groups][ZG slots][R] VT | part_sum [span rows][n_pieces] VT | part_max likewise
constexpr int64_t kMaxChunkRows = 64;    // 34 MB of staging at 128k tokens: stays in the 126 MB L2 between permute and tile kernel
constexpr int64_t kMaxSpanRows = 1024;   // 9.9 MB of pieces per reduction at 128k tokens
template <typename VT, int R> struct Scratch {
    VT* z; VT* part_sum; VT* part_max;
    int64_t chunk_rows, span_rows;  // capacities
    static size_t pad(size_t b) { return (b + 255) & ~size_t(255); }
    static size_t z_bytes(const PlanView& v, int64_t rows) { return pad((size_t)((rows + R - 1) / R) * (size_t)v.ZG * R * sizeof(VT)); }
    static size_t part_bytes(const PlanView& v, int64_t rows) { return pad((size_t)rows * v.n_pieces * sizeof(VT)); }
    static size_t total(const PlanView& v, int64_t chunk, int64_t span) { return z_bytes(v, chunk) + 2 * part_bytes(v, span); }
    static int64_t max_chunk() {
        static const int64_t n = []() { const char* e = getenv("GT_CHUNK_ROWS"); const long x = e ? atol(e) : 0; return x > 0 ? (int64_t)x : kMaxChunkRows; }();
        return n;
    }
    // what a workspace of `bytes` holds (chunk_rows = 0: not even one row)
    Scratch(const PlanView& v, void* base, size_t bytes) {
        int64_t lo = 0, hi = max_chunk();
        while (lo < hi) {  // total() is monotone in the row count
            const int64_t mid = (lo + hi + 1) / 2;
            if (total(v, mid, mid) <= bytes) lo = mid; else hi = mid - 1;
        }
        chunk_rows = lo;
        lo = chunk_rows; hi = std::max<int64_t>(chunk_rows, kMaxSpanRows);
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) / 2;
            if (total(v, chunk_rows, mid) <= bytes) lo = mid; else hi = mid - 1;
        }
        span_rows = lo;
        char* p = static_cast<char*>(base);
        z = reinterpret_cast<VT*>(p);
        p += z_bytes(v, chunk_rows);
        part_sum = reinterpret_cast<VT*>(p);
        p += part_bytes(v, span_rows);
        part_max = reinterpret_cast<VT*>(p);
    }
};

// Resident CTAs per SM of a kernel at a given dynamic shared-memory size (cached: the query is not free).
static int resident_ctas(const void* kernel, int threads, size_t smem) {
    static std::mutex mu;
    static std::map<std::pair<const void*, size_t>, int> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(kernel, smem);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    int n = 0;
    // the query honours the kernel's dynamic shared-memory limit: raise it first, or a size above it reports 0
    if (allow_smem_impl(kernel, smem) != cudaSuccess) (void)cudaGetLastError();
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) {
        (void)cudaGetLastError();
        n = 1;
    }
    cache[key] = n;
    return n;
}
static int sm_count() {
    static std::mutex mu;
    static std::map<int, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(dev);
    if (it != cache.end()) return it->second;
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n;
    return n;
}

// One chunk: permute -> tile kernel.  part_sum / part_max: where this chunk's rows start in the span group's piece arrays.
template <typename VT, int R>
static int launch_mass(const PlanView& v, const void* ws, int in_type, int64_t ld_ws, unsigned flags, VT* z, VT* part_sum, VT* part_max,
                       VT* out_sum, VT* out_max, int64_t ld_out, int rows, unsigned ops, unsigned phases, cudaStream_t st) {
    MassArgs<VT> A;
    A.ws = ws; A.in_type = in_type; A.ld_ws = ld_ws;
    A.log_input = (flags & GT_FLAG_LOG_INPUT) ? 1 : 0; A.dfs_order = (flags & GT_FLAG_DFS_ORDER) ? 1 : 0;
    A.z = z;
    A.out_sum = out_sum; A.out_max = out_max; A.part_sum = part_sum; A.part_max = part_max;
    A.ld_out = ld_out; A.n_rows = rows; A.ops = ops;
    const int RG = (rows + R - 1) / R;
    if (phases & GT_FLAG_PHASE_PERMUTE) {  // one warp per unit of 128 (fp64 rows: 64) positions of a row group
        const int64_t UT = in_type == GT_F64 ? kUnitTokens / 2 : kUnitTokens;
        const int64_t units = (int64_t)RG * ((v.V + UT - 1) / UT);
        const int64_t pgrid = (units + kPermThreads / 32 - 1) / (kPermThreads / 32);
        NvtxRange nv("gt:permute");
        GT_CUDA(launch_pdl(permute_kernel<VT, R>, dim3((unsigned)pgrid), dim3(kPermThreads), 0, st, v, A, (unsigned)units));
    }
    if (phases & GT_FLAG_PHASE_TILE) {
        const size_t smem = mass_smem<VT, R>(v);
        if (smem > 227 * 1024) {
            set_error("tile plan needs %zu bytes of shared memory per CTA (limit 232448): use a smaller tile", smem);
            return GT_ERR_LIMIT;
        }
        GT_CUDA(allow_smem(mass_kernel<VT, R>, smem));
        // persistent grid: one CTA per resident slot; CTA b takes the (row group, tile) pairs b, b + grid, ...
        const int64_t pairs = (int64_t)v.NT * RG;
        const int slots = sm_count() * resident_ctas(reinterpret_cast<const void*>(mass_kernel<VT, R>), kThreads, smem);
        const unsigned grid = (unsigned)std::min<int64_t>(pairs, slots);
        NvtxRange nv("gt:tile");
        GT_CUDA(launch_pdl(mass_kernel<VT, R>, dim3(grid), dim3(kThreads), smem, st, v, A));
    }
    return GT_OK;
}

// The spanning nodes of the `rows` rows whose pieces start at part_sum / part_max and whose outputs start at out_sum / out_max.
template <typename VT>
static int launch_span(const PlanView& v, VT* part_sum, VT* part_max, VT* out_sum, VT* out_max, int64_t ld_out, int rows, unsigned ops,
                       cudaStream_t st) {
    if (v.n_span <= 0 || rows <= 0) return GT_OK;
    MassArgs<VT> A{};
    A.out_sum = out_sum; A.out_max = out_max; A.part_sum = part_sum; A.part_max = part_max;
    A.ld_out = ld_out; A.n_rows = rows; A.ops = ops;
    const int nops = (ops == (unsigned)(GT_OP_SUM | GT_OP_MAX)) ? 2 : 1;
    dim3 sgrid((unsigned)((v.n_span + 255) / 256), (unsigned)std::min(rows, 4096), (unsigned)nops);
    NvtxRange nv("gt:span");
    GT_CUDA(launch_pdl(span_kernel<VT>, sgrid, dim3(256), 0, st, v, A));
    return GT_OK;
}

template <typename VT, int R>
static int reduce_typed(const PlanView& v, const void* ws, int in_type, int64_t n_rows, int64_t ld_ws, void* out_sum,
                        void* out_max, int64_t ld_out, unsigned ops, unsigned flags, void* workspace,
                        size_t workspace_bytes, cudaStream_t st) {
    if (in_type != GT_F32 && in_type != GT_F64 && in_type != GT_F16 && in_type != GT_BF16) {
        set_error("unknown input type %d", in_type);
        return GT_ERR_ARG;
    }
    VT* osum = (ops & GT_OP_SUM) ? static_cast<VT*>(out_sum) : nullptr;
    VT* omax = (ops & GT_OP_MAX) ? static_cast<VT*>(out_max) : nullptr;
    if (v.NT == 0) {  // empty vocabulary: the root is the only node and has no mass
        for (VT* out : {osum, omax})
            if (out) GT_CUDA(cudaMemset2DAsync(out, (size_t)ld_out * sizeof(VT), 0, (size_t)v.N * sizeof(VT), (size_t)n_rows, st));
        return GT_OK;
    }
    // rows per launch: what the caller's scratch can stage (a partial row group is legal: the kernels alias the missing
    // rows to the last valid one)
    const Scratch<VT, R> sc(v, workspace, workspace_bytes);
    if (sc.chunk_rows < 1) {
        set_error("workspace too small: %zu bytes given, one row needs %zu", workspace_bytes, Scratch<VT, R>::total(v, 1, 1));
        return GT_ERR_STATE;
    }
    const unsigned phases = (flags & GT_FLAG_PHASE_MASK) ? (flags & GT_FLAG_PHASE_MASK) : GT_FLAG_PHASE_MASK;
    const size_t in_size = in_type == GT_F64 ? 8 : in_type == GT_F32 ? 4 : 2;
    const int64_t chunk_rows = sc.chunk_rows;
    for (int64_t s0 = 0; s0 < n_rows; s0 += sc.span_rows) {  // span groups
        const int64_t s1 = std::min<int64_t>(n_rows, s0 + sc.span_rows);
        for (int64_t r0 = s0; r0 < s1; r0 += chunk_rows) {  // chunks
            const int rows = (int)std::min<int64_t>(chunk_rows, s1 - r0);
            const void* wsr = static_cast<const char*>(ws) + (size_t)r0 * ld_ws * in_size;
            const size_t poff = (size_t)(r0 - s0) * v.n_pieces;
            const int rc = launch_mass<VT, R>(v, wsr, in_type, ld_ws, flags, sc.z, sc.part_sum + poff, sc.part_max + poff,
                                              osum ? osum + (size_t)r0 * ld_out : nullptr, omax ? omax + (size_t)r0 * ld_out : nullptr,
                                              ld_out, rows, ops, phases, st);
            if (rc != GT_OK) return rc;
        }
        if (phases & GT_FLAG_PHASE_SPAN) {
            const int rc = launch_span<VT>(v, sc.part_sum, sc.part_max, osum ? osum + (size_t)s0 * ld_out : nullptr,
                                           omax ? omax + (size_t)s0 * ld_out : nullptr, ld_out, (int)(s1 - s0), ops, st);
            if (rc != GT_OK) return rc;
        }
    }
    return GT_OK;
}

// ---- read-outs that keep the [B, N] slab on the GPU ---------------------------------------
//
// gather: the caller of the mass path reads a handful of nodes per row (the children of the node a particle stands
// on), so only out[b, k] = mass[b, ids[b, k]] has to cross PCIe.  One thread per (row, k); rows on blockIdx.y.
template <typename VT>
__global__ void __launch_bounds__(256) gather_nodes_kernel(const VT* __restrict__ mass, int64_t ld_mass, int n_rows, int64_t N,
                                                           const int32_t* __restrict__ ids, int n_ids, int64_t ids_ld,
                                                           const int32_t* __restrict__ norm, int log_out,
                                                           VT* __restrict__ out, int64_t ld_out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_ids) return;
    for (int b = blockIdx.y; b < n_rows; b += gridDim.y) {
        const VT* row = mass + (size_t)b * ld_mass;
        const int id = __ldg(ids + (size_t)b * ids_ld + k);
        VT v = (id >= 0 && id < N) ? row[id] : VT(0);
        if (norm) {
            const int z = __ldg(norm + b);
            const VT d = (z >= 0 && z < N) ? row[z] : VT(0);
            v = log_out ? (VT)(log((double)v) - log((double)d)) : v / d;
        } else if (log_out) {
            v = (VT)log((double)v);
        }
        out[(size_t)b * ld_out + k] = v;
    }
}

// token mask: bit i of row b is set iff item i's leaf lies in the subtree of nodes[b], i.e. iff the reference's
// reachability matrix has M[i, nodes[b]] = 1 (parallel.py:33-64).  With leaves in DFS order that is one range test
// on the leaf's DFS rank, so a warp produces one mask word per ballot; kMaskRows rows share every rank load.
constexpr int kMaskThreads = 512, kMaskRows = 16;
__global__ void __launch_bounds__(kMaskThreads) subtree_mask_kernel(PlanView P, const int32_t* __restrict__ nodes, int n_rows,
                                                                    uint32_t* __restrict__ bits, int64_t ld_bits) {
    __shared__ int s_lo[kMaskRows];
    __shared__ unsigned s_len[kMaskRows];
    const int b0 = blockIdx.y * kMaskRows;
    if (threadIdx.x < kMaskRows) {
        const int b = b0 + threadIdx.x;
        int lo = 0; unsigned len = 0;
        if (b < n_rows) {
            const int n = __ldg(nodes + b);
            if (n >= 0 && n < P.N) { lo = __ldg(P.node_lo + n); len = (unsigned)(__ldg(P.node_hi + n) - lo); }
        }
        s_lo[threadIdx.x] = lo; s_len[threadIdx.x] = len;
    }
    const int64_t i = (int64_t)blockIdx.x * kMaskThreads + threadIdx.x;
    const int r = i < P.V ? __ldg(P.leaf_rank + i) : -1;  // -1: fails every range test
    __syncthreads();
    const int64_t w = i >> 5;
    const int rows = min(kMaskRows, n_rows - b0);
    for (int q = 0; q < rows; ++q) {
        const unsigned word = __ballot_sync(0xffffffffu, r >= 0 && (unsigned)(r - s_lo[q]) < s_len[q]);
        if ((threadIdx.x & 31) == 0 && w < ld_bits) bits[(size_t)(b0 + q) * ld_bits + w] = word;
    }
}

}  // namespace gt

extern "C" {

int gt_upload(gt_trie* t, int device) {
    if (!t) { gt::set_error("gt_upload: null trie"); return GT_ERR_ARG; }
    {
        std::lock_guard<std::mutex> lock(t->mu);
        if (t->dev.count(device)) return GT_OK;
    }
    if (!t->plan) {
        const int rc = gt_plan(t, 0);  // takes the lock itself
        if (rc != GT_OK) return rc;
    }
    std::lock_guard<std::mutex> lock(t->mu);
    if (t->dev.count(device)) return GT_OK;  // another thread brought the device up meanwhile
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
    info->tile_leaves = P.T; info->n_tiles = P.NT;
    info->rows_per_item = P.R; info->permute_unit = gt::kUnitTokens;
    info->n_span = (int32_t)P.span_node.size(); info->span_terms = (int64_t)P.n_pieces;
    info->max_levels = P.max_levels; info->max_tile_values = P.max_tile_values;
    info->staged_slots = (int64_t)P.NT * P.T;
    size_t meta = 0;
    for (auto& kv : t->dev) { meta = kv.second->blob_bytes; break; }
    info->meta_bytes = (int64_t)meta;
    return GT_OK;
}

int64_t gt_debug_read_trace(const gt_trie* t, int device, long long* dst, int64_t capacity, int32_t dims[3]) {
    if (!t) { gt::set_error("gt_debug_read_trace: null trie"); return -1; }
    auto it = t->dev.find(device);
    if (it == t->dev.end() || !it->second->trace) { gt::set_error("no trace buffer on device %d (set GT_TRACE=1 before gt_upload)", device); return -1; }
    const int64_t n = (int64_t)gt::kTraceCtas * gt::kTraceItems * gt::kTraceEvents;
    if (dims) { dims[0] = gt::kTraceCtas; dims[1] = gt::kTraceItems; dims[2] = gt::kTraceEvents; }
    if (dst && capacity > 0) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(device);
        cudaDeviceSynchronize();
        const cudaError_t e = cudaMemcpy(dst, it->second->trace, (size_t)std::min(n, capacity) * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaMemset(it->second->trace, 0, (size_t)n * sizeof(long long));
        cudaSetDevice(cur);
        if (e != cudaSuccess) { gt::set_error("trace copy failed: %s", cudaGetErrorString(e)); return -1; }
    }
    return n;
}

size_t gt_workspace_bytes(const gt_trie* t, int64_t max_rows) {
    if (!t || !t->plan || max_rows <= 0) return 0;
    gt::PlanView v{};
    v.ZG = (int64_t)t->plan->NT * t->plan->T; v.n_pieces = t->plan->n_pieces;
    // staging for one chunk + the pieces of one span group, sized for the fp64 pipeline (the fp32 one needs half)
    return gt::Scratch<double, 2>::total(v, std::min<int64_t>(max_rows, gt::Scratch<double, 2>::max_chunk()),
                                         std::min<int64_t>(max_rows, gt::kMaxSpanRows)) + 256;
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
    if (ld_out > ((int64_t)1 << 28)) {  // the kernels index a row group with 32-bit offsets
        gt::set_error("gt_weight_reduce: output row stride %lld exceeds 2^28 elements", (long long)ld_out);
        return GT_ERR_LIMIT;
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
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255)) { gt::set_error("gt_weight_reduce: workspace must be a 256-byte aligned device pointer"); return GT_ERR_ARG; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    gt::NvtxRange nv("gt:weight_reduce");
#define GT_REDUCE(VT, R) gt::reduce_typed<VT, R>(v, ws, in_type, n_rows, ld_ws, out_sum, out_max, ld_out, ops, flags, \
                                                workspace, workspace_bytes, st)
    // 16-byte value slots: four fp32 rows or two fp64 rows per work item
    if (out_type == GT_F32) return GT_REDUCE(float, 4);
    if (out_type == GT_F64) return GT_REDUCE(double, 2);
#undef GT_REDUCE
    gt::set_error("gt_weight_reduce: output type must be GT_F32 or GT_F64");
    return GT_ERR_ARG;
}

int gt_download_rows(void* dst_host, size_t dst_pitch, const void* src_dev, size_t src_pitch, size_t width_bytes, int64_t n_rows,
                     gt_stream stream) {
    if (n_rows < 0 || dst_pitch < width_bytes || src_pitch < width_bytes) { gt::set_error("gt_download_rows: bad size / pitch"); return GT_ERR_ARG; }
    if (n_rows == 0 || width_bytes == 0) return GT_OK;
    if (!dst_host || !src_dev) { gt::set_error("gt_download_rows: null pointer"); return GT_ERR_ARG; }
    GT_CUDA(cudaMemcpy2DAsync(dst_host, dst_pitch, src_dev, src_pitch, width_bytes, (size_t)n_rows, cudaMemcpyDeviceToHost,
                              static_cast<cudaStream_t>(stream)));
    return GT_OK;
}

int gt_gather_nodes(const void* mass, int type, int64_t n_rows, int64_t n_nodes, int64_t ld_mass, const int32_t* node_ids,
                    int64_t n_ids, int64_t ids_ld, const int32_t* norm_node, unsigned flags, void* out, int64_t ld_out,
                    gt_stream stream) {
    if (n_rows < 0 || n_ids < 0 || n_nodes < 0 || ld_mass < n_nodes || ld_out < n_ids || ids_ld < 0 || (flags & ~(unsigned)GT_GATHER_LOG)) {
        gt::set_error("gt_gather_nodes: bad size / stride / flags"); return GT_ERR_ARG;
    }
    if (n_rows == 0 || n_ids == 0) return GT_OK;
    if (!mass || !node_ids || !out) { gt::set_error("gt_gather_nodes: null data pointer"); return GT_ERR_ARG; }
    if (n_rows > INT32_MAX || n_ids > INT32_MAX) { gt::set_error("gt_gather_nodes: too many rows / ids"); return GT_ERR_LIMIT; }
    if (type != GT_F32 && type != GT_F64) { gt::set_error("gt_gather_nodes: type must be GT_F32 or GT_F64"); return GT_ERR_ARG; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)((n_ids + 255) / 256), (unsigned)std::min<int64_t>(n_rows, 32768));
    const int lg = (flags & GT_GATHER_LOG) ? 1 : 0;
    if (type == GT_F32)
        gt::gather_nodes_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(mass), ld_mass, (int)n_rows, n_nodes, node_ids,
                                                            (int)n_ids, ids_ld, norm_node, lg, static_cast<float*>(out), ld_out);
    else
        gt::gather_nodes_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double*>(mass), ld_mass, (int)n_rows, n_nodes, node_ids,
                                                             (int)n_ids, ids_ld, norm_node, lg, static_cast<double*>(out), ld_out);
    GT_CUDA(cudaGetLastError());
    return GT_OK;
}

int gt_subtree_token_mask(const gt_trie* t, const int32_t* nodes, int64_t n_rows, uint32_t* mask_bits, int64_t mask_ld,
                          gt_stream stream) {
    if (!t) { gt::set_error("gt_subtree_token_mask: null trie"); return GT_ERR_ARG; }
    const int64_t words = (t->layout.V + 31) / 32;
    if (n_rows < 0 || mask_ld < words) { gt::set_error("gt_subtree_token_mask: bad n_rows / mask_ld (need >= %lld words)", (long long)words); return GT_ERR_ARG; }
    if (n_rows == 0 || words == 0) return GT_OK;
    if (!nodes || !mask_bits) { gt::set_error("gt_subtree_token_mask: null data pointer"); return GT_ERR_ARG; }
    if (n_rows > (int64_t)65535 * gt::kMaskRows) { gt::set_error("gt_subtree_token_mask: too many rows"); return GT_ERR_LIMIT; }
    int device = -1;
    GT_CUDA(cudaGetDevice(&device));
    auto it = t->dev.find(device);
    if (it == t->dev.end()) { gt::set_error("trie metadata is not resident on device %d (call gt_upload)", device); return GT_ERR_STATE; }
    // the grid covers whole mask words: ceil(V/32) of them, kMaskThreads/32 per CTA
    dim3 grid((unsigned)((words * 32 + gt::kMaskThreads - 1) / gt::kMaskThreads), (unsigned)((n_rows + gt::kMaskRows - 1) / gt::kMaskRows));
    gt::subtree_mask_kernel<<<grid, gt::kMaskThreads, 0, static_cast<cudaStream_t>(stream)>>>(it->second->view, nodes, (int)n_rows, mask_bits, mask_ld);
    GT_CUDA(cudaGetLastError());
    return GT_OK;
}

}  // extern "C"
