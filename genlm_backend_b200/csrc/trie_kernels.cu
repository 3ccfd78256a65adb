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
//   phase 2  tile_kernel    : staged tile -> DFS-ordered leaf values in shared memory (rows of a row group
//                             interleaved per slot) -> aligned-block pyramid (warp shuffles) -> multi-term
//                             ranges (ELL-packed term lists) -> coalesced 128-bit emit of the tile's node-id
//                             interval.  Nodes whose leaf range crosses tiles are written as per-tile pieces.
//   phase 3  span_kernel    : those few nodes, reduced from their pieces (fp64 for sums).
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
    int32_t T, logT, Q, NT, NS, SV;  // SV = value slots per tile in shared memory
    int32_t R, max_tile_nodes, max_tile_ell_rows, max_tile_chunks, max_tile_z, max_seg_recs;
    int64_t V, N, Zrow;
    const int32_t* p1_chunk_ptr; const int4* p1_rec;
    const int32_t* z_tile_off; const uint16_t* p2_slot;
    const int32_t* ell_chunk_ptr; const int2* ell_desc; const uint16_t* ell_terms; const int32_t* ell_row_ptr;
    const int32_t* tile_node_lo; const uint16_t* node_slot;
    const int32_t* piece_ptr; const uint16_t* piece_slot; const int32_t* piece_idx;
    int32_t n_span, n_pieces; const int32_t* span_node; const int32_t* span_pp;
    const int32_t* leaf_rank;  // [V] DFS rank of each item's leaf (inverse of Layout::perm)
    const int32_t* node_lo; const int32_t* node_hi;  // [N] DFS leaf range of every node
    int32_t debug_stop;  // profiling aid (GT_DEBUG_STOP): 0 = normal; 3 = tile_kernel skips the emit stores; 9 = emit only
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

constexpr int kThreads = 512;
constexpr int OP_SUM = 1, OP_MAX = 2;
constexpr size_t kMaxSmem = 200 * 1024;  // dynamic shared memory we are willing to ask for per CTA
constexpr int kSegPad = 8;  // slack so a row segment can be stored at its global 16-byte phase

// ---- programmatic dependent launch ---------------------------------------------------------------------
// Consecutive kernels of one call (permute -> tile -> span -> tile -> span) are launched with the programmatic
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
    pdl_wait();  // the previous call's tile kernels may still be reading z
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
    constexpr int U = 4;  // records in flight per thread
    for (int cb = c0 + threadIdx.x; cb < c1; cb += U * kThreads) {
        int4 rec[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = cb + u * kThreads;
            rec[u] = c < c1 ? __ldg(P.p1_rec + c) : make_int4(-1, -1, -1, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (rec[u].x < 0) continue;
            const unsigned s0 = (unsigned)rec[u].y & 0xFFFFu, s1 = (unsigned)rec[u].y >> 16;
            const unsigned s2 = (unsigned)rec[u].z & 0xFFFFu, s3 = (unsigned)rec[u].z >> 16;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r < nrows) {
                    const VT* sr = seg + r * pitch + phase[r];
                    const VT a = s0 != 0xFFFFu ? sr[s0] : VT(0);
                    const VT b = s1 != 0xFFFFu ? sr[s1] : VT(0);
                    const VT c = s2 != 0xFFFFu ? sr[s2] : VT(0);
                    const VT d = s3 != 0xFFFFu ? sr[s3] : VT(0);
                    store4<VT>(z + (size_t)(b0 + r) * P.Zrow + rec[u].x, a, b, c, d);
                }
            }
        }
    }
}

// ---- phase 2: per-tile pyramid, multi-term ranges, emit, spanning pieces -------------------------------------
//
// Persistent, warp-specialised kernel.  A work item is (tile t, row group g of R rows); items are numbered
// tile-major and every CTA takes one contiguous run of them, so consecutive items of a CTA share the tile metadata,
// which stays in shared memory.  The CTA is split into two groups that run one item apart over a double-buffered
// value array:
//     compute warps  staged rows -> leaf slots -> pyramid -> multi-term ranges           (fill  vals[item & 1])
//     emit warps     node-id interval of the tile -> global memory, spanning-node pieces  (drain vals[item & 1])
// so the output stores -- the HBM-bound part -- stream continuously while the next item is being built.
// Everything read from global memory arrives by bulk copies (cp.async.bulk, the TMA engine) issued one step ahead
// by one thread of the group that consumes it, tracked by mbarriers:
//     A   staged rows of the next item (+ the slot table of its tile if new)   armed after each scatter
//     B   ELL term rows + descriptors of the next tile                         armed after the last ELL of a tile
//     C   emit slots of the next tile                                          armed after the last emit of a tile
//     full[2] / empty[2]   hand-over of the two value arrays between the groups
// Every wait on A/B/C precedes a group barrier and every re-arm follows it, so no thread can still be waiting on
// a phase when the next one completes.

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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

// ---- phase 1, bulk-copy variant: persistent CTAs, rows fetched by the copy engine one group ahead -----------------
// Used when every row segment is 16-byte aligned (row stride and row length multiples of 16 bytes).  The work is the
// NS x n_rows (segment, row) pairs, numbered segment-major; every CTA takes one contiguous, equally long run of them
// -- balance is to the row, not to the row group -- and walks it in groups of up to RP rows of one segment, so the
// segment's records stay in shared memory across consecutive groups and each record read serves RP rows.  Rows land
// in shared memory in their input type (NST stages); conversion / exp happens in the gather.
constexpr int kPermPad = 16;  // slack after every staged row: a row segment lands at its 16-byte phase in global memory
// ALIGNED: every row segment starts and ends on a 16-byte boundary (the common case; no phase, no tail).
template <typename VT, typename IN_T, int RP, int NST, bool LOG, bool ALIGNED>
__global__ void __launch_bounds__(kThreads, 2) permute_bulk_kernel(PlanView P, const IN_T* __restrict__ ws, int64_t ld_ws,
                                                                   VT* __restrict__ z, int n_rows) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int Q = P.Q;
    constexpr int kPadElems = kPermPad / (int)sizeof(IN_T);
    const int pitch = Q + kPadElems;                                                             // elements per staged row
    IN_T* stage = reinterpret_cast<IN_T*>(smem_raw);                                             // [NST][RP][pitch]
    int4* s_rec = reinterpret_cast<int4*>(smem_raw + (size_t)NST * RP * pitch * sizeof(IN_T));   // [max_seg_recs]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_rec) + (size_t)P.max_seg_recs * 16);
    uint64_t* full = bars;          // [NST] rows of a stage have landed
    uint64_t* rec_bar = bars + NST; // records of the current segment have landed

    const int tid = threadIdx.x;
    const int64_t n_units = (int64_t)P.NS * n_rows;
    const int64_t u0 = (int64_t)blockIdx.x * n_units / gridDim.x, u1 = (int64_t)(blockIdx.x + 1) * n_units / gridDim.x;
    if (u0 >= u1) return;
    const int dbg = P.debug_stop;  // profiling aid: 21 = rows are fetched but not gathered / stored, 22 = gather / store only

    // the group that starts at unit u: rows [row, row + n) of segment s
    auto group_at = [&](int64_t u, int& gs, int& grow, int& gn) {
        gs = (int)(u / n_rows);
        grow = (int)(u - (int64_t)gs * n_rows);
        gn = (int)min((int64_t)min(RP, n_rows - grow), u1 - u);
    };
    auto seg_len = [&](int sg) { return (int)min((int64_t)Q, P.V - (int64_t)sg * Q); };
    // A bulk copy wants 16-byte aligned addresses and sizes on both sides, rows need not have them (an odd vocabulary
    // size shifts every row): the copy covers the aligned span around the segment, and the segment's first element
    // lands `phase` bytes into the staged row.  The span ends at most 15 bytes past the segment, inside the same row or
    // the next one -- except for the very last segment of the batch, whose bulk copy stops at the last whole 16 bytes
    // and whose tail elements are loaded one by one (tail fix-up below).
    auto row_ptr = [&](int sg, int r) { return ws + (size_t)r * ld_ws + (size_t)sg * Q; };
    auto phase_bytes = [&](int sg, int r) { return ALIGNED ? 0u : (unsigned)(reinterpret_cast<uintptr_t>(row_ptr(sg, r)) & 15u); };
    auto is_batch_end = [&](int sg, int r) { return !ALIGNED && r == n_rows - 1 && sg == P.NS - 1; };
    auto fetch_rows = [&](int sg, int row, int n, int st) {  // thread 0
        if (dbg == 22) { mbar_expect_tx(full + st, 0); return; }
        if constexpr (ALIGNED) {
            const unsigned bytes = (unsigned)seg_len(sg) * (unsigned)sizeof(IN_T);
            mbar_expect_tx(full + st, (unsigned)n * bytes);
            for (int r = 0; r < n; ++r) bulk_g2s(stage + ((size_t)st * RP + r) * pitch, row_ptr(sg, row + r), bytes, full + st);
            return;
        }
        unsigned span[RP], total = 0;
        for (int r = 0; r < n; ++r) {
            const unsigned need = phase_bytes(sg, row + r) + (unsigned)seg_len(sg) * (unsigned)sizeof(IN_T);
            span[r] = is_batch_end(sg, row + r) ? (need & ~15u) : ((need + 15u) & ~15u);
            total += span[r];
        }
        mbar_expect_tx(full + st, total);
        for (int r = 0; r < n; ++r)
            if (span[r])
                bulk_g2s(stage + ((size_t)st * RP + r) * pitch,
                         reinterpret_cast<const unsigned char*>(row_ptr(sg, row + r)) - phase_bytes(sg, row + r), span[r], full + st);
    };
    auto fetch_recs = [&](int c0, int c1) {  // thread 0
        mbar_expect_tx(rec_bar, (unsigned)(c1 - c0) * 16u);
        if (c1 > c0) bulk_g2s(s_rec, P.p1_rec + c0, (unsigned)(c1 - c0) * 16u, rec_bar);
    };

    int s, row, n;
    group_at(u0, s, row, n);
    int c0 = __ldg(P.p1_chunk_ptr + s), c1 = __ldg(P.p1_chunk_ptr + s + 1), c2 = __ldg(P.p1_chunk_ptr + min(s + 2, P.NS));
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) mbar_init(full + i, 1);
        mbar_init(rec_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fetch_recs(c0, c1);  // plan metadata: may run ahead of the previous kernel's completion
    }
    pdl_wait();  // the previous call's tile kernel may still be reading z; ws may come from a kernel of the caller
    // the thread that issues the fetches runs NST-1 groups ahead
    int64_t fu = u0;
    int fk = 0;
    auto fetch_advance = [&]() {
        if (fu < u1) {
            int fs, frow, fn;
            group_at(fu, fs, frow, fn);
            fetch_rows(fs, frow, fn, fk % NST);
            fu += fn;
        }
        ++fk;
    };
    if (tid == 0)
        for (int i = 0; i < NST - 1; ++i) fetch_advance();
    __syncthreads();

    bool new_seg = true;
    unsigned par_rec = 0;
    int k = 0;
    for (int64_t u = u0; u < u1; ++k) {
        const int st = k % NST;
        const int64_t un = u + n;
        const bool has_next = un < u1;
        int sn = s, rown = row, nn = n;
        if (has_next) group_at(un, sn, rown, nn);
        if (tid == 0) fetch_advance();  // group k + NST - 1: its stage was released by the barrier at the end of group k - 1
        if (new_seg) { mbar_wait(rec_bar, par_rec); par_rec ^= 1u; }
        mbar_wait(full + st, (unsigned)(k / NST) & 1u);

        IN_T* sr0 = stage + (size_t)st * RP * pitch;
        if (is_batch_end(s, row + n - 1) && dbg != 22) {  // tail fix-up: once per launch, in one CTA
            const unsigned ph = phase_bytes(s, row + n - 1), need = ph + (unsigned)seg_len(s) * (unsigned)sizeof(IN_T);
            const int count = (int)((need & 15u) / sizeof(IN_T));           // elements past the last whole 16 bytes
            const int first = seg_len(s) - count;                           // ... and where they start in the segment
            if (tid < count) sr0[(size_t)(n - 1) * pitch + ph / sizeof(IN_T) + first + tid] = row_ptr(s, row + n - 1)[first + tid];
            __syncthreads();
        }
        const int nrec = dbg == 21 ? 0 : c1 - c0;
        VT* zr[RP];
        const IN_T* sr[RP];
#pragma unroll
        for (int r = 0; r < RP; ++r) {
            const int rr = row + min(r, n - 1);
            zr[r] = z + (size_t)rr * P.Zrow;
            sr[r] = sr0 + (size_t)min(r, n - 1) * pitch + phase_bytes(s, rr) / sizeof(IN_T);
        }
        // Padding positions of a record carry source position 0: what lands in z there goes to a trash slot of the
        // tile kernel and is never read, so no position needs a check.
        for (int i = tid; i < nrec; i += kThreads) {
            const int4 rec = s_rec[i];
            const unsigned p0 = (unsigned)rec.y & 0xFFFFu, p1 = (unsigned)rec.y >> 16;
            const unsigned p2 = (unsigned)rec.z & 0xFFFFu, p3 = (unsigned)rec.z >> 16;
#pragma unroll
            for (int r = 0; r < RP; ++r) {
                if (r < n)
                    store4<VT>(zr[r] + rec.x, convert_in<VT, IN_T>(sr[r][p0], LOG), convert_in<VT, IN_T>(sr[r][p1], LOG),
                               convert_in<VT, IN_T>(sr[r][p2], LOG), convert_in<VT, IN_T>(sr[r][p3], LOG));
            }
        }
        __syncthreads();  // this stage (and, on a segment change, the record buffer) may be overwritten
        new_seg = has_next && sn != s;
        if (new_seg) {
            c0 = c1; c1 = c2; c2 = __ldg(P.p1_chunk_ptr + min(sn + 2, P.NS));
            if (tid == 0) fetch_recs(c0, c1);
        }
        s = sn; row = rown; n = nn; u = un;
    }
}

constexpr int kTileThreads = 512;
// trace layout: [kTraceCtas][kTraceItems][kTraceEvents] SM-clock stamps
constexpr int kTraceCtas = 512, kTraceItems = 32, kTraceEvents = 12;
#define GT_TRACE(ev)                                                                                         \
    do {                                                                                                     \
        if (P.trace && tid == 0 && blockIdx.x < kTraceCtas && k < kTraceItems)                                \
            P.trace[((size_t)blockIdx.x * kTraceItems + k) * kTraceEvents + (ev)] = clock64();               \
    } while (0)
#ifndef GT_COMPUTE_WARPS
#define GT_COMPUTE_WARPS 8
#endif
#ifndef GT_ELL_BATCH
#define GT_ELL_BATCH 4  // terms of a multi-term range loaded per step (4, or 8: measured slower, the phase is bound by
                        // shared-memory bandwidth, not by the latency of a batch)
#endif
constexpr int kComputeThreads = 32 * GT_COMPUTE_WARPS;        // warps 0 .. GT_COMPUTE_WARPS-1
constexpr int kEmitThreads = kTileThreads - kComputeThreads;  // the remaining warps

// Shared-memory carve-up of tile_kernel (all sections 16-byte aligned).
struct TileSmem {
    size_t vals, vals_bytes, stage, p2, slots, terms, desc, bars, total;
    __host__ __device__ TileSmem(const PlanView& P, int slot_bytes, int elem_bytes, int R) {
        size_t o = 0;
        vals_bytes = (size_t)(P.SV + 16) * slot_bytes;                      // value slots + one trash slot per bank group
        vals = o;  o += 2 * vals_bytes;                                     // double-buffered between the groups
        stage = o; o += (size_t)R * P.max_tile_z * elem_bytes;              // staged rows of the next item
        p2 = o;    o += (size_t)P.max_tile_z * 2;                           // staged element -> value slot
        slots = o; o += (size_t)P.max_tile_nodes * 2;                       // node -> value slot (emit)
        terms = o; o += (size_t)P.max_tile_ell_rows * 64;                   // ELL term rows
        desc = o;  o += ((size_t)(P.max_tile_chunks + 2) * 8 + 15) & ~size_t(15); // ELL chunk descriptors (+ alignment slack)
        bars = o;  o += 64;                                                 // seven mbarriers
        total = o;
    }
};

__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
template <int ID, int COUNT> __device__ __forceinline__ void group_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
}

// What one launch works on: the staged rows and, per requested reduction, the output slab and the scratch for the
// pieces of spanning nodes.  ops = GT_OP_SUM | GT_OP_MAX; with both, the two reductions of a (tile, row group) are
// consecutive items that share one fetch of the staged rows.
template <typename VT> struct TileArgs {
    const VT* z;
    VT* out_sum; VT* out_max;
    VT* part_sum; VT* part_max;
    int64_t ld_out;
    int n_rows;
    unsigned ops;
};

// ---- compute-group phases (OP is a compile-time parameter; the kernel branches once per phase) ----------------------

// 1. staged rows -> DFS-ordered leaf slots (slot numbers in the table are already swizzled)
template <typename VT, int R, int OP, int NTHREADS>
__device__ __forceinline__ void phase_scatter(VT* vals, const VT* stage, const uint16_t* s_p2, int zpitch, int zn4, int nleaf,
                                              int T, int tid) {
    using RV = RowVec<VT, R>;
    constexpr int B = (int)sizeof(VT) * R;
    constexpr int U = R >= 4 ? 1 : 2;  // quads in flight per thread (a second one clamps to a duplicate of the last quad)
    for (int qb = tid; qb < zn4; qb += U * NTHREADS) {
        uint2 sl[U];
        VT v[U][R][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int q = min(qb + u * NTHREADS, zn4 - 1);
            sl[u] = *reinterpret_cast<const uint2*>(s_p2 + 4 * q);
#pragma unroll
            for (int r = 0; r < R; ++r) load4<VT>(stage + (size_t)r * zpitch + 4 * q, v[u][r][0], v[u][r][1], v[u][r][2], v[u][r][3]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned s4[4] = {sl[u].x & 0xFFFFu, sl[u].x >> 16, sl[u].y & 0xFFFFu, sl[u].y >> 16};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                RV x;
#pragma unroll
                for (int r = 0; r < R; ++r) x.v[r] = v[u][r][e];
                x.store(vals + s4[e] * R);
            }
        }
    }
    for (int i = nleaf + tid; i < T; i += NTHREADS) RV::template ident<OP>().store(vals + swz<B>(i) * R);
    if (tid == 0) RV::template ident<OP>().store(vals + swz<B>(2 * T - 1) * R);  // identity slot (ELL padding, spanning nodes)
}

// 2. pyramid of aligned blocks: level k block i at (swizzled) slot 2T - (T >> (k-1)) + i, levels 1 .. kPyramidTop.
//    Lane u owns LPL = 2^(kPyramidTop-5) consecutive leaves: the first log2(LPL) levels in registers, five more by warp
//    shuffles, so a warp covers 32*LPL leaves and no level needs a cross-warp step (longer aligned blocks are
//    multi-term ranges of the ELL phase).  kPyramidTop = 8: 8 leaves per lane, 256 leaves per warp.
//    The swizzle makes the 16-byte chunk loads and the strided level stores bank-conflict free.
template <typename VT, int R, int OP, int NWARPS>
__device__ __forceinline__ void phase_pyramid(VT* vals, int T, int warp, int lane) {
    using RV = RowVec<VT, R>;
    constexpr int B = (int)sizeof(VT) * R;
    constexpr int SPC = 16 / B;  // slots per 16-byte chunk
    constexpr int LL = kPyramidTop - 5, LPL = 1 << LL;  // in-lane levels, leaves per lane
    static_assert(LL >= 1 && LPL >= SPC, "a lane owns at least one 16-byte chunk of leaves");
    auto level_slot = [&](int k, int i) { return 2 * T - (T >> (k - 1)) + i; };  // level k >= 1, block i
#if GT_PYR_TOP == 8 && !defined(GT_PYR_GENERIC)
    for (int ub = warp * 32; ub < (T >> 3); ub += NWARPS * 32) {  // hand-scheduled form of the loop below for 8 leaves per lane
        const int u = ub + lane;
        RV x[8];
#pragma unroll
        for (int ch = 0; ch < 8 / SPC; ++ch) {
            const int c = (8 * u) / SPC + ch;
            const int cc = c ^ ((c >> 3) & (B / 2 - 1));
            const float4 raw = *reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(vals) + (size_t)cc * 16);
            memcpy(&x[ch * SPC], &raw, 16);
        }
        RV a[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            a[e] = RV::template combine<OP>(x[2 * e], x[2 * e + 1]);
            a[e].store(vals + swz<B>(level_slot(1, 4 * u + e)) * R);
        }
        const RV c0 = RV::template combine<OP>(a[0], a[1]), c1 = RV::template combine<OP>(a[2], a[3]);
        c0.store(vals + swz<B>(level_slot(2, 2 * u)) * R);
        c1.store(vals + swz<B>(level_slot(2, 2 * u + 1)) * R);
        RV y = RV::template combine<OP>(c0, c1);
        y.store(vals + swz<B>(level_slot(3, u)) * R);
#pragma unroll
        for (int j = 1; j <= 5; ++j) {
            y = RV::template combine<OP>(y, y.shfl_down(1 << (j - 1)));
            if ((lane & ((1 << j) - 1)) == 0) y.store(vals + swz<B>(level_slot(3 + j, u >> j)) * R);
        }
    }
    return;
#elif GT_PYR_TOP == 7 && !defined(GT_PYR_GENERIC)
    for (int ub = warp * 32; ub < (T >> 2); ub += NWARPS * 32) {  // hand-scheduled form of the loop below for 4 leaves per lane
        const int u = ub + lane;
        RV x[4];
#pragma unroll
        for (int ch = 0; ch < 4 / SPC; ++ch) {
            const int c = (4 * u) / SPC + ch;
            const int cc = c ^ ((c >> 3) & (B / 2 - 1));
            const float4 raw = *reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(vals) + (size_t)cc * 16);
            memcpy(&x[ch * SPC], &raw, 16);
        }
        const RV a0 = RV::template combine<OP>(x[0], x[1]), a1 = RV::template combine<OP>(x[2], x[3]);
        a0.store(vals + swz<B>(level_slot(1, 2 * u)) * R);
        a1.store(vals + swz<B>(level_slot(1, 2 * u + 1)) * R);
        RV y = RV::template combine<OP>(a0, a1);
        y.store(vals + swz<B>(level_slot(2, u)) * R);
#pragma unroll
        for (int j = 1; j <= 5; ++j) {
            y = RV::template combine<OP>(y, y.shfl_down(1 << (j - 1)));
            if ((lane & ((1 << j) - 1)) == 0) y.store(vals + swz<B>(level_slot(2 + j, u >> j)) * R);
        }
    }
    return;
#endif
    for (int ub = warp * 32; ub < T / LPL; ub += NWARPS * 32) {
        const int u = ub + lane;
        RV x[LPL];
#pragma unroll
        for (int ch = 0; ch < LPL / SPC; ++ch) {
            const int c = (LPL * u) / SPC + ch;
            const int cc = c ^ ((c >> 3) & (B / 2 - 1));
            const float4 raw = *reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(vals) + (size_t)cc * 16);
            memcpy(&x[ch * SPC], &raw, 16);
        }
#pragma unroll
        for (int k = 1; k <= LL; ++k) {  // level k: LPL >> k blocks of this lane, written in place over x[0 ..)
#pragma unroll
            for (int e = 0; e < (LPL >> k); ++e) {
                x[e] = RV::template combine<OP>(x[2 * e], x[2 * e + 1]);
                x[e].store(vals + swz<B>(level_slot(k, (LPL >> k) * u + e)) * R);
            }
        }
        RV y = x[0];
#pragma unroll
        for (int j = 1; j <= 5; ++j) {
            y = RV::template combine<OP>(y, y.shfl_down(1 << (j - 1)));
            if ((lane & ((1 << j) - 1)) == 0) y.store(vals + swz<B>(level_slot(LL + j, u >> j)) * R);
        }
    }
}

// 3. ranges that need more than one block: ELL chunks of 32 ranges, one warp per chunk, term rows read from shared
//    memory (k is a multiple of kEllRowPad: the planner pads rows with the identity slot).  Chunks are sorted by descending
//    term count: rounds alternate direction so that the warps that drew the longest chunks of one round draw the
//    shortest of the next.
template <typename VT, int R, int OP, int NWARPS>
__device__ __forceinline__ void phase_ell(VT* vals, const uint16_t* s_terms, const int2* dsc, int er0, int nchunks, int T,
                                          int warp, int lane) {
    using RV = RowVec<VT, R>;
    for (int base = 0, rr = 0; base < nchunks; base += NWARPS, ++rr) {
        const int c = base + ((rr & 1) ? NWARPS - 1 - warp : warp);
        if (c >= nchunks) continue;
        const int2 d = dsc[c];
        const uint16_t* tp = s_terms + (d.x - er0) * 32 + lane;
        RV acc = RV::template ident<OP>();
        // eight terms at a time while they last (all sixteen shared-memory loads of a batch are independent: the
        // longest ranges of a tile set the length of this phase), then at most one batch of four
        int kb = 0;
#if GT_ELL_BATCH >= 8
        for (; kb + 8 <= d.y; kb += 8) {
            int sl[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) sl[e] = tp[(kb + e) * 32];
            RV v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = RV::load(vals + sl[e] * R);
#pragma unroll
            for (int w = 1; w < 8; w <<= 1)
#pragma unroll
                for (int e = 0; e + w < 8; e += 2 * w) v[e] = RV::template combine<OP>(v[e], v[e + w]);
            acc = RV::template combine<OP>(acc, v[0]);
        }
#endif
        for (; kb + 4 <= d.y; kb += 4) {
            int sl[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) sl[e] = tp[(kb + e) * 32];
#pragma unroll
            for (int e = 0; e < 4; ++e) acc = RV::template combine<OP>(acc, RV::load(vals + sl[e] * R));
        }
#if GT_ELL_ROW_PAD == 1
        if (kb + 2 <= d.y) {  // no padding rows at all: a pair, then a single term
            const int s0 = tp[kb * 32], s1 = tp[(kb + 1) * 32];
            acc = RV::template combine<OP>(acc, RV::template combine<OP>(RV::load(vals + s0 * R), RV::load(vals + s1 * R)));
            kb += 2;
        }
        if (kb < d.y) acc = RV::template combine<OP>(acc, RV::load(vals + (int)tp[kb * 32] * R));
#else
        if (kb < d.y) {  // term rows come in pairs (kEllRowPad = 2): most chunks of a tile hold two-term ranges only
            const int s0 = tp[kb * 32], s1 = tp[(kb + 1) * 32];
            acc = RV::template combine<OP>(acc, RV::template combine<OP>(RV::load(vals + s0 * R), RV::load(vals + s1 * R)));
        }
#endif
        acc.store(vals + (2 * T + c * 32 + lane) * R);
    }
}

template <typename VT, int R>
__global__ void __launch_bounds__(kTileThreads, 2) tile_kernel(PlanView P, TileArgs<VT> A) {
    using RV = RowVec<VT, R>;
    constexpr int B = (int)sizeof(VT) * R;  // bytes per slot
    static_assert(B == 4 || B == 8 || B == 16, "slot must be 4, 8 or 16 bytes");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileSmem L(P, B, (int)sizeof(VT), R);
    VT* stage = reinterpret_cast<VT*>(smem_raw + L.stage);
    uint16_t* s_p2 = reinterpret_cast<uint16_t*>(smem_raw + L.p2);
    uint16_t* s_slots = reinterpret_cast<uint16_t*>(smem_raw + L.slots);
    uint16_t* s_terms = reinterpret_cast<uint16_t*>(smem_raw + L.terms);
    int2* s_desc = reinterpret_cast<int2*>(smem_raw + L.desc);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bars);
    uint64_t* barA = bars;          // rows (+ p2)
    uint64_t* barB = bars + 1;      // terms + descriptors
    uint64_t* barC = bars + 2;      // emit slots
    uint64_t* full = bars + 3;      // [2] value array filled by the compute group
    uint64_t* empty = bars + 5;     // [2] value array drained by the emit group

    const int T = P.T;
    const int n_rows = A.n_rows;
    const int RG = (n_rows + R - 1) / R;
    const int nops = (A.ops == (unsigned)(GT_OP_SUM | GT_OP_MAX)) ? 2 : 1;
    const int first_op = (A.ops & GT_OP_SUM) ? OP_SUM : OP_MAX;
    const int n_items = P.NT * RG * nops;  // item = (tile * RG + row group) * nops + j
    const int i0 = (int)((int64_t)blockIdx.x * n_items / gridDim.x);
    const int i1 = (int)((int64_t)(blockIdx.x + 1) * n_items / gridDim.x);
    if (i0 >= i1) return;
    const int zpitch = P.max_tile_z;
    const int dbg = P.debug_stop;  // profiling aid: 3 = no emit stores, 9 = emit only (compute phases skipped)

    if (threadIdx.x == 0) {
        mbar_init(barA, 1); mbar_init(barB, 1); mbar_init(barC, 1);
        mbar_init(full, kComputeThreads); mbar_init(full + 1, kComputeThreads);
        mbar_init(empty, kEmitThreads); mbar_init(empty + 1, kEmitThreads);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();  // barriers initialised before anyone uses them

    if (threadIdx.x < kComputeThreads) {
        // =========================== compute group ===========================================================
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        constexpr int kWarps = kComputeThreads / 32;
        constexpr int kIssueTid = kComputeThreads - 32;  // lane 0 of the last compute warp issues the fetches: that
                                                         // warp has no share of the pyramid, so nobody waits for it
        // Boundaries of the current tile (index 0..1) and the next one (1..2) in the per-tile prefix arrays; the next
        // tile's are loaded one tile ahead so that no fetch waits for them.
        int zb[3], erb[3], ecb[3];
        auto load_bound = [&](int t, int j) {
            const int tt = min(t, P.NT);
            zb[j] = __ldg(P.z_tile_off + tt); erb[j] = __ldg(P.ell_row_ptr + tt); ecb[j] = __ldg(P.ell_chunk_ptr + tt);
        };
        // staged rows of row group g (+ the tile's slot table).  Rows past the end of the batch alias the last valid
        // row: they compute exactly what that row does, which keeps every loop free of row predicates.
        auto fetch_rows = [&](int g, int zlo, int zn, bool with_p2) {
            const unsigned row_bytes = (unsigned)zn * (unsigned)sizeof(VT);
            mbar_expect_tx(barA, R * row_bytes + (with_p2 ? (unsigned)zn * 2u : 0u));
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int row = min(g * R + r, n_rows - 1);
                bulk_g2s(stage + (size_t)r * zpitch, A.z + (size_t)row * P.Zrow + zlo, row_bytes, barA);
            }
            if (with_p2) bulk_g2s(s_p2, P.p2_slot + zlo, (unsigned)zn * 2u, barA);
        };
        // ELL term rows and chunk descriptors (8 bytes each, copied from the 16-byte aligned pair at or below the
        // first one)
        auto fetch_terms = [&](int er0, int er1, int ec0, int ec1) {
            const int ea = ec0 & ~1;
            const unsigned tb = (unsigned)(er1 - er0) * 64u, db = (unsigned)((ec1 - ea + 1) >> 1) * 16u;
            mbar_expect_tx(barB, tb + db);
            if (tb) bulk_g2s(s_terms, P.ell_terms + (size_t)er0 * 32, tb, barB);
            if (db) bulk_g2s(s_desc, P.ell_desc + ea, db, barB);
        };

        int tg = i0 / nops, j = i0 - tg * nops;  // (tile, row group) index and which reduction of it
        int t = tg / RG, g = tg - t * RG;
        load_bound(t, 0); load_bound(t + 1, 1); load_bound(t + 2, 2);
        // plan metadata may be fetched ahead of the previous kernel's completion; the staged rows may not
        if (tid == kIssueTid) fetch_terms(erb[0], erb[1], ecb[0], ecb[1]);
        pdl_wait();  // z comes from permute_kernel
        pdl_trigger();
        if (tid == kIssueTid) fetch_rows(g, zb[0], zb[1] - zb[0], true);
        bool new_tile = true;
        unsigned parB = 0;
        for (int item = i0; item < i1; ++item) {
            const int k = item - i0;
            VT* vals = reinterpret_cast<VT*>(smem_raw + L.vals + (size_t)(k & 1) * L.vals_bytes);
            const bool is_sum = (j == 0 ? first_op : OP_MAX) == OP_SUM;
            // the item after this one
            int jn = j + 1, tn = t, gn = g;
            if (jn == nops) { jn = 0; if (++gn == RG) { gn = 0; ++tn; } }
            const bool has_next = item + 1 < i1;
            const bool next_rows = has_next && jn == 0;   // the next item needs other staged rows
            const bool next_new = has_next && tn != t;    // ... of another tile
            const int zn4 = (zb[1] - zb[0]) >> 2, nchunks = ecb[1] - ecb[0];
            const int nleaf = (int)min((int64_t)T, P.V - (int64_t)t * T);

            GT_TRACE(0);
            mbar_wait(empty + (k & 1), ((unsigned)(k >> 1) & 1u) ^ 1u);  // the emit group has drained this value array
            GT_TRACE(1);
            mbar_wait(barA, (unsigned)k & 1u);
            GT_TRACE(2);
            if (dbg != 9) {
                if (is_sum) phase_scatter<VT, R, OP_SUM, kComputeThreads>(vals, stage, s_p2, zpitch, zn4, nleaf, T, tid);
                else phase_scatter<VT, R, OP_MAX, kComputeThreads>(vals, stage, s_p2, zpitch, zn4, nleaf, T, tid);
            }
            group_sync<1, kComputeThreads>();
            GT_TRACE(3);
            // the staging buffer (and, on a tile change, the slot table) is free again: fetch the next item's rows,
            // unless it is the other reduction of the same rows (the barrier is armed once per item either way)
            if (tid == kIssueTid) {
                if (next_rows) fetch_rows(gn, next_new ? zb[1] : zb[0], next_new ? zb[2] - zb[1] : zb[1] - zb[0], next_new);
                else mbar_expect_tx(barA, 0);
            }
            GT_TRACE(4);
            if (dbg != 9) {
                if (is_sum) phase_pyramid<VT, R, OP_SUM, kWarps>(vals, T, warp, lane);
                else phase_pyramid<VT, R, OP_MAX, kWarps>(vals, T, warp, lane);
            }
            GT_TRACE(5);
            if (new_tile) { mbar_wait(barB, parB); parB ^= 1u; }  // ELL terms + descriptors of this tile
            group_sync<1, kComputeThreads>();
            GT_TRACE(6);
            if (dbg != 9) {
                const int2* dsc = s_desc + (ecb[0] & 1);
                if (is_sum) phase_ell<VT, R, OP_SUM, kWarps>(vals, s_terms, dsc, erb[0], nchunks, T, warp, lane);
                else phase_ell<VT, R, OP_MAX, kWarps>(vals, s_terms, dsc, erb[0], nchunks, T, warp, lane);
            }
            GT_TRACE(7);
            mbar_arrive(full + (k & 1));  // this thread's share of the value array is complete
            if (next_new) {
                group_sync<1, kComputeThreads>();  // everyone is done with this tile's terms
                if (tid == kIssueTid) fetch_terms(erb[1], erb[2], ecb[1], ecb[2]);
#pragma unroll
                for (int q = 0; q < 2; ++q) { zb[q] = zb[q + 1]; erb[q] = erb[q + 1]; ecb[q] = ecb[q + 1]; }
                load_bound(tn + 1, 2);
            }
            new_tile = next_new;
            t = tn; g = gn; j = jn;
        }
    } else {
        // =========================== emit group ==============================================================
        const int tid = threadIdx.x - kComputeThreads;
        // emit slots of a tile, staged from the 16-byte aligned start at or below its first node
        auto fetch_slots = [&](int n0, int n1) {
            const int na = n0 & ~7;
            const unsigned sb = (unsigned)((n1 - na + 7) >> 3) * 16u;
            mbar_expect_tx(barC, sb);
            bulk_g2s(s_slots, P.node_slot + na, sb, barC);
        };
        int tg = i0 / nops, j = i0 - tg * nops;
        int t = tg / RG, g = tg - t * RG;
        int nb3[3];  // node-id boundaries of this tile and the next (loaded one tile ahead)
#pragma unroll
        for (int q = 0; q < 3; ++q) nb3[q] = __ldg(P.tile_node_lo + min(t + q, P.NT));
        if (tid == 0) fetch_slots(nb3[0], nb3[1]);
        pdl_wait();  // the outputs and piece buffers may still be in use by the previous kernels of the chain
        int pc0 = 0, pc1 = 0, my_pslot = 0, my_pidx = 0;
        bool new_tile = true;
        unsigned parC = 0;
        for (int item = i0; item < i1; ++item) {
            const int k = item - i0;
            const VT* vals = reinterpret_cast<const VT*>(smem_raw + L.vals + (size_t)(k & 1) * L.vals_bytes);
            const bool is_sum = (j == 0 ? first_op : OP_MAX) == OP_SUM;
            VT* out = is_sum ? A.out_sum : A.out_max;
            VT* part = is_sum ? A.part_sum : A.part_max;
            const int n0 = nb3[0], n1 = nb3[1];
            if (new_tile) {
                // this thread's spanning-node piece of the tile, requested long before its first use
                pc0 = __ldg(P.piece_ptr + t); pc1 = __ldg(P.piece_ptr + t + 1);
                if (pc0 + tid < pc1) { my_pslot = __ldg(P.piece_slot + pc0 + tid); my_pidx = __ldg(P.piece_idx + pc0 + tid); }
                mbar_wait(barC, parC); parC ^= 1u;
            }
            int jn = j + 1, tn = t, gn = g;
            if (jn == nops) { jn = 0; if (++gn == RG) { gn = 0; ++tn; } }
            const bool next_new = item + 1 < i1 && tn != t;

            const int b0 = g * R;
            VT* orow[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                orow[r] = out + (size_t)min(b0 + r, n_rows - 1) * A.ld_out;
                asm volatile("" : "+l"(orow[r]));  // keep the row pointers in registers (no rematerialisation per store)
            }
            GT_TRACE(8);
            mbar_wait(full + (k & 1), (unsigned)(k >> 1) & 1u);  // the compute group has filled this value array
            GT_TRACE(9);

            // 4. emit the tile's node-id interval: lane = consecutive node id, so the slot reads of a warp cluster on
            //    a few neighbouring slots (unary chains broadcast) and every store instruction writes 128 contiguous
            //    bytes per row.  The sweep starts at the 128-byte line of row 0 that holds node n0, so with a row
            //    stride that is a multiple of 32 elements every store instruction covers exactly one line.
            //    Spanning nodes inside the interval carry the identity slot: what is written for them here is
            //    overwritten by span_kernel.
            if (dbg != 3) {
#if defined(GT_EMIT_V2) && defined(GT_EMIT_U)
                constexpr int U = GT_EMIT_U;  // nodes per thread and trip of the branch-free variant: 1, 2 or 4
#else
                constexpr int U = 4;
#endif
                const int na = n0 & ~7;
                const int lead = (int)(((reinterpret_cast<uintptr_t>(orow[0]) / sizeof(VT)) + (unsigned)n0) & 31u);
                const unsigned count = (unsigned)(n1 - n0);
                const uint16_t* sl_base = s_slots - na;
                const unsigned char* vbytes = reinterpret_cast<const unsigned char*>(vals);
#ifdef GT_EMIT_V2
                // Branch-free variant, 69 instead of 196 instructions per trip -- and measured SLOWER on B200 (tile
                // kernel 53.7 vs 48.4 us): the stores then leave in bursts, fill the load/store queue that the compute
                // group's shared-memory instructions share, and each compute phase stretches.  Kept for reference.
                for (int nb = n0 - lead + tid; nb < n1; nb += U * kEmitThreads) {
                    // slot reads are unconditional (node ids outside the interval are clamped into it), stores are
                    // predicated: no branch in the loop body
                    RV x[U];
                    bool ok[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int n = nb + u * kEmitThreads;
                        ok[u] = (unsigned)(n - n0) < count;
                        const int nc = min(max(n, n0), n1 - 1);
                        x[u] = RV::load(reinterpret_cast<const VT*>(vbytes + (unsigned)sl_base[nc] * (unsigned)B));
                    }
                    VT* p[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) p[r] = orow[r] + nb;
                    EmitStore<0 * kEmitThreads * (int)sizeof(VT), VT, R>::run(p, x[0], ok[0]);
                    if constexpr (U > 1) EmitStore<1 * kEmitThreads * (int)sizeof(VT), VT, R>::run(p, x[1 % U], ok[1 % U]);
                    if constexpr (U > 2) EmitStore<2 * kEmitThreads * (int)sizeof(VT), VT, R>::run(p, x[2 % U], ok[2 % U]);
                    if constexpr (U > 2) EmitStore<3 * kEmitThreads * (int)sizeof(VT), VT, R>::run(p, x[3 % U], ok[3 % U]);
                }
#else
#ifndef GT_EMIT_STRIDED
                // a warp's U node groups are consecutive and its stores go row by row: U * 128 contiguous bytes per row
                // and warp trip (measured 0.15 us better than groups kEmitThreads nodes apart, stored node by node)
                constexpr int kStep = 32;
                const int nb0 = n0 - lead + (tid >> 5) * (U * 32) + (tid & 31);
#else
                constexpr int kStep = kEmitThreads;
                const int nb0 = n0 - lead + tid;
#endif
                for (int nb = nb0; nb < n1; nb += U * kEmitThreads) {
                    VT* p[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) p[r] = orow[r] + nb;
                    RV x[U];
                    bool ok[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int n = nb + u * kStep;
                        ok[u] = (unsigned)(n - n0) < count;
                        if (ok[u]) x[u] = RV::load(reinterpret_cast<const VT*>(vbytes + (unsigned)sl_base[n] * (unsigned)B));
                    }
#ifndef GT_EMIT_STRIDED
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int u = 0; u < U; ++u)
                            if (ok[u]) __stcs(p[r] + u * kStep, x[u].v[r]);
#else
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (ok[u]) {
#pragma unroll
                            for (int r = 0; r < R; ++r) __stcs(p[r] + u * kStep, x[u].v[r]);
                        }
#endif
                }
#endif
                // 5. pieces of spanning nodes that overlap this tile (reduced by span_kernel, which runs next on the stream)
                for (int i = pc0 + tid; i < pc1; i += kEmitThreads) {
                    const bool mine = i == pc0 + tid;
                    const RV x = RV::load(vals + (mine ? my_pslot : (int)__ldg(P.piece_slot + i)) * R);
                    const int idx = mine ? my_pidx : __ldg(P.piece_idx + i);
#pragma unroll
                    for (int r = 0; r < R; ++r) part[(size_t)min(b0 + r, n_rows - 1) * P.n_pieces + idx] = x.v[r];
                }
            }
            GT_TRACE(10);
            mbar_arrive(empty + (k & 1));  // this thread no longer reads the value array
            if (next_new) {
                group_sync<2, kEmitThreads>();  // everyone is done with this tile's emit slots
                if (tid == 0) fetch_slots(nb3[1], nb3[2]);
                nb3[0] = nb3[1]; nb3[1] = nb3[2];
                nb3[2] = __ldg(P.tile_node_lo + min(tn + 2, P.NT));
            }
            new_tile = next_new;
            t = tn; g = gn; j = jn;
        }
    }
}

// ---- phase 3: nodes whose leaf range crosses tiles, reduced from their per-tile pieces (fp64 for sums) ------------
// One thread per (spanning node, row); consecutive lanes take consecutive spanning nodes of one row, whose pieces
// are adjacent in `part`.  Most spanning nodes have two or three pieces, which a thread loads all at once; the few
// with many (the root has one per tile) are reduced by the whole warp, 32 pieces per step, so that no thread walks a
// long chain of dependent loads.  Runs after tile_kernel in stream order and overwrites the placeholder it emitted.
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
__global__ void __launch_bounds__(256) span_kernel(PlanView P, TileArgs<VT> A) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool live = k < P.n_span;
    // plan metadata first: these loads may run ahead of the tile kernel's completion
    const int q0 = live ? __ldg(P.span_pp + k) : 0, q1 = live ? __ldg(P.span_pp + k + 1) : 0;
    const int node = live ? __ldg(P.span_node + k) : 0;
    pdl_wait();  // pieces and placeholders come from tile_kernel
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
    const size_t o_p1_chunk_ptr = ADDV(P.p1_chunk_ptr), o_p1_rec = ADDV(P.p1_rec);
    const size_t o_z_tile_off = ADDV(P.z_tile_off), o_p2_slot = ADDV(P.p2_slot);
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
    v.p1_chunk_ptr = (const int32_t*)(base + o_p1_chunk_ptr); v.p1_rec = (const int4*)(base + o_p1_rec);
    v.z_tile_off = (const int32_t*)(base + o_z_tile_off); v.p2_slot = (const uint16_t*)(base + o_p2_slot);
    v.ell_chunk_ptr = (const int32_t*)(base + o_ell_chunk_ptr); v.ell_desc = (const int2*)(base + o_ell_desc);
    v.ell_terms = (const uint16_t*)(base + o_ell_terms); v.ell_row_ptr = (const int32_t*)(base + o_ell_row_ptr);
    v.R = P.R; v.max_tile_nodes = P.max_tile_nodes; v.max_tile_ell_rows = P.max_tile_ell_rows;
    v.max_tile_chunks = P.max_tile_chunks; v.max_tile_z = P.max_tile_z;
    v.max_seg_recs = 0;
    for (size_t i = 0; i + 1 < P.p1_chunk_ptr.size(); ++i) v.max_seg_recs = std::max(v.max_seg_recs, P.p1_chunk_ptr[i + 1] - P.p1_chunk_ptr[i]);
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


template <typename VT, int R> static size_t permute_smem(const PlanView& v) { return (size_t)R * (v.Q + kSegPad) * sizeof(VT); }
template <typename VT, int R> static size_t tile_smem(const PlanView& v) { return TileSmem(v, (int)sizeof(VT) * R, (int)sizeof(VT), R).total; }

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

// Scratch layout for one chunk of `rows` rows: z [rows][Zrow] VT | part_sum [rows][n_pieces] VT | part_max likewise
template <typename VT> struct Scratch {
    VT* z; VT* part_sum; VT* part_max;
    static size_t pad(size_t b) { return (b + 255) & ~size_t(255); }
    Scratch(const PlanView& v, void* base, int64_t rows) {
        char* p = static_cast<char*>(base);
        z = reinterpret_cast<VT*>(p);
        p += pad((size_t)rows * v.Zrow * sizeof(VT));
        part_sum = reinterpret_cast<VT*>(p);
        p += pad((size_t)rows * v.n_pieces * sizeof(VT));
        part_max = reinterpret_cast<VT*>(p);
    }
    static size_t total(const PlanView& v, int64_t rows) {
        return pad((size_t)rows * v.Zrow * sizeof(VT)) + 2 * pad((size_t)rows * v.n_pieces * sizeof(VT));
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

template <typename IN_T, int RP, int NST> static size_t permute_bulk_smem(const PlanView& v) {
    return (size_t)NST * RP * (v.Q * sizeof(IN_T) + kPermPad) + (size_t)v.max_seg_recs * 16 + 8 * (NST + 1);
}
// the copy engine moves whole 16-byte units: elements must be naturally aligned and a segment a whole number of units
template <typename IN_T> static bool permute_bulk_ok(const PlanView& v, const void* ws, int64_t ld_ws) {
    (void)ld_ws;
    return (reinterpret_cast<uintptr_t>(ws) % sizeof(IN_T)) == 0 && (((int64_t)v.Q * (int64_t)sizeof(IN_T)) & 15) == 0;
}
template <typename VT, typename IN_T, int RP, int NST, bool LOG, bool ALIGNED>
static int launch_permute_bulk_la(const PlanView& v, const void* ws, int64_t ld_ws, const Scratch<VT>& sc, int rows, cudaStream_t st) {
    const size_t smem = permute_bulk_smem<IN_T, RP, NST>(v);
    GT_CUDA(allow_smem(permute_bulk_kernel<VT, IN_T, RP, NST, LOG, ALIGNED>, smem));
    // one CTA per resident slot; each takes an equal share of the (segment, row) pairs, at least RP of them
    const int64_t groups = ((int64_t)v.NS * rows + RP - 1) / RP;
    const int slots = sm_count() * resident_ctas(reinterpret_cast<const void*>(permute_bulk_kernel<VT, IN_T, RP, NST, LOG, ALIGNED>), kThreads, smem);
    const unsigned grid = (unsigned)std::min<int64_t>(groups, slots);
    GT_CUDA(launch_pdl(permute_bulk_kernel<VT, IN_T, RP, NST, LOG, ALIGNED>, dim3(grid), dim3(kThreads), smem, st, v,
                       static_cast<const IN_T*>(ws), ld_ws, sc.z, rows));
    return GT_OK;
}
template <typename VT, typename IN_T, int RP, int NST>
static int launch_permute_bulk(const PlanView& v, const void* ws, int64_t ld_ws, const Scratch<VT>& sc, int rows,
                               bool log_input, cudaStream_t st) {
    const bool aligned = (reinterpret_cast<uintptr_t>(ws) & 15) == 0 && ((ld_ws * (int64_t)sizeof(IN_T)) & 15) == 0 &&
                         ((v.V * (int64_t)sizeof(IN_T)) & 15) == 0;
    if (aligned)
        return log_input ? launch_permute_bulk_la<VT, IN_T, RP, NST, true, true>(v, ws, ld_ws, sc, rows, st)
                         : launch_permute_bulk_la<VT, IN_T, RP, NST, false, true>(v, ws, ld_ws, sc, rows, st);
    return log_input ? launch_permute_bulk_la<VT, IN_T, RP, NST, true, false>(v, ws, ld_ws, sc, rows, st)
                     : launch_permute_bulk_la<VT, IN_T, RP, NST, false, false>(v, ws, ld_ws, sc, rows, st);
}

template <typename VT, typename IN_T, int R>
static int launch_permute(const PlanView& v, const void* ws, int64_t ld_ws, const Scratch<VT>& sc, int rows,
                          bool log_input, cudaStream_t st) {
    const size_t smem = permute_smem<VT, R>(v);
    GT_CUDA(allow_smem(permute_kernel<VT, IN_T, R>, smem));
    dim3 grid((unsigned)v.NS, (unsigned)((rows + R - 1) / R));
    GT_CUDA(launch_pdl(permute_kernel<VT, IN_T, R>, grid, dim3(kThreads), smem, st, v, static_cast<const IN_T*>(ws), ld_ws,
                       sc.z, rows, log_input ? 1 : 0));
    return GT_OK;
}

template <typename VT, int R>
static int launch_tile(const PlanView& v, const Scratch<VT>& sc, VT* out_sum, VT* out_max, int64_t ld_out, int rows,
                       unsigned ops, unsigned phases, cudaStream_t st) {
    if (v.NT == 0) {  // empty vocabulary: the root is the only node and has no mass
        for (VT* out : {out_sum, out_max})
            if (out) GT_CUDA(cudaMemset2DAsync(out, (size_t)ld_out * sizeof(VT), 0, (size_t)v.N * sizeof(VT), (size_t)rows, st));
        return GT_OK;
    }
    const size_t smem = tile_smem<VT, R>(v);
    if (smem > 227 * 1024) {
        set_error("tile plan needs %zu bytes of shared memory per CTA (limit 232448): use a smaller tile or fewer rows per CTA", smem);
        return GT_ERR_LIMIT;
    }
    GT_CUDA(allow_smem(tile_kernel<VT, R>, smem));
    TileArgs<VT> A;
    A.z = sc.z; A.out_sum = out_sum; A.out_max = out_max; A.part_sum = sc.part_sum; A.part_max = sc.part_max;
    A.ld_out = ld_out; A.n_rows = rows; A.ops = ops;
    const int nops = (ops == (unsigned)(GT_OP_SUM | GT_OP_MAX)) ? 2 : 1;
    // persistent grid: one CTA per resident slot, each takes a contiguous run of (tile, row group, reduction) items
    const int64_t items = (int64_t)v.NT * ((rows + R - 1) / R) * nops;
    const int slots = sm_count() * resident_ctas(reinterpret_cast<const void*>(tile_kernel<VT, R>), kTileThreads, smem);
    const unsigned grid = (unsigned)std::min<int64_t>(items, slots);
    if (phases & GT_FLAG_PHASE_TILE) GT_CUDA(launch_pdl(tile_kernel<VT, R>, dim3(grid), dim3(kTileThreads), smem, st, v, A));
    if (v.n_span > 0 && (phases & GT_FLAG_PHASE_SPAN)) {
        dim3 sgrid((unsigned)((v.n_span + 255) / 256), (unsigned)std::min(rows, 4096), (unsigned)nops);
        GT_CUDA(launch_pdl(span_kernel<VT>, sgrid, dim3(256), 0, st, v, A));
    }
    return GT_OK;
}

template <typename VT, int R>
static int reduce_typed(const PlanView& v, const void* ws, int in_type, int64_t n_rows, int64_t ld_ws, void* out_sum,
                        void* out_max, int64_t ld_out, unsigned ops, unsigned flags, void* workspace,
                        size_t workspace_bytes, cudaStream_t st) {
    // rows per chunk: what the caller's scratch can stage, rounded down to whole row groups when it holds at least
    // one (a partial row group is legal: the kernels alias the missing rows to the last valid one)
    const size_t per_row = (size_t)(v.Zrow + 2 * v.n_pieces) * sizeof(VT);
    int64_t chunk = std::min<int64_t>(n_rows, 32768);
    if (Scratch<VT>::total(v, chunk) > workspace_bytes) {
        chunk = std::min<int64_t>(chunk, (int64_t)(workspace_bytes / std::max<size_t>(per_row, 1)));
        while (chunk > 0 && Scratch<VT>::total(v, chunk) > workspace_bytes) --chunk;
        if (chunk >= R) chunk = (chunk / R) * R;
    }
    if (chunk < 1) {
        set_error("workspace too small: %zu bytes given, one row needs %zu", workspace_bytes, Scratch<VT>::total(v, 1));
        return GT_ERR_STATE;
    }
    const bool log_input = (flags & GT_FLAG_LOG_INPUT) != 0;
    const unsigned phases = (flags & GT_FLAG_PHASE_MASK) ? (flags & GT_FLAG_PHASE_MASK) : GT_FLAG_PHASE_MASK;
    const size_t in_size = in_type == GT_F64 ? 8 : in_type == GT_F32 ? 4 : 2;
    for (int64_t r0 = 0; r0 < n_rows; r0 += chunk) {
        const int rows = (int)std::min<int64_t>(chunk, n_rows - r0);
        const Scratch<VT> sc(v, workspace, rows);
        const void* wsr = static_cast<const char*>(ws) + (size_t)r0 * ld_ws * in_size;
        int rc = GT_OK;
        if (v.NT > 0 && (phases & GT_FLAG_PHASE_PERMUTE)) {
            // rows per CTA in the permute phase: R unless the segment buffer would not fit in shared memory
            constexpr int RP = sizeof(VT) == 4 ? 2 : 1;  // rows per CTA of the permute kernel
            const bool wide = permute_smem<VT, RP>(v) <= kMaxSmem;
            static const bool use_bulk = getenv("GT_NO_BULK_PERMUTE") == nullptr;  // read once: not on the launch path
#ifndef GT_PB_RP
#define GT_PB_RP 2
#endif
#ifndef GT_PB_NST
#define GT_PB_NST 2
#endif
#define GT_PERMUTE(IN_T) ((use_bulk && permute_bulk_ok<IN_T>(v, wsr, ld_ws) && permute_bulk_smem<IN_T, GT_PB_RP, GT_PB_NST>(v) <= 113 * 1024) \
                              ? launch_permute_bulk<VT, IN_T, GT_PB_RP, GT_PB_NST>(v, wsr, ld_ws, sc, rows, log_input, st)      \
                          : wide ? launch_permute<VT, IN_T, RP>(v, wsr, ld_ws, sc, rows, log_input, st)                     \
                                 : launch_permute<VT, IN_T, 1>(v, wsr, ld_ws, sc, rows, log_input, st))
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
        if (phases & (GT_FLAG_PHASE_TILE | GT_FLAG_PHASE_SPAN)) {
            rc = launch_tile<VT, R>(v, sc, (ops & GT_OP_SUM) ? static_cast<VT*>(out_sum) + (size_t)r0 * ld_out : nullptr,
                                    (ops & GT_OP_MAX) ? static_cast<VT*>(out_max) + (size_t)r0 * ld_out : nullptr, ld_out, rows,
                                    ops, phases, st);
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
    if (t->dev.count(device)) return GT_OK;
    if (!t->plan) {
        const int rc = gt_plan(t, 0, 0, 0);
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
    info->rows_per_item = P.R;
    info->n_span = (int32_t)P.span_node.size(); info->span_terms = (int64_t)P.n_pieces;
    info->max_levels = P.max_levels; info->max_tile_values = P.max_tile_values;
    info->staged_row_elems = P.Zrow;
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
    v.Zrow = t->plan->Zrow; v.n_pieces = t->plan->n_pieces;
    return gt::Scratch<double>::total(v, max_rows) + 256;
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
    if (!workspace) { gt::set_error("gt_weight_reduce: null workspace"); return GT_ERR_ARG; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define GT_REDUCE(VT, R) gt::reduce_typed<VT, R>(v, ws, in_type, n_rows, ld_ws, out_sum, out_max, ld_out, ops, flags, \
                                                workspace, workspace_bytes, st)
    // the fp64 pipeline runs half as many rows per CTA, so both pipelines share the plan's slot size
    if (out_type == GT_F32) return v.R == 4 ? GT_REDUCE(float, 4) : GT_REDUCE(float, 2);
    if (out_type == GT_F64) return v.R == 4 ? GT_REDUCE(double, 2) : GT_REDUCE(double, 1);
#undef GT_REDUCE
    gt::set_error("gt_weight_reduce: output type must be GT_F32 or GT_F64");
    return GT_ERR_ARG;
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
