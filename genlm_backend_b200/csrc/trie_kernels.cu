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
    int32_t R, max_tile_nodes, max_tile_ell_rows, max_tile_chunks, max_tile_z;
    int64_t V, N, Zrow;
    const int32_t* p1_chunk_ptr; const int4* p1_rec;
    const int32_t* z_tile_off; const uint16_t* p2_slot;
    const int32_t* ell_chunk_ptr; const int2* ell_desc; const uint16_t* ell_terms; const int32_t* ell_row_ptr;
    const int32_t* tile_node_lo; const uint16_t* node_slot;
    const int32_t* piece_ptr; const uint16_t* piece_slot; const int32_t* piece_idx;
    int32_t n_span, n_pieces; const int32_t* span_node; const int32_t* span_pp;
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
// Persistent kernel.  A work item is (tile t, row group g of R rows); items are numbered tile-major and every CTA
// takes one contiguous run of them, so consecutive items of a CTA share the tile metadata, which stays in shared
// memory.  All bulk traffic to and from global memory goes through the copy engine (cp.async.bulk, TMA), issued by
// one thread, so the LSU pipes only ever see shared-memory work and its latency stays low:
//     in   barrier A  staged rows of the next item (+ the slot table of its tile if new)   armed after each scatter
//          barrier B  ELL term rows + descriptors of the next tile                         armed after the last ELL of a tile
//          barrier C  emit slots of the next tile                                          armed after the last emit of a tile
//     out  the tile's node-id interval is gathered into row-major staging chunks (two buffers) and written by bulk
//          stores (bulk async-groups); only the few nodes outside the 16-byte aligned core use plain stores.
// Every wait on A/B/C precedes a CTA barrier and every re-arm follows it, so no thread can still be waiting on a
// phase when the next one completes.

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
    do {  // try_wait suspends the thread in hardware for a bounded time; loop until the phase has completed
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
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

constexpr int kTileThreads = 512;
constexpr int kOutStageBytes = 16384;  // one output staging buffer (R rows x C nodes)
// trace layout: [kTraceCtas][kTraceItems][kTraceEvents] SM-clock stamps
constexpr int kTraceCtas = 512, kTraceItems = 32, kTraceEvents = 24;
#define GT_TRACE(ev)                                                                                         \
    do {                                                                                                     \
        if (P.trace && tid == 0 && blockIdx.x < kTraceCtas && k < kTraceItems)                                \
            P.trace[((size_t)blockIdx.x * kTraceItems + k) * kTraceEvents + (ev)] = clock64();               \
    } while (0)

// Shared-memory carve-up of tile_kernel (all sections 16-byte aligned).
struct TileSmem {
    size_t vals, stage, p2, slots, terms, desc, ostage, bars, total;
    __host__ __device__ TileSmem(const PlanView& P, int slot_bytes, int elem_bytes, int R) {
        size_t o = 0;
        vals = o;  o += (size_t)(P.SV + 4) * slot_bytes;                    // value slots + trash slot
        stage = o; o += (size_t)R * P.max_tile_z * elem_bytes;              // staged rows of the next item
        p2 = o;    o += (size_t)P.max_tile_z * 2;                           // staged element -> value slot
        slots = o; o += (size_t)P.max_tile_nodes * 2;                       // node -> value slot (emit)
        terms = o; o += (size_t)P.max_tile_ell_rows * 64;                   // ELL term rows
        desc = o;  o += ((size_t)(P.max_tile_chunks + 2) * 8 + 15) & ~size_t(15); // ELL chunk descriptors (+ alignment slack)
        ostage = o; o += 2 * (size_t)kOutStageBytes;                        // output staging, two buffers
        bars = o;  o += 32;                                                 // three mbarriers
        total = o;
    }
};

// bulk store shared -> global (bulk async-group completion); bytes multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores of this thread have finished READING shared memory (the buffers may be rewritten)
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make this thread's shared-memory writes visible to the copy engine (async proxy)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename VT, int R, int OP>
__global__ void __launch_bounds__(kTileThreads, 2) tile_kernel(PlanView P, const VT* __restrict__ z, VT* __restrict__ out,
                                                               int64_t ld_out, VT* __restrict__ part, int n_rows) {
    using RV = RowVec<VT, R>;
    constexpr int B = (int)sizeof(VT) * R;  // bytes per slot
    static_assert(B == 4 || B == 8 || B == 16, "slot must be 4, 8 or 16 bytes");
    constexpr int SPC = 16 / B;             // slots per 16-byte chunk
    constexpr int kWarps = kTileThreads / 32;
    constexpr int C = kOutStageBytes / B;   // nodes per output staging buffer
    constexpr int AL = 16 / (int)sizeof(VT);  // nodes per 16 bytes of an output row
    constexpr int kIssueTid = kTileThreads - 32;          // lane 0 of the last warp issues the input fetches
    constexpr int kStoreWarp0 = kWarps - R;               // lane 0 of warps kStoreWarp0 + r issues the stores of row r
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileSmem L(P, B, (int)sizeof(VT), R);
    VT* vals = reinterpret_cast<VT*>(smem_raw + L.vals);
    VT* stage = reinterpret_cast<VT*>(smem_raw + L.stage);
    uint16_t* s_p2 = reinterpret_cast<uint16_t*>(smem_raw + L.p2);
    uint16_t* s_slots = reinterpret_cast<uint16_t*>(smem_raw + L.slots);
    uint16_t* s_terms = reinterpret_cast<uint16_t*>(smem_raw + L.terms);
    int2* s_desc = reinterpret_cast<int2*>(smem_raw + L.desc);
    VT* ostage = reinterpret_cast<VT*>(smem_raw + L.ostage);  // [2][R][C]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bars);
    uint64_t* barA = bars;      // rows (+ p2)
    uint64_t* barB = bars + 1;  // terms + descriptors
    uint64_t* barC = bars + 2;  // emit slots

    const int T = P.T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int RG = (n_rows + R - 1) / R;
    const int n_items = P.NT * RG;
    const int i0 = (int)((int64_t)blockIdx.x * n_items / gridDim.x);
    const int i1 = (int)((int64_t)(blockIdx.x + 1) * n_items / gridDim.x);
    if (i0 >= i1) return;
    const int zpitch = P.max_tile_z;
    const int dbg = P.debug_stop;  // profiling aid: 3 = no output, 9 = output only (compute phases skipped)
    // bulk stores need 16-byte aligned rows; otherwise every node takes the plain-store path
    const bool bulk_ok = ((ld_out * (int64_t)sizeof(VT)) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    auto level_slot = [&](int k, int i) { return 2 * T - (T >> (k - 1)) + i; };  // level k >= 1, block i

    // Boundaries of the current tile (index 0..1) and the next one (1..2) in the four per-tile prefix arrays; the
    // next tile's are loaded one tile ahead so that no fetch waits for them.
    int zb[3], erb[3], ecb[3], nb3[3];
    auto load_bound = [&](int t, int j) {
        const int tt = min(t, P.NT);
        zb[j] = __ldg(P.z_tile_off + tt); erb[j] = __ldg(P.ell_row_ptr + tt);
        ecb[j] = __ldg(P.ell_chunk_ptr + tt); nb3[j] = __ldg(P.tile_node_lo + tt);
    };

    // ---- asynchronous fetches (thread 0) -----------------------------------------------------------------------
    // Rows past the end of the batch alias the last valid row: they compute exactly what that row does, which keeps
    // every loop free of row predicates; their output is not stored.
    auto fetch_rows = [&](int g, int zlo, int zn, bool with_p2) {
        const unsigned row_bytes = (unsigned)zn * (unsigned)sizeof(VT);
        mbar_expect_tx(barA, R * row_bytes + (with_p2 ? (unsigned)zn * 2u : 0u));
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = min(g * R + r, n_rows - 1);
            bulk_g2s(stage + (size_t)r * zpitch, z + (size_t)row * P.Zrow + zlo, row_bytes, barA);
        }
        if (with_p2) bulk_g2s(s_p2, P.p2_slot + zlo, (unsigned)zn * 2u, barA);
    };
    // ELL term rows and chunk descriptors (8 bytes each, copied from the 16-byte aligned pair at or below the first)
    auto fetch_terms = [&](int er0, int er1, int ec0, int ec1) {
        const int ea = ec0 & ~1;
        const unsigned tb = (unsigned)(er1 - er0) * 64u, db = (unsigned)((ec1 - ea + 1) >> 1) * 16u;
        mbar_expect_tx(barB, tb + db);
        if (tb) bulk_g2s(s_terms, P.ell_terms + (size_t)er0 * 32, tb, barB);
        if (db) bulk_g2s(s_desc, P.ell_desc + ea, db, barB);
    };
    // emit slots, staged from the 16-byte aligned start at or below the tile's first node
    auto fetch_slots = [&](int n0, int n1) {
        const int na = n0 & ~7;
        const unsigned sb = (unsigned)((n1 - na + 7) >> 3) * 16u;
        mbar_expect_tx(barC, sb);
        bulk_g2s(s_slots, P.node_slot + na, sb, barC);
    };

    int t = i0 / RG, g = i0 - t * RG;
    load_bound(t, 0); load_bound(t + 1, 1); load_bound(t + 2, 2);
    if (tid == 0) {
        mbar_init(barA, 1); mbar_init(barB, 1); mbar_init(barC, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fetch_rows(g, zb[0], zb[1] - zb[0], true);
        fetch_terms(erb[0], erb[1], ecb[0], ecb[1]);
        fetch_slots(nb3[0], nb3[1]);
    }
    __syncthreads();  // barriers initialised before anyone waits on them

    bool new_tile = true;
    unsigned parBC = 0;
    int pc0 = 0, pc1 = 0, my_pslot = 0, my_pidx = 0;
    for (int item = i0; item < i1; ++item) {
        const int k = item - i0;
        if (new_tile) {  // this thread's spanning-node piece of the tile, requested a whole item before its first use
            pc0 = __ldg(P.piece_ptr + t); pc1 = __ldg(P.piece_ptr + t + 1);
            if (pc0 + tid < pc1) { my_pslot = __ldg(P.piece_slot + pc0 + tid); my_pidx = __ldg(P.piece_idx + pc0 + tid); }
        }
        int tn = t, gn = g + 1;
        if (gn == RG) { gn = 0; ++tn; }
        const bool has_next = item + 1 < i1;
        const bool next_new = has_next && tn != t;
        const int n0 = nb3[0], n1 = nb3[1];

        // 1. staged rows -> DFS-ordered leaf slots (slot numbers in the table are already swizzled)
        GT_TRACE(0);
        mbar_wait(barA, (unsigned)k & 1u);
        GT_TRACE(1);
        if (dbg != 9) {
            const int zn4 = (zb[1] - zb[0]) >> 2;
            for (int q = tid; q < zn4; q += kTileThreads) {
                const uint2 sl = *reinterpret_cast<const uint2*>(s_p2 + 4 * q);
                VT v[R][4];
#pragma unroll
                for (int r = 0; r < R; ++r) load4<VT>(stage + (size_t)r * zpitch + 4 * q, v[r][0], v[r][1], v[r][2], v[r][3]);
                const unsigned s4[4] = {sl.x & 0xFFFFu, sl.x >> 16, sl.y & 0xFFFFu, sl.y >> 16};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    RV x;
#pragma unroll
                    for (int r = 0; r < R; ++r) x.v[r] = v[r][e];
                    x.store(vals + s4[e] * R);
                }
            }
            const int nleaf = (int)min((int64_t)T, P.V - (int64_t)t * T);
            for (int i = nleaf + tid; i < T; i += kTileThreads) RV::template ident<OP>().store(vals + swz<B>(i) * R);
            if (tid == 0) RV::template ident<OP>().store(vals + swz<B>(2 * T - 1) * R);  // identity slot (ELL padding, spanning nodes)
        }
        __syncthreads();
        GT_TRACE(2);
        // the staging buffer (and, on a tile change, the slot table) is free again: fetch the next item's rows
        // (issued by a warp that has no share of the pyramid, so nobody waits for the issue latency)
        if (tid == kIssueTid) {
            if (has_next) fetch_rows(gn, next_new ? zb[1] : zb[0], next_new ? zb[2] - zb[1] : zb[1] - zb[0], next_new);
            else mbar_expect_tx(barA, 0);
        }
        GT_TRACE(3);

        // 2. pyramid of aligned blocks: level k block i at (swizzled) slot 2T - (T >> (k-1)) + i.
        //    Lane u owns leaves 8u .. 8u+7: levels 1..3 in registers, 4..8 by warp shuffles (256 leaves per warp).
        //    The swizzle makes the 16-byte chunk loads and the strided level stores bank-conflict free.
        if (dbg != 9) {
            for (int ub = warp * 32; ub < (T >> 3); ub += kWarps * 32) {
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
        }
        if (new_tile) mbar_wait(barB, parBC);  // ELL terms + descriptors of this tile
        __syncthreads();
        GT_TRACE(4);

        // 3. ranges that need more than one block: ELL chunks of 32 ranges, one warp per chunk, term rows read from
        //    shared memory (k is a multiple of 4: the planner pads rows with the identity slot)
        if (dbg != 9) {
            const int2* dsc = s_desc + (ecb[0] & 1);
            const int nchunks = ecb[1] - ecb[0];
            for (int c = warp; c < nchunks; c += kWarps) {
                const int2 d = dsc[c];
                const uint16_t* tp = s_terms + (d.x - erb[0]) * 32 + lane;
                RV acc = RV::template ident<OP>();
                for (int kb = 0; kb < d.y; kb += 4) {
                    int sl[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) sl[e] = tp[(kb + e) * 32];
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc = RV::template combine<OP>(acc, RV::load(vals + sl[e] * R));
                }
                acc.store(vals + (2 * T + c * 32 + lane) * R);
            }
        }
        if (new_tile) { mbar_wait(barC, parBC); parBC ^= 1u; }  // emit slots of this tile
        GT_TRACE(22);
        // the previous item's output chunks have left both staging buffers (each issuing thread checks its own stores)
        if (lane == 0 && warp >= kStoreWarp0) bulk_wait_read_all();
        __syncthreads();
        GT_TRACE(5);
        if (tid == 0 && next_new) fetch_terms(erb[1], erb[2], ecb[1], ecb[2]);

        // 4. output.  Core = the 16-byte aligned part [n0a, n1a) of the tile's node-id interval: gathered into
        //    row-major staging chunks (lane = consecutive node id, so the slot reads of a warp cluster on a few
        //    neighbouring slots -- unary chains broadcast -- and the staging writes are conflict-free), then
        //    written by one bulk store per row and chunk.  Spanning nodes inside the interval carry the identity
        //    slot: what is written for them here is overwritten by span_kernel.
        if (dbg != 3) {
            const int b0 = g * R;
            const int nrows_here = min(R, n_rows - b0);
            const uint16_t* sl_base = s_slots - (n0 & ~7);
            const int n0a = bulk_ok ? min(n1, (n0 + AL - 1) & ~(AL - 1)) : n1;
            const int n1a = bulk_ok ? max(n0a, n1 & ~(AL - 1)) : n1;
            // nodes outside the aligned core, and the pieces of spanning nodes that overlap this tile (reduced by
            // span_kernel, which runs next on the stream): plain stores
            auto store_node = [&](int n) {
                const RV x = RV::load(vals + (int)sl_base[n] * R);
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (r < nrows_here) out[(size_t)(b0 + r) * ld_out + n] = x.v[r];
            };
            for (int n = n0 + tid; n < n0a; n += kTileThreads) store_node(n);  // head (everything when !bulk_ok)
            for (int n = n1a + tid; n < n1; n += kTileThreads) store_node(n);  // tail
            for (int i = pc0 + tid; i < pc1; i += kTileThreads) {
                const bool mine = i == pc0 + tid;
                const RV x = RV::load(vals + (mine ? my_pslot : (int)__ldg(P.piece_slot + i)) * R);
                const int idx = mine ? my_pidx : __ldg(P.piece_idx + i);
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (r < nrows_here) part[(size_t)(b0 + r) * P.n_pieces + idx] = x.v[r];
            }
            GT_TRACE(6);
            int buf = 0, ci = 0;
            for (int cs = n0a; cs < n1a; cs += C, buf ^= 1, ++ci) {
                const int cn = min(C, n1a - cs);
                VT* ob = ostage + (size_t)buf * R * C;
                for (int i = tid; i < cn; i += kTileThreads) {
                    const RV x = RV::load(vals + (int)sl_base[cs + i] * R);
#pragma unroll
                    for (int r = 0; r < R; ++r) ob[r * C + i] = x.v[r];
                }
                fence_async_smem();
                if (ci < 4) GT_TRACE(7 + 4 * ci);
                // the chunk after this one reuses the other buffer: its previous store must have been read out
                if (lane == 0 && warp >= kStoreWarp0 && cs + C < n1a) bulk_wait_read_all();
                if (ci < 4) GT_TRACE(8 + 4 * ci);
                __syncthreads();
                if (ci < 4) GT_TRACE(9 + 4 * ci);
                if (lane == 0 && warp >= kStoreWarp0) {  // one row per issuing thread
                    const int r = warp - kStoreWarp0;
                    if (r < nrows_here)
                        bulk_s2g(out + (size_t)(b0 + r) * ld_out + cs, ob + r * C, (unsigned)cn * (unsigned)sizeof(VT));
                    bulk_commit();
                }
                if (ci < 4) GT_TRACE(10 + 4 * ci);
            }
            if (n1a <= n0a) __syncthreads();  // no chunk barrier ran: still separate this item's reads from the next scatter
        } else {
            __syncthreads();
        }
        GT_TRACE(23);
        if (tid == 0 && next_new) fetch_slots(nb3[1], nb3[2]);
        if (next_new) {
#pragma unroll
            for (int j = 0; j < 2; ++j) { zb[j] = zb[j + 1]; erb[j] = erb[j + 1]; ecb[j] = ecb[j + 1]; nb3[j] = nb3[j + 1]; }
            load_bound(tn + 1, 2);
        }
        new_tile = next_new;
        t = tn; g = gn;
    }
    if (lane == 0 && warp >= kStoreWarp0) bulk_wait_all();  // the output is complete in global memory before the CTA retires
}

// ---- phase 3: nodes whose leaf range crosses tiles, reduced from their per-tile pieces (fp64 for sums) ------------
// One thread per (spanning node, row); consecutive lanes take consecutive spanning nodes of one row, whose pieces
// are adjacent in `part`.  Runs after tile_kernel in stream order and overwrites the placeholder it emitted.
template <typename VT, int OP>
__global__ void __launch_bounds__(256) span_kernel(PlanView P, const VT* __restrict__ part, VT* __restrict__ out,
                                                   int64_t ld_out, int n_rows) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P.n_span) return;
    const int q0 = __ldg(P.span_pp + k), q1 = __ldg(P.span_pp + k + 1), node = __ldg(P.span_node + k);
    using AT = typename std::conditional<OP == OP_SUM, double, VT>::type;
    for (int b = blockIdx.y; b < n_rows; b += gridDim.y) {
        const VT* pr = part + (size_t)b * P.n_pieces;
        AT acc = op_ident<OP, AT>();
#pragma unroll 8
        for (int q = q0; q < q1; ++q) acc = op_apply<OP, AT>(acc, (AT)pr[q]);
        out[(size_t)b * ld_out + node] = q1 > q0 ? (VT)acc : VT(0);
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
    v.tile_node_lo = (const int32_t*)(base + o_tile_node_lo); v.node_slot = (const uint16_t*)(base + o_node_slot);
    v.piece_ptr = (const int32_t*)(base + o_piece_ptr); v.piece_slot = (const uint16_t*)(base + o_piece_slot);
    v.piece_idx = (const int32_t*)(base + o_piece_idx);
    v.n_span = (int32_t)P.span_node.size(); v.n_pieces = P.n_pieces;
    v.span_node = (const int32_t*)(base + o_span_node); v.span_pp = (const int32_t*)(base + o_span_pp);
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

// Scratch layout for one chunk of `rows` rows: z [rows][Zrow] VT | part [rows][n_pieces] VT
template <typename VT> struct Scratch {
    VT* z; VT* part;
    Scratch(const PlanView& v, void* base, int64_t rows) {
        char* p = static_cast<char*>(base);
        z = reinterpret_cast<VT*>(p);
        p += (((size_t)rows * v.Zrow * sizeof(VT)) + 255) & ~size_t(255);
        part = reinterpret_cast<VT*>(p);
    }
    static size_t total(const PlanView& v, int64_t rows) {
        return ((((size_t)rows * v.Zrow * sizeof(VT)) + 255) & ~size_t(255)) +
               ((((size_t)rows * v.n_pieces * sizeof(VT)) + 255) & ~size_t(255));
    }
};

template <typename VT, typename IN_T, int R>
static int launch_permute(const PlanView& v, const void* ws, int64_t ld_ws, const Scratch<VT>& sc, int rows,
                          bool log_input, cudaStream_t st) {
    const size_t smem = permute_smem<VT, R>(v);
    GT_CUDA(allow_smem(permute_kernel<VT, IN_T, R>, smem));
    dim3 grid((unsigned)v.NS, (unsigned)((rows + R - 1) / R));
    permute_kernel<VT, IN_T, R><<<grid, kThreads, smem, st>>>(v, static_cast<const IN_T*>(ws), ld_ws, sc.z, rows,
                                                             log_input ? 1 : 0);
    GT_CUDA(cudaGetLastError());
    return GT_OK;
}

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

template <typename VT, int R, int OP>
static int launch_tile(const PlanView& v, const Scratch<VT>& sc, VT* out, int64_t ld_out, int rows, cudaStream_t st) {
    if (v.NT == 0) {  // empty vocabulary: the root is the only node and has no mass
        GT_CUDA(cudaMemset2DAsync(out, (size_t)ld_out * sizeof(VT), 0, (size_t)v.N * sizeof(VT), (size_t)rows, st));
        return GT_OK;
    }
    const size_t smem = tile_smem<VT, R>(v);
    if (smem > 227 * 1024) {
        set_error("tile plan needs %zu bytes of shared memory per CTA (limit 232448): use a smaller tile or fewer rows per CTA", smem);
        return GT_ERR_LIMIT;
    }
    GT_CUDA(allow_smem(tile_kernel<VT, R, OP>, smem));
    // persistent grid: one CTA per resident slot, each takes a contiguous run of (tile, row group) items
    const int64_t items = (int64_t)v.NT * ((rows + R - 1) / R);
    const int slots = sm_count() * resident_ctas(reinterpret_cast<const void*>(tile_kernel<VT, R, OP>), kTileThreads, smem);
    const unsigned grid = (unsigned)std::min<int64_t>(items, slots);
    tile_kernel<VT, R, OP><<<grid, kTileThreads, smem, st>>>(v, sc.z, out, ld_out, sc.part, rows);
    GT_CUDA(cudaGetLastError());
    if (v.n_span > 0) {
        dim3 sgrid((unsigned)((v.n_span + 255) / 256), (unsigned)std::min(rows, 4096));
        span_kernel<VT, OP><<<sgrid, 256, 0, st>>>(v, sc.part, out, ld_out, rows);
        GT_CUDA(cudaGetLastError());
    }
    return GT_OK;
}

template <typename VT, int R>
static int reduce_typed(const PlanView& v, const void* ws, int in_type, int64_t n_rows, int64_t ld_ws, void* out_sum,
                        void* out_max, int64_t ld_out, unsigned ops, unsigned flags, void* workspace,
                        size_t workspace_bytes, cudaStream_t st) {
    // rows per chunk: what the caller's scratch can stage, rounded down to whole row groups when it holds at least
    // one (a partial row group is legal: the kernels alias the missing rows to the last valid one)
    const size_t per_row = (size_t)(v.Zrow + v.n_pieces) * sizeof(VT);
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
#define GT_PERMUTE(IN_T) (wide ? launch_permute<VT, IN_T, RP>(v, wsr, ld_ws, sc, rows, log_input, st) \
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
        if ((ops & GT_OP_SUM) && (phases & GT_FLAG_PHASE_TILE)) {
            rc = launch_tile<VT, R, OP_SUM>(v, sc, static_cast<VT*>(out_sum) + (size_t)r0 * ld_out, ld_out, rows, st);
            if (rc != GT_OK) return rc;
        }
        if ((ops & GT_OP_MAX) && (phases & GT_FLAG_PHASE_TILE)) {
            rc = launch_tile<VT, R, OP_MAX>(v, sc, static_cast<VT*>(out_max) + (size_t)r0 * ld_out, ld_out, rows, st);
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

}  // extern "C"
