// Tile plan: turns the host layout into the static metadata the kernels consume.
//
// Facts used (all consequences of the reference's post-order numbering, base.py:80-83, 236-247):
//   * leaves sorted by node id are in DFS order, so every node owns a contiguous DFS leaf range [lo,hi);
//   * node ids split into consecutive intervals, one per tile of T DFS-ordered leaves: interval t =
//     [id of first leaf of tile t, id of first leaf of tile t+1) holds exactly the nodes whose range
//     lies inside tile t plus the few "spanning" nodes whose range ends in tile t;
//   * a unary node's only child is node id-1 and has the same leaf range.
//
// Mass of a node = reduction over its leaf range.  Ranges are decomposed into aligned power-of-two
// blocks of a per-tile pyramid (level k block i = leaves [i*2^k, (i+1)*2^k) of the tile), so every
// node is a short sum of non-negative terms: no prefix differences, hence no cancellation, and no
// dependency between nodes.
#include "trie_internal.h"

#include <algorithm>
#include <numeric>
#include <unordered_map>
#include <cstring>
#include <cstdlib>

namespace gt {

static inline int ilog2(int32_t x) { int k = 0; while ((1 << (k + 1)) <= x) ++k; return k; }

int build_plan(const Layout& L, int32_t T, int32_t Q, Plan& P) {
    const int64_t V = L.V, N = L.N;
    if (T < 1024 || T > 8192 || (T & (T - 1))) { set_error("tile size must be a power of two in [1024, 8192]"); return GT_ERR_ARG; }
    // Q*8 bytes (one fp64 row segment) must fit in shared memory
    if (Q < 4 || Q > 16384 || (Q & 3)) { set_error("segment size must be a multiple of 4 in [4, 16384]"); return GT_ERR_ARG; }
    const int logT = ilog2(T);
    P.T = T; P.Q = Q;
    P.NT = (int32_t)((V + T - 1) / T);
    P.NS = (int32_t)((V + Q - 1) / Q);
    const int32_t NT = P.NT, NS = P.NS;

    // ---- staging layout: tile-major, inside a tile one run per source segment, padded to 4 -------
    // count[t][s]
    std::vector<int32_t> cnt((size_t)NT * NS, 0);
    for (int64_t r = 0; r < V; ++r) cnt[(size_t)(r / T) * NS + L.perm[(size_t)r] / Q]++;
    std::vector<int64_t> run_off((size_t)NT * NS + 1, 0);  // start of run (t,s) in a staged row
    P.z_tile_off.assign((size_t)NT + 1, 0);
    int64_t z = 0;
    for (int32_t t = 0; t < NT; ++t) {
        P.z_tile_off[t] = (int32_t)z;
        for (int32_t s = 0; s < NS; ++s) {
            run_off[(size_t)t * NS + s] = z;
            z += (cnt[(size_t)t * NS + s] + 3) & ~3;
        }
    }
    P.z_tile_off[NT] = (int32_t)z;
    P.Zrow = z;
    if (z >= (int64_t)1 << 31) { set_error("staged row too long"); return GT_ERR_LIMIT; }

    P.p2_slot.assign((size_t)z, 0xFFFF);
    std::vector<uint16_t> z_src((size_t)z, 0xFFFF);  // position inside the source segment
    {
        std::vector<int64_t> fill(run_off.begin(), run_off.end() - 1);
        for (int64_t r = 0; r < V; ++r) {  // ascending DFS rank inside each run
            const int32_t t = (int32_t)(r / T), pos = L.perm[(size_t)r], s = pos / Q;
            const int64_t at = fill[(size_t)t * NS + s]++;
            P.p2_slot[(size_t)at] = (uint16_t)(r - (int64_t)t * T);
            z_src[(size_t)at] = (uint16_t)(pos - s * Q);
        }
    }
    // phase-1 chunk lists: segment-major
    P.p1_chunk_ptr.assign((size_t)NS + 1, 0);
    P.p1_zoff.clear(); P.p1_src.clear();
    P.p1_zoff.reserve((size_t)z / 4); P.p1_src.reserve((size_t)z);
    for (int32_t s = 0; s < NS; ++s) {
        P.p1_chunk_ptr[s] = (int32_t)P.p1_zoff.size();
        for (int32_t t = 0; t < NT; ++t) {
            const int64_t a = run_off[(size_t)t * NS + s];
            const int64_t n = (cnt[(size_t)t * NS + s] + 3) & ~3;
            for (int64_t k = 0; k < n; k += 4) {
                P.p1_zoff.push_back((int32_t)(a + k));
                for (int e = 0; e < 4; ++e) P.p1_src.push_back(z_src[(size_t)(a + k + e)]);
            }
        }
    }
    P.p1_chunk_ptr[NS] = (int32_t)P.p1_zoff.size();

    // ---- node intervals per tile -------------------------------------------------------------------
    P.tile_node_lo.assign((size_t)NT + 1, (int32_t)N);
    for (int32_t t = 0; t < NT; ++t) P.tile_node_lo[t] = L.leaf_node[(size_t)L.perm[(size_t)t * T]];
    if (NT > 0 && P.tile_node_lo[0] != 0) { set_error("internal: first DFS leaf is not node 0"); return GT_ERR_STATE; }

    // ---- value slots ------------------------------------------------------------------------------
    // tile-local slots: [0,T) leaves, [T,2T) pyramid (level k>=1 block i at 2T-(T>>(k-1))+i), [2T,..) multi-term nodes
    P.node_slot.assign((size_t)N, 0xFFFF);
    P.tile_nleaf.assign((size_t)NT, 0); P.tile_nbranch.assign((size_t)NT, 0);
    for (int32_t t = 0; t < NT; ++t) P.tile_nleaf[t] = (int32_t)std::min<int64_t>(T, V - (int64_t)t * T);

    struct Multi { int32_t node; std::vector<uint16_t> terms; };
    std::vector<std::vector<Multi>> multi((size_t)NT);
    std::vector<uint8_t> spanning((size_t)N, 0);
    auto block_slot = [&](int k, int32_t i) -> uint16_t {
        return (uint16_t)(k == 0 ? i : 2 * T - (T >> (k - 1)) + i);
    };
    std::vector<int32_t> pending_multi_index((size_t)N, -1);  // node -> index in multi[t] (before sorting)
    for (int64_t n = 0; n < N; ++n) {
        const int32_t lo = L.lo[(size_t)n], hi = L.hi[(size_t)n];
        if (hi <= lo) { spanning[(size_t)n] = 1; continue; }  // empty range (root of an empty vocabulary)
        const int32_t t = lo / T;
        if ((hi - 1) / T != t) { spanning[(size_t)n] = 1; continue; }
        if (L.is_leaf[(size_t)n]) { P.node_slot[(size_t)n] = (uint16_t)(lo - t * T); continue; }
        const int32_t deg = L.child_ptr[(size_t)n + 1] - L.child_ptr[(size_t)n];
        if (deg == 1) {  // same range as its child n-1: share the slot (possibly a pending multi-term slot)
            P.node_slot[(size_t)n] = P.node_slot[(size_t)n - 1];
            pending_multi_index[(size_t)n] = pending_multi_index[(size_t)n - 1];
            continue;
        }
        int32_t a = lo - t * T;
        const int32_t b = hi - t * T;
        std::vector<uint16_t> terms;
        while (a < b) {
            int k = a == 0 ? logT : std::min(logT, __builtin_ctz((unsigned)a));
            while (a + (1 << k) > b) --k;
            terms.push_back(block_slot(k, a >> k));
            a += 1 << k;
        }
        if (terms.size() == 1) { P.node_slot[(size_t)n] = terms[0]; continue; }
        pending_multi_index[(size_t)n] = (int32_t)multi[t].size();
        multi[t].push_back(Multi{(int32_t)n, std::move(terms)});
    }

    // order each tile's multi-term nodes by descending term count (lanes of a warp then run the same
    // trip count), assign slots 2T + j
    P.br_ptr.assign((size_t)NT + 1, 0);
    P.br_child_ptr.clear(); P.br_child.clear();
    P.br_child_ptr.push_back(0);
    P.max_tile_values = 2 * T;
    std::vector<std::vector<int32_t>> rank_of((size_t)NT);  // original index -> sorted position
    for (int32_t t = 0; t < NT; ++t) {
        auto& m = multi[t];
        std::vector<int32_t> order(m.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(),
                         [&](int32_t x, int32_t y) { return m[x].terms.size() > m[y].terms.size(); });
        rank_of[t].assign(m.size(), 0);
        P.br_ptr[t] = (int32_t)(P.br_child_ptr.size() - 1);
        for (size_t j = 0; j < order.size(); ++j) {
            rank_of[t][order[j]] = (int32_t)j;
            for (uint16_t s : m[order[j]].terms) P.br_child.push_back(s);
            P.br_child_ptr.push_back((int32_t)P.br_child.size());
        }
        P.tile_nbranch[t] = (int32_t)m.size();
        if (2 * T + (int64_t)m.size() >= 0xFFFF) { set_error("tile value array exceeds 16-bit slots"); return GT_ERR_LIMIT; }
        P.max_tile_values = std::max<int32_t>(P.max_tile_values, 2 * T + (int32_t)m.size());
    }
    P.br_ptr[NT] = (int32_t)(P.br_child_ptr.size() - 1);
    for (int64_t n = 0; n < N; ++n) {
        const int32_t idx = pending_multi_index[(size_t)n];
        if (idx >= 0) {
            const int32_t t = L.lo[(size_t)n] / T;
            P.node_slot[(size_t)n] = (uint16_t)(2 * T + rank_of[t][idx]);
        }
    }
    P.max_levels = logT;

    // ---- spanning nodes: value = reduction over the maximal in-tile nodes below them ---------------
    P.span_node.clear(); P.span_ptr.assign(1, 0); P.span_term.clear();
    std::vector<int32_t> span_index((size_t)N, -1);
    for (int64_t n = 0; n < N; ++n) {  // ascending ids: children before parents
        if (!spanning[(size_t)n]) continue;
        span_index[(size_t)n] = (int32_t)P.span_node.size();
        P.span_node.push_back((int32_t)n);
        for (int32_t p = L.child_ptr[(size_t)n]; p < L.child_ptr[(size_t)n + 1]; ++p) {
            const int32_t c = L.child_idx[(size_t)p];
            if (spanning[(size_t)c]) {
                const int32_t ci = span_index[(size_t)c];
                // copy by index: span_term may reallocate while we append
                for (int32_t q = P.span_ptr[(size_t)ci]; q < P.span_ptr[(size_t)ci + 1]; ++q) {
                    const int32_t term = P.span_term[(size_t)q];
                    P.span_term.push_back(term);
                }
            } else {
                P.span_term.push_back(c);
            }
        }
        P.span_ptr.push_back((int32_t)P.span_term.size());
    }
    return GT_OK;
}

}  // namespace gt

extern "C" {

int gt_plan(gt_trie* t, int32_t tile_leaves, int32_t seg_positions) {
    if (!t) { gt::set_error("gt_plan: null trie"); return GT_ERR_ARG; }
    auto env_int = [](const char* name, int dflt) { const char* s = getenv(name); return s && *s ? atoi(s) : dflt; };
    if (t->plan && tile_leaves <= 0) tile_leaves = t->plan->T;
    if (t->plan && seg_positions <= 0) seg_positions = t->plan->Q;
    if (tile_leaves <= 0) tile_leaves = env_int("GT_TILE_LEAVES", 4096);
    if (seg_positions <= 0) seg_positions = env_int("GT_SEG_POSITIONS", 8192);
    if (t->plan) {
        if (t->plan->T == tile_leaves && t->plan->Q == seg_positions) return GT_OK;
        gt::set_error("gt_plan: a plan with T=%d Q=%d already exists", t->plan->T, t->plan->Q);
        return GT_ERR_STATE;
    }
    std::unique_ptr<gt::Plan> p(new gt::Plan());
    const int rc = gt::build_plan(t->layout, tile_leaves, seg_positions, *p);
    if (rc != GT_OK) return rc;
    t->plan = std::move(p);
    return GT_OK;
}

int64_t gt_export_plan_array(const gt_trie* t, const char* name, void* dst, int64_t capacity, int32_t* elem_size) {
    if (!t || !t->plan || !name) { gt::set_error("gt_export_plan_array: no plan"); return -1; }
    const gt::Plan& P = *t->plan;
    const void* src = nullptr; int64_t n = -1; int32_t es = 4;
#define GT_ARR(field) if (!strcmp(name, #field)) { src = P.field.data(); n = (int64_t)P.field.size(); es = (int32_t)sizeof(P.field[0]); }
    GT_ARR(p1_chunk_ptr) GT_ARR(p1_zoff) GT_ARR(p1_src) GT_ARR(z_tile_off) GT_ARR(p2_slot) GT_ARR(br_ptr)
    GT_ARR(br_child_ptr) GT_ARR(br_child) GT_ARR(tile_node_lo) GT_ARR(node_slot) GT_ARR(span_node) GT_ARR(span_ptr)
    GT_ARR(span_term)
#undef GT_ARR
    if (n < 0) { gt::set_error("gt_export_plan_array: unknown array '%s'", name); return -1; }
    if (elem_size) *elem_size = es;
    if (dst && n > 0) memcpy(dst, src, (size_t)std::min<int64_t>(n, capacity) * es);
    return n;
}

}  // extern "C"
