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
#include <map>
#include <cstring>
#include <cstdlib>

namespace gt {

static inline int ilog2(int32_t x) { int k = 0; while ((1 << (k + 1)) <= x) ++k; return k; }

int build_plan(const Layout& L, int32_t T, int32_t Q, int32_t R, Plan& P) {
    const int64_t V = L.V, N = L.N;
    if ((T != 1024 && T != 2048)) { set_error("tile size must be 1024 or 2048 leaves (two value arrays of a tile must fit in shared memory)"); return GT_ERR_ARG; }
    // Q*8 bytes (one fp64 row segment) must fit in shared memory
    if (Q < 4 || Q > 16384 || (Q & 3)) { set_error("segment size must be a multiple of 4 in [4, 16384]"); return GT_ERR_ARG; }
    if (R != 2 && R != 4) { set_error("rows per CTA must be 2 or 4"); return GT_ERR_ARG; }
    const int logT = ilog2(T);
    // Aligned blocks stop at 2^kPyramidTop leaves: one warp builds all levels of its block with shuffles, so the
    // pyramid needs no cross-warp step; the few ranges longer than that just carry more terms.
    const int kTop = std::min(logT, kPyramidTop);
    P.T = T; P.Q = Q; P.R = R; P.slot_bytes = 4 * R;
    const int32_t SB = P.slot_bytes;
    auto swz = [&](int32_t s) -> uint16_t { return (uint16_t)(s < 2 * T ? swizzle_slot(s, SB) : s); };
    const int bank_mod = 128 / SB;  // slots per 128-byte shared-memory wavefront
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
    // run_len[t][s]: staged length of run (t,s): the count padded to 4; the last run of a tile absorbs what is
    // needed to make the tile's staged range a multiple of 8 elements (16-byte aligned uint16 slot tables, and
    // 16-byte cp.async granules for every row type)
    std::vector<int32_t> run_len((size_t)NT * NS, 0);
    for (int32_t t = 0; t < NT; ++t) {
        P.z_tile_off[t] = (int32_t)z;
        for (int32_t s = 0; s < NS; ++s) {
            run_off[(size_t)t * NS + s] = z;
            int32_t len = (cnt[(size_t)t * NS + s] + 3) & ~3;
            if (s == NS - 1 && ((z + len) & 7)) len += 4;
            run_len[(size_t)t * NS + s] = len;
            z += len;
        }
        P.max_tile_z = std::max<int32_t>(P.max_tile_z, (int32_t)(z - P.z_tile_off[t]));
    }
    P.z_tile_off[NT] = (int32_t)z;
    P.Zrow = z;
    if (z >= (int64_t)1 << 31) { set_error("staged row too long"); return GT_ERR_LIMIT; }

    P.p2_slot.assign((size_t)z, 0xFFFF);
    std::vector<uint16_t> z_src((size_t)z, 0xFFFF);  // position inside the source segment
    {
        // Order inside a run is free (both kernels follow these tables).  The tile kernel scatters element
        // 4*g + e of a run with lane g (for e = 0..3), 16 lanes per shared-memory wavefront, so we give group g
        // elements whose slot is congruent to g mod 16: the 16 lanes of a wavefront then hit 16 different bank
        // groups (8-byte slots).  Leftovers (classes are only roughly balanced) fill the remaining places.
        std::vector<std::vector<std::pair<uint16_t, uint16_t>>> runs((size_t)NT * NS);  // (slot, src)
        for (int64_t r = 0; r < V; ++r) {
            const int32_t t = (int32_t)(r / T), pos = L.perm[(size_t)r], s = pos / Q;
            runs[(size_t)t * NS + s].push_back({(uint16_t)(r - (int64_t)t * T), (uint16_t)(pos - s * Q)});
        }
        for (size_t ri = 0; ri < runs.size(); ++ri) {
            auto& run = runs[ri];
            const size_t n = run.size();
            std::vector<std::vector<std::pair<uint16_t, uint16_t>>> cls((size_t)bank_mod);
            for (auto& e : run) { e.first = swz(e.first); cls[e.first % bank_mod].push_back(e); }
            std::vector<std::pair<uint16_t, uint16_t>> placed(n);
            std::vector<uint8_t> used(n, 0);
            std::vector<size_t> next((size_t)bank_mod, 0);
            for (size_t pos = 0; pos < n; ++pos) {
                const size_t want = (pos / 4) % (size_t)bank_mod;
                if (next[want] < cls[want].size()) { placed[pos] = cls[want][next[want]++]; used[pos] = 1; }
            }
            size_t c = 0;
            for (size_t pos = 0; pos < n; ++pos) {
                if (used[pos]) continue;
                while (next[c] >= cls[c].size()) ++c;
                placed[pos] = cls[c][next[c]++];
            }
            const int64_t at = run_off[ri];
            for (size_t pos = 0; pos < n; ++pos) {
                P.p2_slot[(size_t)(at + (int64_t)pos)] = placed[pos].first;
                z_src[(size_t)(at + (int64_t)pos)] = placed[pos].second;
            }
        }
    }
    // phase-1 records: segment-major
    P.p1_chunk_ptr.assign((size_t)NS + 1, 0);
    P.p1_rec.clear();
    P.p1_rec.reserve((size_t)z);
    for (int32_t s = 0; s < NS; ++s) {
        P.p1_chunk_ptr[s] = (int32_t)(P.p1_rec.size() / 4);
        for (int32_t t = 0; t < NT; ++t) {
            const int64_t a = run_off[(size_t)t * NS + s];
            const int64_t n = run_len[(size_t)t * NS + s];
            for (int64_t k = 0; k < n; k += 4) {
                const uint32_t s0 = z_src[(size_t)(a + k)], s1 = z_src[(size_t)(a + k + 1)];
                const uint32_t s2 = z_src[(size_t)(a + k + 2)], s3 = z_src[(size_t)(a + k + 3)];
                P.p1_rec.push_back((int32_t)(a + k));
                P.p1_rec.push_back((int32_t)(s0 | (s1 << 16)));
                P.p1_rec.push_back((int32_t)(s2 | (s3 << 16)));
                P.p1_rec.push_back(0);
            }
        }
    }
    P.p1_chunk_ptr[NS] = (int32_t)(P.p1_rec.size() / 4);

    // ---- node intervals per tile -------------------------------------------------------------------
    P.tile_node_lo.assign((size_t)NT + 1, (int32_t)N);
    for (int32_t t = 0; t < NT; ++t) P.tile_node_lo[t] = L.leaf_node[(size_t)L.perm[(size_t)t * T]];
    if (NT > 0 && P.tile_node_lo[0] != 0) { set_error("internal: first DFS leaf is not node 0"); return GT_ERR_STATE; }

    // ---- value slots ------------------------------------------------------------------------------
    // A leaf range [a,b) inside tile t resolves to a slot: a leaf, one aligned pyramid block, or a
    // multi-term range (deduplicated per tile) whose slot is known after sorting by term count.
    const uint16_t IDENT = swz(2 * T - 1);
    auto block_slot = [&](int k, int32_t i) -> uint16_t {
        return swz(k == 0 ? i : 2 * T - (T >> (k - 1)) + i);
    };
    struct Multi { std::vector<uint16_t> terms; };
    std::vector<std::vector<Multi>> multi((size_t)NT);
    std::vector<std::map<std::pair<int32_t, int32_t>, int32_t>> multi_index((size_t)NT);
    // returns slot >= 0, or -(1 + index into multi[t])
    auto resolve = [&](int32_t t, int32_t a, int32_t b) -> int32_t {
        std::vector<uint16_t> terms;
        int32_t x = a;
        while (x < b) {
            int k = x == 0 ? kTop : std::min(kTop, __builtin_ctz((unsigned)x));
            while (x + (1 << k) > b) --k;
            terms.push_back(block_slot(k, x >> k));
            x += 1 << k;
        }
        if (terms.size() == 1) return terms[0];
        auto key = std::make_pair(a, b);
        auto it = multi_index[t].find(key);
        if (it != multi_index[t].end()) return -(1 + it->second);
        const int32_t idx = (int32_t)multi[t].size();
        multi_index[t][key] = idx;
        multi[t].push_back(Multi{std::move(terms)});
        return -(1 + idx);
    };

    std::vector<int32_t> node_res((size_t)N, INT32_MIN);  // resolve() result per in-tile node
    std::vector<uint8_t> spanning((size_t)N, 0);
    for (int64_t n = 0; n < N; ++n) {
        const int32_t lo = L.lo[(size_t)n], hi = L.hi[(size_t)n];
        if (hi <= lo) { spanning[(size_t)n] = 1; continue; }  // empty range (root of an empty vocabulary)
        const int32_t t = lo / T;
        if ((hi - 1) / T != t) { spanning[(size_t)n] = 1; continue; }
        node_res[(size_t)n] = resolve(t, lo - t * T, hi - t * T);
    }

    // spanning nodes -> pieces
    struct PieceEmit { int32_t res; int32_t idx; };
    std::vector<std::vector<PieceEmit>> piece_emit((size_t)NT);
    P.span_node.clear(); P.span_pp.assign(1, 0);
    int32_t n_pieces = 0;
    for (int64_t n = 0; n < N; ++n) {
        if (!spanning[(size_t)n]) continue;
        P.span_node.push_back((int32_t)n);
        const int32_t lo = L.lo[(size_t)n], hi = L.hi[(size_t)n];
        if (hi > lo) {
            for (int32_t t = lo / T; t <= (hi - 1) / T; ++t) {
                const int32_t a = std::max(lo, t * T) - t * T, b = std::min<int64_t>(hi, (int64_t)(t + 1) * T) - t * T;
                piece_emit[t].push_back(PieceEmit{resolve(t, a, b), n_pieces++});
            }
        }
        P.span_pp.push_back(n_pieces);
    }
    P.n_pieces = n_pieces;

    // order each tile's multi-term ranges by descending term count (a warp's 32 lanes then share a trip
    // count), pack them as ELL chunks, assign slots 2T + j
    P.ell_chunk_ptr.assign((size_t)NT + 1, 0);
    P.ell_desc.clear(); P.ell_terms.clear();
    P.max_tile_values = 2 * T;
    P.n_multi = 0; P.n_terms = 0;
    std::vector<std::vector<int32_t>> rank_of((size_t)NT);  // original index -> sorted position
    for (int32_t t = 0; t < NT; ++t) {
        auto& m = multi[t];
        std::vector<int32_t> order(m.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(),
                         [&](int32_t x, int32_t y) { return m[x].terms.size() > m[y].terms.size(); });
        rank_of[t].assign(m.size(), 0);
        for (size_t j = 0; j < order.size(); ++j) rank_of[t][order[j]] = (int32_t)j;
        P.ell_chunk_ptr[t] = (int32_t)(P.ell_desc.size() / 2);
        for (size_t j0 = 0; j0 < order.size(); j0 += 32) {
            const size_t cn = std::min<size_t>(32, order.size() - j0);
            const int32_t k_real = (int32_t)m[order[j0]].terms.size();
            const int32_t k = (k_real + 3) & ~3;  // whole batches of 4 term rows; the padding reads the identity slot
            P.ell_desc.push_back((int32_t)(P.ell_terms.size() / 32));
            P.ell_desc.push_back(k);
            // The order in which a range adds its terms is free: per term row pick, lane by lane, a remaining term
            // whose bank group (8-byte slots, 16 lanes per wavefront) is not taken yet in this half-warp.
            std::vector<std::vector<uint16_t>> left(32);
            for (size_t lane = 0; lane < cn; ++lane) left[lane] = m[order[j0 + lane]].terms;
            for (int32_t kk = 0; kk < k; ++kk) {
                uint32_t taken[4] = {0, 0, 0, 0};
                const int lanes_per_wf = bank_mod;  // lanes served by one 128-byte wavefront
                uint16_t row[32];
                for (size_t lane = 0; lane < 32; ++lane) {
                    uint16_t sl = IDENT;
                    auto& rem = left[lane];
                    if (!rem.empty()) {
                        size_t pick = 0;
                        for (size_t q = 0; q < rem.size(); ++q)
                            if (!(taken[lane / lanes_per_wf] & (1u << (rem[q] % bank_mod)))) { pick = q; break; }
                        sl = rem[pick];
                        rem.erase(rem.begin() + (long)pick);
                        taken[lane / lanes_per_wf] |= 1u << (sl % bank_mod);
                    }
                    row[lane] = sl;
                }
                // lanes whose range is exhausted read the identity slot (a broadcast, no conflict)
                for (size_t lane = 0; lane < 32; ++lane) P.ell_terms.push_back(row[lane]);
            }
        }
        P.n_multi += (int64_t)m.size();
        for (auto& e : m) P.n_terms += (int64_t)e.terms.size();
        const int64_t values = 2 * (int64_t)T + (int64_t)((m.size() + 31) / 32) * 32;
        if (values >= 0xFFFF) { set_error("tile value array exceeds 16-bit slots"); return GT_ERR_LIMIT; }
        P.max_tile_values = std::max<int32_t>(P.max_tile_values, (int32_t)values);
    }
    P.ell_chunk_ptr[NT] = (int32_t)(P.ell_desc.size() / 2);
    P.max_levels = kTop;

    auto final_slot = [&](int32_t t, int32_t res) -> uint16_t {
        return res >= 0 ? (uint16_t)res : (uint16_t)(2 * T + rank_of[t][-(res + 1)]);
    };
    P.node_slot.assign((size_t)N, IDENT);  // spanning nodes: identity slot (their emit is overwritten by the piece reduction)
    for (int64_t n = 0; n < N; ++n)
        if (!spanning[(size_t)n]) P.node_slot[(size_t)n] = final_slot(L.lo[(size_t)n] / T, node_res[(size_t)n]);
    P.piece_ptr.assign((size_t)NT + 1, 0);
    P.piece_slot.clear(); P.piece_idx.clear();
    for (int32_t t = 0; t < NT; ++t) {
        P.piece_ptr[t] = (int32_t)P.piece_slot.size();
        for (const PieceEmit& e : piece_emit[t]) {
            P.piece_slot.push_back(final_slot(t, e.res));
            P.piece_idx.push_back(e.idx);
        }
    }
    P.piece_ptr[NT] = (int32_t)P.piece_slot.size();
    P.max_tile_nodes = 0; P.max_tile_ell_rows = 0; P.max_tile_chunks = 0;
    P.ell_row_ptr.assign((size_t)NT + 1, 0);
    for (int32_t t = 0; t < NT; ++t) {
        // the emit stage is copied from the 16-byte aligned start at or below the interval
        const int32_t a = P.tile_node_lo[t] & ~7;
        P.max_tile_nodes = std::max(P.max_tile_nodes, ((P.tile_node_lo[t + 1] - a + 7) & ~7));
        int32_t rows = 0;
        for (int32_t c = P.ell_chunk_ptr[t]; c < P.ell_chunk_ptr[t + 1]; ++c) rows += P.ell_desc[2 * (size_t)c + 1];
        P.max_tile_ell_rows = std::max(P.max_tile_ell_rows, rows);
        P.max_tile_chunks = std::max(P.max_tile_chunks, P.ell_chunk_ptr[t + 1] - P.ell_chunk_ptr[t]);
        P.ell_row_ptr[(size_t)t + 1] = P.ell_row_ptr[(size_t)t] + rows;
    }
    if ((int64_t)P.ell_row_ptr[(size_t)NT] * 32 != (int64_t)P.ell_terms.size()) { set_error("internal: ELL row count"); return GT_ERR_STATE; }
    // staged padding elements land in the trash slot one past the value array (same for every tile)
    P.max_tile_values = (P.max_tile_values + 3) & ~3;
    for (auto& sl : P.p2_slot) if (sl == 0xFFFF) sl = (uint16_t)P.max_tile_values;
    return GT_OK;
}

}  // namespace gt

extern "C" {

int gt_plan(gt_trie* t, int32_t tile_leaves, int32_t seg_positions, int32_t rows_per_cta) {
    if (!t) { gt::set_error("gt_plan: null trie"); return GT_ERR_ARG; }
    auto env_int = [](const char* name, int dflt) { const char* s = getenv(name); return s && *s ? atoi(s) : dflt; };
    if (t->plan && tile_leaves <= 0) tile_leaves = t->plan->T;
    if (t->plan && seg_positions <= 0) seg_positions = t->plan->Q;
    if (t->plan && rows_per_cta <= 0) rows_per_cta = t->plan->R;
    if (tile_leaves <= 0) tile_leaves = env_int("GT_TILE_LEAVES", 1024);
    if (seg_positions <= 0) seg_positions = env_int("GT_SEG_POSITIONS", 4096);
    if (rows_per_cta <= 0) rows_per_cta = env_int("GT_ROWS_PER_CTA", 4);
    if (t->plan) {
        if (t->plan->T == tile_leaves && t->plan->Q == seg_positions && t->plan->R == rows_per_cta) return GT_OK;
        gt::set_error("gt_plan: a plan with T=%d Q=%d R=%d already exists", t->plan->T, t->plan->Q, t->plan->R);
        return GT_ERR_STATE;
    }
    std::unique_ptr<gt::Plan> p(new gt::Plan());
    const int rc = gt::build_plan(t->layout, tile_leaves, seg_positions, rows_per_cta, *p);
    if (rc != GT_OK) return rc;
    t->plan = std::move(p);
    return GT_OK;
}

int64_t gt_export_plan_array(const gt_trie* t, const char* name, void* dst, int64_t capacity, int32_t* elem_size) {
    if (!t || !t->plan || !name) { gt::set_error("gt_export_plan_array: no plan"); return -1; }
    const gt::Plan& P = *t->plan;
    const void* src = nullptr; int64_t n = -1; int32_t es = 4;
#define GT_ARR(field) if (!strcmp(name, #field)) { src = P.field.data(); n = (int64_t)P.field.size(); es = (int32_t)sizeof(P.field[0]); }
    GT_ARR(p1_chunk_ptr) GT_ARR(p1_rec) GT_ARR(z_tile_off) GT_ARR(p2_slot) GT_ARR(ell_chunk_ptr) GT_ARR(ell_desc)
    GT_ARR(ell_terms) GT_ARR(ell_row_ptr) GT_ARR(tile_node_lo) GT_ARR(node_slot) GT_ARR(piece_ptr) GT_ARR(piece_slot) GT_ARR(piece_idx)
    GT_ARR(span_node) GT_ARR(span_pp)
#undef GT_ARR
    if (n < 0) { gt::set_error("gt_export_plan_array: unknown array '%s'", name); return -1; }
    if (elem_size) *elem_size = es;
    if (dst && n > 0) memcpy(dst, src, (size_t)std::min<int64_t>(n, capacity) * es);
    return n;
}

}  // extern "C"
