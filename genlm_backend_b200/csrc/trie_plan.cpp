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

int build_plan(const Layout& L, int32_t T, int32_t R, Plan& P) {
    const int64_t V = L.V, N = L.N;
    if ((T != 1024 && T != 2048)) { set_error("tile size must be 1024 or 2048 leaves (two value arrays of a tile must fit in shared memory)"); return GT_ERR_ARG; }
    if (R != 4) { set_error("rows per work item must be 4 (16-byte value slots: 4 fp32 rows, 2 fp64 rows)"); return GT_ERR_ARG; }
    const int logT = ilog2(T);
    // Aligned blocks stop at 2^kPyramidTop leaves: one warp builds all levels of its block with shuffles, so the
    // pyramid needs no cross-warp step; the few ranges longer than that just carry more terms.
    const int kTop = std::min(logT, kPyramidTop);
    P.T = T; P.R = R; P.slot_bytes = 4 * R;
    const int32_t SB = P.slot_bytes;
    auto swz = [&](int32_t s) -> uint16_t { return (uint16_t)(s < 2 * T ? swizzle_slot(s, SB) : s); };
    const int bank_mod = 128 / SB;  // slots per 128-byte shared-memory wavefront
    P.NT = (int32_t)((V + T - 1) / T);
    const int32_t NT = P.NT;
    if ((int64_t)NT * T >= (int64_t)1 << 31) { set_error("vocabulary too large"); return GT_ERR_LIMIT; }

    // ---- permute table: the item at DFS rank r goes to leaf slot r % T of tile r / T.  The swizzle permutes slots
    // inside aligned groups of 8 chunks only, so applying it to the rank is applying it to the tile-local slot.
    P.leaf_dest.assign((size_t)V, 0);
    for (int64_t r = 0; r < V; ++r) {
        const int32_t t = (int32_t)(r / T);
        P.leaf_dest[(size_t)L.perm[(size_t)r]] = t * T + (int32_t)swz((int32_t)(r - (int64_t)t * T));
    }

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
        // Which lane of its chunk a range takes is free, and decides the bank group of its value slot (slot mod
        // bank_mod = lane mod bank_mod).  The emit phase reads the slots of bank_mod consecutive node ids in one
        // shared-memory wavefront (output rows are line-aligned, so the groups are the aligned ones); leaf and pyramid
        // slots are fixed by arithmetic, so a range takes a lane whose bank group is not used by the other nodes of
        // the groups it is emitted in -- most constrained ranges first.  Measured: 1.21 wavefronts per group read
        // instead of 1.61.  Holes (last chunk only) stay empty lanes.
        const int32_t tn0 = P.tile_node_lo[t], tn1 = P.tile_node_lo[t + 1];
        const int32_t g0 = tn0 / bank_mod, ng = tn1 > tn0 ? (tn1 - 1) / bank_mod - g0 + 1 : 0;
        std::vector<uint32_t> grp_mask((size_t)ng, 0u);            // bank groups taken per emit group
        std::vector<std::vector<int32_t>> grp_of(m.size());        // emit groups each range appears in
        for (int32_t n = tn0; n < tn1; ++n) {
            const int32_t g = n / bank_mod - g0;
            const int32_t res = spanning[(size_t)n] ? (int32_t)IDENT : node_res[(size_t)n];
            if (res >= 0) grp_mask[(size_t)g] |= 1u << (res % bank_mod);
            else {
                auto& v = grp_of[(size_t)(-(res + 1))];
                if (v.empty() || v.back() != g) v.push_back(g);
            }
        }
        const size_t n_chunks = (order.size() + 31) / 32;
        std::vector<int32_t> placed(n_chunks * 32, -1);  // lane position -> range index
        for (size_t c = 0; c < n_chunks; ++c) {
            const size_t j0 = c * 32, cn = std::min<size_t>(32, order.size() - j0);
            auto forbidden = [&](int32_t x) {
                uint32_t f = 0;
                for (int32_t g : grp_of[(size_t)x]) f |= grp_mask[(size_t)g];
                return f;
            };
            std::vector<int32_t> xs(order.begin() + (long)j0, order.begin() + (long)(j0 + cn));
            std::stable_sort(xs.begin(), xs.end(), [&](int32_t x, int32_t y) {
                return __builtin_popcount(forbidden(x)) > __builtin_popcount(forbidden(y));
            });
            uint32_t free_lanes = 0xFFFFFFFFu;
            for (int32_t x : xs) {
                const uint32_t f = forbidden(x);
                int best = -1, best_cnt = -1;
                for (int cg = 0; cg < bank_mod; ++cg) {  // the allowed bank group with the most free lanes
                    if ((f >> cg) & 1u) continue;
                    int cnt = 0;
                    for (int lane = cg; lane < 32; lane += bank_mod) cnt += (free_lanes >> lane) & 1u;
                    if (cnt > best_cnt && cnt > 0) { best = cg; best_cnt = cnt; }
                }
                int lane = -1;
                if (best >= 0) { for (int l = best; l < 32; l += bank_mod) if ((free_lanes >> l) & 1u) { lane = l; break; } }
                else lane = __builtin_ctz(free_lanes);
                free_lanes &= ~(1u << lane);
                placed[j0 + (size_t)lane] = x;
                for (int32_t g : grp_of[(size_t)x]) grp_mask[(size_t)g] |= 1u << (lane % bank_mod);
            }
        }
        rank_of[t].assign(m.size(), 0);
        for (size_t j = 0; j < placed.size(); ++j) if (placed[j] >= 0) rank_of[t][(size_t)placed[j]] = (int32_t)j;
        P.ell_chunk_ptr[t] = (int32_t)(P.ell_desc.size() / 2);
        for (size_t j0 = 0; j0 < placed.size(); j0 += 32) {
            int32_t k_real = 0;
            for (size_t lane = 0; lane < 32; ++lane)
                if (placed[j0 + lane] >= 0) k_real = std::max<int32_t>(k_real, (int32_t)m[(size_t)placed[j0 + lane]].terms.size());
            const int32_t k = (k_real + kEllRowPad - 1) & ~(kEllRowPad - 1);  // the padding reads the identity slot
            P.ell_desc.push_back((int32_t)(P.ell_terms.size() / 32));
            P.ell_desc.push_back(k);
            // The order in which a range adds its terms is free, and a range with fewer terms than the chunk has rows
            // may sit out any of them (it reads the identity slot there).  The bank_mod lanes that share a
            // shared-memory wavefront want different bank groups in every term row: per lane group, a randomised
            // greedy (lanes in random order take a remaining term whose bank group is free in the row, or pass if
            // they can afford to) is repeated from a fixed seed and the arrangement with the fewest wavefronts kept.
            std::vector<uint16_t> rows((size_t)k * 32, IDENT);
            uint32_t rng = 0x9E3779B9u ^ (uint32_t)(t * 7919 + (int32_t)j0);
            auto next_rand = [&]() { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
            for (int l0 = 0; l0 < 32; l0 += bank_mod) {
                std::vector<std::vector<uint16_t>> mine((size_t)bank_mod);
                size_t total_terms = 0;
                for (int q = 0; q < bank_mod; ++q)
                    if (placed[j0 + (size_t)(l0 + q)] >= 0) { mine[(size_t)q] = m[(size_t)placed[j0 + (size_t)(l0 + q)]].terms; total_terms += mine[(size_t)q].size(); }
                if (total_terms == 0) continue;
                std::vector<uint16_t> best_rows;
                int best_cost = INT32_MAX;
                const int tries = 48;
                for (int attempt = 0; attempt < tries && best_cost > k; ++attempt) {
                    std::vector<std::vector<uint16_t>> left = mine;
                    std::vector<uint16_t> cand((size_t)k * (size_t)bank_mod, IDENT);
                    int cost = 0;
                    for (int32_t kk = 0; kk < k; ++kk) {
                        uint32_t taken = 0;
                        int mult[32] = {0};
                        bool ident_read = false;
                        int order_l[32];
                        for (int q = 0; q < bank_mod; ++q) order_l[q] = q;
                        if (attempt > 0)
                            for (int q = bank_mod - 1; q > 0; --q) std::swap(order_l[q], order_l[(int)(next_rand() % (uint32_t)(q + 1))]);
                        for (int oi = 0; oi < bank_mod; ++oi) {
                            const int q = order_l[oi];
                            auto& rem = left[(size_t)q];
                            if (rem.empty()) { ident_read = true; continue; }
                            int pick = -1;
                            for (size_t x = 0; x < rem.size(); ++x)
                                if (!((taken >> (rem[x] % bank_mod)) & 1u)) { pick = (int)x; break; }
                            const bool can_pass = (int32_t)rem.size() < k - kk;
                            if (pick < 0 && can_pass) { ident_read = true; continue; }
                            if (pick < 0) pick = (int)(next_rand() % (uint32_t)rem.size());
                            const uint16_t sl = rem[(size_t)pick];
                            rem.erase(rem.begin() + pick);
                            taken |= 1u << (sl % bank_mod);
                            ++mult[sl % bank_mod];
                            cand[(size_t)kk * (size_t)bank_mod + (size_t)q] = sl;
                        }
                        if (ident_read) ++mult[IDENT % bank_mod];  // one more address in the identity slot's bank group
                        int mx = 1;
                        for (int c = 0; c < bank_mod; ++c) mx = std::max(mx, mult[c]);
                        cost += mx;
                    }
                    if (cost < best_cost) { best_cost = cost; best_rows = cand; }
                }
                for (int32_t kk = 0; kk < k; ++kk)
                    for (int q = 0; q < bank_mod; ++q) rows[(size_t)kk * 32 + (size_t)(l0 + q)] = best_rows[(size_t)kk * (size_t)bank_mod + (size_t)q];
            }
            P.ell_terms.insert(P.ell_terms.end(), rows.begin(), rows.end());
        }
        P.n_multi += (int64_t)m.size();
        for (auto& e : m) P.n_terms += (int64_t)e.terms.size();
        const int64_t values = 2 * (int64_t)T + (int64_t)((m.size() + 31) / 32) * 32;
        if (values >= 0xFFC0 - 64) { set_error("tile value array exceeds 16-bit slots"); return GT_ERR_LIMIT; }
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
    P.max_tile_values = (P.max_tile_values + 3) & ~3;
    return GT_OK;
}

}  // namespace gt

extern "C" {

int gt_plan(gt_trie* t, int32_t tile_leaves) {
    if (!t) { gt::set_error("gt_plan: null trie"); return GT_ERR_ARG; }
    auto env_int = [](const char* name, int dflt) { const char* s = getenv(name); return s && *s ? atoi(s) : dflt; };
    std::lock_guard<std::mutex> lock(t->mu);
    if (t->plan && tile_leaves <= 0) tile_leaves = t->plan->T;
    if (tile_leaves <= 0) tile_leaves = env_int("GT_TILE_LEAVES", 1024);
    if (t->plan) {
        if (t->plan->T == tile_leaves) return GT_OK;
        gt::set_error("gt_plan: a plan with T=%d already exists", t->plan->T);
        return GT_ERR_STATE;
    }
    std::unique_ptr<gt::Plan> p(new gt::Plan());
    const int rc = gt::build_plan(t->layout, tile_leaves, 4, *p);
    if (rc != GT_OK) return rc;
    t->plan = std::move(p);
    return GT_OK;
}

int64_t gt_export_plan_array(const gt_trie* t, const char* name, void* dst, int64_t capacity, int32_t* elem_size) {
    if (!t || !t->plan || !name) { gt::set_error("gt_export_plan_array: no plan"); return -1; }
    const gt::Plan& P = *t->plan;
    const void* src = nullptr; int64_t n = -1; int32_t es = 4;
#define GT_ARR(field) if (!strcmp(name, #field)) { src = P.field.data(); n = (int64_t)P.field.size(); es = (int32_t)sizeof(P.field[0]); }
    GT_ARR(leaf_dest) GT_ARR(ell_chunk_ptr) GT_ARR(ell_desc)
    GT_ARR(ell_terms) GT_ARR(ell_row_ptr) GT_ARR(tile_node_lo) GT_ARR(node_slot) GT_ARR(piece_ptr) GT_ARR(piece_slot) GT_ARR(piece_idx)
    GT_ARR(span_node) GT_ARR(span_pp)
#undef GT_ARR
    if (n < 0) { gt::set_error("gt_export_plan_array: unknown array '%s'", name); return -1; }
    if (elem_size) *elem_size = es;
    if (dst && n > 0) memcpy(dst, src, (size_t)std::min<int64_t>(n, capacity) * es);
    return n;
}

}  // extern "C"
