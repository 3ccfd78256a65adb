// Internal definitions shared by the builder (host C++), the planner and the CUDA side.
#pragma once
#include <cstdint>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>
#include <map>
#include <memory>

#include "../../include/genlm_trie_b200.h"

namespace gt {

void set_error(const char* fmt, ...);

// Host layout in the reference's node-id space (post-order, root = N-1).
struct Layout {
    int64_t V = 0, N = 0, nnz = 0, max_depth = 0;
    std::vector<int32_t> leaf_node;   // [V]
    std::vector<int32_t> parent;      // [N]
    std::vector<int32_t> edge_label;  // [N]
    std::vector<int32_t> child_ptr;   // [N+1]
    std::vector<int32_t> child_idx;   // [N-1]
    std::vector<int32_t> perm;        // [V]  DFS rank -> item position
    std::vector<int32_t> lo, hi;      // [N]  DFS leaf range
    std::vector<uint8_t> is_leaf;     // [N]
};

// ---- tile plan (host copy; see trie_plan.cpp for how it is derived) -------------------------
//
// Source segments: vocabulary positions [s*Q, (s+1)*Q).   Tiles: DFS leaf ranks [t*T, (t+1)*T).
// Staging buffer z (one row = Zrow floats, tile-major): tile t occupies [z_tile_off[t], z_tile_off[t+1]),
// inside it one run per source segment (padded to 4 elements).
struct Plan {
    int32_t T = 0, Q = 0, NT = 0, NS = 0;
    int64_t Zrow = 0;  // floats per staged row (multiple of 4)

    // phase 1 (permute): per segment s, chunks of 4 staged elements.
    //   p1_chunk_ptr[s] .. p1_chunk_ptr[s+1]  : chunk index range of segment s
    //   p1_zoff[c]     : offset (in floats, multiple of 4) of chunk c inside a staged row
    //   p1_src[4*c+k]  : position inside the segment (uint16) feeding element k, 0xFFFF = padding
    std::vector<int32_t> p1_chunk_ptr;  // [NS+1]
    std::vector<int32_t> p1_zoff;       // [n_chunks]
    std::vector<uint16_t> p1_src;       // [4*n_chunks]

    // phase 2 (tile): staged element i of tile t goes to value slot p2_slot[z_tile_off[t] + i]
    // (uint16; 0xFFFF = padding).
    std::vector<int32_t> z_tile_off;    // [NT+1]
    std::vector<uint16_t> p2_slot;      // [Zrow]

    // per-tile value array: slots [0, tile_nleaf) are the tile's leaves in DFS order, then the
    // tile's branching nodes ordered by dependency level.
    std::vector<int32_t> tile_nleaf;    // [NT]
    std::vector<int32_t> tile_nbranch;  // [NT]
    // branching nodes, CSR over all tiles:
    std::vector<int32_t> br_ptr;        // [NT+1]  branching-node index range per tile
    std::vector<int32_t> br_child_ptr;  // [n_br+1] range into br_child (global offsets)
    std::vector<uint16_t> br_child;     // child value slots (tile-local)
    std::vector<int32_t> lvl_ptr;       // [NT+1]  range into lvl_end
    std::vector<int32_t> lvl_end;       // per tile: cumulative end (tile-local branching index) of each level
    int32_t max_levels = 0, max_tile_values = 0;

    // emit: node ids [tile_node_lo[t], tile_node_lo[t+1]) belong to tile t; node_slot[n] is the
    // tile-local value slot, 0xFFFF for spanning nodes (written by the fix-up).
    std::vector<int32_t> tile_node_lo;  // [NT+1]
    std::vector<uint16_t> node_slot;    // [N]

    // spanning nodes (leaf range crosses a tile boundary): value = reduce over frontier node ids.
    std::vector<int32_t> span_node;     // [n_span] node id
    std::vector<int32_t> span_ptr;      // [n_span+1] into span_term
    std::vector<int32_t> span_term;     // frontier node ids (in-tile nodes)
};

struct DevicePlan;  // defined in the CUDA translation unit
void free_device_plan(DevicePlan*);

}  // namespace gt

struct gt_trie {
    gt::Layout layout;
    std::unique_ptr<gt::Plan> plan;
    std::map<int, gt::DevicePlan*> dev;  // device ordinal -> resident metadata
    ~gt_trie();
};

namespace gt {
// trie_plan.cpp
int build_plan(const Layout& L, int32_t T, int32_t Q, Plan& P);
}
