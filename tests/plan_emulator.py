"""CPU emulation of the data flow of csrc/trie_kernels.cu, driven by the plan arrays the C++ planner produced
(gt_export_plan_array).  Test infrastructure: it lets the ``-m "not gpu"`` suite check the planner's metadata
(staging layout, value slots, aligned-block decomposition, spanning fix-up) against the oracle without a GPU.
Float32 arithmetic in the same association order as the kernels (pairwise pyramid, sequential term sums).
"""
import numpy as np


def emulate(engine, ws, op="sum"):
    """ws: float32 [B, V].  Returns float32 [B, N] computed the way permute/tile/span kernels do."""
    info = engine.plan_info()
    A = {name: engine.plan_array(name) for name in (
        "p1_chunk_ptr", "p1_zoff", "p1_src", "z_tile_off", "p2_slot", "br_ptr", "br_child_ptr", "br_child",
        "tile_node_lo", "node_slot", "span_node", "span_ptr", "span_term")}
    T, Q, NT, NS = info["tile_leaves"], info["seg_positions"], info["n_tiles"], info["n_segs"]
    V, N, Zrow = info["n_tokens"], info["n_nodes"], info["staged_row_elems"]
    logT = T.bit_length() - 1
    ws = np.asarray(ws, dtype=np.float32)
    B = ws.shape[0]
    red = np.add if op == "sum" else np.fmax
    ident = np.float32(0.0) if op == "sum" else np.float32(-np.inf)

    # phase 1: permute_kernel
    z = np.full((B, Zrow), np.nan, dtype=np.float32)
    for s in range(NS):
        seg = ws[:, s * Q:min(V, (s + 1) * Q)]
        c0, c1 = A["p1_chunk_ptr"][s], A["p1_chunk_ptr"][s + 1]
        zoff = A["p1_zoff"][c0:c1].astype(np.int64)
        src = A["p1_src"][4 * c0:4 * c1].reshape(-1, 4)
        dst = (zoff[:, None] + np.arange(4)[None, :]).reshape(-1)
        srcf = src.reshape(-1)
        pad = srcf == 0xFFFF
        vals = np.zeros((B, len(srcf)), dtype=np.float32)
        vals[:, ~pad] = seg[:, srcf[~pad]]
        z[:, dst] = vals
    assert not np.isnan(z).any() or np.isnan(ws).any(), "staging row has unwritten elements"

    out = np.full((B, N), np.nan, dtype=np.float32)
    SV = info["max_tile_values"]
    for t in range(NT):
        # phase 2.1: leaves
        vals = np.full((B, SV), np.nan, dtype=np.float32)
        zlo, zhi = A["z_tile_off"][t], A["z_tile_off"][t + 1]
        slot = A["p2_slot"][zlo:zhi]
        keep = slot != 0xFFFF
        vals[:, slot[keep]] = z[:, zlo:zhi][:, keep]
        nleaf = min(T, V - t * T)
        vals[:, nleaf:T] = ident
        assert not np.isnan(vals[:, :nleaf]).any() or np.isnan(ws).any()
        # phase 2.2: pyramid, level k block i at 2T - (T >> (k-1)) + i
        prev = vals[:, :T]
        for k in range(1, logT + 1):
            cur = red(prev[:, 0::2], prev[:, 1::2]).astype(np.float32)
            off = 2 * T - (T >> (k - 1))
            vals[:, off:off + cur.shape[1]] = cur
            prev = cur
        # phase 2.3: multi-term nodes
        j0, j1 = A["br_ptr"][t], A["br_ptr"][t + 1]
        for j in range(j0, j1):
            p0, p1 = A["br_child_ptr"][j], A["br_child_ptr"][j + 1]
            acc = np.full(B, ident, dtype=np.float32)
            for p in range(p0, p1):
                acc = red(acc, vals[:, A["br_child"][p]]).astype(np.float32)
            vals[:, 2 * T + (j - j0)] = acc
        # phase 2.4: emit
        n0, n1 = A["tile_node_lo"][t], A["tile_node_lo"][t + 1]
        sl = A["node_slot"][n0:n1]
        keep = sl != 0xFFFF
        out[:, n0 + np.flatnonzero(keep)] = vals[:, sl[keep]]
    # phase 3: spanning nodes
    for i, node in enumerate(A["span_node"]):
        p0, p1 = A["span_ptr"][i], A["span_ptr"][i + 1]
        terms = A["span_term"][p0:p1]
        if p1 == p0:
            out[:, node] = 0.0
        elif op == "sum":
            out[:, node] = out[:, terms].astype(np.float64).sum(axis=1).astype(np.float32)
        else:
            out[:, node] = np.fmax.reduce(out[:, terms], axis=1)
    return out
