"""CPU emulation of the data flow of csrc/trie_kernels.cu, driven by the plan arrays the C++ planner produced
(gt_export_plan_array).  Test infrastructure: it lets the ``-m "not gpu"`` suite check the planner's metadata
(staging layout, value slots, aligned-block decomposition, spanning fix-up) against the oracle without a GPU.
Float32 arithmetic in the same association order as the kernels (pairwise pyramid, sequential term sums).
"""
import numpy as np


def swizzle_slot(s, slot_bytes):
    """Python twin of gt::swizzle_slot (csrc/trie_internal.h)."""
    cs = {4: 2, 8: 1, 16: 0}[slot_bytes]
    c = s >> cs
    return ((c ^ ((c >> 3) & (slot_bytes // 2 - 1))) << cs) | (s & ((1 << cs) - 1))


def emulate(engine, ws, op="sum"):
    """ws: float32 [B, V].  Returns float32 [B, N] computed the way permute/tile/span kernels do."""
    info = engine.plan_info()
    A = {name: engine.plan_array(name) for name in (
        "p1_chunk_ptr", "p1_rec", "z_tile_off", "p2_slot", "ell_chunk_ptr", "ell_desc", "ell_terms",
        "tile_node_lo", "node_slot", "piece_ptr", "piece_slot", "piece_idx", "span_node", "span_pp")}
    T, Q, NT, NS = info["tile_leaves"], info["seg_positions"], info["n_tiles"], info["n_segs"]
    V, N, Zrow = info["n_tokens"], info["n_nodes"], info["staged_row_elems"]
    logT = T.bit_length() - 1
    phys = swizzle_slot(np.arange(2 * T), 4 * info["rows_per_item"])  # where logical slot s < 2T lives
    ws = np.asarray(ws, dtype=np.float32)
    B = ws.shape[0]
    red = np.add if op == "sum" else np.fmax
    ident = np.float32(0.0) if op == "sum" else np.float32(-np.inf)

    # phase 1: permute_kernel (records of {zoff, src0|src1<<16, src2|src3<<16, 0})
    z = np.full((B, Zrow), np.nan, dtype=np.float32)
    rec = A["p1_rec"].reshape(-1, 4)
    for s in range(NS):
        seg = ws[:, s * Q:min(V, (s + 1) * Q)]
        c0, c1 = A["p1_chunk_ptr"][s], A["p1_chunk_ptr"][s + 1]
        r = rec[c0:c1]
        zoff = r[:, 0].astype(np.int64)
        lohi = r[:, 1:3].astype(np.int64) & 0xFFFFFFFF
        src = np.stack([lohi[:, 0] & 0xFFFF, lohi[:, 0] >> 16, lohi[:, 1] & 0xFFFF, lohi[:, 1] >> 16], axis=1)
        dst = (zoff[:, None] + np.arange(4)[None, :]).reshape(-1)
        srcf = src.reshape(-1)
        pad = srcf == 0xFFFF
        vals = np.zeros((B, len(srcf)), dtype=np.float32)
        vals[:, ~pad] = seg[:, srcf[~pad]]
        z[:, dst] = vals
    assert not np.isnan(z).any() or np.isnan(ws).any(), "staging row has unwritten elements"

    out = np.full((B, N), np.nan, dtype=np.float32)
    part = np.full((B, max(int(A["span_pp"][-1]) if len(A["span_pp"]) else 0, 1)), np.nan, dtype=np.float32)
    if len(A["span_pp"]) == 0 or A["span_pp"][-1] == 0:
        part[:] = 0
    SV = info["max_tile_values"]
    for t in range(NT):
        # phase 2.1: leaves
        vals = np.full((B, SV + 16), np.nan, dtype=np.float32)  # slots SV .. SV+15: trash slots for staged padding
        zlo, zhi = A["z_tile_off"][t], A["z_tile_off"][t + 1]
        slot = A["p2_slot"][zlo:zhi]
        nleaf = min(T, V - t * T)
        assert slot.max() < SV + 16 and np.array_equal(np.sort(slot[slot < SV]), np.sort(phys[:nleaf]))
        vals[:, slot] = z[:, zlo:zhi]
        vals[:, phys[nleaf:T]] = ident
        assert not np.isnan(vals[:, phys[:nleaf]]).any() or np.isnan(ws).any()
        # phase 2.2: pyramid, level k block i at logical slot 2T - (T >> (k-1)) + i (physical: swizzled)
        prev = vals[:, phys[:T]]
        for k in range(1, info["max_levels"] + 1):  # aligned blocks stop at 2^max_levels leaves (a warp's share)
            cur = red(prev[:, 0::2], prev[:, 1::2]).astype(np.float32)
            off = 2 * T - (T >> (k - 1))
            vals[:, phys[off:off + cur.shape[1]]] = cur
            prev = cur
        # phase 2.3: multi-term ranges, ELL chunks of 32 (padding = identity slot 2T-1)
        vals[:, phys[2 * T - 1]] = ident
        c0, c1 = A["ell_chunk_ptr"][t], A["ell_chunk_ptr"][t + 1]
        desc = A["ell_desc"].reshape(-1, 2)
        for c in range(c0, c1):
            off32, k = desc[c]
            terms = A["ell_terms"][32 * off32:32 * (off32 + k)].reshape(k, 32)
            acc = np.full((B, 32), ident, dtype=np.float32)
            for kk in range(k):
                acc = red(acc, vals[:, terms[kk]]).astype(np.float32)
            vals[:, 2 * T + 32 * (c - c0):2 * T + 32 * (c - c0 + 1)] = acc
        # phase 2.4: emit
        n0, n1 = A["tile_node_lo"][t], A["tile_node_lo"][t + 1]
        sl = A["node_slot"][n0:n1]  # spanning nodes carry the identity slot here and are overwritten below
        out[:, n0:n1] = vals[:, sl]
        # phase 2.5: pieces of spanning nodes
        p0, p1 = A["piece_ptr"][t], A["piece_ptr"][t + 1]
        part[:, A["piece_idx"][p0:p1]] = vals[:, A["piece_slot"][p0:p1]]
    # last CTA of the row group: spanning nodes from their pieces (fp64 for sums)
    assert not np.isnan(part).any() or np.isnan(ws).any()
    for i, node in enumerate(A["span_node"]):
        q0, q1 = A["span_pp"][i], A["span_pp"][i + 1]
        if q1 == q0:
            out[:, node] = 0.0
        elif op == "sum":
            out[:, node] = part[:, q0:q1].astype(np.float64).sum(axis=1).astype(np.float32)
        else:
            out[:, node] = np.fmax.reduce(part[:, q0:q1], axis=1)
    return out
