"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by ``genlm_backend_b200``).

CPU restatement of the reference's algorithm for the trie-mass / SMC-sampling hot path:

* ``OracleTrie`` -- pure-Python restatement of the reference trie construction and node numbering
  (``genlm/backend/trie/base.py:13-122, 219-247``) and of the reachability walk
  (``genlm/backend/trie/parallel.py:21-64``);
* ``weight_sum`` / ``weight_max`` -- the numba loops (``base.py:346-393``) restated in C (``trie_oracle.c``),
  float64, single thread per row;
* ``parallel_weight_sum`` / ``parallel_weight_max`` -- the torch formulas of ``parallel.py:92-145`` in numpy
  float32 (secondary check only: the reference's own fp32 SpMM differs from its fp64 path by up to 1e-4 rel);
* ``masked_logsumexp`` / ``masked_probs`` -- float64 restatement of ``README.md:82-87``;
* ``OracleTrie.subtree_token_mask`` / ``gather_nodes`` -- the read-outs (a column of the reachability matrix,
  ``parallel.py:33-64``; indexing the numpy result of ``parallel.py:103,145``).

Parity pin: ``tests/test_oracle.py`` checks every function here against outputs of the reference itself
(``tests/golden/*.npz``, produced by ``tests/golden/make_golden.py`` importing ``/root/reference`` in the dev
container) and against the known answers of the reference's own tests (``tests/test_trie.py:26-85``,
``tests/test_token.py:96-156, 264-312``).  The sampler has no result-pinning test in the reference
("parity unpinned" for the draw; logsumexp is pinned against ``torch.logsumexp`` in float64).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    """Compile ``trie_oracle.c`` (gcc + OpenMP) into ``oracle/liboracle.so``."""
    src = os.path.join(_HERE, "trie_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


_lib = None


def _c():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_batch.restype = ctypes.c_int
        _lib.oracle_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                      ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int]
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def max_threads():
    return int(_c().oracle_max_threads())


class OracleTrie:
    """Restatement of ``TokenCharacterTrie.__init__`` (``base.py:13-93``): insert every item symbol by symbol
    into per-node dicts (insertion-ordered), give every item its own leaf under the key ``(None, idx)``, then
    renumber all nodes by a full post-order walk with children in insertion order (``base.py:80-83, 236-247``).

    ``decode`` items: ``bytes``-like (iterated as ints) or any iterable of hashable labels.  Objects with a
    ``token_id`` attribute (Token) are iterated through ``bytes(item)``.
    """

    def __init__(self, decode):
        children = [{}]
        leaf_of = []
        for idx, item in enumerate(decode):
            word = bytes(item) if isinstance(item, (bytes, bytearray)) else item
            cur = 0
            for letter in word:  # base.py:50-54
                nxt = children[cur].get(letter)
                if nxt is None:
                    nxt = len(children)
                    children[cur][letter] = nxt
                    children.append({})
                cur = nxt
            leaf = len(children)  # base.py:55-61
            children[cur][(None, idx)] = leaf
            children.append({})
            leaf_of.append(leaf)

        # full post-order, children in insertion order (iterative form of base.py:236-247)
        order = {}
        stack = [(0, iter(children[0].values()))]
        while stack:
            node, it = stack[-1]
            child = next(it, None)
            if child is None:
                order[node] = len(order)
                stack.pop()
            else:
                stack.append((child, iter(children[child].values())))

        n = len(children)
        self.children = [None] * n
        for old, kids in enumerate(children):  # base.py:95-108 (_rename)
            self.children[order[old]] = {k: order[c] for k, c in kids.items()}
        self.root = order[0]
        self.idx_to_leaf = np.array([(i, order[x]) for i, x in enumerate(leaf_of)], dtype=np.int32).reshape(-1, 2)
        # jump: sorted child ids per node (base.py:120-122); ordering: internal nodes, post-order (base.py:74, 219-234)
        self.jump = [np.array(sorted(k.values()), dtype=np.int32) for k in self.children]
        self.ordering = np.array([x for x in range(n) if self.children[x]], dtype=np.int64)
        self.jump_ptr = np.zeros(n + 1, dtype=np.int32)
        self.jump_ptr[1:] = np.cumsum([len(j) for j in self.jump])
        self.jump_idx = np.concatenate(self.jump).astype(np.int32) if n > 1 else np.zeros(0, np.int32)
        self.n_nodes = n
        self.n_items = len(leaf_of)

    # ---- parallel.py:21-64 -----------------------------------------------------------------------------
    def reachability(self):
        parent = {}
        for node in range(self.n_nodes):
            for child in self.jump[node]:
                parent[int(child)] = node
        rows, cols = [], []
        for i, node in enumerate(self.idx_to_leaf[:, 1].tolist()):
            rows.append(i)
            cols.append(node)
            cur = node
            while cur in parent:
                cur = parent[cur]
                rows.append(i)
                cols.append(cur)
        return np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64)

    def subtree_token_mask(self, nodes):
        """Column ``nodes[b]`` of the reachability matrix ``M[V, N]`` (``parallel.py:33-64``: ``M[i, n] = 1`` when
        node ``n`` is item ``i``'s leaf or one of its ancestors) as a bool array ``[B, V]``."""
        rows, cols = self.reachability()
        out = np.zeros((len(nodes), self.n_items), dtype=bool)
        for b, n in enumerate(nodes):
            out[b, rows[cols == int(n)]] = True
        return out

    # ---- base.py:346-393 through the C restatement -----------------------------------------------------
    def _batch(self, ws, op, threads=1):
        ws = np.asarray(ws)
        if ws.dtype != np.float64:
            ws = ws.astype(np.float32, copy=False)
        ws = np.ascontiguousarray(ws)
        single = ws.ndim == 1
        if single:
            ws = ws[None, :]
        assert ws.shape[1] == self.n_items, [ws.shape, self.n_items]
        out = np.empty((ws.shape[0], self.n_nodes), dtype=np.float64)
        idx = np.ascontiguousarray(self.idx_to_leaf, dtype=np.int32)
        used = _c().oracle_batch(out.ctypes.data, ws.ctypes.data, int(ws.dtype == np.float64), ws.shape[0], ws.shape[1],
                                 self.n_nodes, idx.ctypes.data, self.n_items, self.jump_ptr.ctypes.data,
                                 self.jump_idx.ctypes.data, self.ordering.ctypes.data, len(self.ordering),
                                 0 if op == "sum" else 1, threads)
        self.last_threads = used
        return out[0] if single else out

    def weight_sum(self, ws, threads=1):
        return self._batch(ws, "sum", threads)

    def weight_max(self, ws, threads=1):
        return self._batch(ws, "max", threads)

    # ---- parallel.py:92-145 in numpy float32 (secondary) ---------------------------------------------------
    def parallel_weight_sum(self, ws):
        rows, cols = self.reachability()
        ws = np.atleast_2d(np.asarray(ws, dtype=np.float32))
        out = np.zeros((ws.shape[0], self.n_nodes), dtype=np.float32)
        for b in range(ws.shape[0]):
            np.add.at(out[b], cols, ws[b, rows])
        return out

    def parallel_weight_max(self, ws):
        rows, cols = self.reachability()
        ws = np.atleast_2d(np.asarray(ws, dtype=np.float32))
        out = np.full((ws.shape[0], self.n_nodes), -np.inf, dtype=np.float32)  # include_self=False
        for b in range(ws.shape[0]):
            np.maximum.at(out[b], cols, ws[b, rows])
        return out


class OracleLayout:
    """The same kernels over layout arrays given directly (used where building the Python trie is too slow,
    e.g. the CPU-baseline timing at 128k tokens): ``idx_to_leaf`` int32[V,2], children CSR, internal nodes."""

    def __init__(self, idx_to_leaf, child_ptr, child_idx):
        self.idx_to_leaf = np.ascontiguousarray(idx_to_leaf, dtype=np.int32)
        self.jump_ptr = np.ascontiguousarray(child_ptr, dtype=np.int32)
        self.jump_idx = np.ascontiguousarray(child_idx, dtype=np.int32)
        self.n_nodes = len(self.jump_ptr) - 1
        self.n_items = len(self.idx_to_leaf)
        self.ordering = np.flatnonzero(np.diff(self.jump_ptr) > 0).astype(np.int64)

    _batch = OracleTrie._batch
    weight_sum = OracleTrie.weight_sum
    weight_max = OracleTrie.weight_max


# ---- SMC row op (README.md:82-87) in float64 ---------------------------------------------------------------
def masked_logsumexp(logps, mask=None, temperature=1.0):
    """``(logps / temperature + mask).logsumexp(-1)`` in float64; ``-inf`` for rows without mass."""
    x = np.asarray(logps, dtype=np.float64) / float(temperature)
    if mask is not None:
        x = x + np.asarray(mask, dtype=np.float64)
    m = np.max(x, axis=-1, keepdims=True)
    safe = np.where(np.isfinite(m), m, 0.0)
    with np.errstate(divide="ignore"):
        return (safe + np.log(np.sum(np.exp(x - safe), axis=-1, keepdims=True)))[..., 0]


def masked_probs(logps, mask=None, temperature=1.0):
    """``(masked - logZ).exp()`` in float64: the categorical the reference hands to ``torch.multinomial``."""
    x = np.asarray(logps, dtype=np.float64) / float(temperature)
    if mask is not None:
        x = x + np.asarray(mask, dtype=np.float64)
    lz = masked_logsumexp(logps, mask, temperature)
    return np.exp(x - lz[..., None])


# ---- read-outs of a mass slab (what a caller of parallel.py:103,145 does on the host with the numpy result) ----------
def gather_nodes(mass, node_ids, normalizer=None, log=False):
    """``mass[b, node_ids[b, k]]`` in float64 (ids outside ``[0, N)`` read as 0), optionally divided by
    ``mass[b, normalizer[b]]``, optionally as logs."""
    mass = np.asarray(mass, dtype=np.float64)
    B, N = mass.shape
    ids = np.asarray(node_ids, dtype=np.int64)
    if ids.ndim == 1:
        ids = np.broadcast_to(ids, (B, len(ids)))
    ok = (ids >= 0) & (ids < N)
    out = np.where(ok, np.take_along_axis(mass, np.where(ok, ids, 0), axis=1), 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        if normalizer is not None:
            z = mass[np.arange(B), np.broadcast_to(np.asarray(normalizer, dtype=np.int64), (B,))][:, None]
            return np.log(out) - np.log(z) if log else out / z
        return np.log(out) if log else out


def unpack_bits(bits, n):
    """int32 / uint32 keep-bitmask ``[..., ceil(n/32)]`` -> bool ``[..., n]`` (bit ``i % 32`` of word ``i // 32``)."""
    w = np.asarray(bits).astype(np.uint32)
    b = (w[..., :, None] >> np.arange(32, dtype=np.uint32)) & 1
    return b.reshape(*w.shape[:-1], -1)[..., :n].astype(bool)
