"""``TokenCharacterTrie``: byte-level trie over a token vocabulary with per-node mass queries.

Drop-in for the reference's ``genlm/backend/trie/base.py`` (same constructor, methods, attributes, node
numbering and errors), rebuilt B200-first:

* the trie is built by the C++ builder behind ``gt_build`` (iterative, flat arrays; reference:
  ``base.py:13-122`` pure-Python dict walk plus recursive renumbering);
* ``weight_sum`` / ``weight_max`` run the sm_100a kernels behind ``gt_weight_reduce`` with the fp64
  pipeline and return ``float64`` like the reference's numba loops (``base.py:346-393``).

The dict / list attributes of the reference (``children``, ``jump``, ``node2prefix`` ...) are
materialised lazily from the flat layout the first time they are read.
"""
import warnings

import numpy as np
import torch

from ..tokenization import Token
from ._engine import TrieEngine, require_cuda

_LEAF_MIN = np.iinfo(np.int32).min


def _encode_vocabulary(decode):
    """Flatten the vocabulary into (symbols int32, offsets int64, word keys, extra edge labels).

    Mirrors the item handling of the reference constructor (``base.py:29-48``): Token -> its bytes with key
    ``(bytes, token_id)``; plain bytes -> itself (DeprecationWarning once); any other iterable is iterated
    and its elements become edge labels as they are.  Byte values map to symbols 0..255; every other
    hashable label gets a symbol >= 256 in order of first appearance.
    """
    V = len(decode)
    keys = [None] * V
    warned = False
    all_bytes = True
    for i, item in enumerate(decode):
        if isinstance(item, Token):
            keys[i] = (item.byte_string, item.token_id)
        elif Token.is_plain_bytes(item):
            if not warned:
                warnings.warn(
                    "Passing plain bytes to TokenCharacterTrie is deprecated. "
                    "Use Token objects from decode_vocab() instead.",
                    DeprecationWarning,
                    stacklevel=3,
                )
                warned = True
            keys[i] = item
        else:
            keys[i] = item
            all_bytes = False

    labels = list(range(256))  # symbol -> edge label object
    offsets = np.zeros(V + 1, dtype=np.int64)
    if all_bytes:
        if V:
            offsets[1:] = np.cumsum(np.fromiter((len(x) for x in decode), dtype=np.int64, count=V))
        symbols = np.frombuffer(b"".join(decode), dtype=np.uint8).astype(np.int32)
    else:
        table = {i: i for i in range(256)}
        syms = []
        for i, item in enumerate(decode):
            word = item.byte_string if isinstance(item, Token) else item
            for letter in word:
                s = table.get(letter)
                if s is None:
                    s = len(labels)
                    table[letter] = s
                    labels.append(letter)
                syms.append(s)
            offsets[i + 1] = len(syms)
        symbols = np.asarray(syms, dtype=np.int32)
    return symbols, offsets, keys, labels


class TokenCharacterTrie:
    """A trie data structure for efficient token-to-character mapping.

    Deviations from the reference class, all on inputs its own tests do not use:

    * there is no CPU path: ``weight_sum`` / ``weight_max`` raise without a CUDA device (the reference's numba loops run
      anywhere); building the trie and every layout attribute work on any host;
    * ``weight_max`` is a true maximum (identity ``-inf``): the reference's numba loop starts every internal node at 0
      (``base.py:387-393``), so it clamps negative weights -- e.g. log-weights -- at 0 where this class returns their
      maximum.  For non-negative weights (probabilities) the results are identical, bit for bit;
    * an empty vocabulary gives mass 0 at the root for both reductions.
    """

    def __init__(self, decode):
        """
        Args:
            decode (list): the token vocabulary.  Each element must be iterable: ``Token`` objects iterate
                their ``byte_string``; other iterables (bytes, an end-of-sequence sentinel) are iterated directly.
        """
        self.decode = decode
        symbols, offsets, keys, labels = _encode_vocabulary(decode)
        self._edge_labels = labels

        # the reference raises on the first repeated word key while inserting (base.py:63-64)
        word2pos = {}
        for i, key in enumerate(keys):
            if key in word2pos:
                raise ValueError(f"Duplicate word in vocabulary: {key}")
            word2pos[key] = i

        self._engine = TrieEngine(symbols, offsets, len(decode))
        lay = self._engine.layout()
        self._layout = lay
        N, V = self._engine.N, self._engine.V

        self.root = N - 1
        # idx_to_leaf[k] = (position in decode, leaf node id)    (base.py:115-117)
        self.idx_to_leaf = np.stack([np.arange(V, dtype=np.int32), lay["leaf_node"]], axis=1) if V else np.zeros((0, 2), np.int32)
        leaf_node = lay["leaf_node"].tolist()
        self.word2leaf = {key: leaf_node[i] for key, i in word2pos.items()}
        self.leaf2word = dict(zip(self.word2leaf.values(), self.word2leaf.keys()))
        # internal nodes in ascending (topological) order    (base.py:74, 119)
        self.ordering = np.flatnonzero(np.diff(lay["child_ptr"]) > 0).astype(np.int64)

        self._children = None
        self._jump = None
        self._node2prefix = None

    # ---- lazily materialised reference attributes ----------------------------------------------------
    def __len__(self):
        return self._engine.N

    @property
    def children(self):
        """``children[node]``: dict edge label -> child id; leaf edges are keyed ``(None, position)``."""
        if self._children is None:
            lay = self._layout
            ptr = lay["child_ptr"].tolist()
            idx = lay["child_idx"].tolist()
            lab = lay["edge_label"].tolist()
            labels = self._edge_labels
            out = []
            for n in range(self._engine.N):
                d = {}
                for c in idx[ptr[n]:ptr[n + 1]]:
                    lc = lab[c]
                    d[(None, -1 - lc) if lc < 0 else labels[lc]] = c
                out.append(d)
            self._children = out
        return self._children

    @property
    def jump(self):
        """``jump[node]``: sorted int32 array of child ids (``base.py:120-122``)."""
        if self._jump is None:
            lay = self._layout
            self._jump = np.split(lay["child_idx"], lay["child_ptr"][1:-1])
        return self._jump

    @property
    def node2prefix(self):
        """``node2prefix[node]``: list of edge labels from the root (a leaf shares its parent's prefix)."""
        if self._node2prefix is None:
            lay = self._layout
            ptr = lay["child_ptr"].tolist()
            idx = lay["child_idx"].tolist()
            lab = lay["edge_label"].tolist()
            labels = self._edge_labels
            prefix = {self.root: []}
            for x in range(self._engine.N - 1, -1, -1):  # parents have larger ids than their children
                px = prefix[x]
                for c in idx[ptr[x]:ptr[x + 1]]:
                    lc = lab[c]
                    prefix[c] = px if lc < 0 else px + [labels[lc]]
            self._node2prefix = prefix
        return self._node2prefix

    # ---- weights ---------------------------------------------------------------------------------------
    def _alloc_weights(self):
        return np.zeros(self._engine.N, dtype=np.float64)

    def _preprocess_ws(self, ws):
        """Token weights -> numpy array (``base.py:132-145``)."""
        if isinstance(ws, torch.Tensor):
            if ws.device.type != "cpu":
                ws = ws.cpu()
            if ws.dtype == torch.bfloat16:
                ws = ws.to(torch.float32)
            ws = ws.numpy()
        return ws

    def _rows_to_device(self, rows):
        """List of 1-D weight vectors -> ``[B, V]`` CUDA tensor (fp16/fp32/fp64 kept, anything else -> fp64)."""
        require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        tens = []
        for ws in rows:
            if isinstance(ws, torch.Tensor):
                t = ws
            else:
                a = np.asarray(ws)
                if a.dtype not in (np.float16, np.float32, np.float64):
                    a = a.astype(np.float64)
                t = torch.from_numpy(np.ascontiguousarray(a))
            if t.dim() != 1 or t.shape[0] != len(self.decode):
                raise AssertionError([tuple(t.shape), len(self.decode)])
            if t.dtype not in (torch.float16, torch.bfloat16, torch.float32, torch.float64):
                t = t.to(torch.float64)
            tens.append(t.to(dev, non_blocking=True))
        dtypes = {t.dtype for t in tens}
        if len(dtypes) > 1:
            tens = [t.to(torch.float64) for t in tens]
        return torch.stack(tens) if tens else torch.empty((0, len(self.decode)), dtype=torch.float64, device=dev)

    def _reduce64(self, rows, op):
        ws = self._rows_to_device(rows)
        out_sum, out_max = self._engine.reduce(ws, (op,), out_dtype=torch.float64)
        out = out_sum if op == "sum" else out_max
        # C-contiguous float64 [B, N] like the reference's arrays: one pitched copy out of the padded device slab
        host = torch.empty((out.shape[0], self._engine.N), dtype=torch.float64, pin_memory=out.shape[0] > 0)
        if out.shape[0]:
            stream = torch.cuda.current_stream(out.device.index)
            self._engine.download(out, host, stream)
            stream.synchronize()
        return host.numpy()

    def weight_sum(self, ws):
        """Sum of the weights of all tokens below each node.  Returns ``float64[num_nodes]``."""
        return self._reduce64([ws], "sum")[0]

    def weight_max(self, ws):
        """Maximum weight among the tokens below each node.  Returns ``float64[num_nodes]``."""
        return self._reduce64([ws], "max")[0]

    def batch_weight_sum(self, ws):
        """Batched ``weight_sum``: ``float64[len(ws), num_nodes]``."""
        return self._reduce64(list(ws), "sum")

    def batch_weight_max(self, ws):
        """Batched ``weight_max``: ``float64[len(ws), num_nodes]``."""
        return self._reduce64(list(ws), "max")

    # ---- visualisation ---------------------------------------------------------------------------------
    def visualize(self, ws=None):
        """Render the trie with Graphviz; ``ws`` optionally gives one value per node (``base.py:249-343``)."""
        try:
            import graphviz
        except ImportError:  # pragma: no cover
            raise ImportError("Please install graphviz: pip install graphviz")  # pragma: no cover

        n_nodes = self._engine.N
        if ws is not None and len(ws) != n_nodes:
            raise ValueError(f"Weight vector length ({len(ws)}) must match number of nodes ({n_nodes})")

        dot = graphviz.Digraph(comment="Token Character Trie")
        dot.attr(rankdir="LR")
        with dot.subgraph(name="cluster_legend") as legend:
            legend.attr(label="Legend", fontsize="10")
            legend.attr("node", fontsize="7", width="0.1", height="0.1")
            legend.node("legend_internal", "Internal Node ID\n'Prefix'\nWeight (if provided)", shape="circle")
            legend.node("legend_leaf", "Complete Token", shape="doublecircle")
            legend.edge("legend_internal", "legend_leaf", label="Token item", fontsize="10")
            legend.attr(rankdir="TB")
            legend.attr(rank="same")

        top = float(max(ws)) if ws is not None and n_nodes else 0.0
        for node in range(n_nodes):
            label = f"{node}\n'{self.node2prefix[node]}'"
            fill = "#ffffff"
            if ws is not None:
                label += f"\n{float(ws[node]):.4f}"
                if top > 0:
                    shade = int(255 * (1 - float(ws[node]) / top))
                    fill = f"#{shade:02x}ff{shade:02x}"
            shape = "doublecircle" if node in self.leaf2word else "circle"
            dot.node(str(node), label, shape=shape, style="filled", fillcolor=fill)
        for node, kids in enumerate(self.children):
            for key, child in kids.items():
                leaf_edge = isinstance(key, tuple) and key[0] is None
                dot.edge(str(node), str(child), label=f"End-of-Token (ID: {key[1]})" if leaf_edge else str(key))
        return dot
