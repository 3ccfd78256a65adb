"""Prints the timeline of tile_kernel's pipeline (GT_TRACE=1): per item, the SM-clock durations between the events
stamped by thread 0.

  compute group leader: 0 start | 1 value array free | 2 rows landed | 3 scatter done | 4 (fetch issued) | 5 pyramid done |
  6 terms ready + barrier | 7 ELL done;   emit group leader: 8 start | 9 value array full | 10 emit done
"""
import ctypes
import os
import sys

os.environ["GT_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from genlm_backend_b200 import ParallelTokenCharacterTrie, _lib
from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows

V, B = 128256, 64
trie = ParallelTokenCharacterTrie(synth_vocab(V))
eng = trie._engine
N = len(trie)
sets = 4
base = dirichlet_rows(B, V, alpha=1.0, seed=1)
ws = [torch.tensor(np.roll(base, k, axis=0)).cuda() for k in range(sets)]
osum = [eng.alloc_out(B, torch.float32, torch.device("cuda", 0)) for _ in range(sets)]
for i in range(8):
    eng.reduce(ws[i % sets], ("sum",), out_sum=osum[i % sets])
torch.cuda.synchronize()
dims = (ctypes.c_int32 * 3)()
n = _lib.lib.gt_debug_read_trace(eng._handle, 0, None, 0, dims)
buf = np.zeros(n, dtype=np.int64)
_lib.lib.gt_debug_read_trace(eng._handle, 0, buf.ctypes.data, n, dims)  # clears: drop the warm-up launches
eng.reduce(ws[0], ("sum",), out_sum=osum[0])
torch.cuda.synchronize()
_lib.lib.gt_debug_read_trace(eng._handle, 0, buf.ctypes.data, n, dims)
tr = buf.reshape(dims[0], dims[1], dims[2]).astype(np.float64)
ev = {"C wait empty": (0, 1), "C wait rows(A)": (1, 2), "C scatter+bar": (2, 3), "C issue fetch": (3, 4), "C pyramid": (4, 5),
      "C wait terms+bar": (5, 6), "C ELL": (6, 7), "C item total": (0, 7),
      "E wait full": (8, 9), "E emit loop": (9, 10)}
valid = tr[:, :, 7] > 0
print("CTAs with items:", int(valid.any(axis=1).sum()), " items traced:", int(valid.sum()))
for first in (True, False):
    sel = valid.copy()
    if first:
        sel[:, 1:] = False
    else:
        sel[:, 0] = False
    if not sel.any():
        continue
    print("first item of a CTA" if first else "later items")
    for name, (a, b) in ev.items():
        ok = sel & (tr[:, :, a] > 0) & (tr[:, :, b] > 0)
        if not ok.any():
            continue
        d = (tr[:, :, b] - tr[:, :, a])[ok]
        print(f"  {name:18s} mean {d.mean():8.0f}  p50 {np.median(d):8.0f}  p90 {np.percentile(d, 90):8.0f} cycles  (n={ok.sum()})")
span = tr[:, :, 10].max(axis=1) - np.where(valid, tr[:, :, 0], np.inf).min(axis=1)
span = span[valid.any(axis=1)]
print(f"CTA lifetime (first event -> last emit): mean {span.mean():.0f}  max {span.max():.0f} cycles")
c = 10
print("CTA 10 timeline (cycles since its first event):")
t0 = tr[c, 0, 0]
for k in range(dims[1]):
    if tr[c, k, 7] > 0:
        print("  item", k, " ".join(f"{int(x - t0):7d}" for x in tr[c, k, :11]))
