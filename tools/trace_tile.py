"""Prints the timeline of tile_kernel's pipeline (GT_TRACE=1): per item, the SM-clock durations between the events
stamped by thread 0.

  0 item start | 1 rows landed (A) | 2 scatter done | 3 next fetch issued | 4 pyramid done | 5 ELL done | 6 output staged + issued
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
ev = {"wait rows(A)": (0, 1), "scatter+bar": (1, 2), "issue fetch": (2, 3), "pyramid+bar": (3, 4), "ELL (own share)": (4, 22),
      "wait_read+bar": (22, 5), "head+pieces": (5, 6),
      "c0 gather": (6, 7), "c0 wait_read": (7, 8), "c0 bar": (8, 9), "c0 issue": (9, 10),
      "c1 gather": (10, 11), "c1 wait_read": (11, 12), "c1 bar": (12, 13), "c1 issue": (13, 14),
      "c2 gather": (14, 15), "c2 wait_read": (15, 16), "c2 bar": (16, 17), "c2 issue": (17, 18),
      "item total": (0, 23)}
valid = tr[:, :, 23] > 0
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
        print(f"  {name:16s} mean {d.mean():8.0f}  p50 {np.median(d):8.0f}  p90 {np.percentile(d, 90):8.0f} cycles  (n={ok.sum()})")
span = tr[:, :, 23].max(axis=1) - np.where(valid, tr[:, :, 0], np.inf).min(axis=1)
span = span[valid.any(axis=1)]
print(f"CTA lifetime (first event -> last output): mean {span.mean():.0f}  max {span.max():.0f} cycles")
