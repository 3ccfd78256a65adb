"""Prints the timeline of mass_kernel's pipelines (GT_TRACE=1): SM-clock stamps per CTA.

  compute group leader, per item:  0 start | 1 rest buffer free | 2 pyramid done | 3 barrier | 4 ELL done
  emit group leader, per item:     8 start | 9 rest buffer full | 10 emit done
  [cta][last item][11]: CTA start

    python tools/trace_tile.py [all|permute|tile]
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

mode = sys.argv[1] if len(sys.argv) > 1 else "all"
PH = {"all": 0, "permute": _lib.GT_FLAG_PHASE_PERMUTE, "tile": _lib.GT_FLAG_PHASE_TILE}[mode]
V, B = 128256, 64
trie = ParallelTokenCharacterTrie(synth_vocab(V))
eng = trie._engine
sets = 4
base = dirichlet_rows(B, V, alpha=1.0, seed=1)
ws = [torch.tensor(np.roll(base, k, axis=0)).cuda() for k in range(sets)]
osum = [eng.alloc_out(B, torch.float32, torch.device("cuda", 0)) for _ in range(sets)]
omax = [eng.alloc_out(B, torch.float32, torch.device("cuda", 0)) for _ in range(sets)]
for i in range(8):
    eng.reduce(ws[i % sets], ("sum", "max"), out_sum=osum[i % sets], out_max=omax[i % sets])
torch.cuda.synchronize()
dims = (ctypes.c_int32 * 3)()
n = _lib.lib.gt_debug_read_trace(eng._handle, 0, None, 0, dims)
buf = np.zeros(n, dtype=np.int64)
_lib.lib.gt_debug_read_trace(eng._handle, 0, buf.ctypes.data, n, dims)  # clears: drop the warm-up launches
eng.reduce(ws[0], ("sum", "max"), out_sum=osum[0], out_max=omax[0], phases=PH)
torch.cuda.synchronize()
_lib.lib.gt_debug_read_trace(eng._handle, 0, buf.ctypes.data, n, dims)
tr = buf.reshape(dims[0], dims[1], dims[2]).astype(np.float64)
start = tr[:, -1, 11]
live = start > 0
print("mode", mode, " CTAs:", int(live.sum()))


def stats(name, d):
    if d.size:
        print(f"  {name:34s} mean {d.mean():8.0f}  p50 {np.median(d):8.0f}  p90 {np.percentile(d, 90):8.0f}  max {d.max():8.0f} cycles (n={d.size})")


def dur(a, b, items=slice(None)):
    x, y = tr[:, items, a], tr[:, items, b]
    ok = (x > 0) & (y > 0)
    return (y - x)[ok]


print("compute / emit groups (per item)")
for name, (a, b) in {"C wait rest buffer": (0, 1), "C pyramid": (1, 2), "C barrier": (2, 3), "C ELL": (3, 4), "C item": (0, 4),
                     "E wait full": (8, 9), "E emit": (9, 10), "E item": (8, 10)}.items():
    stats(name + " (first item)", dur(a, b, slice(0, 1)))
    stats(name + " (later items)", dur(a, b, slice(1, -1)))
print("start-up (cycles since CTA start)")
for name, ev in {"first pair fetched": 0}.items():
    x = tr[:, -1, ev]
    stats(name, (x - start)[live & (x > 0)])
x = tr[:, 0, 1]
stats("compute: first pair landed + rest buffer free", (x - start)[live & (x > 0)])
first_emit = tr[:, 0, 9]
ok = live & (first_emit > 0)
if ok.any():
    stats("CTA start -> first emit begins", (first_emit - start)[ok])
last = tr[:, :-1, 10].max(axis=1)
ok = live & (last > 0)
stats("CTA start -> last event", (last - start)[ok])
t0 = start[live].min()
print(f"grid: first CTA start -> last event anywhere: {last[ok].max() - t0:.0f} cycles; CTA start spread {start[live].max() - t0:.0f}")
c = 10
print("CTA 10 item timeline (cycles since CTA start): ev0 ev1 ev2 ev3 ev4 | ev8 ev9 ev10")
for k in range(dims[1] - 1):
    if tr[c, k, 4] > 0 or tr[c, k, 10] > 0:
        print("  item", k, " ".join(f"{int(tr[c, k, e] - start[c]):7d}" for e in (0, 1, 2, 3, 4, 8, 9, 10)))
