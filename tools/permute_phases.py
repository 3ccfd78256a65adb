"""Permute kernel alone, replayed from a CUDA graph (16 launches per graph).  GT_DEBUG_STOP=21: rows are fetched but not
gathered / stored; 22: gather / store only (no row fetches); unset: the real kernel.

    for m in 0 21 22; do GT_DEBUG_STOP=$m python tools/permute_phases.py; done
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from genlm_backend_b200 import ParallelTokenCharacterTrie, _lib
from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows

V, B = 128256, int(sys.argv[1]) if len(sys.argv) > 1 else 64
trie = ParallelTokenCharacterTrie(synth_vocab(V))
eng = trie._engine
sets = 4
base = dirichlet_rows(B, V, alpha=1.0, seed=1)
ws = [torch.tensor(np.roll(base, k, axis=0)).cuda() for k in range(sets)]
osum = [eng.alloc_out(B, torch.float32, torch.device("cuda", 0)) for _ in range(sets)]
eng.reduce(ws[0], ("sum",), out_sum=osum[0])
torch.cuda.synchronize()


def t(phases, per=16):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for q in range(per):
            eng.reduce(ws[q % sets], ("sum",), out_sum=osum[q % sets], phases=phases)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(20):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (20 * per) * 1e3


print("GT_DEBUG_STOP=%s  batch %d  permute %.2f us" % (os.environ.get("GT_DEBUG_STOP", "0"), B, t(_lib.GT_FLAG_PHASE_PERMUTE)))
