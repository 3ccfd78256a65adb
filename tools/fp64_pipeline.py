"""Device time of the fp64 pipeline (TokenCharacterTrie: fp64 rows in, fp64 node masses out) at BASELINE config 2's shape,
from CUDA-graph replays rotating over buffer sets larger than L2."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from genlm_backend_b200 import TokenCharacterTrie
from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows

V, B = 128256, int(sys.argv[1]) if len(sys.argv) > 1 else 64
trie = TokenCharacterTrie(synth_vocab(V))
eng, N = trie._engine, len(trie)
sets = 3
base = dirichlet_rows(B, V, alpha=1.0, seed=1).astype(np.float64)
ws = [torch.tensor(np.roll(base, k, axis=0)).cuda() for k in range(sets)]
dev = torch.device("cuda", 0)
osum = [eng.alloc_out(B, torch.float64, dev) for _ in range(sets)]
omax = [eng.alloc_out(B, torch.float64, dev) for _ in range(sets)]
for ops in (("sum",), ("sum", "max")):
    def step(k):
        eng.reduce(ws[k], ops, out_dtype=torch.float64, out_sum=osum[k], out_max=omax[k] if len(ops) > 1 else None)
    for k in range(sets):
        step(k)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for k in range(4 * sets):
            step(k % sets)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) / (20 * 4 * sets) * 1e3
    nbytes = B * (8 * V + 8 * N * len(ops))
    print(f"fp64 pipeline, {B} rows x {V} tokens, {'+'.join(ops)}: {us:7.1f} us per step, {B / us * 1e6:10.0f} distributions/s, "
          f"{nbytes / us * 1e-3:6.0f} GB/s algorithmic ({nbytes / us * 1e-3 / 6551.7:.2f} of the measured HBM peak)")
