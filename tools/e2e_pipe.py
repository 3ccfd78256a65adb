"""End-to-end rate of the reference-shaped API (pinned host rows in, numpy out) for several pipeline slicings.
    python tools/e2e_pipe.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from genlm_backend_b200 import ParallelTokenCharacterTrie
from genlm_backend_b200.trie import parallel as par
from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows

V, B = 128256, 64
trie = ParallelTokenCharacterTrie(synth_vocab(V))
N = len(trie)
base = dirichlet_rows(B, V, alpha=1.0, seed=1)
host = [torch.tensor(np.roll(base, k, axis=0)).pin_memory() for k in range(2)]


def rate(n=12):
    for i in range(4):
        s, m = trie.batch_weight_sum_max(host[i % 2])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        s, m = trie.batch_weight_sum_max(host[i % 2])
        _ = float(s[0, N - 1]) + float(m[B - 1, N - 1])
    torch.cuda.synchronize()
    return B * n / (time.perf_counter() - t0)


for rep in range(2):
    for first, rows, slots in ((32, 32, 3), (8, 32, 3), (16, 32, 3), (8, 16, 3), (8, 16, 4), (4, 8, 4), (64, 64, 2)):
        par._PIPE_FIRST, par._PIPE_ROWS, par._PIPE_SLOTS = first, rows, slots
        trie._streams = {}
        print(f"first {first:3d} rows {rows:3d} slots {slots}: {rate():9.0f} distributions/s", flush=True)
