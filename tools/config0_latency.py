"""BASELINE config 0: weight_sum + weight_max of ONE distribution over a GPT-2-sized byte vocabulary (50,257 tokens),
through the reference-shaped classes (host tensor in, numpy out).  Per-call wall-clock latency, median of many calls.

    python tools/config0_latency.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from genlm_backend_b200 import ParallelTokenCharacterTrie, TokenCharacterTrie
from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows

V = 50257
dec = synth_vocab(V)
ws = torch.tensor(dirichlet_rows(1, V, alpha=1.0, seed=1)[0])
ws_dev = ws.cuda()


def lat(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return np.median(ts) * 1e6, np.percentile(ts, 90) * 1e6


for cls in (ParallelTokenCharacterTrie, TokenCharacterTrie):
    t0 = time.perf_counter()
    trie = cls(dec)
    build = time.perf_counter() - t0
    print(f"{cls.__name__}: build {build * 1e3:.0f} ms, N = {len(trie)}")
    for name, fn in (("weight_sum(cpu tensor)", lambda: trie.weight_sum(ws)), ("weight_max(cpu tensor)", lambda: trie.weight_max(ws)),
                     ("weight_sum(cuda tensor)", lambda: trie.weight_sum(ws_dev)), ("weight_max(cuda tensor)", lambda: trie.weight_max(ws_dev)),
                     ("sum + max (cuda tensor)", lambda: (trie.weight_sum(ws_dev), trie.weight_max(ws_dev))),
                     ("sum + max (cpu tensor)", lambda: (trie.weight_sum(ws), trie.weight_max(ws)))):
        m, p90 = lat(fn)
        print(f"  {name:26s} median {m:8.1f} us   p90 {p90:8.1f} us")
