"""BASELINE config 3: AsyncTokenCharacterTrie autobatched weight_sum, R concurrent requests at 128,256 tokens.
Each request is one float32 row (a CPU tensor, as the reference's callers pass it); every future resolves to its own
float32 numpy row of node masses.  Wall-clock rate of the whole gather, best of a few repetitions.

    python tools/async_bench.py [requests] [devices]
"""
import asyncio
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from genlm_backend_b200 import AsyncTokenCharacterTrie
from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows

V = 128256
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ndev = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kw = {"devices": list(range(ndev))} if ndev > 1 else {}
at = AsyncTokenCharacterTrie.from_vocab(synth_vocab(V), backend="parallel", **kw)
N = len(at.trie)
base = dirichlet_rows(min(R, 256), V, alpha=1.0, seed=1)
rows = [torch.tensor(base[i % len(base)]) for i in range(R)]
dev_rows = [r.cuda() for r in rows]


async def once(reqs):
    t0 = time.perf_counter()
    out = await asyncio.gather(*[at.weight_sum(r) for r in reqs])
    dt = time.perf_counter() - t0
    assert len(out) == len(reqs) and out[0].shape == (N,)
    return dt, float(out[-1][N - 1])


async def main():
    for name, reqs in (("CPU rows", rows), ("CUDA rows", dev_rows)):
        best = 1e9
        for rep in range(4):
            dt, root = await once(reqs)
            best = min(best, dt)
        print(f"{name}: {R} concurrent weight_sum requests on {ndev} GPU(s): best {best * 1e3:8.1f} ms  "
              f"{R / best:9.0f} distributions/s  (root mass {root:.4f})", flush=True)
    await at.cleanup()


asyncio.run(main())
