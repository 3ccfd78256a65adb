"""Where a one-row weight_sum call spends its time (BASELINE config 0 shape: 50,257 tokens): host time of each stage of
ParallelTokenCharacterTrie._small_to_host, and the device time of the three kernels."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from genlm_backend_b200 import ParallelTokenCharacterTrie
from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows

V = 50257
trie = ParallelTokenCharacterTrie(synth_vocab(V))
eng, N = trie._engine, len(trie)
row = torch.tensor(dirichlet_rows(1, V, alpha=1.0, seed=1)[0])
drow = row.cuda().reshape(1, -1)
for _ in range(20):
    trie.weight_sum(row)
torch.cuda.synchronize()
st = torch.cuda.current_stream()
acc = {k: [] for k in ("to_device", "reduce", "pinned_alloc", "download", "sync", "total_cpu_row", "total_cuda_row")}
for _ in range(300):
    t0 = time.perf_counter()
    x = row.reshape(1, -1).to("cuda", non_blocking=True)
    t1 = time.perf_counter()
    o, _ = eng.reduce(x, ("sum",))
    t2 = time.perf_counter()
    h = torch.empty((1, N), dtype=torch.float32, pin_memory=True)
    t3 = time.perf_counter()
    eng.download(o, h, st)
    t4 = time.perf_counter()
    st.synchronize()
    t5 = time.perf_counter()
    for k, v in zip(("to_device", "reduce", "pinned_alloc", "download", "sync"), (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
        acc[k].append(v)
    t0 = time.perf_counter(); trie.weight_sum(row); acc["total_cpu_row"].append(time.perf_counter() - t0)
    t0 = time.perf_counter(); trie.weight_sum(drow[0]); acc["total_cuda_row"].append(time.perf_counter() - t0)
for k, v in acc.items():
    print(f"{k:16s} median {np.median(v) * 1e6:7.1f} us")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200):
    eng.reduce(drow, ("sum",))
b.record()
torch.cuda.synchronize()
print(f"kernels (back to back, device time per call) {a.elapsed_time(b) / 200 * 1e3:7.1f} us")
h = torch.empty((1, N), dtype=torch.float32, pin_memory=True)
o, _ = eng.reduce(drow, ("sum",))
torch.cuda.synchronize()
a.record()
for _ in range(200):
    eng.download(o, h, st)
b.record()
torch.cuda.synchronize()
print(f"download of {N * 4 / 1e3:.0f} KB (back to back, device time per copy) {a.elapsed_time(b) / 200 * 1e3:7.1f} us")
