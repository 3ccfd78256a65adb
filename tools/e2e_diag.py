"""Where does the host<->device time of the reference-shaped API go?  (diagnostic, run on the GPU box)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from genlm_backend_b200 import ParallelTokenCharacterTrie
from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows

V, B = 128256, 64
trie = ParallelTokenCharacterTrie(synth_vocab(V))
N = len(trie)
host = torch.tensor(dirichlet_rows(B, V, alpha=1.0, seed=1)).pin_memory()


def t(fn, n=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


dev_in = torch.empty((B, V), device="cuda")
dev_out = torch.empty((B, N), device="cuda")
pin_out = torch.empty((B, N), pin_memory=True)
print("H2D 33MB pinned  ms", t(lambda: dev_in.copy_(host, non_blocking=True)), "GB/s", B * V * 4 / 1e6 / t(lambda: dev_in.copy_(host, non_blocking=True)))
print("D2H 88MB pinned  ms", t(lambda: pin_out.copy_(dev_out, non_blocking=True)), "GB/s", B * N * 4 / 1e6 / t(lambda: pin_out.copy_(dev_out, non_blocking=True)))
print("alloc pinned 88MB ms", t(lambda: torch.empty((B, N), pin_memory=True)))
print("alloc+free pageable 88MB + touch ms", t(lambda: np.empty((B, N), np.float32).fill(0)))
print("D2H 88MB to pageable (.cpu()) ms", t(lambda: dev_out.cpu()))
print("api sum      ms", t(lambda: trie.batch_weight_sum(host)))
print("api sum+max  ms", t(lambda: trie.batch_weight_sum_max(host)))
print("api sum from device rows ms", t(lambda: trie.batch_weight_sum(dev_in)))
s0 = torch.cuda.Stream(); s1 = torch.cuda.Stream()
def duplex():
    with torch.cuda.stream(s0):
        dev_in.copy_(host, non_blocking=True)
    with torch.cuda.stream(s1):
        pin_out.copy_(dev_out, non_blocking=True)
print("duplex H2D 33MB + D2H 88MB ms", t(duplex))

# ---- bench-like loop: results retained, inputs alternate; with and without an NVML sampling thread -------------
import threading
host2 = [host, torch.tensor(dirichlet_rows(B, V, alpha=1.0, seed=2)).pin_memory()]
def loop(n=10):
    keep = None
    for i in range(4):
        keep = trie.batch_weight_sum_max(host2[i % 2])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        sums, maxes = trie.batch_weight_sum_max(host2[i % 2])
        _ = float(sums[0, N - 1]) + float(maxes[B - 1, N - 1])
    return (time.perf_counter() - t0) / n * 1e3
print("bench-like e2e loop, no sampler ms/step", loop())
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
stop = threading.Event()
def sampler(period):
    while not stop.is_set():
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        time.sleep(period)
t0 = time.perf_counter(); pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); print("one nvml clock query ms", (time.perf_counter() - t0) * 1e3)
for period in (0.02, 0.2):
    stop.clear(); th = threading.Thread(target=sampler, args=(period,), daemon=True); th.start()
    print(f"bench-like e2e loop, nvml sampler every {period}s ms/step", loop())
    stop.set(); th.join()
