"""Throughput of the fused masked logsumexp + sampling kernel (BASELINE.json config 4 shape per GPU: 512 rows x 128,256
vocab).  Device time from CUDA events over a CUDA graph, rotating over two 263 MB logit buffers (> L2)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from genlm_backend_b200 import smc

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=512)
ap.add_argument("--vocab", type=int, default=128256)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
B, V = args.rows, args.vocab
torch.manual_seed(0)
sets = [torch.log_softmax(torch.randn(B, V, device="cuda"), dim=-1) for _ in range(2)]
shared = (torch.rand(V, device="cuda") < 0.5).float().log()
per_row = (torch.rand(B, V, device="cuda") < 0.5).float().log()
bool_mask = torch.rand(V, device="cuda") < 0.5
bool_rows = torch.rand(B, V, device="cuda") < 0.5


def pack_bits(keep):
    """bool [..., V] -> int32 [..., ceil(V/32)] keep-bitmask (bit i of word w = element 32*w + i)"""
    pad = (-keep.shape[-1]) % 32
    k = torch.nn.functional.pad(keep, (0, pad)).view(*keep.shape[:-1], -1, 32).to(torch.int64)
    w = (k << torch.arange(32, device=keep.device, dtype=torch.int64)).sum(-1)
    return torch.where(w >= 2**31, w - 2**32, w).to(torch.int32)


W = (V + 31) // 32
cases = {"no mask": (None, 4 * V), "shared additive fp32 mask": (shared, 4 * V), "shared bool mask": (bool_mask, 4 * V),
         "shared bit mask": (pack_bits(bool_mask), 4 * V),
         "per-row additive fp32 mask": (per_row, 8 * V), "per-row bool mask": (bool_rows, 5 * V),
         "per-row bit mask": (pack_bits(bool_rows), 4 * V + 4 * W)}
for dtype in (torch.float32, torch.bfloat16):
    data = [s.to(dtype) for s in sets]
    for name, (mask, bytes_per_row) in cases.items():
        bpr = bytes_per_row - 4 * V + data[0].element_size() * V
        for i in range(3):
            smc.masked_logsumexp_sample(data[i % 2], mask, seed=i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for k in range(2):
                smc.masked_logsumexp_sample(data[k], mask, seed=k)
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / (2 * args.iters) * 1e3
        print(f"{str(dtype):15s} {name:28s} {us:8.1f} us  {B / us * 1e6:12.0f} rows/s  {B * bpr / us * 1e-3:7.0f} GB/s")
