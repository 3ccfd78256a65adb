"""Launches the sampler kernel variants once each (after a warm-up) for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:lse_sample -c 8 -o gpurun_out/prof_sampler python tools/profile_sampler.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from genlm_backend_b200 import smc

B, V = 512, 128256
torch.manual_seed(0)
logp = torch.log_softmax(torch.randn(B, V, device="cuda"), dim=-1)
keep = torch.rand(B, V, device="cuda") < 0.5
pad = (-V) % 32
k = torch.nn.functional.pad(keep, (0, pad)).view(B, -1, 32).to(torch.int64)
w = (k << torch.arange(32, device="cuda", dtype=torch.int64)).sum(-1)
bits = torch.where(w >= 2**31, w - 2**32, w).to(torch.int32)
premasked = torch.where(keep, logp, torch.full_like(logp, float("-inf")))
for rep in range(2):
    smc.masked_logsumexp_sample(logp, None, seed=1)          # no mask
    smc.masked_logsumexp_sample(premasked, None, seed=1)     # no mask, half the entries already -inf
    smc.masked_logsumexp_sample(logp, bits, seed=1)          # per-row bit mask
    smc.masked_logsumexp_sample(logp, keep, seed=1)          # per-row bool mask
torch.cuda.synchronize()
