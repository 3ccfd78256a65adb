"""bf16 rows through the sampler kernel for an ncu capture (no mask, per-row bit mask)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from genlm_backend_b200 import smc

B, V = 512, 128256
torch.manual_seed(0)
logp = torch.log_softmax(torch.randn(B, V, device="cuda"), dim=-1).to(torch.bfloat16)
for rep in range(3):
    smc.masked_logsumexp_sample(logp, None, seed=1)
torch.cuda.synchronize()
