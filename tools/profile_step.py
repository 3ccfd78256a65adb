"""Small driver for ncu: a few direct (un-graphed) steps of the bench workload, plus the sampler kernel."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from genlm_backend_b200 import ParallelTokenCharacterTrie, smc
from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows, logsoftmax_rows

ap = argparse.ArgumentParser()
ap.add_argument("--vocab", type=int, default=128256)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--sampler-rows", type=int, default=0)
ap.add_argument("--phases", default="all", help="comma list of launch groups per step: all, permute, tile")
args = ap.parse_args()

trie = ParallelTokenCharacterTrie(synth_vocab(args.vocab))
N = len(trie)
sets = 4
base = dirichlet_rows(args.batch, args.vocab, alpha=1.0, seed=1)
ws = [torch.tensor(np.roll(base, k, axis=0)).cuda() for k in range(sets)]
osum = [trie._engine.alloc_out(args.batch, torch.float32, torch.device("cuda", 0)) for _ in range(sets)]
omax = [trie._engine.alloc_out(args.batch, torch.float32, torch.device("cuda", 0)) for _ in range(sets)]
from genlm_backend_b200 import _lib

PH = {"all": 0, "permute": _lib.GT_FLAG_PHASE_PERMUTE, "tile": _lib.GT_FLAG_PHASE_TILE, "span": _lib.GT_FLAG_PHASE_SPAN}
for i in range(args.steps):
    k = i % sets
    for ph in args.phases.split(","):
        trie._engine.reduce(ws[k], ("sum", "max"), out_sum=osum[k], out_max=omax[k], phases=PH[ph])
torch.cuda.synchronize()
if args.sampler_rows:
    logp = torch.tensor(logsoftmax_rows(args.sampler_rows, args.vocab, seed=0)).cuda()
    mask = (torch.rand(args.vocab, device="cuda") < 0.5).float().log()
    for i in range(3):
        smc.masked_logsumexp_sample(logp, mask, seed=i)
    torch.cuda.synchronize()
print("done", N)
