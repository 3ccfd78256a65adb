import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from genlm_backend_b200 import ParallelTokenCharacterTrie, _lib
from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows
V, B = 128256, 64
trie = ParallelTokenCharacterTrie(synth_vocab(V)); eng = trie._engine
sets = 4
base = dirichlet_rows(B, V, alpha=1.0, seed=1)
ws = [torch.tensor(np.roll(base, k, axis=0)).cuda() for k in range(sets)]
osum = [eng.alloc_out(B, torch.float32, torch.device("cuda", 0)) for _ in range(sets)]
for k in range(sets): eng.reduce(ws[k], ("sum",), out_sum=osum[k])
torch.cuda.synchronize()
def t(phases):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for q in range(16): eng.reduce(ws[q % sets], ("sum",), out_sum=osum[q % sets], phases=phases)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(20): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 320 * 1e3
print(os.environ.get("GT_LIB_NAME"), os.environ.get("GT_SEG_POSITIONS"), "P %.2f  PTS(sum) %.2f" % (t(_lib.GT_FLAG_PHASE_PERMUTE), t(0)))
