"""Where does the step time go between the kernels?  Times sub-chains of one step (P = permute, T = tile kernel for both
reductions) replayed from CUDA graphs that hold `per_graph` steps each, and the full step launched
directly on the stream, so that the per-graph launch bubble and the inter-kernel gaps can be separated.

    python tools/chain_times.py [batch]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from genlm_backend_b200 import ParallelTokenCharacterTrie, _lib
from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows

V = int(os.environ.get("GT_VOCAB", "128256"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
trie = ParallelTokenCharacterTrie(synth_vocab(V))
eng = trie._engine
dev = torch.device("cuda", 0)
sets = 4
base = dirichlet_rows(B, V, alpha=1.0, seed=1)
ws = [torch.tensor(np.roll(base, k, axis=0)).cuda() for k in range(sets)]
osum = [eng.alloc_out(B, torch.float32, dev) for _ in range(sets)]
omax = [eng.alloc_out(B, torch.float32, dev) for _ in range(sets)]
P, T, S = _lib.GT_FLAG_PHASE_PERMUTE, _lib.GT_FLAG_PHASE_TILE, _lib.GT_FLAG_PHASE_SPAN


def step(k, phases=0, ops=("sum", "max")):
    eng.reduce(ws[k], ops, out_sum=osum[k], out_max=omax[k], phases=phases)


for k in range(sets):
    step(k)
torch.cuda.synchronize()


def timed(fn, n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(0)
    torch.cuda.synchronize()
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def graph_time(phases, per_graph, ops=("sum", "max"), steps=400):
    graphs = []
    for j in range(sets // min(per_graph, sets) if per_graph < sets else 1):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for q in range(per_graph):
                step((j * per_graph + q) % sets, phases, ops)
        graphs.append(g)
    return timed(lambda i: graphs[i % len(graphs)].replay(), max(1, steps // per_graph)) / per_graph


names = {P: "P", T: "T", S: "S", P | T: "PT", T | S: "TS", P | T | S: "PTS"}
print(f"batch {B}: microseconds per step")
print(f"{'chain':6s} {'1/graph':>9s} {'4/graph':>9s} {'16/graph':>9s}")
for ph in (P, T, S, P | T, T | S, P | T | S):
    print(f"{names[ph]:6s} " + " ".join(f"{graph_time(ph, n):9.2f}" for n in (1, 4, 16)))
print(f"PTS direct launches (no graph): {timed(lambda i: step(i % sets), 400):9.2f}")
print(f"sum only, PTS 1/graph {graph_time(0, 1, ('sum',)):.2f}  4/graph {graph_time(0, 4, ('sum',)):.2f}")


# ---- one step split into `parts` row blocks that run on their own streams (fork / join inside the graph) ------------
def split_time(parts, per_graph=4, steps=400):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    blocks = [(i * B // parts, (i + 1) * B // parts) for i in range(parts)]

    def split_step(k):
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        for si, (lo, hi) in enumerate(blocks):
            st = streams[si]
            st.wait_event(fork)
            with torch.cuda.stream(st):
                eng.reduce(ws[k][lo:hi], ("sum", "max"), out_sum=osum[k][lo:hi], out_max=omax[k][lo:hi])
            join = torch.cuda.Event()
            join.record(st)
            cur.wait_event(join)

    for k in range(sets):
        split_step(k)
    torch.cuda.synchronize()
    graphs = []
    for j in range(max(1, sets // per_graph)):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for q in range(per_graph):
                split_step((j * per_graph + q) % sets)
        graphs.append(g)
    return timed(lambda i: graphs[i % len(graphs)].replay(), max(1, steps // per_graph)) / per_graph


for parts in (2, 4):
    print(f"PTS, {parts} row blocks on {parts} streams: 1/graph {split_time(parts, 1):.2f}  4/graph {split_time(parts, 4):.2f}")
