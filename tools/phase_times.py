"""Times the tile kernel truncated after each phase (GT_DEBUG_STOP) to see where its time goes."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    from genlm_backend_b200 import ParallelTokenCharacterTrie, _lib
    from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows
    V, B = 128256, 64
    trie = ParallelTokenCharacterTrie(synth_vocab(V)); N = len(trie); eng = trie._engine
    sets = 4
    base = dirichlet_rows(B, V, alpha=1.0, seed=1)
    ws = [torch.tensor(np.roll(base, k, axis=0)).cuda() for k in range(sets)]
    osum = [eng.alloc_out(B, torch.float32, torch.device("cuda", 0)) for _ in range(sets)]
    for k in range(sets):
        eng.reduce(ws[k], ("sum",), out_sum=osum[k])
    torch.cuda.synchronize()
    graphs = []
    for k in range(sets):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            eng.reduce(ws[k], ("sum",), out_sum=osum[k], phases=_lib.GT_FLAG_PHASE_TILE)
        graphs.append(g)
    for g in graphs: g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(200): graphs[i % sets].replay()
    b.record(); torch.cuda.synchronize()
    print(json.dumps({"stop": os.environ.get("GT_DEBUG_STOP", "0"), "us": a.elapsed_time(b) / 200 * 1e3}))
else:
    for stop in ["0", "1", "2", "3", "9"]:
        env = dict(os.environ, GT_DEBUG_STOP=stop)
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-500:])
