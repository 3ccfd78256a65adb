"""ORACLE -- TEST INFRASTRUCTURE ONLY.  The synthetic inputs of SURVEY.md section 8(d) (byte vocabularies shaped like BPE
vocabularies, Dirichlet rows), restated here with numpy only so that the reference arm of bench.py can produce the
bench workload without importing the product package.  ``tests/test_oracle.py`` checks that these are the same
vocabularies and rows as ``genlm_backend_b200.synthetic``."""
import numpy as np


def synth_vocab_bytes(V, seed=0, max_len=32):
    """V distinct byte strings: the 256 single bytes, then concatenations of earlier entries."""
    rng = np.random.default_rng(seed)
    toks = [bytes([i]) for i in range(256)]
    seen = set(toks)
    while len(toks) < V:
        n = len(toks)
        a = toks[rng.integers(0, n)]
        u = rng.random()
        b = toks[rng.integers(0, 256)] if u < 0.75 else toks[rng.integers(0, min(n, 2048))]
        c = a + b
        if len(c) <= max_len and c not in seen:
            seen.add(c)
            toks.append(c)
    return toks[:V]


def dirichlet_rows(B, V, alpha=0.1, seed=1, dtype=np.float32):
    rng = np.random.default_rng(seed)
    out = np.empty((B, V), dtype=dtype)
    step = max(1, (1 << 24) // max(V, 1))
    for r0 in range(0, B, step):
        r1 = min(B, r0 + step)
        out[r0:r1] = rng.dirichlet(np.full(V, alpha), size=r1 - r0).astype(dtype)
    return out
