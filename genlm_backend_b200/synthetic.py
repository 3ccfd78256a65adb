"""Deterministic synthetic inputs (there is no network for tokenizers or datasets): byte vocabularies shaped
like BPE vocabularies, Dirichlet weight rows, log-softmax rows and masks.  Used by tests and bench.py.
The generator is the one SURVEY.md section 8(d) fixes, so node counts are reproducible
(V=50257 -> N=134729, V=128256 -> N=345180, V=151665 -> N=407861 at seed 0).
"""
import numpy as np

from .tokenization import Token


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


def synth_vocab(V, seed=0, max_len=32):
    return [Token(i, b) for i, b in enumerate(synth_vocab_bytes(V, seed, max_len))]


def dirichlet_rows(B, V, alpha=0.1, seed=1, dtype=np.float32):
    """``B`` rows ~ Dirichlet(alpha): alpha=0.1 gives exact zeros and 1e-30-scale masses after the fp32 cast."""
    rng = np.random.default_rng(seed)
    out = np.empty((B, V), dtype=dtype)
    step = max(1, (1 << 24) // max(V, 1))
    for r0 in range(0, B, step):
        r1 = min(B, r0 + step)
        out[r0:r1] = rng.dirichlet(np.full(V, alpha), size=r1 - r0).astype(dtype)
    return out


def logsoftmax_rows(B, V, seed=0, dtype=np.float32):
    """``log_softmax`` of standard-normal logits, the shape of ``next_token_logprobs`` rows."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, V))
    x = x - x.max(axis=1, keepdims=True)
    x = x - np.log(np.exp(x).sum(axis=1, keepdims=True))
    return x.astype(dtype)


def bernoulli_log_mask(B, V, p=0.5, seed=1):
    """Additive masks in {0, -inf} (``log`` of a boolean mask, README.md:59-65)."""
    rng = np.random.default_rng(seed)
    keep = rng.random((B, V)) < p
    return np.where(keep, 0.0, -np.inf).astype(np.float32)
