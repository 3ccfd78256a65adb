"""Generates tests/golden/*.npz by running the REFERENCE implementation (imported from /root/reference) on
seeded inputs.  Run in the dev container only:  python tests/golden/make_golden.py
The fixtures are what the GPU-box tests compare against (the reference cannot travel there).

Imports genlm.backend.trie without executing genlm/backend/__init__.py (which pulls vLLM): stub parent packages
are placed in sys.modules so only the trie and token modules are loaded, unmodified.
"""
import os
import sys
import types
import hashlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def load_reference():
    for name, path in [("genlm", "genlm"), ("genlm.backend", "genlm/backend")]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF, path)]
            sys.modules[name] = m
    import genlm.backend.tokenization.token as tok

    tk = types.ModuleType("genlm.backend.tokenization")
    tk.__path__ = [os.path.join(REF, "genlm/backend/tokenization")]
    tk.Token = tok.Token
    sys.modules["genlm.backend.tokenization"] = tk
    import genlm.backend.trie as trie

    return trie, tok.Token


def flat(items):
    """(bytes blob, lengths) for a list of byte strings."""
    return np.frombuffer(b"".join(items), dtype=np.uint8), np.array([len(x) for x in items], dtype=np.int32)


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def layout_arrays(t):
    jump = [np.asarray(j, dtype=np.int32) for j in t.jump]
    ptr = np.zeros(len(jump) + 1, dtype=np.int32)
    ptr[1:] = np.cumsum([len(j) for j in jump])
    return {
        "n_nodes": np.int64(len(t.children)),
        "root": np.int64(t.root),
        "idx_to_leaf": np.asarray(t.idx_to_leaf, dtype=np.int32),
        "ordering": np.asarray(t.ordering, dtype=np.int64),
        "jump_ptr": ptr,
        "jump_idx": np.concatenate(jump).astype(np.int32) if len(jump) > 1 else np.zeros(0, np.int32),
    }


def main():
    from genlm_backend_b200.synthetic import synth_vocab_bytes, dirichlet_rows, logsoftmax_rows, bernoulli_log_mask

    trie, Token = load_reference()
    Seq, Par = trie.TokenCharacterTrie, trie.ParallelTokenCharacterTrie

    # 1. the reference's own toy fixture (tests/test_trie.py:16-18) with its batch rows (tests/test_trie.py:120-126)
    toy = [b"a", b"b", b"ab", b"<eos>"]
    ws = np.array([[0.1, 0.2, 0.2, 0.5], [0, 0.3, 0.6, 0.1], [0.99, 0.01, 0, 0]], dtype=np.float32)
    dec = [Token(i, b) for i, b in enumerate(toy)]
    s, p = Seq(dec), Par(dec, device="cpu")
    blob, lens = flat(toy)
    rows, cols = p.src_indices.numpy(), p.dst_indices.numpy()
    np.savez_compressed(
        os.path.join(HERE, "toy.npz"), blob=blob, lens=lens, ws=ws,
        seq_sum=s.batch_weight_sum(torch.tensor(ws)), seq_max=s.batch_weight_max(torch.tensor(ws)),
        par_sum=p.batch_weight_sum(torch.tensor(ws)), par_max=p.batch_weight_max(torch.tensor(ws)),
        reach_rows=rows, reach_cols=cols, positions=p.positions.numpy(),
        node2prefix_nodes=np.array(list(s.node2prefix.keys()), dtype=np.int64),
        node2prefix_lens=np.array([len(v) for v in s.node2prefix.values()], dtype=np.int64),
        node2prefix_flat=np.array([x for v in s.node2prefix.values() for x in v], dtype=np.int64),
        **layout_arrays(s),
    )

    # 2. duplicates + empty token + a 128-byte token + unsorted insertion order
    rng = np.random.default_rng(7)
    long_tok = bytes(rng.integers(0, 256, size=128, dtype=np.uint8).tolist())
    edge = [b"ab", b"", b"ab", b"a", long_tok, b"abc", b"b", long_tok[:64], b"ab", b"\x00", b"\xff\xfe"]
    dec = [Token(i, b) for i, b in enumerate(edge)]
    s, p = Seq(dec), Par(dec, device="cpu")
    ws = rng.dirichlet(np.full(len(edge), 0.5), size=4).astype(np.float32)
    ws[1, 2] = 0.0
    blob, lens = flat(edge)
    np.savez_compressed(
        os.path.join(HERE, "edge.npz"), blob=blob, lens=lens, ws=ws,
        seq_sum=s.batch_weight_sum(torch.tensor(ws)), seq_max=s.batch_weight_max(torch.tensor(ws)),
        par_sum=p.batch_weight_sum(torch.tensor(ws)), par_max=p.batch_weight_max(torch.tensor(ws)),
        reach_rows=p.src_indices.numpy(), reach_cols=p.dst_indices.numpy(),
        **layout_arrays(s),
    )

    # 3. non-byte sentinel item + plain bytes items (base.py:34-48): the second fixture of SURVEY appendix A
    class EOS:
        def __iter__(self):
            return iter([self])

        def __repr__(self):
            return "EOS"

    eos = EOS()
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dec = [Token(0, b"ab"), Token(1, b""), Token(2, b"ab"), Token(3, b"a"), eos, b"plain"]
        s = Seq(dec)
    ws = np.array([[0.1, 0.2, 0.3, 0.25, 0.15, 0.0], [0.0, 0.0, 0.5, 0.0, 0.25, 0.25]], dtype=np.float32)
    np.savez_compressed(
        os.path.join(HERE, "sentinel.npz"), ws=ws,
        seq_sum=s.batch_weight_sum(torch.tensor(ws)), seq_max=s.batch_weight_max(torch.tensor(ws)),
        # edge labels with the sentinel encoded as -1, leaf edges as -(2+idx)
        children_ptr=np.cumsum([0] + [len(c) for c in s.children]).astype(np.int32),
        children_key=np.array([(-(2 + k[1]) if isinstance(k, tuple) else (-1 if k is eos else k)) for c in s.children for k in c], dtype=np.int64),
        children_val=np.array([v for c in s.children for v in c.values()], dtype=np.int64),
        **layout_arrays(s),
    )

    # 4. medium synthetic vocabulary: full layout + full results for three Dirichlet(0.1) rows
    V = 3000
    toks = synth_vocab_bytes(V, seed=0)
    dec = [Token(i, b) for i, b in enumerate(toks)]
    s, p = Seq(dec), Par(dec, device="cpu")
    ws = dirichlet_rows(3, V, alpha=0.1, seed=1)
    blob, lens = flat(toks)
    np.savez_compressed(
        os.path.join(HERE, "synth3000.npz"), blob=blob, lens=lens, ws=ws,
        seq_sum=s.batch_weight_sum(torch.tensor(ws)), seq_max=s.batch_weight_max(torch.tensor(ws)),
        par_sum=p.batch_weight_sum(torch.tensor(ws)), par_max=p.batch_weight_max(torch.tensor(ws)),
        reach_digest=digest(np.stack([p.src_indices.numpy(), p.dst_indices.numpy()])), nnz=np.int64(len(p.src_indices)),
        **layout_arrays(s),
    )

    # 5. BASELINE config sizes: layout digests + sampled node values for one Dirichlet(0.1) row (full arrays are MBs)
    for V in (50257, 128256):
        toks = synth_vocab_bytes(V, seed=0)
        dec = [Token(i, b) for i, b in enumerate(toks)]
        s = Seq(dec)
        ws = dirichlet_rows(2, V, alpha=0.1, seed=1)
        ssum = s.batch_weight_sum(torch.tensor(ws))
        smax = s.batch_weight_max(torch.tensor(ws))
        N = len(s.children)
        pick = np.sort(np.random.default_rng(3).choice(N, size=4096, replace=False))
        pick[-1] = s.root
        probe = np.random.default_rng(4).standard_normal(N)
        lay = layout_arrays(s)
        np.savez_compressed(
            os.path.join(HERE, f"synth{V}.npz"), n_nodes=np.int64(N), root=np.int64(s.root),
            nnz=np.int64(sum(len(t) + 2 for t in toks)),
            idx_to_leaf_digest=digest(lay["idx_to_leaf"]), ordering_digest=digest(lay["ordering"]),
            jump_ptr_digest=digest(lay["jump_ptr"]), jump_idx_digest=digest(lay["jump_idx"]),
            pick=pick, seq_sum_pick=ssum[:, pick], seq_max_pick=smax[:, pick],
            seq_sum_probe=ssum @ probe, seq_max_probe=smax @ probe,
            seq_max_digest=digest(smax.astype(np.float32)),
        )
        print("V", V, "N", N)

    # 6. sampler: logsumexp pinned against torch (float64) and a torch.multinomial histogram for a two-sample test
    V = 1000
    logp = logsoftmax_rows(4, V, seed=0)
    mask = bernoulli_log_mask(4, V, p=0.5, seed=1)
    masked = torch.tensor(logp, dtype=torch.float64) + torch.tensor(mask, dtype=torch.float64)
    logZ = masked.logsumexp(-1).numpy()
    logZ_T = (torch.tensor(logp, dtype=torch.float64) / 0.7 + torch.tensor(mask, dtype=torch.float64)).logsumexp(-1).numpy()
    g = torch.Generator().manual_seed(123)
    probs = (masked[0] - masked[0].logsumexp(-1)).exp().to(torch.float32)
    draws = torch.multinomial(probs.expand(2000, V), 100, replacement=True, generator=g).reshape(-1)
    counts = torch.bincount(draws, minlength=V).numpy()
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), logp=logp, mask=mask, logZ=logZ, logZ_T07=logZ_T,
                        multinomial_counts_row0=counts)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
