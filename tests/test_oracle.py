"""Pins the oracle (oracle/__init__.py, oracle/trie_oracle.c) against the reference: golden outputs produced by the
reference itself (tests/golden/make_golden.py) and the known answers of the reference's own tests."""
import hashlib

import numpy as np
import pytest

import oracle
from genlm_backend_b200.synthetic import synth_vocab_bytes, dirichlet_rows
from helpers import load_golden, unflat


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def test_known_answers_of_reference_tests():
    """tests/test_trie.py:26-85 of the reference (toy vocabulary, sum and max)."""
    t = oracle.OracleTrie([b"a", b"b", b"ab", b"<eos>"])
    ws = np.array([0.1, 0.2, 0.2, 0.5], dtype=np.float32)
    s, m = t.weight_sum(ws), t.weight_max(ws)
    prefix = {t.root: b""}
    for x in reversed(range(t.n_nodes)):
        for k, y in t.children[x].items():
            prefix[y] = prefix[x] if isinstance(k, tuple) else prefix[x] + bytes([k])
    leaves = set(t.idx_to_leaf[:, 1].tolist())
    want_sum = {b"": 1, b"a": 0.3, b"b": 0.2, b"ab": 0.2, b"<": 0.5, b"<e": 0.5, b"<eo": 0.5, b"<eos": 0.5, b"<eos>": 0.5}
    want_max = {**want_sum, b"": 0.5, b"a": 0.2}
    leaf_want = {b"a": 0.1, b"b": 0.2, b"ab": 0.2, b"<eos>": 0.5}
    for node, p in prefix.items():
        if node in leaves:
            assert np.isclose(s[node], leaf_want[p], rtol=1e-5) and np.isclose(m[node], leaf_want[p], rtol=1e-5)
        else:
            assert np.isclose(s[node], want_sum[p], rtol=1e-5, atol=1e-8), (p, s[node])
            assert np.isclose(m[node], want_max[p], rtol=1e-5, atol=1e-8), (p, m[node])


def test_duplicate_byte_strings_known_answers():
    """tests/test_token.py:96-118, 264-312 of the reference."""
    t = oracle.OracleTrie([b"a", b"hello", b"hello", b"world"])
    s = t.weight_sum(np.array([0.1, 0.3, 0.5, 0.1], dtype=np.float32))
    l1, l2 = t.idx_to_leaf[1][1], t.idx_to_leaf[2][1]
    assert l1 != l2 and np.isclose(s[l1], 0.3, rtol=1e-5) and np.isclose(s[l2], 0.5, rtol=1e-5)
    t = oracle.OracleTrie([b"ab", b"ab", b"ac", b"b"])
    m = t.weight_max(np.array([0.1, 0.9, 0.3, 0.2], dtype=np.float32))
    assert np.isclose(m[t.root], 0.9, rtol=1e-5)
    t = oracle.OracleTrie([b"ab", b"ab", b"c"])
    s = t.weight_sum(np.array([0.3, 0.5, 0.2], dtype=np.float32))
    assert np.isclose(s[t.root], 1.0, rtol=1e-5)


@pytest.mark.parametrize("name", ["toy", "edge", "synth3000"])
def test_against_reference_outputs(name):
    g = load_golden(name)
    t = oracle.OracleTrie(unflat(g["blob"], g["lens"]))
    assert t.n_nodes == int(g["n_nodes"]) and t.root == int(g["root"])
    assert np.array_equal(t.idx_to_leaf, g["idx_to_leaf"])
    assert np.array_equal(t.ordering, g["ordering"])
    assert np.array_equal(t.jump_ptr, g["jump_ptr"]) and np.array_equal(t.jump_idx, g["jump_idx"])
    # numba path: bit-identical (same loop order, float64)
    assert np.array_equal(t.weight_sum(g["ws"]), g["seq_sum"])
    assert np.array_equal(t.weight_max(g["ws"]), g["seq_max"])
    assert np.array_equal(t.weight_sum(g["ws"], threads=0), g["seq_sum"])  # OpenMP over rows changes nothing
    # torch path restated in numpy fp32: same values up to fp32 summation order
    np.testing.assert_allclose(t.parallel_weight_sum(g["ws"]), g["par_sum"], rtol=1e-5, atol=1e-8)
    assert np.array_equal(t.parallel_weight_max(g["ws"]), g["par_max"])
    rows, cols = t.reachability()
    if "reach_rows" in g:
        assert np.array_equal(rows, g["reach_rows"]) and np.array_equal(cols, g["reach_cols"])
    else:
        assert np.array_equal(digest(np.stack([rows, cols])), g["reach_digest"])


def test_baseline_size_gpt2(golden_dir):
    """BASELINE config 1 (V = 50,257): layout digests and sampled outputs of the reference."""
    g = load_golden("synth50257")
    toks = synth_vocab_bytes(50257)
    t = oracle.OracleTrie(toks)
    assert t.n_nodes == int(g["n_nodes"]) == 134729
    assert np.array_equal(digest(t.idx_to_leaf), g["idx_to_leaf_digest"])
    assert np.array_equal(digest(t.ordering), g["ordering_digest"])
    assert np.array_equal(digest(t.jump_ptr), g["jump_ptr_digest"])
    assert np.array_equal(digest(t.jump_idx), g["jump_idx_digest"])
    ws = dirichlet_rows(2, 50257, alpha=0.1, seed=1)
    s, m = t.weight_sum(ws), t.weight_max(ws)
    assert np.array_equal(s[:, g["pick"]], g["seq_sum_pick"])
    assert np.array_equal(m[:, g["pick"]], g["seq_max_pick"])
    assert np.array_equal(digest(m.astype(np.float32)), g["seq_max_digest"])


def test_masked_logsumexp_pinned_against_torch():
    g = load_golden("sampler")
    np.testing.assert_allclose(oracle.masked_logsumexp(g["logp"], g["mask"]), g["logZ"], rtol=1e-12)
    np.testing.assert_allclose(oracle.masked_logsumexp(g["logp"], g["mask"], temperature=0.7), g["logZ_T07"], rtol=1e-12)
    p = oracle.masked_probs(g["logp"], g["mask"])
    np.testing.assert_allclose(p.sum(-1), 1.0, rtol=1e-12)
    assert (p[~np.isfinite(g["mask"])] == 0).all()
    allmasked = np.full((1, 5), -np.inf)
    assert oracle.masked_logsumexp(np.zeros((1, 5)), allmasked)[0] == -np.inf


@pytest.mark.parametrize("name", ["toy", "edge", "synth3000"])
def test_torch_reference_restatement_pinned(name):
    """oracle/torch_ref.py (what bench.py times as `reference_gpu`) against the reference's own torch path outputs."""
    import torch

    from oracle.torch_ref import TorchReferenceTrie

    g = load_golden(name)
    t = oracle.OracleTrie(unflat(g["blob"], g["lens"]))
    rows, cols = t.reachability()
    ref = TorchReferenceTrie(t.idx_to_leaf, rows, cols, t.n_nodes, "cpu")
    ws = torch.tensor(g["ws"], dtype=torch.float32)
    np.testing.assert_allclose(ref.batch_weight_sum(ws), g["par_sum"], rtol=1e-6, atol=1e-10)  # same library kernels: summation order only
    assert np.array_equal(ref.batch_weight_max(ws), g["par_max"])


def test_torch_smc_step_restatement():
    import torch

    from oracle.torch_ref import smc_step

    g = load_golden("sampler")
    logZ, tok = smc_step(torch.tensor(g["logp"], dtype=torch.float64), torch.tensor(g["mask"], dtype=torch.float64),
                         generator=torch.Generator().manual_seed(0))
    finite = np.isfinite(g["logZ"])
    np.testing.assert_allclose(logZ.numpy()[finite], g["logZ"][finite], rtol=1e-12)
    keep = np.isfinite(g["mask"]) if g["mask"].ndim == 2 else np.broadcast_to(np.isfinite(g["mask"]), g["logp"].shape)
    assert all(keep[b, int(t)] for b, t in enumerate(tok) if finite[b])


def test_oracle_synth_matches_product_generator():
    from oracle import synth
    from genlm_backend_b200 import synthetic

    assert synth.synth_vocab_bytes(3000, seed=3) == synthetic.synth_vocab_bytes(3000, seed=3)
    assert np.array_equal(synth.dirichlet_rows(5, 777, alpha=0.3, seed=9), synthetic.dirichlet_rows(5, 777, alpha=0.3, seed=9))
