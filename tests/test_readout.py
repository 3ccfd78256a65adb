"""Read-outs that keep the mass slab on the GPU: gather of a few nodes per row and the
subtree token mask.

CPU part: the oracle's restatements are pinned to the reference's own reachability arrays (``src_indices`` /
``dst_indices`` of ``ParallelTokenCharacterTrie``, ``tests/golden/{toy,edge}.npz``).  GPU part (``-m gpu``): the CUDA
kernels behind ``gt_gather_nodes`` / ``gt_subtree_token_mask`` against the oracle; bit masks and indices bit-exact,
gathered masses exactly the slab's values, ratios / logs to fp32 rounding (tolerances below).
"""
import numpy as np
import pytest
import torch

import oracle
from genlm_backend_b200 import ParallelTokenCharacterTrie, smc
from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows
from helpers import load_golden, unflat, tokens


# ---- oracle pinned to the reference (CPU) ---------------------------------------------------------------------
@pytest.mark.parametrize("name", ["toy", "edge"])
def test_oracle_subtree_mask_is_reference_reachability_column(name):
    g = load_golden(name)
    ot = oracle.OracleTrie(tokens(unflat(g["blob"], g["lens"])))
    V, N = ot.n_items, ot.n_nodes
    M = np.zeros((V, N), dtype=bool)  # the reference's M (parallel.py:52-64) from its own index arrays
    M[g["reach_rows"], g["reach_cols"]] = True
    have = ot.subtree_token_mask(np.arange(N))
    assert np.array_equal(have, M.T)


def test_oracle_toy_known_masks():
    # node 3 = prefix "a" -> tokens a(0), ab(2); node 12 = root -> everything; node 6 = leaf of <eos>
    g = load_golden("toy")
    ot = oracle.OracleTrie(tokens(unflat(g["blob"], g["lens"])))
    m = ot.subtree_token_mask([3, 12, 6, 5])
    assert m.tolist() == [[True, False, True, False], [True] * 4, [False, False, False, True], [False, True, False, False]]


def test_oracle_gather_and_unpack():
    mass = np.array([[0.1, 0.2, 0.7], [0.0, 0.5, 0.5]])
    assert np.allclose(oracle.gather_nodes(mass, [2, 0, -1]), [[0.7, 0.1, 0.0], [0.5, 0.0, 0.0]])
    assert np.allclose(oracle.gather_nodes(mass, [[1], [2]], normalizer=[2, 1]), [[0.2 / 0.7], [1.0]])
    with np.errstate(divide="ignore"):
        assert np.allclose(oracle.gather_nodes(mass, [0], log=True), np.log([[0.1], [0.0]]))
    bits = np.array([[5, 1]], dtype=np.int32)
    assert oracle.unpack_bits(bits, 34)[0].nonzero()[0].tolist() == [0, 2, 32]


# ---- CUDA kernels against the oracle (GPU) --------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["toy", "edge", "synth3000"])
def test_subtree_mask_matches_oracle_all_nodes(name):
    g = load_golden(name)
    dec = tokens(unflat(g["blob"], g["lens"]))
    trie = ParallelTokenCharacterTrie(dec)
    ot = oracle.OracleTrie(dec)
    nodes = np.arange(len(trie)) if len(trie) <= 64 else np.random.default_rng(0).integers(0, len(trie), 300)
    nodes = np.concatenate([nodes, [-1, len(trie)]])  # out-of-range ids give empty masks
    bits = trie.subtree_token_mask(torch.tensor(nodes))
    assert bits.dtype == torch.int32 and bits.shape == (len(nodes), (len(dec) + 31) // 32)
    have = oracle.unpack_bits(bits.cpu().numpy(), len(dec))
    want = np.concatenate([ot.subtree_token_mask(nodes[:-2]), np.zeros((2, len(dec)), bool)])
    assert np.array_equal(have, want)
    # padding bits of the last word are never set
    assert np.array_equal(oracle.unpack_bits(bits.cpu().numpy(), bits.shape[1] * 32)[:, len(dec):].any(axis=1), np.zeros(len(nodes), bool))


@pytest.mark.gpu
def test_subtree_mask_full_size_properties():
    V = 128256
    trie = ParallelTokenCharacterTrie(synth_vocab(V))
    lay = trie._layout
    rng = np.random.default_rng(3)
    nodes = np.concatenate([[trie.root], rng.integers(0, len(trie), 511)])
    have = oracle.unpack_bits(trie.subtree_token_mask(torch.tensor(nodes)).cpu().numpy(), V)
    # a node's tokens are the DFS leaf range [lo, hi) read through perm (DFS rank -> item position)
    assert np.array_equal(have.sum(axis=1), (lay["hi"] - lay["lo"])[nodes])
    assert have[0].all()  # the root reaches every token
    for b in range(1, 40):
        want = np.zeros(V, bool)
        want[lay["perm"][lay["lo"][nodes[b]]:lay["hi"][nodes[b]]]] = True
        assert np.array_equal(have[b], want)
    # a child's mask is contained in its parent's
    par = lay["parent"][nodes[1:200]]
    pm = oracle.unpack_bits(trie.subtree_token_mask(torch.tensor(par)).cpu().numpy(), V)
    assert not (have[1:200] & ~pm).any()


@pytest.mark.gpu
def test_mask_feeds_sampler():
    """Mass under a node == exp(logZ) of the fused masked logsumexp with that node's token mask (the loop the two
    halves of the path close: trie mass <-> masked sampling), and every draw lies under the node."""
    V, B = 3000, 64
    g = load_golden("synth3000")
    dec = tokens(unflat(g["blob"], g["lens"]))
    trie = ParallelTokenCharacterTrie(dec)
    p = dirichlet_rows(B, V, alpha=1.0, seed=5)
    nodes = np.random.default_rng(1).integers(0, len(trie), B)
    nodes[0] = trie.root
    bits = trie.subtree_token_mask(torch.tensor(nodes))
    logp = torch.tensor(np.log(p)).cuda()
    logZ, tok = smc.masked_logsumexp_sample(logp, bits, seed=7)
    mass = trie.batch_weight_sum_at(torch.tensor(p), torch.tensor(nodes)[:, None])[:, 0]
    nz = mass > 0
    np.testing.assert_allclose(np.exp(logZ.cpu().numpy()[nz].astype(np.float64)), mass[nz], rtol=2e-5)
    keep = oracle.unpack_bits(bits.cpu().numpy(), V)
    t = tok.cpu().numpy()
    assert (t[nz] >= 0).all() and keep[np.flatnonzero(nz), t[nz]].all()
    assert (t[~nz] == -1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_gather_nodes_matches_oracle(dtype):
    g = load_golden("synth3000")
    dec = tokens(unflat(g["blob"], g["lens"]))
    trie = ParallelTokenCharacterTrie(dec)
    B, N = 37, len(trie)
    ws = dirichlet_rows(B, len(dec), alpha=0.1, seed=2)
    sums = trie.batch_weight_sum_tensor(torch.tensor(ws)).to(dtype)
    host = sums.cpu().numpy()
    rng = np.random.default_rng(4)
    ids = rng.integers(-1, N + 1, size=(B, 19))
    norm = rng.integers(0, N, size=B)
    norm[:5] = trie.root
    have = trie.gather_nodes(sums, torch.tensor(ids)).cpu().numpy()
    assert have.dtype == host.dtype and np.array_equal(have, oracle.gather_nodes(host, ids).astype(host.dtype))  # exact copy
    shared = trie.gather_nodes(sums, torch.tensor(ids[0])).cpu().numpy()
    assert np.array_equal(shared, oracle.gather_nodes(host, ids[0]).astype(host.dtype))
    tol = 3e-7 if dtype == torch.float32 else 1e-14  # one fp32 (fp64) division / log rounding
    with np.errstate(divide="ignore", invalid="ignore"):
        want = oracle.gather_nodes(host, ids, normalizer=norm)
        got = trie.gather_nodes(sums, torch.tensor(ids), normalizer=torch.tensor(norm)).cpu().numpy().astype(np.float64)
        ok = np.isfinite(want)
        np.testing.assert_allclose(got[ok], want[ok], rtol=tol)
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want))
        wantl = oracle.gather_nodes(host, ids, normalizer=norm, log=True)
        gotl = trie.gather_nodes(sums, torch.tensor(ids), normalizer=torch.tensor(norm), log=True).cpu().numpy().astype(np.float64)
        ok = np.isfinite(wantl)
        np.testing.assert_allclose(gotl[ok], wantl[ok], rtol=tol, atol=2e-5 if dtype == torch.float32 else 1e-12)
        assert np.array_equal(np.isneginf(gotl), np.isneginf(wantl))


@pytest.mark.gpu
def test_batch_weight_at_equals_full_slab():
    V = 50257
    trie = ParallelTokenCharacterTrie(synth_vocab(V))
    ws = torch.tensor(dirichlet_rows(8, V, alpha=1.0, seed=9))
    ids = torch.tensor(np.random.default_rng(0).integers(0, len(trie), size=(8, 257)))
    full_sum, full_max = trie.batch_weight_sum(ws), trie.batch_weight_max(ws)
    assert np.array_equal(trie.batch_weight_sum_at(ws, ids), np.take_along_axis(full_sum, ids.numpy(), axis=1))
    assert np.array_equal(trie.batch_weight_max_at(ws, ids), np.take_along_axis(full_max, ids.numpy(), axis=1))
    # children of the root: the next-byte distribution, normalised by the root mass
    kids = torch.tensor(trie.jump[trie.root])
    cond = trie.batch_weight_sum_at(ws, kids, normalizer=trie.root)
    np.testing.assert_allclose(cond.sum(axis=1), 1.0, rtol=1e-5)


@pytest.mark.gpu
def test_batch_weight_sum_max_at_one_pass():
    g = load_golden("synth3000")
    dec = tokens(unflat(g["blob"], g["lens"]))
    trie = ParallelTokenCharacterTrie(dec)
    ws = torch.tensor(dirichlet_rows(9, len(dec), alpha=0.3, seed=8))
    ids = torch.tensor(np.random.default_rng(2).integers(0, len(trie), size=(9, 33)))
    s_at, m_at = trie.batch_weight_sum_max_at(ws, ids)
    assert np.array_equal(s_at, np.take_along_axis(trie.batch_weight_sum(ws), ids.numpy(), axis=1))
    assert np.array_equal(m_at, np.take_along_axis(trie.batch_weight_max(ws), ids.numpy(), axis=1))
