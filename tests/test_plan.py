"""Host-side tests of the tile planner (csrc/trie_plan.cpp): the kernels' data flow is emulated in numpy from the
real plan arrays (tests/plan_emulator.py) and compared with the oracle.  No GPU needed."""
import numpy as np
import pytest

import oracle
from genlm_backend_b200 import TokenCharacterTrie, Token
from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows
from helpers import rel_err
from plan_emulator import emulate, swizzle_slot


def oracle_for(trie):
    lay = trie._layout
    return oracle.OracleLayout(trie.idx_to_leaf, lay["child_ptr"], lay["child_idx"])


@pytest.mark.parametrize("V,T", [(1, 1024), (5, 1024), (700, 1024), (3000, 1024), (3000, 2048), (20011, 2048),
                                 (20011, 1024), (50257, 1024)])
def test_emulated_kernels_match_oracle(V, T):
    trie = TokenCharacterTrie(synth_vocab(max(V, 256), seed=2)[-V:])
    trie._engine.plan(T)
    ws = dirichlet_rows(3, V, alpha=0.1, seed=1)
    o = oracle_for(trie)
    have = emulate(trie._engine, ws, "sum")
    assert not np.isnan(have).any()  # every node written, and none of them from a leaf slot past the vocabulary
    r, z = rel_err(have, o.weight_sum(ws))
    assert r <= 2e-6 and z == 0.0
    assert np.array_equal(emulate(trie._engine, ws, "max"), o.weight_max(ws).astype(np.float32))


def test_plan_invariants():
    V, T = 20011, 1024
    trie = TokenCharacterTrie(synth_vocab(V, seed=4))
    eng = trie._engine
    eng.plan(T)
    info = eng.plan_info()
    lay = trie._layout
    assert info["n_tiles"] == -(-V // T) and info["staged_slots"] == info["n_tiles"] * T
    assert info["rows_per_item"] == 4 and info["permute_unit"] % 32 == 0
    # the permute table sends the item at DFS rank r to leaf slot swizzle(r % T) of tile r // T: a bijection onto the
    # first V leaf slots in DFS order
    dest = eng.plan_array("leaf_dest").astype(np.int64)
    rank = np.empty(V, dtype=np.int64)
    rank[lay["perm"]] = np.arange(V)
    assert np.array_equal(dest, (rank // T) * T + swizzle_slot(rank % T, 16))
    assert len(np.unique(dest)) == V
    # node intervals partition the id space; spanning nodes are exactly those without a slot
    lo = eng.plan_array("tile_node_lo")
    assert lo[0] == 0 and lo[-1] == len(trie) and (np.diff(lo) > 0).all()
    slot = eng.plan_array("node_slot")
    spanning = (lay["lo"] // T) != ((lay["hi"] - 1) // T)
    ident = swizzle_slot(2 * T - 1, 4 * info["rows_per_item"])
    assert np.array_equal(slot == ident, spanning)  # spanning nodes point at the identity slot
    assert np.array_equal(np.sort(eng.plan_array("span_node")), np.flatnonzero(spanning))
    # every spanning node has one piece per tile it overlaps, each written by exactly one tile
    pp, sn = eng.plan_array("span_pp"), eng.plan_array("span_node")
    for i, n in enumerate(sn):
        assert pp[i + 1] - pp[i] == (lay["hi"][n] - 1) // T - lay["lo"][n] // T + 1
    pidx = eng.plan_array("piece_idx")
    assert np.array_equal(np.sort(pidx), np.arange(pp[-1]))
    assert info["n_span"] == int(spanning.sum()) and info["span_terms"] == pp[-1]
    # ELL padding uses the identity slot; real terms never point at it or past the tile's value array
    terms = eng.plan_array("ell_terms")
    assert terms.max() <= 2 * T - 1 and (terms != ident).sum() > 0
    rows = eng.plan_array("ell_row_ptr")
    assert rows[0] == 0 and rows[-1] * 32 == len(terms) and (eng.plan_array("ell_desc").reshape(-1, 2)[:, 1] > 0).all()


def test_plan_parameter_validation():
    trie = TokenCharacterTrie([Token(0, b"a")])
    from genlm_backend_b200._lib import GtError

    for T in (1000, 512, 4096):
        with pytest.raises(GtError):
            trie._engine.plan(T)
    trie._engine.plan(2048)
    trie._engine.plan()  # the default resolves to the existing plan
    with pytest.raises(GtError):
        trie._engine.plan(1024)
