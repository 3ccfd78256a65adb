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


@pytest.mark.parametrize("V,T,Q,R", [(1, 1024, 4, 2), (5, 1024, 4, 4), (700, 1024, 128, 2), (3000, 1024, 512, 4),
                                     (3000, 2048, 8192, 2), (20011, 2048, 8192, 2), (20011, 1024, 1000, 4),
                                     (20011, 1024, 16384, 2), (20011, 2048, 4096, 4), (50257, 1024, 4096, 4)])
def test_emulated_kernels_match_oracle(V, T, Q, R):
    trie = TokenCharacterTrie(synth_vocab(max(V, 256), seed=2)[-V:])
    trie._engine.plan(T, Q, R)
    ws = dirichlet_rows(3, V, alpha=0.1, seed=1)
    o = oracle_for(trie)
    have = emulate(trie._engine, ws, "sum")
    assert not np.isnan(have).any()  # every node written exactly through one of the three phases
    r, z = rel_err(have, o.weight_sum(ws))
    assert r <= 2e-6 and z == 0.0
    assert np.array_equal(emulate(trie._engine, ws, "max"), o.weight_max(ws).astype(np.float32))


def test_plan_invariants():
    V, T, Q = 20011, 1024, 1000
    trie = TokenCharacterTrie(synth_vocab(V, seed=4))
    eng = trie._engine
    eng.plan(T, Q)
    info = eng.plan_info()
    lay = trie._layout
    assert info["n_tiles"] == -(-V // T) and info["n_segs"] == -(-V // Q)
    assert info["staged_row_elems"] % 4 == 0 and info["staged_row_elems"] >= V
    # staging is a permutation of the row plus padding; a padding element is one the tile kernel sends to a trash slot
    # (past the value array), whatever source position its record names
    rec, cptr = eng.plan_array("p1_rec").reshape(-1, 4), eng.plan_array("p1_chunk_ptr")
    zoff = rec[:, 0]
    p2 = eng.plan_array("p2_slot").astype(np.int64)
    seen = np.zeros(V, dtype=np.int64)
    for s in range(info["n_segs"]):
        r = rec[cptr[s]:cptr[s + 1]]
        lohi = r[:, 1:3].astype(np.int64) & 0xFFFFFFFF
        e = np.stack([lohi[:, 0] & 0xFFFF, lohi[:, 0] >> 16, lohi[:, 1] & 0xFFFF, lohi[:, 1] >> 16], axis=1).reshape(-1)
        dst = (r[:, 0].astype(np.int64)[:, None] + np.arange(4)[None, :]).reshape(-1)
        real = p2[dst] < info["max_tile_values"]
        assert (e < min(Q, V - s * Q)).all()  # every named position, padding included, lies inside the segment
        np.add.at(seen, e[real] + s * Q, 1)
    assert (seen == 1).all()
    assert len(np.unique(zoff)) == len(zoff) and (zoff % 4 == 0).all()
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

    for T, Q, R in [(1000, 8192, 2), (512, 8192, 2), (4096, 8192, 2), (2048, 6, 2), (2048, 32768, 2), (2048, 4096, 3)]:
        with pytest.raises(GtError):
            trie._engine.plan(T, Q, R)
    trie._engine.plan(2048, 4096, 2)
    trie._engine.plan()  # defaults resolve to the existing plan
    with pytest.raises(GtError):
        trie._engine.plan(1024, 4096)
