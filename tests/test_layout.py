"""Host-side tests (no GPU): the C++ builder reproduces the reference's node numbering and attributes, the
Python classes mirror the reference's interface and errors (tests/test_trie.py:282-323, tests/test_token.py)."""
import copy
import hashlib
import pickle
import warnings

import numpy as np
import pytest
import torch

import oracle
from genlm_backend_b200 import Token, TokenCharacterTrie, ParallelTokenCharacterTrie, AsyncTokenCharacterTrie
from genlm_backend_b200.synthetic import synth_vocab, synth_vocab_bytes
from helpers import load_golden, unflat, tokens, EOS


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def jump_csr(trie):
    lay = trie._layout
    return lay["child_ptr"], lay["child_idx"]


@pytest.mark.parametrize("name", ["toy", "edge", "synth3000"])
def test_layout_equals_reference(name):
    g = load_golden(name)
    trie = ParallelTokenCharacterTrie(tokens(unflat(g["blob"], g["lens"])), device="cpu")
    assert len(trie.children) == int(g["n_nodes"]) and trie.root == int(g["root"])
    assert trie.idx_to_leaf.dtype == np.int32 and np.array_equal(trie.idx_to_leaf, g["idx_to_leaf"])
    assert np.array_equal(trie.ordering, g["ordering"])
    ptr, idx = jump_csr(trie)
    assert np.array_equal(ptr, g["jump_ptr"]) and np.array_equal(idx, g["jump_idx"])
    assert all(np.array_equal(j, idx[ptr[n]:ptr[n + 1]]) for n, j in enumerate(trie.jump))
    rows, cols = trie.src_indices.numpy(), trie.dst_indices.numpy()
    if "reach_rows" in g:
        assert np.array_equal(rows, g["reach_rows"]) and np.array_equal(cols, g["reach_cols"])
    else:
        assert len(rows) == int(g["nnz"]) and np.array_equal(digest(np.stack([rows, cols])), g["reach_digest"])
    assert np.array_equal(trie.positions.numpy(), np.arange(len(trie.decode)))


def test_toy_attributes_match_reference_exactly():
    """SURVEY appendix A / golden toy fixture: children dicts, word2leaf, node2prefix (incl. key order), M."""
    g = load_golden("toy")
    trie = ParallelTokenCharacterTrie(tokens(unflat(g["blob"], g["lens"])), device="cpu")
    assert trie.children == [{}, {}, {(None, 2): 1}, {(None, 0): 0, 98: 2}, {}, {(None, 1): 4}, {}, {(None, 3): 6},
                             {62: 7}, {115: 8}, {111: 9}, {101: 10}, {97: 3, 98: 5, 60: 11}]
    assert [list(c) for c in trie.children][3] == [(None, 0), 98]  # insertion order of edges
    assert trie.word2leaf == {(b"a", 0): 0, (b"b", 1): 4, (b"ab", 2): 1, (b"<eos>", 3): 6}
    assert trie.leaf2word == {0: (b"a", 0), 4: (b"b", 1), 1: (b"ab", 2), 6: (b"<eos>", 3)}
    n2p = trie.node2prefix
    assert list(n2p.keys()) == g["node2prefix_nodes"].tolist()
    flat, at = g["node2prefix_flat"].tolist(), 0
    for node, n in zip(g["node2prefix_nodes"].tolist(), g["node2prefix_lens"].tolist()):
        assert n2p[node] == flat[at:at + n]
        at += n
    assert n2p[0] is n2p[3]  # a leaf shares its parent's prefix list (base.py:88-90)
    M = trie.M
    assert M.layout == torch.sparse_csr and tuple(M.shape) == (4, 13)
    assert M.crow_indices().tolist() == [0, 3, 6, 10, 17]
    assert M.col_indices().tolist() == [0, 3, 12, 4, 5, 12, 1, 2, 3, 12, 6, 7, 8, 9, 10, 11, 12]
    assert trie._build_parent_map()[0] == 3 and 12 not in trie._build_parent_map()


def test_sentinel_plain_bytes_and_empty_token():
    g = load_golden("sentinel")
    eos = EOS()
    with pytest.warns(DeprecationWarning, match="Passing plain bytes to TokenCharacterTrie is deprecated"):
        trie = TokenCharacterTrie([Token(0, b"ab"), Token(1, b""), Token(2, b"ab"), Token(3, b"a"), eos, b"plain"])
    assert len(trie.children) == int(g["n_nodes"])
    assert np.array_equal(trie.idx_to_leaf, g["idx_to_leaf"]) and np.array_equal(trie.ordering, g["ordering"])
    keys = [(-(2 + k[1]) if isinstance(k, tuple) else (-1 if k is eos else k)) for c in trie.children for k in c]
    assert keys == g["children_key"].tolist()
    assert [v for c in trie.children for v in c.values()] == g["children_val"].tolist()
    assert trie.word2leaf[eos] == trie.idx_to_leaf[4][1] and trie.word2leaf[b"plain"] == trie.idx_to_leaf[5][1]
    assert trie.node2prefix[trie.word2leaf[eos]] == [eos]
    assert trie.node2prefix[trie.word2leaf[(b"", 1)]] == []


@pytest.mark.parametrize("V", [50257, 128256])
def test_baseline_sizes_layout_digests(V):
    g = load_golden(f"synth{V}")
    trie = TokenCharacterTrie(synth_vocab(V))
    assert len(trie) == int(g["n_nodes"]) and trie.root == int(g["root"]) and trie._engine.nnz == int(g["nnz"])
    ptr, idx = jump_csr(trie)
    assert np.array_equal(digest(trie.idx_to_leaf), g["idx_to_leaf_digest"])
    assert np.array_equal(digest(trie.ordering), g["ordering_digest"])
    assert np.array_equal(digest(ptr), g["jump_ptr_digest"]) and np.array_equal(digest(idx), g["jump_idx_digest"])


def test_builder_equals_oracle_on_random_vocabularies():
    rng = np.random.default_rng(0)
    for trial in range(20):
        V = int(rng.integers(1, 400))
        alphabet = int(rng.integers(1, 6))
        items = [bytes(rng.integers(0, alphabet, size=int(rng.integers(0, 9)), dtype=np.uint8).tolist()) for _ in range(V)]
        trie = TokenCharacterTrie(tokens(items))  # duplicates by bytes are fine: ids differ
        o = oracle.OracleTrie(items)
        assert trie.children == o.children and trie.root == o.root
        assert np.array_equal(trie.idx_to_leaf, o.idx_to_leaf) and np.array_equal(trie.ordering, o.ordering)
        lay = trie._layout
        for n in range(len(trie)):  # subtree(n) = ids [n - size + 1, n]; leaves of n = DFS ranks [lo, hi)
            kids = lay["child_idx"][lay["child_ptr"][n]:lay["child_ptr"][n + 1]]
            if len(kids):
                assert lay["lo"][n] == lay["lo"][kids[0]] and lay["hi"][n] == lay["hi"][kids[-1]]
                assert (np.diff(kids) > 0).all() and kids[-1] == n - 1
        assert np.array_equal(lay["leaf_node"][lay["perm"]], np.sort(lay["leaf_node"]))


def test_long_token_needs_no_recursion():
    items = [bytes([1]) * 5000, bytes([1]) * 4999 + bytes([2])]
    trie = TokenCharacterTrie(tokens(items))
    assert len(trie) == 5000 + 1 + 1 + 2 and trie._engine.nnz == 5002 + 5002


def test_empty_vocabulary():
    trie = TokenCharacterTrie([])
    assert len(trie) == 1 and trie.root == 0 and trie.children == [{}] and trie.idx_to_leaf.shape == (0, 2)


def test_duplicate_word_error():
    with pytest.warns(DeprecationWarning):
        with pytest.raises(ValueError, match="Duplicate word in vocabulary"):
            TokenCharacterTrie(decode=[b"hello", b"world", b"hello"])
    with pytest.raises(ValueError, match="Duplicate word in vocabulary"):
        TokenCharacterTrie(decode=[Token(0, b"test"), Token(1, b"other"), Token(0, b"test")])


def test_non_token_iterables_warn_and_key_by_object():
    decode = [Token(0, b"hello"), b"world", Token(2, b"test"), b"data"]
    with pytest.warns(DeprecationWarning, match="Passing plain bytes to TokenCharacterTrie is deprecated"):
        trie = TokenCharacterTrie(decode=decode)
    assert (b"hello", 0) in trie.word2leaf and b"world" in trie.word2leaf
    assert (b"test", 2) in trie.word2leaf and b"data" in trie.word2leaf


def test_duplicate_byte_strings_get_distinct_leaves():
    vocab = [Token(5, b"test"), Token(10, b"test")]
    trie = TokenCharacterTrie(decode=vocab)
    assert (b"test", 5) in trie.word2leaf and (b"test", 10) in trie.word2leaf
    assert trie.word2leaf[(b"test", 5)] != trie.word2leaf[(b"test", 10)]
    assert len(trie.idx_to_leaf) == 2 and set(trie.leaf2word) == set(trie.idx_to_leaf[:, 1].tolist())


def test_parallel_device_argument():
    vocab = [Token(0, b"a"), Token(1, b"b"), Token(2, b"c")]
    with pytest.raises(ValueError):
        ParallelTokenCharacterTrie(decode=vocab, device="invalid")
    with pytest.raises(ValueError):
        ParallelTokenCharacterTrie(decode=vocab, device="cuda:1")
    with pytest.raises(TypeError):
        ParallelTokenCharacterTrie(decode=vocab, device="cpu", bogus=1)
    trie = ParallelTokenCharacterTrie(decode=vocab, device="cpu")
    assert trie.device == "cpu"
    processed = trie._preprocess_ws(np.array([[0.5, 0.5, 0.5], [0.1, 0.5, 0.5]]))
    assert isinstance(processed, torch.Tensor) and processed.device.type == trie.device and processed.dtype == torch.float32
    processed = trie._preprocess_ws([[0.5, 0.5, 0.5], [0.1, 0.5, 0.5]])
    assert processed.dtype == torch.float32 and tuple(processed.shape) == (2, 3)
    with pytest.raises(AssertionError):
        trie._preprocess_ws([[0.5, 0.5]])
    seq = TokenCharacterTrie(decode=vocab)
    assert isinstance(seq._preprocess_ws(torch.zeros(3)), np.ndarray)


def test_no_cpu_fallback_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    trie = ParallelTokenCharacterTrie([Token(0, b"a"), Token(1, b"b")], device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        trie.weight_sum(torch.tensor([0.5, 0.5]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TokenCharacterTrie([Token(0, b"a")]).weight_max([1.0])
    from genlm_backend_b200 import smc

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        smc.masked_logsumexp_sample(torch.zeros(4))


def test_visualize_contract():
    trie = TokenCharacterTrie([Token(0, b"a"), Token(1, b"b")])
    try:
        import graphviz  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            trie.visualize()
        return
    trie.visualize()
    trie.visualize(torch.tensor([0.1] * len(trie.children)))
    with pytest.raises(ValueError):
        trie.visualize(torch.tensor([0.1] * (len(trie.children) + 1)))


# ---- Token (tests/test_token.py of the reference) ---------------------------------------------------------------
def test_token_semantics():
    a, b, c = Token(1, b"hello"), Token(2, b"hello"), Token(1, b"other")
    assert a != b and a == c and hash(a) == hash(c) and len({a, b, c}) == 2
    assert a == b"hello" and bytes(a) == b"hello" and a.byte_string == b"hello" and type(a.byte_string) is bytes
    assert a < b and b > a and a <= c and a >= c
    assert not (a > b"hello") and a >= b"hello"
    assert b"".join([a, b]) == b"hellohello" and a.decode() == "hello" and a[0] == 104
    assert pickle.loads(pickle.dumps(a)).token_id == 1 and copy.deepcopy(b).token_id == 2
    assert Token.as_bytes(a) == b"hello" and Token.as_bytes(b"x") == b"x"
    assert Token.is_plain_bytes(b"x") and not Token.is_plain_bytes(a)
    assert repr(a) == "Token(token_id=1, byte_string=b'hello')"
    with pytest.raises(TypeError):
        Token("1", b"x")
    with pytest.raises(TypeError):
        Token(1, "x")


def test_async_invalid_backend():
    with pytest.raises(ValueError):
        AsyncTokenCharacterTrie.from_vocab(["a", "b", "c"], backend="invalid")
