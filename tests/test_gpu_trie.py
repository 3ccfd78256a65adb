"""GPU parity tests for the trie mass path (run with -m gpu on the B200 box).

Oracle: the C restatement of the reference's numba loops (oracle/trie_oracle.c, base.py:346-393) and golden
outputs of the reference itself (tests/golden).  Bars: weight_max and all indexing bit-exact; weight_sum within
fp32 relative 1e-5 of the fp64 path (north star) -- we assert the much tighter 2e-6 the design guarantees, and
exact zeros.
"""
import hashlib
import warnings

import numpy as np
import pytest
import torch

import oracle
from genlm_backend_b200 import Token, TokenCharacterTrie, ParallelTokenCharacterTrie
from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows
from helpers import load_golden, unflat, tokens, rel_err, EOS

pytestmark = pytest.mark.gpu

SUM_RTOL = 2e-6  # design bound (fp32 pyramid + short fp32 term sums); the contract is 1e-5


def oracle_for(trie):
    lay = trie._layout
    return oracle.OracleLayout(trie.idx_to_leaf, lay["child_ptr"], lay["child_idx"])


def check_against(trie, ws, want_sum, want_max, f64=False):
    """ws: numpy [B, V]; wants: float64 [B, N] from the oracle / reference."""
    have_sum = trie.batch_weight_sum(torch.tensor(ws))
    have_max = trie.batch_weight_max(torch.tensor(ws))
    assert have_sum.shape == want_sum.shape and have_max.shape == want_max.shape
    assert have_sum.dtype == (np.float64 if f64 else np.float32)
    r, z = rel_err(have_sum, want_sum)
    assert r <= (1e-12 if f64 else SUM_RTOL), r
    assert z == 0.0
    want_max = want_max if f64 else want_max.astype(np.float32)
    assert np.array_equal(have_max, want_max)


@pytest.mark.parametrize("name", ["toy", "edge", "synth3000"])
@pytest.mark.parametrize("cls", [ParallelTokenCharacterTrie, TokenCharacterTrie])
def test_golden_reference_outputs(name, cls):
    g = load_golden(name)
    trie = cls(tokens(unflat(g["blob"], g["lens"])))
    f64 = cls is TokenCharacterTrie
    check_against(trie, g["ws"], g["seq_sum"], g["seq_max"], f64=f64)
    if not f64:  # the reference's own fp32 torch path, to its own tolerance (tests/test_trie.py:108)
        np.testing.assert_allclose(trie.batch_weight_sum(torch.tensor(g["ws"])), g["par_sum"], rtol=1e-5, atol=1e-8)
        assert np.array_equal(trie.batch_weight_max(torch.tensor(g["ws"])), g["par_max"])


def test_golden_sentinel_and_plain_bytes():
    g = load_golden("sentinel")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dec = [Token(0, b"ab"), Token(1, b""), Token(2, b"ab"), Token(3, b"a"), EOS(), b"plain"]
        trie = ParallelTokenCharacterTrie(dec)
    check_against(trie, g["ws"], g["seq_sum"], g["seq_max"])


@pytest.mark.parametrize("V", [50257, 128256])
def test_golden_baseline_sizes(V):
    """BASELINE.json configs 1 and 2: sampled node values, probe dot products and the max digest of the reference."""
    g = load_golden(f"synth{V}")
    trie = ParallelTokenCharacterTrie(synth_vocab(V))
    assert len(trie) == int(g["n_nodes"]) and trie.root == int(g["root"])
    ws = dirichlet_rows(2, V, alpha=0.1, seed=1)
    hs = trie.batch_weight_sum(torch.tensor(ws))
    hm = trie.batch_weight_max(torch.tensor(ws))
    pick = g["pick"]
    r, z = rel_err(hs[:, pick], g["seq_sum_pick"])
    assert r <= SUM_RTOL and z == 0.0
    assert np.array_equal(hm[:, pick], g["seq_max_pick"].astype(np.float32))
    probe = np.random.default_rng(4).standard_normal(len(trie))
    np.testing.assert_allclose(hs.astype(np.float64) @ probe, g["seq_sum_probe"], rtol=1e-5)
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(hm).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, g["seq_max_digest"])  # weight_max bit-exact over all 2 x N values
    # sequential class at full size: float64 end to end
    seq = TokenCharacterTrie(synth_vocab(V))
    h64 = seq.batch_weight_sum(torch.tensor(ws))
    r, z = rel_err(h64[:, pick], g["seq_sum_pick"])
    assert r <= 1e-12 and z == 0.0


@pytest.mark.parametrize("V,T", [(300, 1024), (1024, 1024), (5000, 1024), (5000, 2048), (20011, 2048), (20011, 1024), (20480, 1024)])
def test_plan_shapes_against_oracle(V, T):
    """Different vocabulary / tile sizes (more spanning nodes, partial and exactly full tiles, odd V: unaligned rows)."""
    trie = ParallelTokenCharacterTrie(synth_vocab(V, seed=3))
    trie._engine.plan(T)
    o = oracle_for(trie)
    ws = dirichlet_rows(5, V, alpha=0.1, seed=2)
    check_against(trie, ws, o.weight_sum(ws), o.weight_max(ws))


@pytest.mark.parametrize("B", [0, 1, 2, 3, 7, 64, 131])
def test_batch_sizes(B):
    V = 4099
    trie = ParallelTokenCharacterTrie(synth_vocab(V, seed=5))
    o = oracle_for(trie)
    ws = dirichlet_rows(max(B, 1), V, alpha=0.3, seed=B)[:B]
    hs = trie.batch_weight_sum(torch.tensor(ws))
    hm = trie.batch_weight_max(torch.tensor(ws))
    assert hs.shape == (B, len(trie)) and hm.shape == (B, len(trie))
    if B:
        r, z = rel_err(hs, o.weight_sum(ws))
        assert r <= SUM_RTOL and z == 0.0
        assert np.array_equal(hm, o.weight_max(ws).astype(np.float32))


def test_small_workspace_chunks_the_batch():
    """The C side chunks a batch that does not fit the caller's scratch."""
    from genlm_backend_b200 import _lib

    V, B = 3001, 37
    trie = ParallelTokenCharacterTrie(synth_vocab(V, seed=6))
    eng = trie._engine
    ws = torch.tensor(dirichlet_rows(B, V, alpha=0.2, seed=9), device="cuda")
    eng.ensure_device(0)
    out = torch.empty((B, len(trie)), dtype=torch.float32, device="cuda")
    need_one = int(_lib.lib.gt_workspace_bytes(eng._handle, 1))
    work = torch.empty(need_one // 2 * 5 // 2, dtype=torch.uint8, device="cuda")  # ~5 float rows
    _lib.check(_lib.lib.gt_weight_reduce(eng._handle, ws.data_ptr(), _lib.GT_F32, B, ws.stride(0), out.data_ptr(), None,
                                         _lib.GT_F32, out.stride(0), _lib.GT_OP_SUM, 0, work.data_ptr(), work.numel(),
                                         torch.cuda.current_stream().cuda_stream), "gt_weight_reduce")
    torch.cuda.synchronize()
    r, z = rel_err(out.cpu().numpy(), oracle_for(trie).weight_sum(ws.cpu().numpy()))
    assert r <= SUM_RTOL and z == 0.0


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float64])
def test_input_dtypes(dtype):
    V = 6007
    trie = ParallelTokenCharacterTrie(synth_vocab(V, seed=8))
    o = oracle_for(trie)
    ws = torch.tensor(dirichlet_rows(3, V, alpha=0.5, seed=4)).to(dtype)
    as32 = ws.to(torch.float32).numpy()  # the kernel converts each element to fp32 first
    hs, hm = trie.batch_weight_tensor(ws.cuda(), ops=("sum", "max"))
    r, z = rel_err(hs.cpu().numpy(), o.weight_sum(as32))
    assert r <= SUM_RTOL and z == 0.0
    assert np.array_equal(hm.cpu().numpy(), o.weight_max(as32).astype(np.float32))
    # the sequential class keeps fp64 inputs in fp64
    if dtype == torch.float64:
        seq = TokenCharacterTrie(synth_vocab(V, seed=8))
        w64 = np.random.default_rng(0).dirichlet(np.full(V, 0.5), size=2)
        r, z = rel_err(seq.batch_weight_sum(w64), o.weight_sum(w64))
        assert r <= 1e-12 and z == 0.0
        assert np.array_equal(seq.batch_weight_max(w64), o.weight_max(w64))


def test_log_input_fuses_exp():
    V = 5003
    trie = ParallelTokenCharacterTrie(synth_vocab(V, seed=10))
    ws = dirichlet_rows(3, V, alpha=0.1, seed=11)
    with np.errstate(divide="ignore"):
        logw = np.log(ws)  # -inf for exact zeros
    hs, hm = trie.batch_weight_tensor(torch.tensor(logw).cuda(), ops=("sum", "max"), log_input=True)
    back = np.exp(logw.astype(np.float64))
    o = oracle_for(trie)
    r, z = rel_err(hs.cpu().numpy(), o.weight_sum(back))
    assert r <= 1e-5 and z == 0.0  # expf adds ~1e-7 per leaf on top of the reduction error
    # expf results below the smallest normal fp32 are denormals with few mantissa bits: absolute bound there
    np.testing.assert_allclose(hm.cpu().numpy(), o.weight_max(back), rtol=1e-6, atol=1e-37)
    assert (hm.cpu().numpy()[o.weight_max(back) == 0] == 0).all()


def test_reference_input_forms_and_errors():
    """tests/test_trie.py:210-264 (lists, numpy, tuples of rows) and parallel.py:73 (length assert)."""
    dec = [Token(0, b"a"), Token(1, b"b"), Token(2, b"ab"), Token(3, b"<eos>")]
    par = ParallelTokenCharacterTrie(dec)
    seq = TokenCharacterTrie(dec)
    want = [0.1, 0.2, 0.2, 0.3, 0.2, 0.2, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 1.0]
    for ws in ([0.1, 0.2, 0.2, 0.5], np.array([0.1, 0.2, 0.2, 0.5]), torch.tensor([0.1, 0.2, 0.2, 0.5]),
               torch.tensor([0.1, 0.2, 0.2, 0.5], device="cuda")):
        np.testing.assert_allclose(par.weight_sum(ws), want, rtol=1e-6)
        np.testing.assert_allclose(seq.weight_sum(ws), want, rtol=1e-6)
        assert par.weight_sum(ws).dtype == np.float32 and seq.weight_sum(ws).dtype == np.float64
    rows = (torch.tensor([0.1, 0.2, 0.2, 0.5]), torch.tensor([0.0, 0.3, 0.6, 0.1]))  # what the async wrapper passes
    assert par.batch_weight_max(rows).shape == (2, 13)
    assert seq.batch_weight_max(rows).shape == (2, 13)
    with pytest.raises(AssertionError):
        par.weight_sum(torch.tensor([0.1, 0.2, 0.2]))
    with pytest.raises(AssertionError):
        seq.weight_sum(torch.tensor([0.1, 0.2, 0.2]))
    processed = par._preprocess_ws(np.array([[0.5] * 4, [0.1] * 4]))
    assert isinstance(processed, torch.Tensor) and processed.device.type == par.device and processed.dtype == torch.float32


def test_properties_at_full_size():
    """Size-independent properties at BASELINE config 5's vocabulary (151,665 tokens, rows not 16-byte aligned)."""
    V, B = 151665, 6
    trie = ParallelTokenCharacterTrie(synth_vocab(V))
    assert len(trie) == 407861
    lay = trie._layout
    a = torch.tensor(dirichlet_rows(B, V, alpha=1.0, seed=21), device="cuda")
    b = torch.tensor(dirichlet_rows(B, V, alpha=0.1, seed=22), device="cuda")
    sa, ma = trie.batch_weight_tensor(a, ops=("sum", "max"))
    sb, mb = trie.batch_weight_tensor(b, ops=("sum", "max"))
    sab = trie.batch_weight_sum_tensor(a + b)
    torch.cuda.synchronize()
    # leaves carry their token's weight bit-exactly, for both ops
    leaf = torch.tensor(lay["leaf_node"].astype(np.int64), device="cuda")
    assert torch.equal(sa[:, leaf], a) and torch.equal(ma[:, leaf], a) and torch.equal(mb[:, leaf], b)
    # root: total mass / global max
    np.testing.assert_allclose(sa[:, trie.root].cpu().numpy(), a.double().sum(1).cpu().numpy(), rtol=1e-6)
    assert torch.equal(mb[:, trie.root], b.max(1).values)
    # linearity of the sum
    np.testing.assert_allclose(sab.cpu().numpy(), (sa + sb).cpu().numpy(), rtol=1e-5, atol=1e-30)
    # every internal node: sum == sum of children (fp64 check), max == max of children (exact)
    ptr = torch.tensor(lay["child_ptr"].astype(np.int64), device="cuda")
    idx = torch.tensor(lay["child_idx"].astype(np.int64), device="cuda")
    owner = torch.repeat_interleave(torch.arange(len(trie), device="cuda"), ptr[1:] - ptr[:-1])
    internal = (ptr[1:] - ptr[:-1]) > 0
    for s, m in ((sa, ma), (sb, mb)):
        tot = torch.zeros((B, len(trie)), dtype=torch.float64, device="cuda").index_add_(1, owner, s[:, idx].double())
        err = ((tot - s.double()).abs() / tot.clamp_min(1e-300))[:, internal]
        assert float(err.max()) <= 4e-6
        assert bool(((tot == 0) == (s == 0))[:, internal].all())
        mx = torch.full((B, len(trie)), -1.0, device="cuda").scatter_reduce_(1, owner.expand(B, -1), m[:, idx], reduce="amax")
        assert torch.equal(mx[:, internal], m[:, internal])
    # idempotence: max over the max-leaves again
    assert torch.equal(trie.batch_weight_max_tensor(ma[:, leaf]), ma)


def test_multi_device_row_sharding_matches_single():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    V, B = 9001, 13
    dec = synth_vocab(V, seed=12)
    ws = torch.tensor(dirichlet_rows(B, V, alpha=0.2, seed=13))
    one = ParallelTokenCharacterTrie(dec, devices=[0])
    many = ParallelTokenCharacterTrie(dec, devices=list(range(n)))
    assert np.array_equal(one.batch_weight_sum(ws), many.batch_weight_sum(ws))
    assert np.array_equal(one.batch_weight_max(ws), many.batch_weight_max(ws))


def test_async_autobatch_1024_requests_full_size():
    """BASELINE config 3: AsyncTokenCharacterTrie autobatching 1,024 concurrent weight_sum requests at 128,256 tokens
    (reference: async_impl.py:46-137, tests/test_trie.py:157-207).  Every future gets its own row's result: a sample of
    rows against the oracle, all rows through the root / leaf properties."""
    import asyncio

    from genlm_backend_b200 import AsyncTokenCharacterTrie

    V, R = 128256, 1024
    at = AsyncTokenCharacterTrie.from_vocab(synth_vocab(V), backend="parallel")
    trie = at.trie
    ws = dirichlet_rows(R, V, alpha=0.1, seed=31)
    rows = [torch.tensor(w) for w in ws]

    async def main():
        sums = await asyncio.gather(*[at.weight_sum(r) for r in rows])
        maxes = await asyncio.gather(*[at.weight_max(r) for r in rows[:64]])
        await at.cleanup()
        return sums, maxes

    sums, maxes = asyncio.run(main())
    assert len(sums) == R and all(s.shape == (len(trie),) and s.dtype == np.float32 for s in sums)
    leaf = trie._layout["leaf_node"]
    o = oracle_for(trie)
    pick = [0, 1, 511, 1023]
    want = o.weight_sum(ws[pick].astype(np.float64))
    for j, i in enumerate(pick):
        r, z = rel_err(sums[i], want[j])
        assert r <= SUM_RTOL and z == 0.0
    for i in range(R):  # each request got its own row back
        assert np.array_equal(sums[i][leaf], ws[i])
        assert abs(float(sums[i][trie.root]) - float(ws[i].astype(np.float64).sum())) <= 1e-6
    wm = o.weight_max(ws[:4].astype(np.float64)).astype(np.float32)
    for i in range(4):
        assert np.array_equal(maxes[i], wm[i])
    for i in range(64):
        assert float(maxes[i][trie.root]) == float(ws[i].max())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16, torch.float64])
@pytest.mark.parametrize("V", [5003, 4097, 12289])
def test_rows_of_any_alignment(V, dtype):
    """Row starts that are not 16-byte aligned (odd vocabulary sizes, column-offset views, 2- / 4- / 8-byte elements):
    the bulk-copy permute fetches the aligned span around each segment and fixes up the batch's last elements."""
    trie = ParallelTokenCharacterTrie(synth_vocab(V, seed=V % 13))
    o = oracle_for(trie)
    B = 7
    base = dirichlet_rows(B, V + 9, alpha=0.3, seed=V % 7)
    for off in (0, 1, 3, 6):  # first element of row 0 at byte offset off * itemsize of a 512-byte aligned buffer
        wide = torch.tensor(base).to(dtype).cuda()
        view = wide[:, off:off + V]
        assert view.stride(0) == V + 9
        back = view.to(torch.float64).cpu().numpy()
        hs, hm = trie.batch_weight_tensor(view, ops=("sum", "max"))
        r, z = rel_err(hs.cpu().numpy(), o.weight_sum(back))
        assert r <= SUM_RTOL and z == 0.0, (off, r, z)
        assert np.array_equal(hm.cpu().numpy(), o.weight_max(back).astype(np.float32)), off
        # a batch that ends exactly at the end of its allocation: the last segment's tail is loaded element by element
        tight = view.contiguous()[1:].clone()
        hs2 = trie.batch_weight_sum_tensor(tight)
        r, z = rel_err(hs2.cpu().numpy(), o.weight_sum(back[1:]))
        assert r <= SUM_RTOL and z == 0.0, (off, "tight", r, z)


@pytest.mark.parametrize("V,B,dtype", [(4100, 200, torch.float32), (4104, 131, torch.bfloat16), (4104, 150, torch.float16),
                                       (4098, 140, torch.float64), (20480, 129, torch.float32), (4099, 140, torch.float32),
                                       (4101, 133, torch.float32), (2050, 1100, torch.float32)])
def test_batches_above_the_scratch_rows(V, B, dtype):
    """Batches larger than the engine's scratch (64 rows of staging, spanning-node pieces of 1,024 rows): processed chunk
    by chunk on the C side, one span kernel per span group (B = 1,100: two span groups).  Aligned and unaligned rows,
    every input type, both reductions, partial last chunk and row group, against the oracle."""
    trie = ParallelTokenCharacterTrie(synth_vocab(V, seed=V % 11))
    o = oracle_for(trie)
    pad = (-V) % 8 if V == 4101 else 0  # row stride 4104 elements: aligned rows whose length is not a multiple of 16 bytes
    wide = torch.tensor(dirichlet_rows(B, V + pad, alpha=0.3, seed=B)).to(dtype).cuda()
    ws = wide[:, :V]
    back = ws.to(torch.float64).cpu().numpy()
    hs, hm = trie.batch_weight_tensor(ws, ops=("sum", "max"))
    r, z = rel_err(hs.cpu().numpy(), o.weight_sum(back))
    assert r <= SUM_RTOL and z == 0.0, (r, z)
    assert np.array_equal(hm.cpu().numpy(), o.weight_max(back).astype(np.float32))
    # the float64 pipeline (two rows per work item) through the same chunking
    if dtype == torch.float64:
        seq = TokenCharacterTrie(synth_vocab(V, seed=V % 11))
        h64 = seq.batch_weight_sum(back)
        r, z = rel_err(h64, o.weight_sum(back))
        assert r <= 1e-12 and z == 0.0
        assert np.array_equal(seq.batch_weight_max(back), o.weight_max(back))


@pytest.mark.parametrize("B", [1, 5, 9, 40, 100])
def test_host_results_are_c_contiguous(B):
    """The reference returns ``masses.cpu().numpy()``: C-contiguous [B, N] arrays (parallel.py:103,145).  Ours come out of
    row-padded device slabs through one pitched copy per slice; B <= 8 takes the single-stream latency path."""
    V = 3001
    par = ParallelTokenCharacterTrie(synth_vocab(V, seed=2))
    seq = TokenCharacterTrie(synth_vocab(V, seed=2))
    o = oracle_for(par)
    ws = dirichlet_rows(B, V, alpha=0.3, seed=B)
    want_s, want_m = o.weight_sum(ws), o.weight_max(ws)
    for x in (torch.tensor(ws), torch.tensor(ws).cuda(), [torch.tensor(w) for w in ws]):
        for got, want, exact in ((par.batch_weight_sum(x), want_s, False), (par.batch_weight_max(x), want_m, True)):
            assert got.flags["C_CONTIGUOUS"] and got.shape == (B, len(par)) and got.dtype == np.float32
            if exact:
                assert np.array_equal(got, want.astype(np.float32))
            else:
                r, z = rel_err(got, want)
                assert r <= SUM_RTOL and z == 0.0
    s64 = seq.batch_weight_sum(ws)
    assert s64.flags["C_CONTIGUOUS"] and s64.dtype == np.float64 and rel_err(s64, want_s)[0] <= 1e-12
    both = par.batch_weight_sum_max(torch.tensor(ws))
    assert both[0].flags["C_CONTIGUOUS"] and both[1].flags["C_CONTIGUOUS"]
    rows = par.batch_weight_rows(torch.tensor(ws), "sum")
    assert len(rows) == B and all(r.flags["C_CONTIGUOUS"] and r.shape == (len(par),) for r in rows)
    assert max(r.base.shape[0] if r.base is not None and r.base.ndim == 2 else 1 for r in rows) <= 32  # a kept row pins <= 32 rows
    assert rel_err(np.stack(rows), want_s)[0] <= SUM_RTOL
    one = par.weight_sum(torch.tensor(ws[0]))
    assert one.flags["C_CONTIGUOUS"] and one.shape == (len(par),) and rel_err(one, want_s[0])[0] <= SUM_RTOL
    assert np.array_equal(par.weight_max(ws[0].tolist()), want_m[0].astype(np.float32))


def test_concurrent_streams_and_threads_do_not_share_scratch():
    """The engine keeps one scratch buffer per (device, stream): calls issued on different CUDA streams, and from
    different host threads, run their permute / tile / span kernels concurrently without touching each other's staging."""
    import threading

    V, B = 20011, 48
    trie = ParallelTokenCharacterTrie(synth_vocab(V, seed=14))
    o = oracle_for(trie)
    batches = [torch.tensor(dirichlet_rows(B, V, alpha=0.3, seed=60 + i)).cuda() for i in range(4)]
    want = [(o.weight_sum(x.cpu().numpy()), o.weight_max(x.cpu().numpy()).astype(np.float32)) for x in batches]
    trie.batch_weight_tensor(batches[0], ops=("sum", "max"))  # plan upload, kernel attributes
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in batches]
    results = [None] * len(batches)
    for rep in range(3):  # interleaved launches on four streams, no synchronisation in between
        for i, (x, st) in enumerate(zip(batches, streams)):
            with torch.cuda.stream(st):
                results[i] = trie.batch_weight_tensor(x, ops=("sum", "max"))
    torch.cuda.synchronize()
    assert len(trie._engine._workspaces) >= len(streams)
    for i, (s, m) in enumerate(results):
        r, z = rel_err(s.cpu().numpy(), want[i][0])
        assert r <= SUM_RTOL and z == 0.0, (i, r, z)
        assert np.array_equal(m.cpu().numpy(), want[i][1]), i

    out = [None] * len(batches)

    def worker(i):
        with torch.cuda.stream(streams[i]):
            for _ in range(3):
                out[i] = trie.batch_weight_sum(batches[i])  # host-returning path, its own pipeline streams per call

    # the host-returning path shares one set of pipeline streams and staging buffers per trie: a lock serialises callers
    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(batches))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for i in range(len(batches)):
        assert rel_err(out[i], want[i][0])[0] <= SUM_RTOL


@pytest.mark.parametrize("V,B,dtype", [(5003, 7, torch.float32), (20480, 70, torch.float32), (4099, 5, torch.bfloat16), (3001, 3, torch.float64)])
def test_rows_given_in_dfs_order(V, B, dtype):
    """GT_FLAG_DFS_ORDER: rows whose columns are already in DFS leaf order (ws[:, dfs_token_order]) give bit-identical
    results to the same weights in vocabulary order, for both pipelines, with and without the fused exp."""
    trie = ParallelTokenCharacterTrie(synth_vocab(V, seed=V % 5))
    ws = torch.tensor(dirichlet_rows(B, V, alpha=0.3, seed=B)).to(dtype).cuda()
    order = trie.dfs_token_order.cuda()
    assert sorted(order.tolist()) == list(range(V))
    s0, m0 = trie.batch_weight_tensor(ws, ops=("sum", "max"))
    s1, m1 = trie.batch_weight_tensor(ws[:, order].contiguous(), ops=("sum", "max"), dfs_order=True)
    assert torch.equal(s0, s1) and torch.equal(m0, m1)
    o = oracle_for(trie)
    r, z = rel_err(s1.cpu().numpy(), o.weight_sum(ws.to(torch.float64).cpu().numpy()))
    assert r <= SUM_RTOL and z == 0.0
    logw = ws.to(torch.float32).log()
    l0 = trie.batch_weight_sum_tensor(logw, log_input=True)
    l1 = trie.batch_weight_tensor(logw[:, order].contiguous(), ops=("sum",), log_input=True, dfs_order=True)[0]
    assert torch.equal(l0, l1)
    if dtype == torch.float64:
        seq = TokenCharacterTrie(synth_vocab(V, seed=V % 5))
        a, _ = seq._engine.reduce(ws, ("sum",), out_dtype=torch.float64)
        b, _ = seq._engine.reduce(ws[:, order].contiguous(), ("sum",), out_dtype=torch.float64, dfs_order=True)
        assert torch.equal(a, b)


def test_special_values_and_degenerate_vocabularies():
    """inf / NaN / denormal / all-zero weights and one-token / empty vocabularies against the oracle (the numba loops'
    semantics: sums propagate NaN and inf, `max(total, c)` keeps `total` when `c` is NaN, base.py:346-393)."""
    V = 3001
    par = ParallelTokenCharacterTrie(synth_vocab(V, seed=21))
    o = oracle_for(par)
    ws = dirichlet_rows(6, V, alpha=0.3, seed=3).astype(np.float32)
    ws[1, 17] = np.inf
    ws[2, 1234] = np.nan
    ws[3] = 0.0
    ws[4] *= 1e-42  # denormals
    ws[5, ::2] = 0.0
    want_s, want_m = o.weight_sum(ws), o.weight_max(ws)
    hs, hm = par.batch_weight_sum(torch.tensor(ws)), par.batch_weight_max(torch.tensor(ws))
    fin = np.isfinite(want_s)
    assert np.array_equal(np.isnan(hs), np.isnan(want_s)) and np.array_equal(np.isposinf(hs), np.isposinf(want_s))
    r, z = rel_err(hs[fin], want_s[fin])
    assert r <= SUM_RTOL and z == 0.0
    ok_rows = [0, 1, 3, 4, 5]
    assert np.array_equal(hm[ok_rows], want_m[ok_rows].astype(np.float32))
    # NaN weights: both of the reference's paths ignore or propagate them differently (numba's `max(total, c)` starts at 0
    # and skips a NaN child, torch's amax propagates it); ours is fmax over the leaf range -- a NaN only where every leaf
    # under the node is NaN (here: the leaf and the unary chain above it), the oracle's value everywhere else
    unaffected = ~np.isnan(hm[2])
    assert np.array_equal(hm[2][unaffected], want_m[2][unaffected].astype(np.float32))
    assert np.isnan(hm[2]).sum() >= 1 and np.isnan(hs[2]).sum() >= np.isnan(hm[2]).sum()
    assert (hs[3] == 0).all() and (hm[3] == 0).all()
    # one token, and a vocabulary whose tokens share one long chain
    for dec in ([Token(0, b"a")], [Token(i, b"x" * (i + 1)) for i in range(40)]):
        t = ParallelTokenCharacterTrie(dec)
        oo = oracle_for(t)
        w = np.random.default_rng(1).random((3, len(dec))).astype(np.float32)
        assert rel_err(t.batch_weight_sum(torch.tensor(w)), oo.weight_sum(w))[0] <= SUM_RTOL
        assert np.array_equal(t.batch_weight_max(torch.tensor(w)), oo.weight_max(w).astype(np.float32))
    # the empty vocabulary: the root is the only node and has no mass (base.py:124-130 allocates zeros)
    empty = ParallelTokenCharacterTrie([])
    assert len(empty) == 1
    out = empty.batch_weight_sum(torch.zeros((2, 0)))
    assert out.shape == (2, 1) and (out == 0).all()
