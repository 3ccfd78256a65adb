"""Multi-rank host logic on CPU: two gloo processes shard a batch by rows exactly as bench.py / the multi-device API
do, compute their block (the oracle stands in for the GPU kernels here: there is no GPU in the CPU suite), and
all-gather the node masses.  The result must equal the unsharded computation bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from genlm_backend_b200.sharding import row_block, row_blocks, all_gather_rows  # noqa: E402


def test_row_blocks_partition_every_batch():
    for n in [0, 1, 2, 3, 7, 64, 1000, 1024, 8192]:
        for world in [1, 2, 3, 4, 8]:
            blocks = row_blocks(n, world)
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        row_block(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from genlm_backend_b200 import TokenCharacterTrie
    from genlm_backend_b200.synthetic import synth_vocab, dirichlet_rows

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        V = 700
        trie = TokenCharacterTrie(synth_vocab(V, seed=3))  # host builder only
        lay = trie._layout
        o = oracle.OracleLayout(trie.idx_to_leaf, lay["child_ptr"], lay["child_idx"])
        ws = dirichlet_rows(n_rows, V, alpha=0.3, seed=5)  # every rank derives the same global batch
        lo, hi = row_block(n_rows, world, rank)
        local = torch.from_numpy(o.weight_sum(ws[lo:hi]) if hi > lo else np.zeros((0, len(trie))))
        full = all_gather_rows(local, n_rows)
        want = torch.from_numpy(o.weight_sum(ws)) if n_rows else torch.zeros((0, len(trie)), dtype=torch.float64)
        ok = full.shape == want.shape and torch.equal(full, want)
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
            f.write("ok" if ok else f"mismatch {tuple(full.shape)} {tuple(want.shape)}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [8, 5, 1])
def test_two_ranks_shard_rows_and_gather(tmp_path, n_rows):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_rows, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"rank{r}.txt").read_text() == "ok"
