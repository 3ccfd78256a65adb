"""Row sharding of a batch over ranks / devices, and the optional gather of node masses.

The hot path has no cross-row term (one SMC particle per row; reference: ``batch_weight_*`` is row-independent,
``genlm/backend/trie/parallel.py:92-145``), so multi-GPU execution is a contiguous split of the rows with no collective on
the data path.  ``all_gather_rows`` is the optional exchange of BASELINE.json config 5 ("NCCL all-gather of node
masses"): it is never called by the kernels or the timed benchmark path.
"""
import torch
import torch.distributed as dist


def row_block(n_rows, world, rank):
    """Contiguous block ``[lo, hi)`` of ``n_rows`` rows owned by ``rank`` of ``world``: sizes differ by at most one,
    lower ranks take the larger blocks, ranks beyond the row count get an empty block."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    base, extra = divmod(max(int(n_rows), 0), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def row_blocks(n_rows, world):
    return [row_block(n_rows, world, r) for r in range(world)]


def all_gather_rows(local, n_rows, group=None):
    """Gather row blocks (``local``: this rank's ``[hi-lo, N]`` tensor, split as ``row_block`` does) into the full
    ``[n_rows, N]`` tensor on every rank.  Uneven blocks are padded to the largest block for the collective.
    Works with NCCL (CUDA tensors, NVLink/NVSwitch) and gloo (CPU tensors, used by the tests)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = row_block(n_rows, world, rank)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {hi - lo}")
    biggest = -(-n_rows // world) if n_rows else 0
    n_cols = local.shape[1]
    send = local
    if local.shape[0] != biggest:
        send = local.new_zeros((biggest, n_cols))
        send[: local.shape[0]] = local
    gathered = local.new_empty((world * biggest, n_cols))
    dist.all_gather_into_tensor(gathered, send.contiguous(), group=group)
    if world * biggest == n_rows:
        return gathered
    parts = []
    for r, (a, b) in enumerate(row_blocks(n_rows, world)):
        parts.append(gathered[r * biggest: r * biggest + (b - a)])
    return torch.cat(parts, dim=0)
