"""ORACLE -- TEST INFRASTRUCTURE / BASELINE ONLY (never imported by ``genlm_backend_b200``).

Torch restatement of the reference's *library-kernel* path, i.e. what the product replaces on the same GPU:

* ``TorchReferenceTrie`` -- ``ParallelTokenCharacterTrie`` of ``genlm/backend/trie/parallel.py``: the reachability
  matrix ``M`` as sparse CSR (``:55-64``), ``batch_weight_sum`` = ``torch.sparse.mm(ws[:, positions], M)`` (``:92-103``),
  ``batch_weight_max`` = ``zeros.scatter_reduce_(amax, include_self=False)`` over the (leaf, ancestor) pairs
  (``:120-145``), both with and without the ``.cpu().numpy()`` the reference ends on;
* ``smc_step`` -- the SMC particle step of ``README.md:82-87`` batched over particles:
  ``masked = logps + mask; logZ = masked.logsumexp(-1); tok = multinomial((masked - logZ).exp(), 1)``.

Pinned by ``tests/test_oracle.py`` against the outputs of the reference itself (``par_sum`` / ``par_max`` of
``tests/golden/*.npz``).  bench.py times these on the B200 as ``reference_gpu`` / ``sampler.reference``.
"""
import numpy as np
import torch


class TorchReferenceTrie:
    def __init__(self, idx_to_leaf, reach_rows, reach_cols, n_nodes, device):
        self.device = torch.device(device)
        idx_to_leaf = np.asarray(idx_to_leaf)
        self.n_nodes = int(n_nodes)
        self.positions = torch.tensor(idx_to_leaf[:, 0], dtype=torch.long, device=self.device)  # parallel.py:17-19
        self.src_indices = torch.tensor(np.asarray(reach_rows), dtype=torch.long, device=self.device)  # parallel.py:55-56
        self.dst_indices = torch.tensor(np.asarray(reach_cols), dtype=torch.long, device=self.device)
        indices = torch.stack([self.src_indices, self.dst_indices])
        values = torch.ones(indices.shape[1], device=self.device)
        self.M = torch.sparse_coo_tensor(indices, values, (len(idx_to_leaf), self.n_nodes)).to_sparse_csr()  # parallel.py:58-64

    def batch_weight_sum_tensor(self, ws):
        return torch.sparse.mm(ws[:, self.positions], self.M)  # parallel.py:102

    def batch_weight_max_tensor(self, ws):
        leaf_weights = ws[:, self.positions]  # parallel.py:133-145
        batch_size = leaf_weights.shape[0]
        result = torch.zeros((batch_size, self.n_nodes), device=self.device)
        result.scatter_reduce_(dim=1, index=self.dst_indices.expand(batch_size, -1), src=leaf_weights[:, self.src_indices],
                               reduce="amax", include_self=False)
        return result

    def batch_weight_sum(self, ws):
        return self.batch_weight_sum_tensor(ws.to(device=self.device, dtype=torch.float32)).cpu().numpy()  # parallel.py:103

    def batch_weight_max(self, ws):
        return self.batch_weight_max_tensor(ws.to(device=self.device, dtype=torch.float32)).cpu().numpy()  # parallel.py:145


def smc_step(logps, mask=None, generator=None):
    """README.md:82-87 for a batch of particles: returns (logZ [B], token ids [B])."""
    masked = logps if mask is None else logps + mask
    logZ = masked.logsumexp(dim=-1)
    tok = torch.multinomial((masked - logZ[:, None]).exp(), 1, generator=generator)[:, 0]
    return logZ, tok
