"""Device-side driver shared by the trie classes: owns the C handle, keeps the plan metadata resident
per GPU, caches scratch buffers, and launches ``gt_weight_reduce`` on torch's current stream.

PyTorch is used for device memory, streams and host<->device copies only; every reduction runs in the
hand-written kernels behind the C ABI (``csrc/trie_kernels.cu``).  There is no CPU path.
"""
import ctypes
import os

import numpy as np
import torch

from .. import _lib
from .._lib import lib, check

_IN_TYPES = {
    torch.float32: _lib.GT_F32,
    torch.float64: _lib.GT_F64,
    torch.float16: _lib.GT_F16,
    torch.bfloat16: _lib.GT_BF16,
}
_OUT_TYPES = {torch.float32: _lib.GT_F32, torch.float64: _lib.GT_F64}
OPS = {"sum": _lib.GT_OP_SUM, "max": _lib.GT_OP_MAX}

# rows the per-stream scratch is sized for: gt_workspace_bytes caps the staging part at one 64-row chunk (34 MB at 128k
# tokens: it stays in L2 between the permute and the tile kernel) and keeps spanning-node pieces for up to 1,024 rows, so
# that one span kernel serves a whole large batch
_WORKSPACE_ROWS = int(os.environ.get("GT_WORKSPACE_ROWS", "1024"))
_MAX_WORKSPACES = 16  # per engine: (device, stream) pairs we keep scratch for


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "genlm_backend_b200 computes trie masses with CUDA kernels only (sm_100a); "
            "no CUDA device is available and there is no CPU fallback"
        )


class TrieEngine:
    """Owns a ``gt_trie*`` and everything resident on the GPUs for it."""

    def __init__(self, symbols, offsets, n_tokens):
        symbols = np.ascontiguousarray(symbols, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        handle = ctypes.c_void_p()
        check(lib.gt_build(symbols.ctypes.data, offsets.ctypes.data, int(n_tokens), ctypes.byref(handle)), "gt_build")
        self._handle = handle
        self.V = int(lib.gt_num_tokens(handle))
        self.N = int(lib.gt_num_nodes(handle))
        self.nnz = int(lib.gt_num_reach(handle))
        self._uploaded = set()
        self._workspaces = {}
        self._ld32 = None

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                lib.gt_free(h)
            except Exception:
                pass
            self._handle = None

    # ---- host layout ---------------------------------------------------------------------------------
    def layout(self):
        V, N = self.V, self.N
        out = {
            "leaf_node": np.empty(V, np.int32), "parent": np.empty(N, np.int32), "edge_label": np.empty(N, np.int32),
            "child_ptr": np.empty(N + 1, np.int32), "child_idx": np.empty(max(N - 1, 0), np.int32),
            "perm": np.empty(V, np.int32), "lo": np.empty(N, np.int32), "hi": np.empty(N, np.int32),
        }
        check(lib.gt_export_layout(self._handle, *[a.ctypes.data for a in out.values()]), "gt_export_layout")
        return out

    def reachability(self):
        rows = np.empty(self.nnz, np.int64)
        cols = np.empty(self.nnz, np.int64)
        check(lib.gt_export_reachability(self._handle, rows.ctypes.data, cols.ctypes.data), "gt_export_reachability")
        return rows, cols

    def plan(self, tile_leaves=0):
        check(lib.gt_plan(self._handle, tile_leaves), "gt_plan")

    def plan_info(self):
        self.plan()
        info = _lib.PlanInfo()
        check(lib.gt_get_plan_info(self._handle, ctypes.byref(info)), "gt_get_plan_info")
        return {name: getattr(info, name) for name, _ in info._fields_}

    def plan_array(self, name):
        self.plan()
        es = ctypes.c_int32()
        n = lib.gt_export_plan_array(self._handle, name.encode(), None, 0, ctypes.byref(es))
        if n < 0:
            check(1, "gt_export_plan_array")
        arr = np.empty(n, np.uint16 if es.value == 2 else np.int32)
        lib.gt_export_plan_array(self._handle, name.encode(), arr.ctypes.data, n, ctypes.byref(es))
        return arr

    # ---- device ----------------------------------------------------------------------------------------
    def ensure_device(self, index):
        if index not in self._uploaded:
            check(lib.gt_upload(self._handle, int(index)), "gt_upload")
            self._uploaded.add(index)

    def _workspace(self, index, stream, rows):
        """Scratch for launches on ``stream`` (a raw ``cudaStream_t``) of device ``index``.  One buffer per
        (device, stream): calls on one stream are ordered, so they may share it; calls on different streams (or
        threads using different streams) never do."""
        need = int(lib.gt_workspace_bytes(self._handle, min(max(rows, 1), _WORKSPACE_ROWS)))
        key = (index, int(stream or 0))
        buf = self._workspaces.get(key)
        if buf is None or buf.numel() < need:
            if buf is None and len(self._workspaces) >= _MAX_WORKSPACES:
                self._workspaces.pop(next(iter(self._workspaces)))  # the caching allocator keeps the block stream-ordered
            # allocated on `stream` (the caller made it current), so a later reuse of the block is ordered after our kernels
            buf = torch.empty(max(need, 4096), dtype=torch.uint8, device=torch.device("cuda", index))
            self._workspaces[key] = buf
        return buf

    def row_stride(self, dtype=torch.float32):
        """Row stride (elements) of the output slabs the engine allocates: N rounded up to a whole number of
        128-byte lines, so every row starts line-aligned and the emit stores of all rows of a row group are
        line-aligned together."""
        per_line = 128 // torch.empty((), dtype=dtype).element_size()
        return (self.N + per_line - 1) // per_line * per_line

    def alloc_out(self, B, dtype, device):
        """``[B, N]`` view of a ``[B, row_stride]`` device slab (rows are padded to 128-byte lines)."""
        ld = self.row_stride(dtype)
        return torch.empty((max(B, 1), ld), dtype=dtype, device=device)[:B, : self.N]

    def reduce(self, ws, ops, out_dtype=torch.float32, log_input=False, out_sum=None, out_max=None, phases=0, dfs_order=False):
        """Launch the mass kernels for a ``[B, V]`` CUDA tensor on its device's current stream.

        Returns ``(out_sum, out_max)`` device tensors of shape ``[B, N]`` (``None`` for an op not asked for).
        Nothing is synchronised here.  ``dfs_order``: column ``r`` of ``ws`` is the weight of item ``perm[r]`` (the rows
        are already in DFS leaf order, ``GT_FLAG_DFS_ORDER``).
        """
        require_cuda()
        if not (isinstance(ws, torch.Tensor) and ws.is_cuda and ws.dim() == 2):
            raise ValueError("reduce expects a 2-D CUDA tensor")
        if ws.shape[1] != self.V:
            raise AssertionError([ws.shape[1], self.V])
        if ws.dtype not in _IN_TYPES:
            ws = ws.to(torch.float32)
        if ws.shape[0] > 1 and ws.shape[1] > 1 and ws.stride(1) != 1:
            ws = ws.contiguous()
        elif ws.shape[1] > 1 and ws.stride(1) != 1:
            ws = ws.contiguous()
        B = ws.shape[0]
        index = ws.device.index
        self.ensure_device(index)
        opmask = 0
        for op in ops:
            opmask |= OPS[op]

        def _out(t):
            if t is None:
                return self.alloc_out(B, out_dtype, ws.device)
            if not (t.is_cuda and t.device == ws.device and t.dtype == out_dtype and t.shape == (B, self.N) and (t.stride(1) == 1 or self.N <= 1)):
                raise ValueError("out tensor must be a [B, N] tensor of the output dtype on the input's device")
            return t

        out_sum = _out(out_sum) if opmask & _lib.GT_OP_SUM else None
        out_max = _out(out_max) if opmask & _lib.GT_OP_MAX else None
        if B == 0:
            return out_sum, out_max
        ld_out = (out_sum if out_sum is not None else out_max).stride(0) if B > 1 else self.N
        if out_sum is not None and out_max is not None and B > 1 and out_sum.stride(0) != out_max.stride(0):
            raise ValueError("out_sum and out_max must share a row stride")
        ld_ws = ws.stride(0) if B > 1 else max(self.V, 1)
        with torch.cuda.device(index):
            stream = torch.cuda.current_stream(index).cuda_stream
            work = self._workspace(index, stream, B)
            check(
                lib.gt_weight_reduce(
                    self._handle, ws.data_ptr(), _IN_TYPES[ws.dtype], B, ld_ws,
                    out_sum.data_ptr() if out_sum is not None else None,
                    out_max.data_ptr() if out_max is not None else None,
                    _OUT_TYPES[out_dtype], ld_out, opmask,
                    (_lib.GT_FLAG_LOG_INPUT if log_input else 0) | (_lib.GT_FLAG_DFS_ORDER if dfs_order else 0) | int(phases),
                    work.data_ptr(), work.numel(), stream,
                ),
                "gt_weight_reduce",
            )
        return out_sum, out_max

    def reduce_raw(self, ws, opmask, log_input, index, stream_ptr):
        """``reduce`` without the checks, for the few-row latency path: ``ws`` is a contiguous float32 / fp16 / bf16
        ``[B, V]`` tensor on device ``index``, which is the current device; launches on the raw stream ``stream_ptr``.
        Returns the padded ``[B, row_stride]`` float32 slabs ``(sum, max)`` (``None`` for an op not in ``opmask``)."""
        B = ws.shape[0]
        if index not in self._uploaded:
            self.ensure_device(index)
        ld = self._ld32
        if ld is None:
            ld = self._ld32 = self.row_stride(torch.float32)
        out_sum = torch.empty((B, ld), dtype=torch.float32, device=ws.device) if opmask & 1 else None
        out_max = torch.empty((B, ld), dtype=torch.float32, device=ws.device) if opmask & 2 else None
        work = self._workspace(index, stream_ptr, B)
        rc = lib.gt_weight_reduce(
            self._handle, ws.data_ptr(), _IN_TYPES[ws.dtype], B, self.V,
            out_sum.data_ptr() if out_sum is not None else None, out_max.data_ptr() if out_max is not None else None,
            _lib.GT_F32, ld, opmask, _lib.GT_FLAG_LOG_INPUT if log_input else 0, work.data_ptr(), work.numel(), stream_ptr)
        if rc:
            check(rc, "gt_weight_reduce")
        return out_sum, out_max

    def download_raw(self, slab, host_out, stream_ptr):
        """Pitched D2H copy of a padded float32 slab from ``reduce_raw`` into a C-contiguous ``[B, N]`` pinned tensor."""
        rc = lib.gt_download_rows(host_out.data_ptr(), self.N * 4, slab.data_ptr(), slab.stride(0) * 4, self.N * 4, slab.shape[0], stream_ptr)
        if rc:
            check(rc, "gt_download_rows")

    def download(self, dev_out, host_out, stream):
        """Pitched D2H copy (``gt_download_rows``) of a ``[rows, N]`` device result (row stride = the slab's padded
        stride) into a C-contiguous ``[rows, N]`` page-locked host tensor, on ``stream``.  The caller keeps both alive
        until it has synchronised the stream."""
        rows = dev_out.shape[0]
        if rows == 0 or self.N == 0:
            return
        es = dev_out.element_size()
        check(
            lib.gt_download_rows(host_out.data_ptr(), host_out.stride(0) * es if rows > 1 else self.N * es, dev_out.data_ptr(),
                                 dev_out.stride(0) * es if rows > 1 else self.N * es, self.N * es, rows, stream.cuda_stream),
            "gt_download_rows",
        )

    # ---- read-outs that keep the [B, N] slab on the GPU --------------------------------------------------------
    def gather_nodes(self, mass, node_ids, normalizer=None, log=False):
        """``out[b, k] = mass[b, node_ids[b, k]]`` (``node_ids`` 1-D: shared by all rows), optionally divided by
        ``mass[b, normalizer[b]]`` and / or returned as logs.  Everything stays on ``mass``'s device."""
        require_cuda()
        if not (isinstance(mass, torch.Tensor) and mass.is_cuda and mass.dim() == 2 and mass.dtype in _OUT_TYPES):
            raise ValueError("mass must be a 2-D float32 / float64 CUDA tensor")
        if mass.shape[1] != self.N:
            raise AssertionError([mass.shape[1], self.N])
        if mass.shape[1] > 1 and mass.stride(1) != 1:
            mass = mass.contiguous()
        B, dev = mass.shape[0], mass.device
        ids = torch.as_tensor(node_ids).to(device=dev, dtype=torch.int32)
        if ids.dim() == 1:
            ids, ids_ld = ids.contiguous(), 0
        elif ids.dim() == 2 and ids.shape[0] == B:
            ids = ids.contiguous()
            ids_ld = ids.shape[1]
        else:
            raise ValueError("node_ids must be [K] or [B, K]")
        K = ids.shape[-1]
        norm = None
        if normalizer is not None:
            norm = torch.as_tensor(normalizer).to(device=dev, dtype=torch.int32)
            norm = norm.expand(B).contiguous() if norm.dim() == 0 else norm.contiguous()
            if norm.shape != (B,):
                raise ValueError("normalizer must be a node id or one node id per row")
        out = torch.empty((B, K), dtype=mass.dtype, device=dev)
        if B and K:
            with torch.cuda.device(dev.index):
                check(
                    lib.gt_gather_nodes(
                        mass.data_ptr(), _OUT_TYPES[mass.dtype], B, self.N, mass.stride(0) if B > 1 else self.N,
                        ids.data_ptr(), K, ids_ld, norm.data_ptr() if norm is not None else None,
                        _lib.GT_GATHER_LOG if log else 0, out.data_ptr(), K, torch.cuda.current_stream(dev.index).cuda_stream,
                    ),
                    "gt_gather_nodes",
                )
        return out

    def subtree_token_mask(self, nodes, device=None):
        """Keep-bitmask ``int32[B, ceil(V/32)]`` of the tokens under each node (bit ``i % 32`` of word ``i // 32``):
        the layout ``smc.masked_logsumexp_sample`` takes as a bit mask."""
        require_cuda()
        nodes = torch.as_tensor(nodes)
        if device is None:
            device = nodes.device if nodes.is_cuda else torch.device("cuda", torch.cuda.current_device())
        device = torch.device(device)
        nodes = nodes.to(device=device, dtype=torch.int32).reshape(-1).contiguous()
        B, W = nodes.shape[0], (self.V + 31) // 32
        self.ensure_device(device.index)
        bits = torch.empty((B, W), dtype=torch.int32, device=device)
        if B and W:
            with torch.cuda.device(device.index):
                check(
                    lib.gt_subtree_token_mask(self._handle, nodes.data_ptr(), B, bits.data_ptr(), W,
                                              torch.cuda.current_stream(device.index).cuda_stream),
                    "gt_subtree_token_mask",
                )
        return bits
