"""Fused SMC particle step: masked logsumexp + one categorical draw per row, in one kernel launch.

Replaces the per-particle torch sequence of the reference idiom (``README.md:82-91``; temperature variant
``genlm/backend/llm/base.py:131-146``)::

    masked = logps + mask            # mask in {0, -inf}
    logZ   = masked.logsumexp(-1)    # particle weight increment
    tok    = torch.multinomial((masked - logZ).exp(), 1).item()

``masked_logsumexp_sample`` does this for a whole ``[B, V]`` batch of particles without a host sync per row.
Draws use one Philox uniform per row through an exact inverse CDF: distribution-equal to
``torch.multinomial``, not stream-equal.
"""
import weakref

import torch

from . import _lib
from ._lib import lib, check
from .trie._engine import require_cuda

_IN_TYPES = {torch.float32: _lib.GT_F32, torch.float64: _lib.GT_F64, torch.float16: _lib.GT_F16, torch.bfloat16: _lib.GT_BF16}


# Shared additive {0, -inf} masks (the reference idiom: `valid_ids = torch.tensor([...], dtype=torch.float).log()`, built once
# and added to every particle's row, README.md:59-70) are packed into keep-bitmasks once and cached: the kernel then reads
# V/8 mask bytes per row instead of 4V (shared fp32 mask: 0.79 of the HBM peak at 512 rows, bit mask: 0.87).  An entry
# belongs to one live tensor object (weak reference: a new tensor that reuses the address or id of a dead one never
# matches) at one version (torch bumps it on every in-place write), so a mask that is written to is packed again.
_BITMASK_CACHE = {}  # id(mask) -> (weakref to the mask, its _version, bits or None)
_BITMASK_CACHE_MAX = 16


def _pack_keep_bits(keep):
    """bool ``[V]`` -> int32 ``[ceil(V/32)]`` (bit ``i % 32`` of word ``i // 32``)."""
    V = keep.shape[0]
    pad = (-V) % 32
    k = torch.nn.functional.pad(keep, (0, pad)).view(-1, 32).to(torch.int64)
    w = (k << torch.arange(32, device=keep.device, dtype=torch.int64)).sum(-1)
    return torch.where(w >= 2**31, w - 2**32, w).to(torch.int32)


def _cached_bits(mask):
    """Cache entry of this very tensor at its current version, or ``None``."""
    hit = _BITMASK_CACHE.get(id(mask))
    if hit is not None and hit[0]() is mask and hit[1] == mask._version:
        return hit
    return None


def _shared_additive_as_bits(mask):
    """int32 keep-bitmask of a 1-D additive float mask whose values are all 0 or -inf, else ``None``."""
    hit = _cached_bits(mask)
    if hit is None:
        keep = mask == 0
        exact = bool((keep | (mask == float("-inf"))).all())  # one sync per distinct mask
        for k in [k for k, h in _BITMASK_CACHE.items() if h[0]() is None]:  # entries of dead tensors
            del _BITMASK_CACHE[k]
        if len(_BITMASK_CACHE) >= _BITMASK_CACHE_MAX:
            _BITMASK_CACHE.pop(next(iter(_BITMASK_CACHE)))
        hit = (weakref.ref(mask), mask._version, _pack_keep_bits(keep) if exact else None)
        _BITMASK_CACHE[id(mask)] = hit
    return hit[2]


def _mask_args(mask, B, V, device):
    if mask is None:
        return None, _lib.GT_MASK_NONE, 0, None
    if not isinstance(mask, torch.Tensor):
        mask = torch.as_tensor(mask)
    mask = mask.to(device)
    if mask.dtype == torch.bool:
        kind, m = _lib.GT_MASK_BOOL_U8, mask.view(torch.uint8)  # same bytes: no conversion pass
    elif mask.dtype == torch.uint8:
        kind, m = _lib.GT_MASK_BOOL_U8, mask
    elif mask.dtype == torch.int32 and mask.shape[-1] == (V + 31) // 32:
        kind, m = _lib.GT_MASK_BITS_U32, mask
    elif mask.is_floating_point():
        bits = None
        if mask.dim() == 1 and mask.shape[0] == V and not torch.cuda.is_current_stream_capturing():
            bits = _shared_additive_as_bits(mask)
        elif mask.dim() == 1 and mask.shape[0] == V:
            hit = _cached_bits(mask)
            bits = hit[2] if hit else None  # during graph capture: only what was packed before
        if bits is not None:
            kind, m = _lib.GT_MASK_BITS_U32, bits
        else:
            kind, m = _lib.GT_MASK_ADD_F32, mask.to(torch.float32)
    else:
        raise ValueError(f"unsupported mask dtype {mask.dtype}")
    width = (V + 31) // 32 if kind == _lib.GT_MASK_BITS_U32 else V
    if m.dim() == 1:
        if m.shape[0] != width:
            raise ValueError(f"mask length {m.shape[0]} does not match {width}")
        m = m.contiguous()
        ld = 0
    elif m.dim() == 2:
        if m.shape != (B, width):
            raise ValueError(f"mask shape {tuple(m.shape)} does not match {(B, width)}")
        if m.stride(1) != 1:
            m = m.contiguous()
        ld = m.stride(0) if B > 1 else width
    else:
        raise ValueError("mask must be 1-D (shared) or 2-D (per row)")
    return m, kind, ld, m


def masked_logsumexp_sample(logps, mask=None, temperature=1.0, seed=0, offset=0, check_valid=False):
    """Masked logsumexp and one categorical draw per row.

    Args:
        logps (torch.Tensor): ``[B, V]`` (or ``[V]``) log-probabilities / logits on a CUDA device; fp32, fp16,
            bf16 or fp64.
        mask: ``None``; an additive float mask (``{0, -inf}``, the reference idiom); a bool / uint8 keep-mask;
            or an int32 bitmask of ``ceil(V/32)`` words.  1-D masks are shared by all rows.
        temperature (float): logits are divided by it before masking (``llm/base.py:136``).
        seed, offset (int): Philox key and counter base; row ``b`` uses counter ``offset + b``.
        check_valid (bool): synchronise and raise ``RuntimeError`` like ``torch.multinomial`` when a row has no
            mass or a NaN.

    Returns:
        ``(logZ float32[B], tokens int32[B])`` on the device; rows without mass give ``(-inf, -1)``.
    """
    require_cuda()
    if not isinstance(logps, torch.Tensor) or not logps.is_cuda:
        raise ValueError("logps must be a CUDA tensor")
    squeeze = logps.dim() == 1
    if squeeze:
        logps = logps.unsqueeze(0)
    if logps.dim() != 2:
        raise ValueError("logps must be 1-D or 2-D")
    if logps.dtype not in _IN_TYPES:
        logps = logps.to(torch.float32)
    if logps.shape[1] > 1 and logps.stride(1) != 1:
        logps = logps.contiguous()
    B, V = logps.shape
    dev = logps.device
    m, kind, mask_ld, _keep = _mask_args(mask, B, V, dev)
    logZ = torch.empty(B, dtype=torch.float32, device=dev)
    tok = torch.empty(B, dtype=torch.int32, device=dev)
    if B:
        with torch.cuda.device(dev.index):
            check(
                lib.gt_lse_sample(
                    logps.data_ptr(), _IN_TYPES[logps.dtype], B, V, logps.stride(0) if B > 1 else V,
                    m.data_ptr() if m is not None else None, kind, mask_ld, float(temperature),
                    int(seed) & (2**64 - 1), int(offset) & (2**64 - 1), logZ.data_ptr(), tok.data_ptr(),
                    torch.cuda.current_stream(dev.index).cuda_stream,
                ),
                "gt_lse_sample",
            )
    if check_valid and B:
        bad = tok < 0
        if bool(bad.any()):
            if bool(torch.isnan(logZ[bad]).any()):
                raise RuntimeError("probability tensor contains either `inf`, `nan` or element < 0")
            raise RuntimeError("invalid multinomial distribution (sum of probabilities <= 0)")
    if squeeze:
        return logZ[0], tok[0]
    return logZ, tok
