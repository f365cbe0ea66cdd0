"""Process-wide runtime state of the B200 path: compute dtype and the 16-bit weight-operand cache.

The residual stream and the master weights stay fp32 (the reference trains in fp32,
P/config/r2r_GOAT_pretrain.json:24 ``"fp16": false``).  GEMM / attention operands are

  * ``torch.float32``  -- parity mode (1e-5 vs the fp32 oracle): SIMT kernels, no operand copies;
  * ``torch.bfloat16`` / ``torch.float16`` -- tensor-core mode (1e-3): tcgen05 kernels read 16-bit
    operand copies of the weights.  Those copies come from (a) the flat shadow buffer the fused
    optimizer maintains (``engine.FlatParams``: zero extra traffic, the AdamW kernel writes them), or
    (b) a per-parameter cache keyed on the parameter's version counter (re-cast after any in-place
    update, e.g. a stock torch optimizer or ``load_state_dict``).
"""
import contextlib

import torch

from . import ops

_compute_dtype = torch.bfloat16

# Bumped by every in-place parameter update that bypasses torch's version counters (the fused optimizer kernel writes
# through raw pointers): part of every operand-cache key below and of modules.KVCache's tags, so a cast made before the
# update is never reused after it.
_generation = [0]


def generation():
    return _generation[0]


def bump_generation():
    _generation[0] += 1


def compute_dtype():
    return _compute_dtype


def set_compute_dtype(dtype):
    global _compute_dtype
    if dtype not in (torch.float32, torch.float16, torch.bfloat16):
        raise TypeError("compute dtype must be float32, float16 or bfloat16")
    _compute_dtype = dtype


@contextlib.contextmanager
def compute(dtype):
    """``with runtime.compute(torch.float32): ...`` -- scoped compute dtype."""
    prev = _compute_dtype
    set_compute_dtype(dtype)
    try:
        yield
    finally:
        set_compute_dtype(prev)


def _adjacent(ts):
    """True when the tensors are contiguous and laid out back to back in one storage."""
    for a, b in zip(ts[:-1], ts[1:]):
        if not (a.is_contiguous() and b.is_contiguous()):
            return False
        if a.dtype != b.dtype or a.data_ptr() + a.numel() * a.element_size() != b.data_ptr():
            return False
        if a.untyped_storage().data_ptr() != b.untyped_storage().data_ptr():
            return False  # neighbours in the caching allocator, not one buffer
    return ts[-1].is_contiguous()


def _fused_view(ts):
    """[sum rows, cols] view over back-to-back tensors (first dims concatenated)."""
    first = ts[0]
    rows = sum(t.shape[0] for t in ts)
    shape = (rows,) + tuple(first.shape[1:])
    stride = first.stride()
    return first.as_strided(shape, stride)


def wc(param, cdt=None):
    """Operand copy of one parameter in the compute dtype (the parameter itself in fp32 mode)."""
    cdt = cdt or _compute_dtype
    p = param.detach()
    if cdt == torch.float32:
        return p if p.is_contiguous() else p.contiguous()
    sh = getattr(param, "_goat_shadow", None)
    if sh is not None and sh.dtype == cdt:
        return sh
    key = (param._version, _generation[0], p.data_ptr())
    ent = getattr(param, "_goat_cast", None)
    if ent is not None and ent[0] == key and ent[1].dtype == cdt and ent[1].device == p.device:
        return ent[1]
    t = ops.cast(p.contiguous(), cdt)
    param._goat_cast = (key, t)
    return t


def wc_cat(params, cdt=None):
    """Operand copy of several parameters concatenated along dim 0 (fused QKV weights / biases).
    Zero-copy when they already sit back to back (engine.FlatParams lays them out that way)."""
    cdt = cdt or _compute_dtype
    if cdt != torch.float32:
        shs = [getattr(p, "_goat_shadow", None) for p in params]
        if all(s is not None and s.dtype == cdt for s in shs) and _adjacent(shs):
            return _fused_view(shs)
    else:
        ds = [p.detach() for p in params]
        if _adjacent(ds):
            return _fused_view(ds)
    key = tuple(p._version for p in params) + (cdt, _generation[0], params[0].data_ptr())
    ent = getattr(params[0], "_goat_cat", None)
    if ent is not None and ent[0] == key and ent[2] == tuple(id(p) for p in params) and ent[1].device == params[0].device:
        return ent[1]
    cat = torch.cat([p.detach() for p in params], dim=0)
    t = cat if cdt == torch.float32 else ops.cast(cat, cdt)
    params[0]._goat_cat = (key, t, tuple(id(p) for p in params))
    return t
