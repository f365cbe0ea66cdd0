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


# ---------------------------------------------------------------------------------------------
# Split weights.  A 16-bit copy of a weight carries 11 (fp16) / 8 (bf16) mantissa bits, and that rounding error is
# SYSTEMATIC: the same for every token, so attention / pooling do not average it out -- it alone costs the CFP embeddings
# ~9e-4 of the 1e-3 budget (profiles/r02_fp16_error_budget.txt).  With split weights every 16-bit weight buffer is
# allocated as [hi | lo] (lo = round(W - hi)); goat_gemm runs the K loop twice over the activations (hi, then lo) for
# FORWARD GEMMs (K-major weight operand), so the weight enters with ~22 bits.  Backward GEMMs read hi only.
# ops.gemm finds the lo half of any view into a registered buffer by pointer arithmetic: no plumbing through the
# autograd functions.
# ---------------------------------------------------------------------------------------------
_weight_split = True
_split_buffers = {}     # base data_ptr of a [hi | lo] buffer -> (bytes of the hi half, the tensor kept alive)


def weight_split():
    return _weight_split


def set_weight_split(on):
    """Forward GEMMs use hi + lo 16-bit weight operands (default) or the plain 16-bit copy (~8 % faster step, forward
    error ~1.6x larger).  Set before FlatParams / the first forward."""
    global _weight_split
    _weight_split = bool(on)
    bump_generation()


def register_split_buffer(buf, hi_numel):
    """buf: 1-D 16-bit tensor of 2 * hi_numel elements, [hi | lo]"""
    _split_buffers[buf.data_ptr()] = (hi_numel * buf.element_size(), buf)
    if len(_split_buffers) > 4096:                 # per-parameter cast caches of long-dead models
        for k in list(_split_buffers)[:2048]:
            del _split_buffers[k]


def lo_pointer(t):
    """device address of the lo twin of a view into a registered [hi | lo] buffer (None if t is not such a view)"""
    if not _weight_split or not _split_buffers:
        return None
    base = t.untyped_storage().data_ptr()
    ent = _split_buffers.get(base)
    if ent is None:
        return None
    off = t.data_ptr() - base
    if off < 0 or off + (t.numel() and 1) > ent[0]:
        return None
    return t.data_ptr() + ent[0]


def _cast_split(p32, cdt):
    """[hi | lo] 16-bit copy of a contiguous fp32 tensor -> the hi view (shape of p32); lo is found via lo_pointer"""
    n = p32.numel()
    npad = (n + 7) // 8 * 8
    buf = torch.empty(2 * npad, device=p32.device, dtype=cdt)
    hi = buf[:n].view(p32.shape)
    ops.split_cast(p32.reshape(-1), buf[:n], buf[npad:npad + n])
    register_split_buffer(buf, npad)
    return hi


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
    t = _cast_split(p.contiguous(), cdt) if _weight_split else ops.cast(p.contiguous(), cdt)
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
    t = cat if cdt == torch.float32 else (_cast_split(cat, cdt) if (_weight_split and cat.dim() == 2) else ops.cast(cat, cdt))
    params[0]._goat_cat = (key, t, tuple(id(p) for p in params))
    return t
