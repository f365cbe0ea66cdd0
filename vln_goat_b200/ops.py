"""Thin tensor-level wrappers around the C ABI (no autograd here; see functional.py).

Everything takes CUDA tensors, enqueues on torch's current stream, and allocates outputs /
workspaces through torch's caching allocator (the library itself allocates nothing).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ACT_DGELU, ACT_DRELU, ACT_GELU, ACT_NONE, ACT_RELU, ACT_TANH, BF16, F16, F32  # noqa: F401

# kernels launched by this library since import (bench.py reports the per-step delta as gpu_launches)
LAUNCHES = [0]
# when set to a list, every gemm() call appends (M, N, K, a_mn, b_mn, dtype code, ran_on_simt, accumulate, act, has_bias,
# has_res, out_is_fp32, has_dropout, has_out2) -- bench.py's roofline pass re-times exactly these launches
GEMM_LOG = None
# same for the attention core: ("fwd" | "bwd", B, Nq, Nk, has_bias, drop_p) per call
ATTN_LOG = None

_DT = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}


def dt(t):
    try:
        return _DT[t if isinstance(t, torch.dtype) else t.dtype]
    except KeyError:
        raise TypeError("unsupported dtype %s" % (t if isinstance(t, torch.dtype) else t.dtype))


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """torch's current stream as a raw handle (torch.cuda.current_stream() costs ~15 us of Python per call)."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vln_goat_b200 ops need CUDA tensors (no CPU fallback)")


def _rowmajor2d(t, name):
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError("%s must be 2-D with unit inner stride, got shape %s strides %s" % (name, tuple(t.shape), t.stride()))
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def gemm(A, B, *, a_mn=False, b_mn=False, bias=None, res=None, aux_in=None, aux_out=None, out=None, out_dtype=None,
         out2=None, act=ACT_NONE, alpha=1.0, drop_p=0.0, drop_seed=0, seed_ptr=None, force_simt=False,
         accumulate=False):
    """out[M,N] = epilogue(alpha * A B^T).  A: [M,K] (a_mn False) or [K,M] (a_mn True);
    B: [N,K] (b_mn False) or [K,N] (b_mn True).  See goat_gemm in include/goat_sm100.h."""
    _req_cuda(A, B, bias, res, aux_in, aux_out, out, out2)
    lda = _rowmajor2d(A, "A")
    ldb = _rowmajor2d(B, "B")
    M, K = (A.shape[1], A.shape[0]) if a_mn else (A.shape[0], A.shape[1])
    N, Kb = (B.shape[1], B.shape[0]) if b_mn else (B.shape[0], B.shape[1])
    if K != Kb or A.dtype != B.dtype:
        raise ValueError("gemm: A %s / B %s mismatch" % (tuple(A.shape), tuple(B.shape)))
    if out is None:
        if accumulate:
            out = torch.zeros((M, N), device=A.device, dtype=torch.float32)
        else:
            out = torch.empty((M, N), device=A.device, dtype=out_dtype or A.dtype)
    if (not accumulate and A.dtype == torch.float32 and K >= 256 and M <= 64 and N <= 64
            and act == ACT_NONE and res is None and drop_p == 0.0 and out2 is None and aux_out is None and alpha == 1.0
            and out.dtype == torch.float32 and out.is_contiguous()):
        # ONE output tile and a long reduction on the fp32 SIMT kernel (the B x B InfoNCE similarities, 1-wide heads on the
        # [CLS] rows): start from the bias and let the split-K slices add their partial products.  Everything larger keeps
        # the single-pass kernel, so the fp32 parity mode stays bit-reproducible for the transformer blocks.
        if bias is None:
            out.zero_()
        else:
            out.copy_(bias.unsqueeze(0).expand(M, N))
        bias, accumulate = None, True
    a = _lib.GemmArgs()
    a.M, a.N, a.K = M, N, K
    a.dtype = dt(A)
    a.a_mn_major, a.b_mn_major = int(a_mn), int(b_mn)
    a.lda, a.ldb = lda, ldb
    a.A, a.B = A.data_ptr(), B.data_ptr()
    a.bias = None if bias is None else bias.data_ptr()
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N or not bias.is_contiguous()):
        raise ValueError("gemm: bias must be contiguous fp32 [N]")
    if res is not None:
        if res.dtype != torch.float32 or tuple(res.shape) != (M, N):
            raise ValueError("gemm: res must be fp32 [M,N]")
        a.res, a.ldres = res.data_ptr(), _rowmajor2d(res, "res")
    for name, t in (("aux_in", aux_in), ("aux_out", aux_out)):
        if t is not None:
            if t.dtype != A.dtype or tuple(t.shape) != (M, N):
                raise ValueError("gemm: %s must have the operand dtype and shape [M,N]" % name)
            setattr(a, name, t.data_ptr())
            a.ldaux = _rowmajor2d(t, name)
    if tuple(out.shape) != (M, N):
        raise ValueError("gemm: out shape %s != (%d,%d)" % (tuple(out.shape), M, N))
    a.out, a.ldc, a.out_dtype = out.data_ptr(), _rowmajor2d(out, "out"), dt(out)
    if out2 is not None:
        if out2.dtype != A.dtype or tuple(out2.shape) != (M, N):
            raise ValueError("gemm: out2 must have the operand dtype and shape [M,N]")
        a.out2, a.ldc2 = out2.data_ptr(), _rowmajor2d(out2, "out2")
    a.act, a.alpha, a.drop_p, a.drop_seed = act, alpha, drop_p, drop_seed
    a.drop_seed_ptr = None if seed_ptr is None else seed_ptr.data_ptr()
    a.force_simt = int(force_simt)
    a.accumulate = int(accumulate)
    umma_ok = a.dtype != F32 and K >= 16 and lda % 8 == 0 and ldb % 8 == 0 and not force_simt
    if umma_ok and not b_mn and not accumulate:
        # forward GEMM against a split weight ([hi | lo] buffer registered in runtime): add the lo term in the same launch
        from . import runtime
        lo = runtime.lo_pointer(B)
        if lo is not None:
            a.B_lo = lo
    _lib.check(_lib.lib().goat_gemm(C.byref(a), _stream()), "goat_gemm")
    LAUNCHES[0] += 1
    if GEMM_LOG is not None:
        GEMM_LOG.append((M, N, K, int(a_mn), int(b_mn), a.dtype, int(not umma_ok), int(accumulate), int(act), int(bias is not None),
                         int(res is not None), int(out.dtype == torch.float32), int(drop_p > 0.0), int(out2 is not None),
                         int(bool(a.B_lo))))
    return out


def _tok3(t, name, heads):
    """[B, N, >=heads*64] view with unit inner stride -> (ld, sb)."""
    if t.dim() != 3 or t.stride(2) != 1 or t.shape[2] != heads * 64:
        raise ValueError("%s must be [B,N,heads*64] with unit inner stride, got %s / %s" % (name, tuple(t.shape), t.stride()))
    return t.stride(1), t.stride(0)


def _attn_args(q, k, v, o, heads, kmask, bias, scale, lse, drop_p, drop_seed, seed_ptr=None, force_simt=False):
    B, Nq, _ = q.shape
    Nk = k.shape[1]
    a = _lib.AttnArgs()
    a.B, a.heads, a.Nq, a.Nk, a.D = B, heads, Nq, Nk, 64
    a.dtype = dt(q)
    a.Q, a.K, a.V, a.O = q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr()
    a.ldq, a.sbq = _tok3(q, "q", heads)
    a.ldk, a.sbk = _tok3(k, "k", heads)
    a.ldv, a.sbv = _tok3(v, "v", heads)
    a.ldo, a.sbo = _tok3(o, "o", heads)
    if kmask is not None:
        if kmask.dtype != torch.float32 or tuple(kmask.shape) != (B, Nk) or not kmask.is_contiguous():
            raise ValueError("attn: kmask must be contiguous fp32 [B,Nk]")
        a.kmask = kmask.data_ptr()
    if bias is not None:
        if bias.dtype != torch.float32 or tuple(bias.shape) != (B, Nq, Nk) or not bias.is_contiguous():
            raise ValueError("attn: bias must be contiguous fp32 [B,Nq,Nk]")
        a.bias = bias.data_ptr()
    a.scale = scale
    a.lse = lse.data_ptr()
    a.drop_p, a.drop_seed = drop_p, drop_seed
    a.drop_seed_ptr = None if seed_ptr is None else seed_ptr.data_ptr()
    a.force_simt = int(force_simt)
    return a


def attn_fwd(q, k, v, heads, kmask=None, bias=None, scale=0.125, drop_p=0.0, drop_seed=0, seed_ptr=None,
             force_simt=False):
    """-> (O [B,Nq,heads*64] same dtype, lse [B,heads,Nq] fp32)"""
    _req_cuda(q, k, v, kmask, bias)
    B, Nq, _ = q.shape
    o = torch.empty((B, Nq, heads * 64), device=q.device, dtype=q.dtype)
    lse = torch.empty((B, heads, Nq), device=q.device, dtype=torch.float32)
    a = _attn_args(q, k, v, o, heads, kmask, bias, scale, lse, drop_p, drop_seed, seed_ptr, force_simt)
    _lib.check(_lib.lib().goat_attn_core_fwd(C.byref(a), _stream()), "goat_attn_core_fwd")
    LAUNCHES[0] += 1
    if ATTN_LOG is not None:
        ATTN_LOG.append(("fwd", B, Nq, k.shape[1], int(bias is not None), float(drop_p)))
    return o, lse


def attn_bwd(do, q, k, v, o, lse, heads, dq, dk, dv, kmask=None, bias=None, scale=0.125, drop_p=0.0, drop_seed=0,
             seed_ptr=None, want_dbias=False, force_simt=False):
    """dq/dk/dv: preallocated views with the same strides as q/k/v.  -> dbias [B,Nq,Nk] fp32 or None"""
    _req_cuda(do, q, k, v, o, lse, dq, dk, dv)
    a = _attn_args(q, k, v, o, heads, kmask, bias, scale, lse, drop_p, drop_seed, seed_ptr, force_simt)
    for name, t, ref in (("dq", dq, q), ("dk", dk, k), ("dv", dv, v), ("do", do, o)):
        if t.stride() != ref.stride() or t.shape != ref.shape or t.dtype != ref.dtype:
            raise ValueError("attn_bwd: %s must match the layout of its forward tensor" % name)
    a.dO, a.dQ, a.dK, a.dV = do.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    dbias = None
    if want_dbias:
        dbias = torch.zeros((q.shape[0], q.shape[1], k.shape[1]), device=q.device, dtype=torch.float32)
        a.dbias = dbias.data_ptr()
    _lib.check(_lib.lib().goat_attn_core_bwd(C.byref(a), _stream()), "goat_attn_core_bwd")
    if ATTN_LOG is not None:
        ATTN_LOG.append(("bwd", q.shape[0], q.shape[1], k.shape[1], int(bias is not None), float(drop_p)))
    # one kernel on the tcgen05 paths (16-bit, Nq <= 128), two (dQ, then dK/dV) on the fp32-math SIMT path
    LAUNCHES[0] += 1 if (q.dtype != torch.float32 and q.shape[1] <= 128 and not force_simt) else 2
    return dbias


def layernorm_fwd(x, gamma, beta, eps, want32=True, dtype16=None):
    """x [M,H] contiguous -> (y32 or None, y16 or None, mean [M], rstd [M])"""
    _req_cuda(x, gamma, beta)
    if x.dim() != 2 or not x.is_contiguous():
        raise ValueError("layernorm_fwd: x must be contiguous [M,H]")
    M, H = x.shape
    y32 = torch.empty((M, H), device=x.device, dtype=torch.float32) if want32 else None
    y16 = torch.empty((M, H), device=x.device, dtype=dtype16) if dtype16 is not None else None
    mean = torch.empty((M,), device=x.device, dtype=torch.float32)
    rstd = torch.empty((M,), device=x.device, dtype=torch.float32)
    rc = _lib.lib().goat_layernorm_fwd(_p(x), dt(x), _p(gamma), _p(beta), eps, _p(y32), _p(y16),
                                       dt(dtype16) if dtype16 is not None else F16, _p(mean), _p(rstd), M, H, _stream())
    _lib.check(rc, "goat_layernorm_fwd")
    LAUNCHES[0] += 1
    return y32, y16, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dres=None, want32=True, dtype16=None, drop_p=0.0, drop_seed=0,
                  seed_ptr=None, want_colsum=False, dgamma_out=None, dbeta_out=None, dcol_out=None, accumulate=False):
    """-> (dx32 or None, dx16 or None, dgamma [H], dbeta [H], dcolsum [H] or None)
    *_out: optional contiguous fp32 [H] destinations (e.g. flat gradient views) written instead of new tensors.
    accumulate=True (H == 768): dgamma / dbeta / dcolsum are ADDED to the destinations by the one kernel (fp32 atomics,
    no finalize launch); destinations that are not given are fresh zeros."""
    _req_cuda(dy, x, gamma, mean, rstd, dres)
    M, H = x.shape
    if dy.dtype != torch.float32 or not dy.is_contiguous() or tuple(dy.shape) != (M, H):
        raise ValueError("layernorm_bwd: dy must be contiguous fp32 [M,H]")
    dev = x.device
    dx32 = torch.empty((M, H), device=dev, dtype=torch.float32) if want32 else None
    dx16 = torch.empty((M, H), device=dev, dtype=dtype16) if dtype16 is not None else None
    for o in (dgamma_out, dbeta_out, dcol_out):
        if o is not None and (o.dtype != torch.float32 or o.numel() != H or not o.is_contiguous()):
            raise ValueError("layernorm_bwd: *_out must be contiguous fp32 [H]")
    accumulate = accumulate and H == 768
    new = torch.zeros if accumulate else torch.empty
    dgamma = dgamma_out if dgamma_out is not None else new((H,), device=dev, dtype=torch.float32)
    dbeta = dbeta_out if dbeta_out is not None else new((H,), device=dev, dtype=torch.float32)
    dcol = (dcol_out if dcol_out is not None else new((H,), device=dev, dtype=torch.float32)) \
        if want_colsum else None
    if accumulate:
        rc = _lib.lib().goat_layernorm_bwd_acc(_p(dy), _p(x), dt(x), _p(gamma), _p(mean), _p(rstd), _p(dres), _p(dx32),
                                               _p(dx16), dt(dtype16) if dtype16 is not None else F16, drop_p, drop_seed,
                                               _p(seed_ptr), _p(dgamma), _p(dbeta), _p(dcol), M, H, _stream())
        _lib.check(rc, "goat_layernorm_bwd_acc")
        LAUNCHES[0] += 1
        return dx32, dx16, dgamma, dbeta, dcol
    ws = torch.empty((_lib.lib().goat_layernorm_bwd_workspace_bytes(M, H),), device=dev, dtype=torch.uint8)
    rc = _lib.lib().goat_layernorm_bwd(_p(dy), _p(x), dt(x), _p(gamma), _p(mean), _p(rstd), _p(dres), _p(dx32), _p(dx16),
                                       dt(dtype16) if dtype16 is not None else F16, drop_p, drop_seed, _p(seed_ptr),
                                       _p(dgamma), _p(dbeta), _p(dcol), _p(ws), M, H, _stream())
    _lib.check(rc, "goat_layernorm_bwd")
    LAUNCHES[0] += 2
    return dx32, dx16, dgamma, dbeta, dcol


def colsum(x, out=None, accumulate=False):
    """x [M,N] (unit inner stride) -> fp32 [N].  accumulate=True: out += column sums in one kernel (fp32 atomics; out
    zero-initialised or holding an earlier partial sum); a fresh zeroed tensor when out is None."""
    _req_cuda(x, out)
    ld = _rowmajor2d(x, "x")
    M, N = x.shape
    if out is None:
        out = (torch.zeros if accumulate else torch.empty)((N,), device=x.device, dtype=torch.float32)
    elif out.dtype != torch.float32 or out.numel() != N or not out.is_contiguous():
        raise ValueError("colsum: out must be contiguous fp32 [N]")
    if accumulate:
        vec = 16 // x.element_size()
        if N % vec == 0 and ld % vec == 0 and x.data_ptr() % 16 == 0:
            _lib.check(_lib.lib().goat_colsum_acc(_p(x), dt(x), M, N, ld, _p(out), _stream()), "goat_colsum_acc")
            LAUNCHES[0] += 1
            return out
        out.add_(colsum(x))     # odd widths (1-wide heads, 7/14-wide position features): two-kernel path
        return out
    ws = torch.empty((_lib.lib().goat_colsum_workspace_bytes(M, N),), device=x.device, dtype=torch.uint8)
    _lib.check(_lib.lib().goat_colsum(_p(x), dt(x), M, N, ld, _p(out), _p(ws), _stream()), "goat_colsum")
    LAUNCHES[0] += 2
    return out


def split_cast(x32, hi, lo):
    """hi = x32 rounded to 16 bit, lo = (x32 - hi) rounded: the two terms of a split weight operand (goat_gemm B / B_lo)"""
    if x32.dtype != torch.float32 or not x32.is_contiguous() or hi.dtype != lo.dtype or hi.numel() != x32.numel() \
            or lo.numel() != x32.numel():
        raise ValueError("split_cast: need a contiguous fp32 source and two 16-bit outputs of its size")
    _lib.check(_lib.lib().goat_split_cast(_p(x32), _p(hi), _p(lo), dt(hi), x32.numel(), _stream()), "goat_split_cast")
    LAUNCHES[0] += 1
    return hi, lo


def cast(src, dtype, out=None, drop_p=0.0, drop_seed=0, seed_ptr=None):
    """contiguous tensor -> new tensor (or `out`) of `dtype`; optional dropout mask by linear element index"""
    _req_cuda(src, out)
    if not src.is_contiguous():
        raise ValueError("cast: src must be contiguous")
    if out is None:
        out = torch.empty(src.shape, device=src.device, dtype=dtype)
    elif not out.is_contiguous() or out.numel() != src.numel():
        raise ValueError("cast: bad out")
    _lib.check(_lib.lib().goat_dropout_cast(_p(src), dt(src), _p(out), dt(out), src.numel(), drop_p, drop_seed,
                                            _p(seed_ptr), _stream()), "goat_dropout_cast")
    LAUNCHES[0] += 1
    return out


# --------------------------------------------------------------------------------------
# heads.cu: pooling, dictionary sums, door gate, cross-entropy, gather-reduce, embeddings (all fp32)
# --------------------------------------------------------------------------------------
def _f32c(t, name, shape=None):
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError("%s must be a contiguous fp32 tensor" % name)
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("%s must have shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
    return t


def _nvalid(t):
    if t is not None and (t.dtype != torch.int32 or t.numel() != 1 or not t.is_cuda):
        raise ValueError("n_valid must be a CUDA int32 tensor with one element")
    return t


def attn_pool_fwd(x, w, bias, mode, n_valid=None):
    """x [B,N,H], w [H] (+ bias [1] in mode 0) -> (out [B,H], a [B,N], s [B,N] or None).  See goat_attn_pool_fwd."""
    _req_cuda(x, w, bias)
    _nvalid(n_valid)
    B, N, H = x.shape
    _f32c(x, "x"); _f32c(w, "w"); _f32c(bias, "bias")
    if w.numel() != H:
        raise ValueError("attn_pool: w must have H elements")
    out = torch.empty((B, H), device=x.device, dtype=torch.float32)
    a = torch.empty((B, N), device=x.device, dtype=torch.float32)
    s = torch.empty((B, N), device=x.device, dtype=torch.float32) if mode == 0 else None
    _lib.check(_lib.lib().goat_attn_pool_fwd(_p(x), _p(w), _p(bias), mode, B, N, H, _p(out), _p(a), _p(s), _p(n_valid),
                                             _stream()), "goat_attn_pool_fwd")
    LAUNCHES[0] += 1
    return out, a, s


def attn_pool_bwd(dout, x, w, a, s, out, mode, dw, db, n_valid=None):
    """-> dx [B,N,H]; dw [H] / db [1] are accumulated into (pass zeroed or flat-gradient tensors)."""
    _req_cuda(dout, x, w, a, dw, db)
    B, N, H = x.shape
    _f32c(dout, "dout", (B, H)); _f32c(dw, "dw"); _f32c(db, "db")
    dx = torch.empty_like(x)
    _lib.check(_lib.lib().goat_attn_pool_bwd(_p(dout), _p(x), _p(w), _p(a), _p(s), _p(out), mode, B, N, H, _p(dx), _p(dw),
                                             _p(db), _p(_nvalid(n_valid)), _stream()), "goat_attn_pool_bwd")
    LAUNCHES[0] += 1
    return dx


def wsum_fwd(x, p):
    """x [B,N,H] fp32, p [B,N] fp32 -> out [B,H]"""
    _req_cuda(x, p)
    B, N, H = x.shape
    _f32c(x, "x"); _f32c(p, "p", (B, N))
    out = torch.empty((B, H), device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().goat_wsum_fwd(_p(x), _p(p), B, N, H, _p(out), _stream()), "goat_wsum_fwd")
    LAUNCHES[0] += 1
    return out


def wsum_bwd(dout, p, N):
    _req_cuda(dout, p)
    B, H = dout.shape
    _f32c(dout, "dout"); _f32c(p, "p", (B, N))
    dx = torch.empty((B, N, H), device=dout.device, dtype=torch.float32)
    _lib.check(_lib.lib().goat_wsum_bwd(_p(dout), _p(p), B, N, H, _p(dx), _stream()), "goat_wsum_bwd")
    LAUNCHES[0] += 1
    return dx


def door_gate_fwd(aug, ori, wa, ba, wo, bo):
    """aug, ori [M,H]; wa, wo [H]; ba, bo [1] -> (out [M,H], gate [M])"""
    _req_cuda(aug, ori, wa, ba, wo, bo)
    M, H = aug.shape
    _f32c(aug, "aug"); _f32c(ori, "ori", (M, H)); _f32c(wa, "wa"); _f32c(wo, "wo"); _f32c(ba, "ba"); _f32c(bo, "bo")
    out = torch.empty_like(aug)
    gate = torch.empty((M,), device=aug.device, dtype=torch.float32)
    _lib.check(_lib.lib().goat_door_gate_fwd(_p(aug), _p(ori), _p(wa), _p(ba), _p(wo), _p(bo), M, H, _p(out), _p(gate),
                                             _stream()), "goat_door_gate_fwd")
    LAUNCHES[0] += 1
    return out, gate


def door_gate_bwd(dout, aug, ori, wa, wo, gate, dwa, dwo, dba, dbo):
    """-> (daug, dori); dwa/dwo [H], dba/dbo [1] accumulated into."""
    _req_cuda(dout, aug, ori, wa, wo, gate, dwa, dwo, dba, dbo)
    M, H = aug.shape
    _f32c(dout, "dout", (M, H))
    for n, t in (("dwa", dwa), ("dwo", dwo), ("dba", dba), ("dbo", dbo)):
        _f32c(t, n)
    daug = torch.empty_like(aug)
    dori = torch.empty_like(ori)
    _lib.check(_lib.lib().goat_door_gate_bwd(_p(dout), _p(aug), _p(ori), _p(wa), _p(wo), _p(gate), M, H, _p(daug), _p(dori),
                                             _p(dwa), _p(dwo), _p(dba), _p(dbo), _stream()), "goat_door_gate_bwd")
    LAUNCHES[0] += 1
    return daug, dori


def xent_fwd(logits, labels, ignore_index=-100):
    """logits fp32 [M,N] (any strides), labels int64 [M] -> (loss [M], lse [M])"""
    _req_cuda(logits, labels)
    if logits.dim() != 2 or logits.dtype != torch.float32:
        raise ValueError("xent: logits must be 2-D fp32")
    if labels.dtype != torch.int64 or labels.numel() != logits.shape[0] or not labels.is_contiguous():
        raise ValueError("xent: labels must be contiguous int64 [M]")
    M, N = logits.shape
    loss = torch.empty((M,), device=logits.device, dtype=torch.float32)
    lse = torch.empty((M,), device=logits.device, dtype=torch.float32)
    _lib.check(_lib.lib().goat_xent_fwd(_p(logits), logits.stride(0), logits.stride(1), _p(labels), M, N, ignore_index,
                                        _p(loss), _p(lse), _stream()), "goat_xent_fwd")
    LAUNCHES[0] += 1
    return loss, lse


def xent_bwd(dloss, logits, labels, lse, ignore_index=-100, out=None, accumulate=False):
    """-> dlogits, shaped and strided like ``out`` (default: new contiguous [M,N])"""
    _req_cuda(dloss, logits, labels, lse, out)
    M, N = logits.shape
    _f32c(dloss, "dloss", (M,))
    if out is None:
        out = torch.empty((M, N), device=logits.device, dtype=torch.float32)
        accumulate = False
    elif tuple(out.shape) != (M, N) or out.dtype != torch.float32:
        raise ValueError("xent_bwd: bad out")
    _lib.check(_lib.lib().goat_xent_bwd(_p(dloss), _p(logits), logits.stride(0), logits.stride(1), _p(labels), _p(lse), M, N,
                                        ignore_index, _p(out), out.stride(0), out.stride(1), int(accumulate), _stream()),
               "goat_xent_bwd")
    LAUNCHES[0] += 1
    return out


def xent_chunk_fwd(logits, labels, c0, first, run_max, run_sum, picked):
    """one column chunk [c0, c0+Nc) of the logits (fp32 [M,Nc], unit inner stride): update the running row statistics"""
    _req_cuda(logits, labels, run_max, run_sum, picked)
    M, Nc = logits.shape
    if logits.dtype != torch.float32 or logits.stride(1) != 1:
        raise ValueError("xent_chunk_fwd: logits must be fp32 with unit inner stride")
    _lib.check(_lib.lib().goat_xent_chunk_fwd(_p(logits), logits.stride(0), _p(labels), M, Nc, int(c0), int(first),
                                              _p(run_max), _p(run_sum), _p(picked), _stream()), "goat_xent_chunk_fwd")
    LAUNCHES[0] += 1


def xent_chunk_bwd(dloss, logits, labels, lse, c0, ignore_index, n_classes, out):
    """out [M,Nc] (any dtype, unit inner stride) = dloss (softmax - onehot) for the chunk starting at class c0"""
    _req_cuda(dloss, logits, labels, lse, out)
    M, Nc = logits.shape
    _lib.check(_lib.lib().goat_xent_chunk_bwd(_p(dloss), _p(logits), logits.stride(0), _p(labels), _p(lse), M, Nc, int(c0),
                                              int(ignore_index), int(n_classes), _p(out), dt(out), out.stride(0), _stream()),
               "goat_xent_chunk_bwd")
    LAUNCHES[0] += 1


def segment_reduce_fwd(src, idx, mean):
    """src [R0,H] fp32, idx int32 [R,K] (-1 = empty) -> out [R,H]"""
    _req_cuda(src, idx)
    _f32c(src, "src")
    if idx.dtype != torch.int32 or idx.dim() != 2 or not idx.is_contiguous():
        raise ValueError("segment_reduce: idx must be contiguous int32 [R,K]")
    R, K = idx.shape
    H = src.shape[1]
    out = torch.empty((R, H), device=src.device, dtype=torch.float32)
    _lib.check(_lib.lib().goat_segment_reduce_fwd(_p(src), _p(idx), R, K, H, int(mean), _p(out), _stream()),
               "goat_segment_reduce_fwd")
    LAUNCHES[0] += 1
    return out


def segment_reduce_bwd(dout, idx, mean, n_src, out=None):
    """-> dsrc [n_src,H]; with ``out`` (contiguous fp32 [n_src,H], e.g. a flat-gradient view) the result is ADDED to it"""
    _req_cuda(dout, idx, out)
    R, K = idx.shape
    H = dout.shape[1]
    _f32c(dout, "dout", (R, H))
    if out is None:
        dsrc = torch.zeros((n_src, H), device=dout.device, dtype=torch.float32)
    else:
        dsrc = _f32c(out, "out", (n_src, H))
    _lib.check(_lib.lib().goat_segment_reduce_bwd(_p(dout), _p(idx), R, K, H, int(mean), _p(dsrc), _stream()),
               "goat_segment_reduce_bwd")
    LAUNCHES[0] += 1
    return dsrc


def act_grad(dy, ref, act, out_dtype):
    """dy fp32 (contiguous) * act'(ref) -> new tensor of out_dtype.  ref: forward output (RELU / TANH, fp32) or the
    stored pre-activation (GELU, fp32 / 16-bit); None for ACT_NONE (a plain cast)."""
    _req_cuda(dy, ref)
    if dy.dtype != torch.float32 or not dy.is_contiguous():
        raise ValueError("act_grad: dy must be contiguous fp32")
    if ref is not None and (not ref.is_contiguous() or ref.numel() != dy.numel()):
        raise ValueError("act_grad: ref must be contiguous with dy's element count")
    out = torch.empty(dy.shape, device=dy.device, dtype=out_dtype)
    _lib.check(_lib.lib().goat_act_grad(_p(dy), _p(ref), dt(ref) if ref is not None else F32, int(act), _p(out), dt(out),
                                        dy.numel(), _stream()), "goat_act_grad")
    LAUNCHES[0] += 1
    return out


def gather_rows(src, idx):
    """src [R0, ...] (contiguous, any dtype, row size a multiple of 16 bytes), idx int32 [R] (-1 = zero row) -> [R, ...]"""
    _req_cuda(src, idx)
    if not src.is_contiguous() or idx.dtype != torch.int32 or idx.dim() != 1 or not idx.is_contiguous():
        raise ValueError("gather_rows: src must be contiguous and idx a contiguous int32 vector")
    row_bytes = (src.numel() // src.shape[0]) * src.element_size()
    out = torch.empty((idx.shape[0],) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
    _lib.check(_lib.lib().goat_gather_rows(_p(src), _p(idx), idx.shape[0], row_bytes, _p(out), _stream()), "goat_gather_rows")
    LAUNCHES[0] += 1
    return out


def act_fwd(x, act):
    """fp32 contiguous -> act(x) fp32 (exact erf GELU / ReLU / tanh)"""
    _req_cuda(x)
    _f32c(x, "x")
    out = torch.empty_like(x)
    _lib.check(_lib.lib().goat_act_fwd(_p(x), int(act), _p(out), x.numel(), _stream()), "goat_act_fwd")
    LAUNCHES[0] += 1
    return out


def sprel_fwd(d, w, b):
    """d fp32 (any shape, contiguous), w / b fp32 one element each -> d * w + b"""
    _req_cuda(d, w, b)
    _f32c(d, "d"); _f32c(w, "w"); _f32c(b, "b")
    out = torch.empty_like(d)
    _lib.check(_lib.lib().goat_sprel_fwd(_p(d), _p(w), _p(b), _p(out), d.numel(), _stream()), "goat_sprel_fwd")
    LAUNCHES[0] += 1
    return out


def sprel_bwd(dout, d, dw, db):
    """dw[0] += sum(dout * d), db[0] += sum(dout)"""
    _req_cuda(dout, d, dw, db)
    _f32c(dout, "dout"); _f32c(d, "d"); _f32c(dw, "dw"); _f32c(db, "db")
    _lib.check(_lib.lib().goat_sprel_bwd(_p(dout), _p(d), _p(dw), _p(db), d.numel(), _stream()), "goat_sprel_bwd")
    LAUNCHES[0] += 1


def embed_fwd(ids, word, pos, type_):
    """ids int64 [B,L] -> [B*L,H] fp32 = word[ids] + pos[arange(L)] + type[0]"""
    _req_cuda(ids, word, pos, type_)
    if ids.dtype != torch.int64 or ids.dim() != 2 or not ids.is_contiguous():
        raise ValueError("embed: ids must be contiguous int64 [B,L]")
    B, L = ids.shape
    H = word.shape[1]
    _f32c(word, "word"); _f32c(pos, "pos"); _f32c(type_, "type")
    if L > pos.shape[0]:
        raise ValueError("embed: sequence length %d exceeds the %d learned positions" % (L, pos.shape[0]))
    out = torch.empty((B * L, H), device=ids.device, dtype=torch.float32)
    _lib.check(_lib.lib().goat_embed_fwd(_p(ids), _p(word), _p(pos), _p(type_), B * L, L, H, _p(out), _stream()),
               "goat_embed_fwd")
    LAUNCHES[0] += 1
    return out


def embed_bwd(dout, ids, dword, dpos, dtype_, padding_idx=-1):
    """accumulates into the given gradient tables (any may be None); row ``padding_idx`` of word / pos gets nothing"""
    _req_cuda(dout, ids, dword, dpos, dtype_)
    B, L = ids.shape
    H = dout.shape[1]
    _f32c(dout, "dout", (B * L, H)); _f32c(dword, "dword"); _f32c(dpos, "dpos"); _f32c(dtype_, "dtype")
    _lib.check(_lib.lib().goat_embed_bwd(_p(dout), _p(ids), B * L, L, H, int(padding_idx), _p(dword), _p(dpos), _p(dtype_),
                                         _stream()), "goat_embed_bwd")
    LAUNCHES[0] += 1
