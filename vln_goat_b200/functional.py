"""Autograd composites over the C-ABI primitives: one Function per transformer block.

Every Function works on 2-D token-major activations [tokens, hidden].  The residual stream stays
fp32 (``x32``); GEMM operands are the compute dtype (fp16 / bf16 copies written by the LayerNorm
kernel, or the fp32 tensor itself in fp32 parity mode).  Backward passes are hand-written from the
same primitives (dgrad / wgrad GEMMs with MN-major operands, recompute-based attention backward).

Reference semantics (P/ = pretrain_src/, M/ = map_nav_src/ of CrystalSixone/VLN-GOAT):
  AttnBlockFn  = BertAttention / RobertaAttention        P/model/Bert_backbone.py:199-342, :387-543
  FFNBlockFn   = BertIntermediate + BertOutput           P/model/Bert_backbone.py:345-370
  PanoLayerFn  = TransformerEncoderLayer.forward_pre     P/model/transformer.py:170-182
  LinearFn     = nn.Linear (+ ReLU / tanh / GELU)        heads, poolers, position embeddings
  LayerNormFn  = nn.LayerNorm
"""
import itertools
from collections import namedtuple

import torch

from . import ops

_seed_counter = itertools.count(1)


_seed_ptr = None


def seed_ptr():
    """Optional device word (uint64 as int64 tensor [1]) every dropout kernel adds to its seed: bump it between
    CUDA-graph replays so a captured step draws fresh masks (engine.TrainStep does)."""
    return _seed_ptr


def set_seed_ptr(t):
    global _seed_ptr
    if t is not None and (t.dtype != torch.int64 or t.numel() != 1 or not t.is_cuda):
        raise ValueError("seed_ptr must be a CUDA int64 tensor with one element")
    _seed_ptr = t


def next_seed():
    """A fresh host-side dropout seed (kernels add the optional device-side step counter to it)."""
    return (next(_seed_counter) * 0x9E3779B1) & 0x7FFFFFFFFFFFFFFF


def _c(x32, x16, cdt):
    """compute-dtype view of an activation: fp32 mode uses the fp32 tensor itself"""
    if cdt == torch.float32:
        return x32
    if x16 is not None:
        return x16
    return ops.cast(x32, cdt)


def _d16(cdt):
    return None if cdt == torch.float32 else cdt


def _split_rows(t, params):
    out, r = [], 0
    for p in params:
        out.append(t[r:r + p.shape[0]])
        r += p.shape[0]
    return tuple(out)


def _mark(params):
    for p in params:
        if p is not None:
            p._goat_fresh = False      # diagnostic only (engine.FlatParams.unwritten)


def _grad_into(params, fn, can_acc):
    """Produce the gradient of one parameter, or of several whose rows are concatenated (fused QKV).

    ``fn(out)`` computes it; with ``out`` given it ADDS the result to ``out`` (a zero-initialised tensor is allocated
    when None).  ONE policy for every parameter that lives in an engine.FlatParams buffer: gradients are ACCUMULATED
    into the flat gradient views, which are zero at step start (the optimizer kernel clears them; FlatParams.zero_grad()
    does it explicitly) -- several backward passes between two optimizer steps therefore sum, as torch's ``.grad``
    does under gradient accumulation (P/train_r2r_goat.py:322-327).  None is returned for autograd in that case;
    otherwise the tensors are returned for autograd to accumulate the usual way."""
    dsts = [getattr(p, "_goat_grad", None) for p in params]
    if all(d is not None for d in dsts):
        from . import runtime
        if can_acc and (len(dsts) == 1 or runtime._adjacent(dsts)) and dsts[0].is_contiguous():
            fn(dsts[0] if len(dsts) == 1 else runtime._fused_view(dsts))
        else:
            t = fn(None)
            for d, part in zip(dsts, _split_rows(t, params)):
                d.add_(part)
        _mark(params)
        return (None,) * len(params)
    t = fn(None)
    return _split_rows(t, params) if len(params) > 1 else (t,)


def _wgrad(params, dy_c, x_c):
    """dW = dY^T X for weight(s) [N,K] (rows of several weights concatenated)."""
    # split-K + atomic accumulate: into the flat gradient view, or into fresh zeros
    return _grad_into(params, lambda out: ops.gemm(dy_c, x_c, a_mn=True, b_mn=True, out=out, accumulate=True), True)


def _bgrad(params, dy_c):
    """db = column sums of dY."""
    if params[0] is None:
        return (None,) * len(params)
    # one atomic-accumulate kernel straight into the flat gradient view, or into fresh zeros
    return _grad_into(params, lambda out: ops.colsum(dy_c, out=out, accumulate=True), True)


def _vgrad(param, t):
    """a gradient vector that a kernel already produced (LayerNorm dgamma/dbeta, LN-fused bias column sums)"""
    if param is None:
        return None
    return _grad_into((param,), lambda out: t, False)[0]


def _adst(param):
    """flat-gradient destination a kernel may ACCUMULATE into, else None"""
    if param is None:
        return None
    d = getattr(param, "_goat_grad", None)
    if d is None or not d.is_contiguous():
        return None
    return d


def _ln_bwd(dy32, x, gamma, mean, rstd, dres, cdt16, p, seed, seed_ptr, gamma_p, beta_p, bias_p=None):
    """LayerNorm backward with dgamma / dbeta (/ the producing Linear's bias grad) landing in the flat gradient
    buffer when there is one.  -> (dx32, dx16-or-None, dgamma, dbeta, dbias) with None for in-place grads."""
    ag, ab, ac = _adst(gamma_p), _adst(beta_p), _adst(bias_p)
    if x.shape[1] == 768 and ag is not None and ab is not None and (bias_p is None or ac is not None):
        # one kernel: column sums added into the flat gradient views with atomics (no finalize launch)
        dx32, dx16, _, _, _ = ops.layernorm_bwd(dy32, x, gamma, mean, rstd, dres, True, cdt16, p, seed, seed_ptr,
                                                want_colsum=bias_p is not None, dgamma_out=ag, dbeta_out=ab,
                                                dcol_out=ac, accumulate=True)
        _mark((gamma_p, beta_p, bias_p))
        return dx32, dx16, None, None, None
    dx32, dx16, dg, db, dcol = ops.layernorm_bwd(dy32, x, gamma, mean, rstd, dres, True, cdt16, p, seed, seed_ptr,
                                                 want_colsum=bias_p is not None)
    return dx32, dx16, _vgrad(gamma_p, dg), _vgrad(beta_p, db), None if bias_p is None else _vgrad(bias_p, dcol)


def _ln_bwd_split(dy32, pre, gamma, mean, rstd, cdt, p, seed, seed_ptr, gamma_p, beta_p, bias_p):
    """LN backward of a post-LN block: (residual-path grad fp32, GEMM-operand grad in the compute dtype with the
    hidden-dropout mask applied, dgamma, dbeta, bias grad of the producing Linear)."""
    if cdt == torch.float32 and p == 0.0:
        d32, _, dg, db, dcol = _ln_bwd(dy32, pre, gamma, mean, rstd, None, None, 0.0, 0, None, gamma_p, beta_p, bias_p)
        return d32, d32, dg, db, dcol
    return _ln_bwd(dy32, pre, gamma, mean, rstd, None, cdt, p, seed, seed_ptr, gamma_p, beta_p, bias_p)


AttnCfg = namedtuple("AttnCfg", "B Nq Nk heads eps attn_p hid_p seed seed_ptr cross cdt")


class AttnBlockFn(torch.autograd.Function):
    """y = LN(dropout(ctx W_o^T + b_o) + x),  ctx = softmax(q k^T / 8 + kmask + bias) v
    self-attention: q,k,v from x (one fused [3H,H] projection); cross: q from x, k,v from kv."""

    @staticmethod
    def forward(ctx, x32, x16, kv32, kv16, kmask, bias, Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma, beta, Wqkv_c, bqkv,
                Wo_c, cfg, kvp32=None, kvp16=None):
        """kvp32 / kvp16: optional K|V projection of the keys/values already computed by KVProjFn (rollout-level
        cache, modules.KVCache): the projection GEMM is skipped here and its gradient is handed back as the (fp32)
        gradient of kvp32, so autograd sums it over every step that used the cached projection."""
        ctx.set_materialize_grads(False)   # no zero-filled gradient for the non-differentiable 16-bit copy
        H = x32.shape[1]
        cdt = cfg.cdt
        xc = _c(x32, x16, cdt)
        ctx.cached_kv = cfg.cross and kvp32 is not None
        if ctx.cached_kv:
            kvc = None
            q2 = ops.gemm(xc, Wqkv_c[:H], bias=bqkv[:H], out_dtype=cdt)            # [M,H]
            kvp = kvp32 if cdt == torch.float32 else kvp16                          # [Mk,2H]
            k2, v2 = kvp[:, :H], kvp[:, H:]
            qkv = q2
        elif not cfg.cross:
            qkv = ops.gemm(xc, Wqkv_c, bias=bqkv, out_dtype=cdt)                   # [M,3H]
            q2, k2, v2 = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
            kvc = None
            kvp = None
        else:
            kvc = _c(kv32, kv16, cdt)
            q2 = ops.gemm(xc, Wqkv_c[:H], bias=bqkv[:H], out_dtype=cdt)            # [M,H]
            kvp = ops.gemm(kvc, Wqkv_c[H:], bias=bqkv[H:], out_dtype=cdt)          # [Mk,2H]
            k2, v2 = kvp[:, :H], kvp[:, H:]
            qkv = q2
        B, Nq, Nk = cfg.B, cfg.Nq, cfg.Nk
        q3 = q2.unflatten(0, (B, Nq))
        k3 = k2.unflatten(0, (B, Nk))
        v3 = v2.unflatten(0, (B, Nk))
        o3, lse = ops.attn_fwd(q3, k3, v3, cfg.heads, kmask, bias, 0.125, cfg.attn_p, cfg.seed, cfg.seed_ptr)
        o2 = o3.view(B * Nq, H)
        pre = ops.gemm(o2, Wo_c, bias=bo, res=x32, out_dtype=torch.float32, drop_p=cfg.hid_p, drop_seed=cfg.seed + 1,
                       seed_ptr=cfg.seed_ptr)
        y32, y16, mean, rstd = ops.layernorm_fwd(pre, gamma, beta, cfg.eps, True, _d16(cdt))
        ctx.cfg = cfg
        ctx.params = (Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma, beta)
        ctx.save_for_backward(xc, kvc, qkv, kvp, o2, lse, pre, mean, rstd, kmask, bias, Wqkv_c, Wo_c, gamma)
        if y16 is not None:
            ctx.mark_non_differentiable(y16)
        return y32, y16

    @staticmethod
    def backward(ctx, dy32, _dy16):
        cfg = ctx.cfg
        xc, kvc, qkv, kvp, o2, lse, pre, mean, rstd, kmask, bias, Wqkv_c, Wo_c, gamma = ctx.saved_tensors
        Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma_p, beta_p = ctx.params
        cdt = cfg.cdt
        H = pre.shape[1]
        B, Nq, Nk = cfg.B, cfg.Nq, cfg.Nk
        dy32 = torch.zeros_like(pre) if dy32 is None else dy32.contiguous()
        # LN backward: dpre32 feeds the residual path, dpre_c (dropout-masked) feeds the out-proj GEMMs
        dpre32, dpre_c, dgamma, dbeta, dbo = _ln_bwd_split(dy32, pre, gamma, mean, rstd, cdt, cfg.hid_p, cfg.seed + 1,
                                                           cfg.seed_ptr, gamma_p, beta_p, bo)
        dWo, = _wgrad((Wo,), dpre_c, o2)                                                    # [H,H]
        do2 = ops.gemm(dpre_c, Wo_c, b_mn=True, out_dtype=cdt)                              # [M,H]
        want_dbias = bias is not None and ctx.needs_input_grad[5]
        if not cfg.cross:
            dqkv = torch.empty_like(qkv)
            q2, k2, v2 = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
            dq2, dk2, dv2 = dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:]
        else:
            dqkv = torch.empty_like(qkv)          # dq only
            dkvp = torch.empty_like(kvp)
            q2, k2, v2 = qkv, kvp[:, :H], kvp[:, H:]
            dq2, dk2, dv2 = dqkv, dkvp[:, :H], dkvp[:, H:]
        dbias = ops.attn_bwd(do2.view(B, Nq, H), q2.unflatten(0, (B, Nq)), k2.unflatten(0, (B, Nk)),
                             v2.unflatten(0, (B, Nk)), o2.view(B, Nq, H), lse, cfg.heads,
                             dq2.unflatten(0, (B, Nq)), dk2.unflatten(0, (B, Nk)), dv2.unflatten(0, (B, Nk)),
                             kmask, bias, 0.125, cfg.attn_p, cfg.seed, cfg.seed_ptr, want_dbias=want_dbias)
        dkv32 = None
        dkvp32 = None
        if ctx.cached_kv:
            dWq, = _wgrad((Wq,), dqkv, xc)
            dbq, = _bgrad((bq,), dqkv)
            dWk = dWv = dbk = dbv = None
            dx32 = ops.gemm(dqkv, Wqkv_c[:H], b_mn=True, res=dpre32, out_dtype=torch.float32)
            dkvp32 = dkvp if cdt == torch.float32 else ops.cast(dkvp, torch.float32)   # summed over steps in fp32
        elif not cfg.cross:
            dWq, dWk, dWv = _wgrad((Wq, Wk, Wv), dqkv, xc)                                   # [3H,H]
            dbq, dbk, dbv = _bgrad((bq, bk, bv), dqkv)
            dx32 = ops.gemm(dqkv, Wqkv_c, b_mn=True, res=dpre32, out_dtype=torch.float32)   # [M,H]
        else:
            dWq, = _wgrad((Wq,), dqkv, xc)                                                   # [H,H]
            dWk, dWv = _wgrad((Wk, Wv), dkvp, kvc)                                           # [2H,H]
            dbq, = _bgrad((bq,), dqkv)
            dbk, dbv = _bgrad((bk, bv), dkvp)
            dx32 = ops.gemm(dqkv, Wqkv_c[:H], b_mn=True, res=dpre32, out_dtype=torch.float32)
            if ctx.needs_input_grad[2]:
                dkv32 = ops.gemm(dkvp, Wqkv_c[H:], b_mn=True, out_dtype=torch.float32)      # [Mk,H]
        return (dx32, None, dkv32, None, None, dbias, dWq, dbq, dWk, dbk, dWv, dbv,
                dWo, dbo, dgamma, dbeta, None, None, None, None, dkvp32, None)


class KVProjFn(torch.autograd.Function):
    """[K | V] = kv W_kv^T + b_kv of one cross-attention, as its own autograd node: a rollout computes it ONCE per
    layer for the instruction embeddings and every navigation step reuses it (the reference re-projects
    the same text 15 x 2 x 3 times per rollout, M/r2r/agent.py:575-590 -> P/model/Bert_backbone.py:221-224).
    Returns (kvp32 fp32 [Mk,2H], kvp16 16-bit copy or None).  The steps' gradients arrive summed (fp32) and the
    projection's dgrad / wgrad GEMMs and bias sums run once."""

    @staticmethod
    def forward(ctx, kv32, kv16, Wk, bk, Wv, bv, Wkv_c, bkv, cdt):
        ctx.set_materialize_grads(False)
        kvc = _c(kv32, kv16, cdt)
        kvp16 = None
        if cdt == torch.float32:
            kvp32 = ops.gemm(kvc, Wkv_c, bias=bkv, out_dtype=torch.float32)
        else:
            kvp16 = torch.empty((kvc.shape[0], Wkv_c.shape[0]), device=kvc.device, dtype=cdt)
            kvp32 = ops.gemm(kvc, Wkv_c, bias=bkv, out_dtype=torch.float32, out2=kvp16)
            ctx.mark_non_differentiable(kvp16)
        ctx.cdt = cdt
        ctx.params = (Wk, bk, Wv, bv)
        ctx.save_for_backward(kvc, Wkv_c)
        return kvp32, kvp16

    @staticmethod
    def backward(ctx, dkvp32, _d16):
        if dkvp32 is None:
            return (None,) * 9
        kvc, Wkv_c = ctx.saved_tensors
        Wk, bk, Wv, bv = ctx.params
        dkvp_c = dkvp32.contiguous() if ctx.cdt == torch.float32 else ops.cast(dkvp32.contiguous(), ctx.cdt)
        dWk, dWv = _wgrad((Wk, Wv), dkvp_c, kvc)
        dbk, dbv = _bgrad((bk, bv), dkvp_c)
        dkv32 = ops.gemm(dkvp_c, Wkv_c, b_mn=True, out_dtype=torch.float32) if ctx.needs_input_grad[0] else None
        return dkv32, None, dWk, dbk, dWv, dbv, None, None, None


FFNCfg = namedtuple("FFNCfg", "eps hid_p seed seed_ptr cdt")


class FFNBlockFn(torch.autograd.Function):
    """y = LN(dropout(gelu(x W1^T + b1) W2^T + b2) + x)"""

    @staticmethod
    def forward(ctx, x32, x16, W1, b1, W2, b2, gamma, beta, W1_c, W2_c, cfg):
        ctx.set_materialize_grads(False)
        cdt = cfg.cdt
        xc = _c(x32, x16, cdt)
        M = xc.shape[0]
        F = W1_c.shape[0]
        z = torch.empty((M, F), device=xc.device, dtype=cdt)
        h = ops.gemm(xc, W1_c, bias=b1, act=ops.ACT_GELU, aux_out=z, out_dtype=cdt)
        pre = ops.gemm(h, W2_c, bias=b2, res=x32, out_dtype=torch.float32, drop_p=cfg.hid_p, drop_seed=cfg.seed,
                       seed_ptr=cfg.seed_ptr)
        y32, y16, mean, rstd = ops.layernorm_fwd(pre, gamma, beta, cfg.eps, True, _d16(cdt))
        ctx.cfg = cfg
        ctx.params = (W1, b1, W2, b2, gamma, beta)
        ctx.save_for_backward(xc, z, h, pre, mean, rstd, W1_c, W2_c, gamma)
        if y16 is not None:
            ctx.mark_non_differentiable(y16)
        return y32, y16

    @staticmethod
    def backward(ctx, dy32, _dy16):
        cfg = ctx.cfg
        xc, z, h, pre, mean, rstd, W1_c, W2_c, gamma = ctx.saved_tensors
        cdt = cfg.cdt
        W1, b1, W2, b2, gamma_p, beta_p = ctx.params
        dy32 = torch.zeros_like(pre) if dy32 is None else dy32.contiguous()
        dpre32, dpre_c, dgamma, dbeta, db2 = _ln_bwd_split(dy32, pre, gamma, mean, rstd, cdt, cfg.hid_p,
                                                           cfg.seed, cfg.seed_ptr, gamma_p, beta_p, b2)
        dW2, = _wgrad((W2,), dpre_c, h)                                                      # [H,F]
        dz = ops.gemm(dpre_c, W2_c, b_mn=True, act=ops.ACT_DGELU, aux_in=z, out_dtype=cdt)   # [M,F]
        db1, = _bgrad((b1,), dz)
        dW1, = _wgrad((W1,), dz, xc)                                                         # [F,H]
        dx32 = ops.gemm(dz, W1_c, b_mn=True, res=dpre32, out_dtype=torch.float32)           # [M,H]
        return dx32, None, dW1, db1, dW2, db2, dgamma, dbeta, None, None, None


PanoCfg = namedtuple("PanoCfg", "B N heads eps hid_p seed seed_ptr cdt")


class PanoLayerFn(torch.autograd.Function):
    """pre-LN encoder layer:  y = x + drop(OutProj(Attn(LN1 x)));  out = y + drop(W2 drop(gelu(W1 LN2 y)))"""

    @staticmethod
    def forward(ctx, x32, kmask, Win, bin_, Wout, bout, W1, b1, W2, b2, g1, be1, g2, be2, Win_c, Wout_c, W1_c, W2_c, cfg):
        cdt = cfg.cdt
        H = x32.shape[1]
        B, N = cfg.B, cfg.N
        f32 = cdt == torch.float32
        a32, a16, mean1, rstd1 = ops.layernorm_fwd(x32, g1, be1, cfg.eps, f32, _d16(cdt))
        x2c = a32 if f32 else a16
        qkv = ops.gemm(x2c, Win_c, bias=bin_, out_dtype=cdt)
        q3 = qkv[:, :H].unflatten(0, (B, N))
        k3 = qkv[:, H:2 * H].unflatten(0, (B, N))
        v3 = qkv[:, 2 * H:].unflatten(0, (B, N))
        o3, lse = ops.attn_fwd(q3, k3, v3, cfg.heads, kmask, None, 0.125, cfg.hid_p, cfg.seed, cfg.seed_ptr)
        o2 = o3.view(B * N, H)
        y32 = ops.gemm(o2, Wout_c, bias=bout, res=x32, out_dtype=torch.float32, drop_p=cfg.hid_p, drop_seed=cfg.seed + 1,
                       seed_ptr=cfg.seed_ptr)
        b32, b16, mean2, rstd2 = ops.layernorm_fwd(y32, g2, be2, cfg.eps, f32, _d16(cdt))
        y2c = b32 if f32 else b16
        F = W1_c.shape[0]
        z = torch.empty((B * N, F), device=x32.device, dtype=cdt)
        h = ops.gemm(y2c, W1_c, bias=b1, act=ops.ACT_GELU, aux_out=z, out_dtype=cdt, drop_p=cfg.hid_p,
                     drop_seed=cfg.seed + 2, seed_ptr=cfg.seed_ptr)
        out32 = ops.gemm(h, W2_c, bias=b2, res=y32, out_dtype=torch.float32, drop_p=cfg.hid_p, drop_seed=cfg.seed + 3,
                         seed_ptr=cfg.seed_ptr)
        ctx.cfg = cfg
        ctx.params = (Win, bin_, Wout, bout, W1, b1, W2, b2, g1, be1, g2, be2)
        ctx.save_for_backward(x32, x2c, qkv, o2, lse, y32, y2c, z, h, mean1, rstd1, mean2, rstd2, kmask, Win_c, Wout_c,
                              W1_c, W2_c, g1, g2)
        return out32

    @staticmethod
    def backward(ctx, dout32):
        cfg = ctx.cfg
        (x32, x2c, qkv, o2, lse, y32, y2c, z, h, mean1, rstd1, mean2, rstd2, kmask, Win_c, Wout_c, W1_c, W2_c, g1,
         g2) = ctx.saved_tensors
        cdt = cfg.cdt
        H = x32.shape[1]
        B, N = cfg.B, cfg.N
        p, sp = cfg.hid_p, cfg.seed_ptr
        dout32 = dout32.contiguous()
        # FFN half
        dout_c = ops.cast(dout32, cdt, drop_p=p, drop_seed=cfg.seed + 3, seed_ptr=sp) if (p > 0 or cdt != torch.float32) \
            else dout32
        Win, bin_, Wout, bout, W1, b1, W2, b2, g1p, be1p, g2p, be2p = ctx.params
        db2, = _bgrad((b2,), dout_c)
        dW2, = _wgrad((W2,), dout_c, h)
        dz = ops.gemm(dout_c, W2_c, b_mn=True, act=ops.ACT_DGELU, aux_in=z, out_dtype=cdt, drop_p=p,
                      drop_seed=cfg.seed + 2, seed_ptr=sp)
        db1, = _bgrad((b1,), dz)
        dW1, = _wgrad((W1,), dz, y2c)
        dy2 = ops.gemm(dz, W1_c, b_mn=True, out_dtype=torch.float32)
        dy32, _, dg2, dbe2, _ = _ln_bwd(dy2, y32, g2, mean2, rstd2, dout32, None, 0.0, 0, None, g2p, be2p)
        # attention half
        dy_c = ops.cast(dy32, cdt, drop_p=p, drop_seed=cfg.seed + 1, seed_ptr=sp) if (p > 0 or cdt != torch.float32) \
            else dy32
        dbout, = _bgrad((bout,), dy_c)
        dWout, = _wgrad((Wout,), dy_c, o2)
        do2 = ops.gemm(dy_c, Wout_c, b_mn=True, out_dtype=cdt)
        dqkv = torch.empty_like(qkv)
        ops.attn_bwd(do2.view(B, N, H), qkv[:, :H].unflatten(0, (B, N)), qkv[:, H:2 * H].unflatten(0, (B, N)),
                     qkv[:, 2 * H:].unflatten(0, (B, N)), o2.view(B, N, H), lse, cfg.heads,
                     dqkv[:, :H].unflatten(0, (B, N)), dqkv[:, H:2 * H].unflatten(0, (B, N)),
                     dqkv[:, 2 * H:].unflatten(0, (B, N)), kmask, None, 0.125, p, cfg.seed, sp)
        dbin, = _bgrad((bin_,), dqkv)
        dWin, = _wgrad((Win,), dqkv, x2c)
        dx2 = ops.gemm(dqkv, Win_c, b_mn=True, out_dtype=torch.float32)
        dx32, _, dg1, dbe1, _ = _ln_bwd(dx2, x32, g1, mean1, rstd1, dy32, None, 0.0, 0, None, g1p, be1p)
        return (dx32, None, dWin, dbin, dWout, dbout, dW1, db1, dW2, db2, dg1, dbe1, dg2, dbe2, None, None, None, None,
                None)


def _pad8(n):
    return (n + 7) // 8 * 8


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b), act in {none, relu, tanh, gelu}; fp32 in / fp32 out, GEMM in the compute dtype.
    Shapes the tcgen05 path cannot take (K or N tiny: 7/14-wide position features, 1-wide heads) run on
    the fp32 SIMT kernel regardless of the compute dtype.  A wide output whose width is not a multiple of 8 (the
    50265-wide vocabulary projection) is produced into a row-padded buffer and returned as a view: the leading
    dimension stays a multiple of 8, so the epilogue stores vectors and the backward GEMMs (whose K / M dimension is
    that width) can read the 16-bit gradient through TMA instead of falling to the SIMT kernel."""

    @staticmethod
    def forward(ctx, x32, x16, W, b, W_c, act, cdt):
        small = (W.shape[1] % 8 != 0) or W.shape[1] < 16 or W.shape[0] < 8
        if small:
            cdt = torch.float32
            W_c = W
        xc = _c(x32, x16, cdt)
        M, N = xc.shape[0], W.shape[0]
        aux = None
        out = None
        if N % 8 != 0 and N >= 256 and cdt != torch.float32:
            out = torch.empty((M, _pad8(N)), device=xc.device, dtype=torch.float32)[:, :N]
        if act == ops.ACT_GELU:
            # fp32 pre-activation kept for backward, exact erf GELU on it: a fused 16-bit epilogue would round the
            # pre-activation to 16 bits before the GELU (fine inside an FFN, whose output is 16-bit anyway; a visible
            # error for these fp32-out head transforms)
            aux = ops.gemm(xc, W_c, bias=b, out_dtype=torch.float32)
            y = ops.act_fwd(aux, act)
        else:
            y = ops.gemm(xc, W_c, bias=b, act=act, aux_out=aux, out=out, out_dtype=torch.float32)
        ctx.act, ctx.cdt = act, cdt
        ctx.params = (W, b)
        ctx.has_bias = b is not None
        ctx.save_for_backward(xc, W_c, y if act in (ops.ACT_RELU, ops.ACT_TANH) else None, aux)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, W_c, y, aux = ctx.saved_tensors
        cdt = ctx.cdt
        M, N = dy.shape
        padded = (dy.dim() == 2 and dy.stride(1) == 1 and dy.stride(0) == _pad8(N) and dy.stride(0) != N and
                  ctx.act == ops.ACT_NONE)
        if padded:
            # gradient of a row-padded output (see forward): convert the whole padded block, keep the leading dimension
            Np = dy.stride(0)
            base = dy.as_strided((M, Np), (Np, 1))
            dyc = (base if cdt == torch.float32 else ops.cast(base, cdt))[:, :N]
        else:
            dy = dy.contiguous()
            if ctx.act in (ops.ACT_RELU, ops.ACT_TANH):
                dyc = ops.act_grad(dy, y, ctx.act, cdt)          # dy * act'(y), converted to the operand dtype, one kernel
            elif ctx.act == ops.ACT_GELU:
                dyc = ops.act_grad(dy, aux, ctx.act, cdt)
            else:
                dyc = dy if cdt == torch.float32 else ops.cast(dy, cdt)
        W, b = ctx.params
        db, = _bgrad((b,), dyc)
        dW, = _wgrad((W,), dyc, xc)
        dx = ops.gemm(dyc, W_c, b_mn=True, out_dtype=torch.float32) if ctx.needs_input_grad[0] else None
        return dx, None, dW, db, None, None, None


class LayerNormFn(torch.autograd.Function):
    """y = LN(x) with fp32 statistics; returns (y32, y16-or-None)."""

    @staticmethod
    def forward(ctx, x32, gamma, beta, eps, cdt):
        ctx.set_materialize_grads(False)
        y32, y16, mean, rstd = ops.layernorm_fwd(x32, gamma, beta, eps, True, _d16(cdt))
        ctx.params = (gamma, beta)
        ctx.save_for_backward(x32, gamma, mean, rstd)
        if y16 is not None:
            ctx.mark_non_differentiable(y16)
        return y32, y16

    @staticmethod
    def backward(ctx, dy, _d16):
        x32, gamma, mean, rstd = ctx.saved_tensors
        dy = torch.zeros_like(x32) if dy is None else dy
        dx32, _, dg, db, _ = _ln_bwd(dy.contiguous(), x32, gamma, mean, rstd, None, None, 0.0, 0, None, ctx.params[0],
                                     ctx.params[1])
        return dx32, dg, db, None, None


# --------------------------------------------------------------------------------------
# heads.cu composites: pooling, dictionary sums, door gate, cross-entropy, gather-reduce, embeddings
# --------------------------------------------------------------------------------------
def _acc_dst(param, shape=None):
    """Destination a kernel can ACCUMULATE a parameter gradient into: the parameter's flat-gradient view when it
    lives in an engine.FlatParams buffer (zero at step start, see _grad_into), else fresh zeros that are handed back
    to autograd.  -> (tensor, returned_to_autograd)"""
    if param is None:
        return None, False
    d = getattr(param, "_goat_grad", None)
    if d is not None and d.is_contiguous():
        param._goat_fresh = False
        return d, False
    return torch.zeros(param.shape if shape is None else shape, device=param.device, dtype=torch.float32), True


class AttnPoolFn(torch.autograd.Function):
    """mode 0: out = sum_n softmax_n(tanh(x_n . w + b)) x_n  (adaptive panorama fusion, P/model/vilmodel_goat.py:354-362)
    mode 1: out = tanh(sum_n softmax_n(tanh(x_n) . w) x_n)    (CFP pooling, P/model/pretrain_goat.py:502-515)"""

    @staticmethod
    def forward(ctx, x, w, b, mode, n_valid=None):
        """n_valid: optional CUDA int32 [1] -- pool over the first n_valid tokens only (see goat_attn_pool_fwd)"""
        x = x.contiguous()
        wv = w.detach().reshape(-1).contiguous()
        out, a, s = ops.attn_pool_fwd(x, wv, None if b is None else b.detach().contiguous(), mode, n_valid)
        ctx.mode = mode
        ctx.params = (w, b)
        ctx.n_valid = n_valid
        ctx.save_for_backward(x, wv, a, s, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, wv, a, s, out = ctx.saved_tensors
        w, b = ctx.params
        dw, ret_w = _acc_dst(w)
        db, ret_b = _acc_dst(b)
        dx = ops.attn_pool_bwd(dout.contiguous(), x, wv, a, s, out, ctx.mode, dw.view(-1), None if db is None else db.view(-1),
                               ctx.n_valid)
        return dx, (dw if ret_w else None), (db if ret_b else None), None, None


class WSumFn(torch.autograd.Function):
    """out[b] = sum_n p[b,n] x[b,n]   (BACL dictionary expectation, M/models/vilmodel_GOAT.py:664-665)"""

    @staticmethod
    def forward(ctx, x, p):
        x = x.contiguous()
        p = p.to(torch.float32).reshape(x.shape[0], x.shape[1]).contiguous()
        ctx.save_for_backward(p)
        ctx.N = x.shape[1]
        return ops.wsum_fwd(x, p)

    @staticmethod
    def backward(ctx, dout):
        p, = ctx.saved_tensors
        return ops.wsum_bwd(dout.contiguous(), p, ctx.N), None


class DoorGateFn(torch.autograd.Function):
    """g = sigmoid(aug . wa + ba + ori . wo + bo); out = g aug + (1-g) ori   (M/models/vilmodel_GOAT.py:145-148, :548-552)"""

    @staticmethod
    def forward(ctx, aug, ori, wa, ba, wo, bo):
        shp = aug.shape
        a2 = aug.reshape(-1, shp[-1]).contiguous()
        o2 = ori.reshape(-1, shp[-1]).contiguous()
        wav, wov = wa.detach().reshape(-1).contiguous(), wo.detach().reshape(-1).contiguous()
        out, gate = ops.door_gate_fwd(a2, o2, wav, ba.detach().contiguous(), wov, bo.detach().contiguous())
        ctx.params = (wa, ba, wo, bo)
        ctx.shp = shp
        ctx.save_for_backward(a2, o2, wav, wov, gate)
        return out.view(shp)

    @staticmethod
    def backward(ctx, dout):
        a2, o2, wav, wov, gate = ctx.saved_tensors
        wa, ba, wo, bo = ctx.params
        dwa, r1 = _acc_dst(wa)
        dba, r2 = _acc_dst(ba)
        dwo, r3 = _acc_dst(wo)
        dbo, r4 = _acc_dst(bo)
        daug, dori = ops.door_gate_bwd(dout.reshape(a2.shape).contiguous(), a2, o2, wav, wov, gate, dwa.view(-1), dwo.view(-1),
                                       dba.view(-1), dbo.view(-1))
        return (daug.view(ctx.shp), dori.view(ctx.shp), dwa if r1 else None, dba if r2 else None, dwo if r3 else None,
                dbo if r4 else None)


class XentFn(torch.autograd.Function):
    """F.cross_entropy(logits, labels, reduction='none') with ignore_index; logits may be a strided view (sim.T)."""

    @staticmethod
    def forward(ctx, logits, labels, ignore_index):
        if logits.dtype != torch.float32:
            logits = logits.float()
        labels = labels.contiguous()
        loss, lse = ops.xent_fwd(logits, labels, ignore_index)
        ctx.ignore_index = ignore_index
        ctx.save_for_backward(logits, labels, lse)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        logits, labels, lse = ctx.saved_tensors
        out = None
        if logits.stride(1) == 1 and logits.stride(0) != logits.shape[1]:
            # row-padded logits (LinearFn): the gradient keeps the padded leading dimension; the pad columns are zeroed
            # because the consumer converts the whole padded block
            Np = logits.stride(0)
            out = torch.zeros((logits.shape[0], Np), device=logits.device, dtype=torch.float32)[:, :logits.shape[1]]
        return ops.xent_bwd(dloss.contiguous(), logits, labels, lse, ctx.ignore_index, out=out), None, None


class VocabXentFn(torch.autograd.Function):
    """cross_entropy(h W^T + b, labels) per row WITHOUT materialising the [n, vocab] logits (the reference
    writes them as fp32 and reads them back three times, P/model/Bert_backbone.py:826 + P/model/pretrain_goat.py:209-224).
    The class axis is walked in chunks of ``CHUNK`` columns: forward = chunk GEMM (tcgen05, fp32 out into ONE reused,
    L2-resident buffer) + a running (max, sum exp, picked logit) update; backward recomputes each chunk's logits, turns
    them into the 16-bit gradient chunk, and feeds the dgrad (accumulated over chunks) / wgrad / bias-gradient kernels."""

    CHUNK = 6144      # 24 tiles of 256 columns x 3 row tiles of a 640-row batch = 72 pair tiles: one wave of the 74 pairs

    @staticmethod
    def forward(ctx, h32, W, b, labels, ignore_index, W_c, cdt):
        n, V = h32.shape[0], W.shape[0]
        hc = _c(h32.contiguous(), None, cdt)
        labels = labels.contiguous()
        dev = h32.device
        m = torch.empty(n, device=dev, dtype=torch.float32)
        l = torch.empty(n, device=dev, dtype=torch.float32)
        picked = torch.empty(n, device=dev, dtype=torch.float32)
        C_ = min(VocabXentFn.CHUNK, _pad8(V))
        buf = torch.empty((n, C_), device=dev, dtype=torch.float32)
        for c0 in range(0, V, C_):
            nc = min(C_, V - c0)
            lg = ops.gemm(hc, W_c[c0:c0 + nc], bias=None if b is None else b.detach()[c0:c0 + nc], out=buf[:, :nc],
                          out_dtype=torch.float32)
            ops.xent_chunk_fwd(lg, labels, c0, c0 == 0, m, l, picked)
        lse = m + torch.log(l)
        valid = labels != ignore_index
        loss = torch.where(valid, lse - picked, torch.zeros_like(lse))
        ctx.ignore_index, ctx.cdt, ctx.chunk = ignore_index, cdt, C_
        ctx.params = (W, b)
        ctx.save_for_backward(hc, W_c, labels, lse)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        hc, W_c, labels, lse = ctx.saved_tensors
        W, b = ctx.params
        cdt, C_ = ctx.cdt, ctx.chunk
        n, V = hc.shape[0], W.shape[0]
        dev = hc.device
        dloss = dloss.contiguous()
        buf = torch.empty((n, C_), device=dev, dtype=torch.float32)
        d16 = torch.empty((n, C_), device=dev, dtype=cdt)
        dh = torch.zeros((n, hc.shape[1]), device=dev, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        gW = getattr(W, "_goat_grad", None)
        gb = getattr(b, "_goat_grad", None) if b is not None else None
        dW = None if gW is not None else torch.zeros(W.shape, device=dev, dtype=torch.float32)
        db = None if (gb is not None or b is None) else torch.zeros(b.shape, device=dev, dtype=torch.float32)
        for c0 in range(0, V, C_):
            nc = min(C_, V - c0)
            lg = ops.gemm(hc, W_c[c0:c0 + nc], bias=None if b is None else b.detach()[c0:c0 + nc], out=buf[:, :nc],
                          out_dtype=torch.float32)
            dl = d16[:, :nc]
            ops.xent_chunk_bwd(dloss, lg, labels, lse, c0, ctx.ignore_index, V, dl)
            if dh is not None:
                ops.gemm(dl, W_c[c0:c0 + nc], b_mn=True, out=dh, accumulate=True)            # dh += dlogits_c W_c
            ops.gemm(dl, hc, a_mn=True, b_mn=True, out=(gW if gW is not None else dW)[c0:c0 + nc], accumulate=True)
            if b is not None:
                ops.colsum(dl, out=(gb if gb is not None else db)[c0:c0 + nc], accumulate=True)
        _mark((W, b))
        return dh, dW, db, None, None, None, None


class SegmentReduceFn(torch.autograd.Function):
    """out[r] = sum / mean over k of src[idx[r,k]] (idx -1 = empty)."""

    @staticmethod
    def forward(ctx, src, idx, mean):
        src = src.contiguous()
        ctx.mean, ctx.n_src = mean, src.shape[0]
        ctx.save_for_backward(idx)
        return ops.segment_reduce_fwd(src, idx, mean)

    @staticmethod
    def backward(ctx, dout):
        idx, = ctx.saved_tensors
        return ops.segment_reduce_bwd(dout.contiguous(), idx, ctx.mean, ctx.n_src), None, None


class EmbedFn(torch.autograd.Function):
    """word[ids] + pos[arange(L)] + type[0]  -> [B*L, H]   (P/model/Bert_backbone.py:87-116, before LayerNorm)"""

    @staticmethod
    def forward(ctx, ids, word, pos, type_, padding_idx=-1):
        ids = ids.contiguous()
        ctx.params = (word, pos, type_)
        ctx.padding_idx = -1 if padding_idx is None else int(padding_idx)
        ctx.save_for_backward(ids)
        return ops.embed_fwd(ids, word.detach(), pos.detach(), type_.detach())

    @staticmethod
    def backward(ctx, dout):
        ids, = ctx.saved_tensors
        word, pos, type_ = ctx.params
        need = ctx.needs_input_grad
        dw, rw = _acc_dst(word) if need[1] else (None, False)
        dp, rp = _acc_dst(pos) if need[2] else (None, False)
        dt_, rt = _acc_dst(type_) if need[3] else (None, False)
        ops.embed_bwd(dout.contiguous(), ids, dw, dp, None if dt_ is None else dt_[0], ctx.padding_idx)
        return None, (dw if rw else None), (dp if rp else None), (dt_ if rt else None), None


class GatherRowsFn(torch.autograd.Function):
    """table[ids]  (nn.Embedding of the global-map step ids, P/model/vilmodel_goat.py:478-480): one gather kernel; the
    table gradient is accumulated into the flat gradient view (or returned to autograd)."""

    @staticmethod
    def forward(ctx, ids, table):
        shp = ids.shape
        idx = ids.reshape(-1, 1).to(torch.int32).contiguous()
        ctx.params = (table,)
        ctx.save_for_backward(idx)
        return ops.segment_reduce_fwd(table.detach().contiguous(), idx, False).view(shp + (table.shape[1],))

    @staticmethod
    def backward(ctx, dout):
        idx, = ctx.saved_tensors
        table, = ctx.params
        dst, ret = _acc_dst(table)
        ops.segment_reduce_bwd(dout.reshape(idx.shape[0], -1).contiguous(), idx, False, table.shape[0], out=dst)
        return None, (dst if ret else None)


class SprelFn(torch.autograd.Function):
    """sprel_linear = nn.Linear(1, 1) over the pairwise distances: d * w + b  (P/model/vilmodel_goat.py:499-501).  The
    distances are data; dw / db are two scalar reductions accumulated by one kernel."""

    @staticmethod
    def forward(ctx, d, w, b):
        d = d.to(torch.float32).contiguous()
        ctx.params = (w, b)
        ctx.save_for_backward(d)
        return ops.sprel_fwd(d, w.detach().reshape(1), b.detach().reshape(1))

    @staticmethod
    def backward(ctx, dout):
        d, = ctx.saved_tensors
        w, b = ctx.params
        dw, rw = _acc_dst(w)
        db, rb = _acc_dst(b)
        ops.sprel_bwd(dout.contiguous(), d, dw.view(-1), db.view(-1))
        return None, (dw if rw else None), (db if rb else None)


class DropoutFn(torch.autograd.Function):
    """nn.Dropout on an fp32 tensor with the counter-based mask (regenerated in backward from the seed)."""

    @staticmethod
    def forward(ctx, x, p, seed, seed_ptr):
        x = x.contiguous()
        ctx.p, ctx.seed, ctx.seed_ptr = p, seed, seed_ptr
        return ops.cast(x, torch.float32, drop_p=p, drop_seed=seed, seed_ptr=seed_ptr)

    @staticmethod
    def backward(ctx, dy):
        return ops.cast(dy.contiguous(), torch.float32, drop_p=ctx.p, drop_seed=ctx.seed, seed_ptr=ctx.seed_ptr), None, None, None


def dropout(x, p, training):
    if not training or p <= 0.0:
        return x
    return DropoutFn.apply(x, float(p), next_seed(), seed_ptr())
