"""GPU parity of the C-ABI primitives (goat_gemm, goat_attn_core_*, goat_layernorm_*, goat_adamw_step) against
fp64 torch / the CPU oracle, and the tcgen05 kernels against the SIMT kernels on the same inputs.

Tolerances: fp32 paths 1e-5 relative to the output scale; 16-bit paths compare against an fp64 reference computed
from the SAME rounded 16-bit inputs, so what is bounded is the accumulation / output rounding: 2e-3 (fp16 out),
1.6e-2 (bf16 out), 1e-5 (fp32 out of a tensor-core GEMM); all relative to max(1, max|ref|).
"""
import pytest
import torch

from oracle import goat_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(got, ref):
    ref = ref.double().cpu()
    return (got.double().cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1.0)


OUT_TOL = {torch.float32: 2e-5, torch.float16: 2e-3, torch.bfloat16: 1.6e-2}


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 768, 768), (2368, 2304, 768), (5120, 768, 3072), (77, 200, 72),
                                   (2368, 768, 2368), (8, 8, 16), (768, 768, 5120), (768, 3072, 2368)])
def test_gemm_umma_vs_fp64_and_simt(dtype, a_mn, b_mn, M, N, K):
    from vln_goat_b200 import ops
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("MN-major operands need a leading dimension that is a multiple of 8")
    torch.manual_seed(M + N + K)
    A = (torch.randn(M, K, device=DEV) * 0.5).to(dtype)
    B = (torch.randn(N, K, device=DEV) * 0.5).to(dtype)
    ref = A.double() @ B.double().t()
    Ain = A.t().contiguous() if a_mn else A
    Bin = B.t().contiguous() if b_mn else B
    out = ops.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    out_s = ops.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, force_simt=True)
    assert rel(out, ref) < 2e-5
    assert rel(out_s, ref) < 2e-5
    out16 = ops.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn)
    assert rel(out16, ref) < OUT_TOL[dtype]
    # accumulate mode (split-K + fp32 atomics): twice into the same zero-initialised buffer
    acc = ops.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn, accumulate=True)
    ops.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn, accumulate=True, out=acc, alpha=0.5)
    assert rel(acc, 1.5 * ref) < 2e-5


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_gemm_epilogues(dtype):
    from vln_goat_b200 import ops
    torch.manual_seed(5)
    M, N, K = 300, 768, 768
    A = (torch.randn(M, K, device=DEV) * 0.5).to(dtype)
    W = (torch.randn(N, K, device=DEV) * 0.05).to(dtype)
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    acc = A.double() @ W.double().t()
    out = ops.gemm(A, W, bias=bias, res=res, alpha=0.5, out_dtype=torch.float32)
    assert rel(out, 0.5 * acc + bias.double() + res.double()) < 2e-5
    z = torch.empty(M, N, device=DEV, dtype=dtype)
    h = ops.gemm(A, W, bias=bias, act=ops.ACT_GELU, aux_out=z)
    zr = acc + bias.double()
    assert rel(z, zr) < OUT_TOL[dtype]
    assert rel(h, O.gelu_erf(z.double())) < OUT_TOL[dtype]     # GELU of the ROUNDED pre-activation (what backward sees)
    g = ops.gemm(A, W, act=ops.ACT_DGELU, aux_in=z, out_dtype=torch.float32)
    zz = z.double().requires_grad_(True)
    O.gelu_erf(zz).sum().backward()
    assert rel(g, acc * zz.grad) < 1e-4
    r = ops.gemm(A, W, bias=bias, act=ops.ACT_RELU, out_dtype=torch.float32)
    assert rel(r, torch.relu(zr)) < 2e-5
    t = ops.gemm(A, W, bias=bias, act=ops.ACT_TANH, out_dtype=torch.float32)
    assert rel(t, torch.tanh(zr)) < 2e-5
    # dropout: same mask on the tcgen05 and SIMT kernels, keep fraction ~ 1-p, kept values scaled by 1/(1-p)
    d1 = ops.gemm(A, W, out_dtype=torch.float32, drop_p=0.1, drop_seed=77)
    d2 = ops.gemm(A, W, out_dtype=torch.float32, drop_p=0.1, drop_seed=77, force_simt=True)
    assert torch.equal(d1 == 0, d2 == 0)
    keep = (d1 != 0).double().mean().item()
    assert abs(keep - 0.9) < 0.01
    kept = d1 != 0
    assert rel(d1[kept], (acc / 0.9)[kept.cpu()] if not acc.is_cuda else (acc / 0.9)[kept]) < 2e-5


@pytest.mark.parametrize("M,N,K", [(5120, 3072, 768), (2368, 2304, 768), (6912, 3072, 768)])
def test_gemm_tail_wave_split_keeps_every_epilogue_operand_aligned(M, N, K):
    """Shapes whose last wave of 256 x 256 tiles is mostly empty run as two launches (row tiles that fill whole waves, then
    the remaining rows as 256 x 128 tiles; gemm_umma.cu).  The second launch must see its rows of the residual, the GELU
    pre-activation copies and the SAME dropout mask (indexed by the global row): compare with the SIMT kernel."""
    from vln_goat_b200 import ops
    torch.manual_seed(M + N)
    dtype = torch.float16
    A = (torch.randn(M, K, device=DEV) * 0.5).to(dtype)
    W = (torch.randn(N, K, device=DEV) * 0.05).to(dtype)
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    acc = A.double() @ W.double().t()
    out2 = torch.empty(M, N, device=DEV, dtype=dtype)
    out = ops.gemm(A, W, bias=bias, res=res, out_dtype=torch.float32, out2=out2, drop_p=0.1, drop_seed=4242)
    ref = ops.gemm(A, W, bias=bias, res=res, out_dtype=torch.float32, drop_p=0.1, drop_seed=4242, force_simt=True)
    assert rel(out, ref) < 2e-5
    assert rel(out2, ref) < OUT_TOL[dtype]
    mask = ops.cast(torch.ones(M, N, device=DEV), torch.float32, drop_p=0.1, drop_seed=4242) != 0
    plain = acc + bias.double() + res.double()
    assert rel(out[~mask], (bias.double() * 0 + res.double())[~mask]) < 2e-5     # dropped: only the residual survives
    z = torch.empty(M, N, device=DEV, dtype=dtype)
    h = ops.gemm(A, W, bias=bias, act=ops.ACT_GELU, aux_out=z)
    assert rel(z, acc + bias.double()) < OUT_TOL[dtype]
    assert rel(h, O.gelu_erf(z.double())) < OUT_TOL[dtype]
    g = ops.gemm(A, W, act=ops.ACT_DGELU, aux_in=z)
    zz = z.double().requires_grad_(True)
    O.gelu_erf(zz).sum().backward()
    assert rel(g, acc * zz.grad) < OUT_TOL[dtype]
    del plain


def test_dropout_mask_is_shared_by_gemm_layernorm_and_cast():
    """Hidden dropout is applied in a GEMM epilogue (forward) and regenerated by the LayerNorm backward / cast kernels
    from (seed, linear element index): all three must draw the same keep mask, at the requested rate."""
    from vln_goat_b200 import ops
    torch.manual_seed(9)
    M, H = 300, 768
    A = torch.randn(M, H, device=DEV).to(torch.bfloat16)
    W = torch.randn(H, H, device=DEV).to(torch.bfloat16)
    for seed in (9, 123456789012345):
        d = ops.gemm(A, W, out_dtype=torch.float32, drop_p=0.1, drop_seed=seed)
        mask = d != 0
        assert abs(mask.double().mean().item() - 0.9) < 0.01
        c = ops.cast(torch.ones(M, H, device=DEV), torch.float32, drop_p=0.1, drop_seed=seed)
        assert torch.equal(mask, c != 0)
        x = torch.randn(M, H, device=DEV)
        g, b = torch.ones(H, device=DEV), torch.zeros(H, device=DEV)
        _, _, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-5, True, torch.bfloat16)
        dy = torch.randn(M, H, device=DEV)
        dx32, dx16, _, _, _ = ops.layernorm_bwd(dy, x, g, mean, rstd, None, True, torch.bfloat16, 0.1, seed)
        live = dx32.abs() > 1e-3          # where the unmasked gradient is clearly non-zero
        assert torch.equal((dx16 != 0)[live], mask[live])
        # masks of different seeds differ
        d2 = ops.gemm(A, W, out_dtype=torch.float32, drop_p=0.1, drop_seed=seed + 1)
        assert (mask != (d2 != 0)).double().mean().item() > 0.1


def test_gemm_fp32_simt_small_shapes():
    from vln_goat_b200 import ops
    torch.manual_seed(6)
    for (M, N, K) in ((300, 768, 7), (300, 1, 768), (37, 768, 14), (5, 3, 2)):
        A = torch.randn(M, K, device=DEV)
        W = torch.randn(N, K, device=DEV)
        b = torch.randn(N, device=DEV)
        out = ops.gemm(A, W, bias=b)
        assert rel(out, A.double() @ W.double().t() + b.double()) < 1e-5
    # weight gradients of the narrow layers: dW[N,K'] = dY^T X with the reduction over all tokens -- split-K + atomics
    for (T_, N, Kp) in ((8064, 768, 7), (1536, 1, 768), (2368, 768, 14), (100, 5, 3)):
        dy = torch.randn(T_, N, device=DEV)
        x = torch.randn(T_, Kp, device=DEV)
        acc0 = torch.randn(N, Kp, device=DEV)
        out = acc0.clone()
        ops.gemm(dy, x, a_mn=True, b_mn=True, out=out, accumulate=True)
        assert rel(out, acc0.double() + dy.double().t() @ x.double()) < 1e-5


def _attn_case(B, Nq, Nk, dtype, sprel, neg_inf=False, seed=0):
    torch.manual_seed(seed)
    H = 768
    q = torch.randn(B, Nq, H).to(dtype)
    k = torch.randn(B, Nk, H).to(dtype)
    v = torch.randn(B, Nk, H).to(dtype)
    lens = torch.randint(1, Nk + 1, (B,))
    lens[0] = Nk
    valid = O.gen_seq_masks(lens, Nk)
    kmask = torch.zeros(B, Nk).masked_fill(~valid, float("-inf")) if neg_inf else (1.0 - valid.float()) * -10000.0
    bias = torch.randn(B, Nq, Nk) if sprel else None
    w = torch.randn(B, Nq, H).to(dtype)
    return q, k, v, kmask, bias, w


def _attn_ref(q, k, v, kmask, bias, w, heads=12):
    q64, k64, v64 = (t.double().requires_grad_(True) for t in (q, k, v))
    b64 = bias.double().requires_grad_(True) if bias is not None else None
    mask = kmask.double()[:, None, None, :]
    if b64 is not None:
        mask = mask + b64[:, None]
    ref = O.attn_core(q64, k64, v64, mask, heads)
    (ref * w.double()).sum().backward()
    return ref.detach(), q64.grad, k64.grad, v64.grad, (b64.grad if b64 is not None else None)


ATT_SHAPES = [(2, 36, 80, False), (3, 12, 12, True), (2, 80, 80, False), (1, 37, 512, False), (2, 128, 200, True),
              (2, 130, 70, False), (1, 1, 1, False), (2, 38, 24, False), (4, 60, 129, True), (2, 37, 35, False),
              # long instructions (RxR 300 tokens, 512-token stress C5, 200-token R2R fine-tune text as query): the
              # tensor-core kernels tile the query axis (128-row tiles) and chunk the keys (128-key chunks)
              (2, 512, 512, False), (4, 300, 300, False), (2, 200, 37, False), (1, 257, 130, True)]


@pytest.mark.parametrize("B,Nq,Nk,sprel", ATT_SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_attention_core_vs_fp64(B, Nq, Nk, sprel, dtype):
    """fp32 -> SIMT kernels; 16-bit -> tcgen05 kernels (pipelined for Nq, Nk <= 128; query-tiled / key-chunked beyond)."""
    from vln_goat_b200 import ops
    q, k, v, kmask, bias, w = _attn_case(B, Nq, Nk, dtype, sprel)
    ref, rq, rk, rv_, rb = _attn_ref(q, k, v, kmask, bias, w)
    qd, kd, vd = q.to(DEV), k.to(DEV), v.to(DEV)
    bd = bias.to(DEV) if sprel else None
    o, lse = ops.attn_fwd(qd, kd, vd, 12, kmask.to(DEV), bd)
    tol = OUT_TOL[dtype]
    assert rel(o, ref) < tol
    dq, dk, dv = torch.empty_like(qd), torch.empty_like(kd), torch.empty_like(vd)
    db = ops.attn_bwd(w.to(DEV), qd, kd, vd, o, lse, 12, dq, dk, dv, kmask.to(DEV), bd, want_dbias=sprel)
    # backward operands (P, dS) are themselves rounded to 16 bits on the tensor-core path
    btol = tol if dtype == torch.float32 else 2 * tol
    assert rel(dq, rq) < btol
    assert rel(dk, rk) < btol
    assert rel(dv, rv_) < btol
    if sprel:
        assert rel(db, rb) < btol


@pytest.mark.parametrize("B,Nq,Nk,sprel", [(2, 37, 80, False), (2, 80, 200, True), (3, 38, 39, False), (2, 300, 300, False),
                                           (1, 200, 37, True)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_attention_tc_matches_simt_with_dropout(B, Nq, Nk, sprel, dtype):
    """Same counter-based dropout mask in both kernel families: outputs agree to 16-bit rounding."""
    from vln_goat_b200 import ops
    q, k, v, kmask, bias, w = (t.to(DEV) if t is not None else None for t in _attn_case(B, Nq, Nk, dtype, sprel, seed=3))
    kw = dict(drop_p=0.1, drop_seed=1234)
    o1, l1 = ops.attn_fwd(q, k, v, 12, kmask, bias, **kw)
    o2, l2 = ops.attn_fwd(q, k, v, 12, kmask, bias, force_simt=True, **kw)
    assert rel(o1, o2) < OUT_TOL[dtype]
    assert rel(l1, l2) < 1e-4
    g1 = [torch.empty_like(t) for t in (q, k, v)]
    g2 = [torch.empty_like(t) for t in (q, k, v)]
    b1 = ops.attn_bwd(w, q, k, v, o2, l2, 12, *g1, kmask, bias, want_dbias=sprel, **kw)
    b2 = ops.attn_bwd(w, q, k, v, o2, l2, 12, *g2, kmask, bias, want_dbias=sprel, force_simt=True, **kw)
    for a, b in zip(g1, g2):
        assert rel(a, b) < 2 * OUT_TOL[dtype]
    if sprel:
        assert rel(b1, b2) < 2 * OUT_TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attention_strided_qkv_and_neg_inf_padding(dtype):
    """pano-encoder style call: q/k/v are column slices of one [B,N,3H] buffer, -inf key padding."""
    from vln_goat_b200 import ops
    torch.manual_seed(9)
    B, N, H = 2, 36, 768
    qkv = torch.randn(B, N, 3 * H).to(dtype)
    lens = torch.tensor([36, 29])
    km = torch.zeros(B, N).masked_fill(~O.gen_seq_masks(lens, N), float("-inf"))
    c = qkv.double()
    ref = O.attn_core(c[:, :, :H], c[:, :, H:2 * H], c[:, :, 2 * H:], km.double()[:, None, None, :], 12)
    d = qkv.to(DEV)
    o, _ = ops.attn_fwd(d[:, :, :H], d[:, :, H:2 * H], d[:, :, 2 * H:], 12, km.to(DEV))
    assert rel(o, ref) < OUT_TOL[dtype]


@pytest.mark.parametrize("M,H", [(300, 768), (5120, 768), (7, 768), (64, 1024)])
@pytest.mark.parametrize("eps", [1e-12, 1e-5])
def test_layernorm_fwd_bwd(M, H, eps):
    from vln_goat_b200 import ops
    torch.manual_seed(M)
    x = torch.randn(M, H) * 2 + 0.3
    g = 1 + 0.1 * torch.randn(H)
    b = 0.1 * torch.randn(H)
    dy = torch.randn(M, H)
    dres = torch.randn(M, H)
    x64 = x.double().requires_grad_(True)
    g64, b64 = g.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = O.layernorm(x64, g64, b64, eps)
    (ref * dy.double()).sum().backward()
    y32, y16, mean, rstd = ops.layernorm_fwd(x.to(DEV), g.to(DEV), b.to(DEV), eps, True, torch.float16)
    assert rel(y32, ref.detach()) < 1e-5
    assert rel(y16, ref.detach()) < 2e-3
    dx32, dx16, dg, db, dcol = ops.layernorm_bwd(dy.to(DEV), x.to(DEV), g.to(DEV), mean, rstd, dres.to(DEV), True,
                                                 torch.float16, want_colsum=True)
    assert rel(dx32, x64.grad + dres.double()) < 1e-5
    assert rel(dx16, x64.grad) < 2e-3
    assert rel(dg, g64.grad) < 1e-5
    assert rel(db, b64.grad) < 1e-5
    assert rel(dcol, dx16.double().sum(0)) < 1e-5
    if H != 768:
        return
    # accumulate variant (H == 768: one kernel, atomics into destinations that already hold something)
    base = torch.full((H,), 0.5, device=DEV)
    ag, ab, ac = base.clone(), base.clone(), base.clone()
    ex32, ex16, _, _, _ = ops.layernorm_bwd(dy.to(DEV), x.to(DEV), g.to(DEV), mean, rstd, dres.to(DEV), True, torch.float16,
                                            want_colsum=True, dgamma_out=ag, dbeta_out=ab, dcol_out=ac, accumulate=True)
    assert torch.equal(ex32, dx32) and torch.equal(ex16, dx16)
    assert rel(ag - 0.5, g64.grad) < 1e-5
    assert rel(ab - 0.5, b64.grad) < 1e-5
    assert rel(ac - 0.5, dx16.double().sum(0)) < 1e-5


def test_colsum_and_cast():
    from vln_goat_b200 import ops
    torch.manual_seed(1)
    x = torch.randn(5120, 768, device=DEV).half()
    assert rel(ops.colsum(x), x.double().sum(0)) < 1e-5
    y = torch.randn(300, 2304, device=DEV)
    assert rel(ops.colsum(y[:, 768:1536]), y[:, 768:1536].double().sum(0)) < 1e-5
    z = torch.randn(1000, 33, device=DEV)
    assert torch.equal(ops.cast(z, torch.bfloat16), z.bfloat16())
    # one-kernel atomic variant: adds to what the destination holds; odd widths take the two-kernel path
    for t in (x, y[:, 768:1536], torch.randn(37, 3072, device=DEV).bfloat16(), z):
        dst = torch.full((t.shape[1],), 0.25, device=DEV)
        ops.colsum(t, out=dst, accumulate=True)
        assert rel(dst - 0.25, t.double().sum(0)) < 1e-5


@pytest.mark.parametrize("shadow", [None, torch.bfloat16])
def test_fused_adamw_matches_reference_numerics(shadow):
    """FlatParams.adamw_step (clip + AdamW, 2 launches) vs the per-tensor restatement of P/optim/adamw.py, 3 steps."""
    from vln_goat_b200 import engine
    torch.manual_seed(0)

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(37, 19)
            self.LayerNorm = torch.nn.LayerNorm(19)
            self.b = torch.nn.Linear(19, 5, bias=False)

    m = Toy().to(DEV)
    ref_p = {n: p.detach().cpu().clone() for n, p in m.named_parameters()}
    ref_m = {n: torch.zeros_like(p) for n, p in ref_p.items()}
    ref_v = {n: torch.zeros_like(p) for n, p in ref_p.items()}
    flat = engine.FlatParams(m, shadow_dtype=shadow)
    opt = dict(lr=1e-2, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01, max_grad_norm=0.5)
    for t in range(1, 4):
        grads = {n: torch.randn_like(p) * 0.3 for n, p in ref_p.items()}
        for n, p in zip(flat.names, flat.params):
            p._goat_grad.copy_(grads[n])
        flat.adamw_step(grad_scale=1.0, **opt)
        assert float(flat.g.abs().max()) == 0.0   # zero_grad fused into the optimizer kernel
        norm, coef = O.clip_grad_norm(list(grads.values()), opt["max_grad_norm"])
        for n in ref_p:
            wd = 0.0 if any(nd in n for nd in engine.NO_DECAY) else opt["weight_decay"]
            O.adamw_step(ref_p[n], grads[n] * coef, ref_m[n], ref_v[n], t, opt["lr"], opt["betas"], opt["eps"], wd)
        torch.cuda.synchronize()
        assert abs(flat.grad_norm.item() - norm.item()) < 1e-5 * norm.item()
        for n, p in zip(flat.names, flat.params):
            assert rel(p.detach(), ref_p[n]) < 1e-5, (t, n)
            if shadow is not None:
                assert torch.equal(p._goat_shadow.cpu(), p.detach().cpu().to(shadow))


def test_attention_ignores_stale_tmem_columns():
    """Regression: TMEM columns beyond the score MMA's N keep whatever an earlier kernel left there (NaN after a GEMM
    on NaN inputs); the softmax sweep must never read them (it once zeroed whole rows through a NaN row sum)."""
    from vln_goat_b200 import ops
    torch.manual_seed(4)
    nan = torch.full((1024, 256), float("nan"), device=DEV, dtype=torch.bfloat16)
    q, k, v, kmask, bias, w = (t.to(DEV) if t is not None else None
                               for t in _attn_case(16, 12, 12, torch.bfloat16, True, seed=4))
    ref, _ = ops.attn_fwd(q, k, v, 12, kmask, bias, force_simt=True)
    for _ in range(3):
        ops.gemm(nan, nan, out_dtype=torch.float32)       # poisons the accumulators' TMEM columns on every SM
        o, _ = ops.attn_fwd(q, k, v, 12, kmask, bias)
        assert torch.isfinite(o.float()).all()
        assert rel(o, ref) < OUT_TOL[torch.bfloat16]
