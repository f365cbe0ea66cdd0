"""On-device diagnostics for the primitive kernels (prints, never asserts).

    python tests/gpu_diag.py            # runs every group, each in its own subprocess with a timeout
    python tests/gpu_diag.py gemm_kk    # one group in-process

A protocol bug in the tcgen05 kernel traps after ~2 s (see mbar_wait) and poisons the CUDA
context, so groups are isolated in subprocesses.  Used while bringing kernels up; the asserting
tests live in tests/test_gpu_*.py.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _summ(name, got, ref, grid=4):
    import torch
    got = got.double().cpu()
    ref = ref.double().cpu()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-30
    bad = (err > 1e-2 * scale).double().mean().item()
    idx = err.argmax().item()
    pos = (idx // ref.shape[1], idx % ref.shape[1]) if ref.dim() == 2 else idx
    print("  %-44s max_abs_err %.3e (ref max %.3e) frac_bad %.4f argmax %s nan %d" % (
        name, err.max().item(), scale, bad, pos, int(torch.isnan(got).sum().item())), flush=True)
    if bad > 0 and ref.dim() == 2:
        M, N = ref.shape
        rows = []
        for i in range(grid):
            r = []
            for j in range(grid):
                blk = err[i * M // grid:(i + 1) * M // grid, j * N // grid:(j + 1) * N // grid]
                r.append("%.1e" % (blk.max().item() if blk.numel() else 0))
            rows.append(" ".join(r))
        print("    block max err grid:\n      " + "\n      ".join(rows), flush=True)


def gemm_group(a_mn, b_mn):
    import torch
    from vln_goat_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    shapes = [(128, 128, 64), (128, 128, 128), (256, 256, 768), (300, 768, 768), (2368, 2304, 768), (5120, 768, 3072),
              (77, 200, 72), (2368, 768, 2368)]
    for dtype in (torch.float16, torch.bfloat16):
        for (M, N, K) in shapes:
            A = (torch.randn(M, K, device=dev) * 0.5).to(dtype)
            B = (torch.randn(N, K, device=dev) * 0.5).to(dtype)
            ref = A.double() @ B.double().t()
            Ain = A.t().contiguous() if a_mn else A
            Bin = B.t().contiguous() if b_mn else B
            if (a_mn and M % 8) or (b_mn and N % 8):
                continue
            t0 = time.time()
            out = ops.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
            torch.cuda.synchronize()
            _summ("umma %s M%d N%d K%d (%.0f ms)" % (str(dtype)[6:], M, N, K, (time.time() - t0) * 1e3), out, ref)
            out_s = ops.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, force_simt=True)
            torch.cuda.synchronize()
            _summ("simt %s M%d N%d K%d" % (str(dtype)[6:], M, N, K), out_s, ref)


def gemm_epilogue():
    import torch
    from vln_goat_b200 import ops
    torch.manual_seed(1)
    dev = "cuda"
    M, N, K = 300, 768, 768
    for dtype in (torch.float16, torch.bfloat16):
        A = (torch.randn(M, K, device=dev) * 0.5).to(dtype)
        B = (torch.randn(N, K, device=dev) * 0.05).to(dtype)
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev)
        acc = A.double() @ B.double().t()
        # bias + residual, fp32 out + 16-bit copy
        out2 = torch.empty(M, N, device=dev, dtype=dtype)
        out = ops.gemm(A, B, bias=bias, res=res, out_dtype=torch.float32, out2=out2, alpha=0.5)
        ref = 0.5 * acc + bias.double() + res.double()
        _summ("bias+res alpha f32 %s" % dtype, out, ref)
        _summ("  out2 copy", out2, ref.to(dtype))
        # GELU with aux_out, 16-bit out
        z = torch.empty(M, N, device=dev, dtype=dtype)
        h = ops.gemm(A, B, bias=bias, act=ops.ACT_GELU, aux_out=z, out_dtype=dtype)
        zref = (acc + bias.double())
        _summ("gelu aux_out(z) %s" % dtype, z, zref)
        zr = z.double()
        _summ("gelu out %s" % dtype, h, zr * 0.5 * (1 + torch.erf(zr / 2 ** 0.5)))
        # DGELU
        g = ops.gemm(A, B, act=ops.ACT_DGELU, aux_in=z, out_dtype=dtype)
        cdf = 0.5 * (1 + torch.erf(zr / 2 ** 0.5))
        pdf = torch.exp(-0.5 * zr * zr) / (2 * 3.141592653589793) ** 0.5
        _summ("dgelu %s" % dtype, g, acc * (cdf + zr * pdf))
        # relu / drelu / tanh via SIMT and UMMA
        r = ops.gemm(A, B, bias=bias, act=ops.ACT_RELU, out_dtype=dtype)
        _summ("relu %s" % dtype, r, torch.relu(zref))
        dr = ops.gemm(A, B, act=ops.ACT_DRELU, aux_in=r, out_dtype=torch.float32)
        _summ("drelu %s" % dtype, dr, acc * (r.double() > 0))
        # dropout statistics
        d = ops.gemm(A, B, bias=bias, out_dtype=torch.float32, drop_p=0.1, drop_seed=123)
        keep = (d != 0).double().mean().item()
        d2 = ops.gemm(A, B, bias=bias, out_dtype=torch.float32, drop_p=0.1, drop_seed=123, force_simt=True)
        print("  dropout keep frac %.4f (expect 0.9); umma/simt masks equal: %s" % (
            keep, bool(((d != 0) == (d2 != 0)).all().item())), flush=True)
    # fp32 SIMT path
    A = torch.randn(M, K, device=dev)
    B = torch.randn(N, K, device=dev) * 0.05
    out = ops.gemm(A, B, bias=bias, res=res)
    _summ("fp32 simt bias+res", out, A.double() @ B.double().t() + bias.double() + res.double())
    out = ops.gemm(A.t().contiguous(), B.t().contiguous(), a_mn=True, b_mn=True)
    _summ("fp32 simt mn/mn", out, A.double() @ B.double().t())
    A7 = torch.randn(M, 7, device=dev)
    B7 = torch.randn(N, 7, device=dev)
    _summ("fp32 simt K=7", ops.gemm(A7, B7), A7.double() @ B7.double().t())
    W1 = torch.randn(1, K, device=dev)
    _summ("fp32 simt N=1", ops.gemm(A, W1), A.double() @ W1.double().t())


def attention():
    import torch
    from oracle import goat_oracle as O
    from vln_goat_b200 import ops
    torch.manual_seed(2)
    dev = "cuda"
    for dtype, tag in ((torch.float32, "f32"), (torch.float16, "f16"), (torch.bfloat16, "bf16")):
        for (B, Nq, Nk, sprel) in ((2, 36, 80, False), (3, 12, 12, True), (2, 80, 80, False), (1, 37, 512, False),
                                   (2, 130, 70, False), (1, 1, 1, False)):
            heads, H = 12, 768
            qkv_q = (torch.randn(B, Nq, H) * 1.0).to(dtype)
            kk = (torch.randn(B, Nk, H) * 1.0).to(dtype)
            vv = (torch.randn(B, Nk, H) * 1.0).to(dtype)
            lens = torch.randint(1, Nk + 1, (B,))
            lens[0] = Nk
            kmask = (1.0 - O.gen_seq_masks(lens, Nk).float()) * -10000.0
            bias = torch.randn(B, Nq, Nk) if sprel else None
            q64, k64, v64 = (t.double().requires_grad_(True) for t in (qkv_q, kk, vv))
            b64 = bias.double().requires_grad_(True) if sprel else None
            mask = kmask.double()[:, None, None, :]
            if sprel:
                mask = mask + b64[:, None]
            ref = O.attn_core(q64, k64, v64, mask, heads)
            w = torch.randn(B, Nq, H).to(dtype)
            (ref * w.double()).sum().backward()
            q, k, v = qkv_q.to(dev), kk.to(dev), vv.to(dev)
            o, lse = ops.attn_fwd(q, k, v, heads, kmask.to(dev), bias.to(dev) if sprel else None)
            _summ("attn fwd %s B%d Nq%d Nk%d sprel%d" % (tag, B, Nq, Nk, sprel), o.reshape(B * Nq, H), ref.reshape(B * Nq, H))
            dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
            db = ops.attn_bwd(w.to(dev), q, k, v, o, lse, heads, dq, dk, dv, kmask.to(dev),
                              bias.to(dev) if sprel else None, want_dbias=sprel)
            _summ("  dq", dq.reshape(B * Nq, H), q64.grad.reshape(B * Nq, H))
            _summ("  dk", dk.reshape(B * Nk, H), k64.grad.reshape(B * Nk, H))
            _summ("  dv", dv.reshape(B * Nk, H), v64.grad.reshape(B * Nk, H))
            if sprel:
                _summ("  dbias", db.reshape(B * Nq, Nk), b64.grad.reshape(B * Nq, Nk))
    # strided QKV buffer, -inf key padding (pano encoder style)
    B, N, H, heads = 2, 36, 768, 12
    qkv = torch.randn(B, N, 3 * H, device=dev)
    lens = torch.tensor([36, 29])
    km = torch.zeros(B, N).masked_fill(O.gen_seq_masks(lens, N).logical_not(), float("-inf"))
    o, lse = ops.attn_fwd(qkv[:, :, :H], qkv[:, :, H:2 * H], qkv[:, :, 2 * H:], heads, km.to(dev))
    c = qkv.cpu().double()
    ref = O.attn_core(c[:, :, :H], c[:, :, H:2 * H], c[:, :, 2 * H:], km.double()[:, None, None, :], heads)
    _summ("attn fwd f32 strided qkv, -inf pad", o.reshape(B * N, H), ref.reshape(B * N, H))


def layernorm():
    import torch
    from vln_goat_b200 import ops
    torch.manual_seed(3)
    dev = "cuda"
    for (M, H) in ((300, 768), (5120, 768), (7, 768), (64, 3072 // 3)):
        x = torch.randn(M, H, device=dev) * 2 + 0.3
        g = 1 + 0.1 * torch.randn(H, device=dev)
        b = 0.1 * torch.randn(H, device=dev)
        for eps in (1e-12, 1e-5):
            y32, y16, mean, rstd = ops.layernorm_fwd(x, g, b, eps, True, torch.float16)
            xd = x.double().cpu().requires_grad_(True)
            gd = g.double().cpu().requires_grad_(True)
            bd = b.double().cpu().requires_grad_(True)
            ref = torch.nn.functional.layer_norm(xd, (H,), gd, bd, eps)
            _summ("ln fwd M%d H%d eps%g" % (M, H, eps), y32, ref)
            _summ("  y16", y16, ref)
            dy = torch.randn(M, H, device=dev)
            dres = torch.randn(M, H, device=dev)
            (ref * dy.double().cpu()).sum().backward()
            dx32, dx16, dg, dbt, dcol = ops.layernorm_bwd(dy, x, g, mean, rstd, dres, True, torch.float16, want_colsum=True)
            _summ("  dx32(+dres)", dx32, xd.grad + dres.double().cpu())
            _summ("  dx16", dx16, xd.grad)
            _summ("  dgamma", dg[None], gd.grad[None])
            _summ("  dbeta", dbt[None], bd.grad[None])
            _summ("  dcolsum", dcol[None], dx16.double().cpu().sum(0)[None])
    x = torch.randn(1000, 2304, device=dev).half()
    _summ("colsum f16", ops.colsum(x)[None], x.double().cpu().sum(0)[None])
    xs = torch.randn(1000, 2304, device=dev)
    _summ("colsum f32 strided", ops.colsum(xs[:, 768:1536])[None], xs[:, 768:1536].double().cpu().sum(0)[None])
    _summ("cast f32->bf16", ops.cast(xs, torch.bfloat16), xs.bfloat16())


GROUPS = {
    "gemm_kk": lambda: gemm_group(False, False),
    "gemm_kmn": lambda: gemm_group(False, True),
    "gemm_mnmn": lambda: gemm_group(True, True),
    "gemm_mnk": lambda: gemm_group(True, False),
    "gemm_epilogue": gemm_epilogue,
    "attention": attention,
    "layernorm": layernorm,
}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        for g in sys.argv[1:]:
            print("== %s" % g, flush=True)
            GROUPS[g]()
    else:
        for g in GROUPS:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), g], timeout=240)
                print("== %s exit %d (%.0f s)" % (g, r.returncode, time.time() - t0), flush=True)
            except subprocess.TimeoutExpired:
                print("== %s TIMEOUT" % g, flush=True)
