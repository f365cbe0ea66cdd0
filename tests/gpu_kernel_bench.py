"""Scratch micro-benchmark (not a test): the step's dominant kernel shapes, timed alone with CUDA events.
Run under ncu for the per-kernel captures in profiles/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vln_goat_b200 import ops

dev = "cuda"
dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[os.environ.get("DT", "fp16")]
reps = int(os.environ.get("REPS", "20"))
warm = int(os.environ.get("WARM", "3"))


use_graph = os.environ.get("GRAPH", "1") == "1" and warm > 0


only = [x for x in os.environ.get("ONLY", "").split(",") if x]


def timeit(name, fn, flops=None, bytes_=None):
    if only and not any(x in name for x in only):
        return
    """Device time per call.  The launches are captured in a CUDA graph (reps calls per replay) so that the
    Python / ctypes / tensor-map-encode host cost of each call (tens of microseconds) does not hide the kernel time."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if use_graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
    else:
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    extra = ""
    if flops:
        extra += "  %.0f TFLOP/s" % (flops / us / 1e6)
    if bytes_:
        extra += "  %.0f GB/s" % (bytes_ / us / 1e3)
    print("%-52s %8.1f us%s" % (name, us, extra), flush=True)


M = 5120
x = torch.randn(M, 768, device=dev).to(dt)
x32 = torch.randn(M, 768, device=dev)
h = torch.randn(M, 3072, device=dev).to(dt)
def _w(n, k):
    """a weight operand; SPLIT=1: a [hi | lo] split weight (forward GEMMs then run the K loop twice, runtime.py)"""
    w32 = torch.randn(n, k, device=dev) * 0.02
    if os.environ.get("SPLIT", "0") == "1":
        from vln_goat_b200 import runtime
        return runtime._cast_split(w32, dt)
    return w32.to(dt)


Wqkv = _w(2304, 768)
Wo = _w(768, 768)
W1 = _w(3072, 768)
W2 = _w(768, 3072)
b3 = torch.zeros(2304, device=dev); b1 = torch.zeros(3072, device=dev); b0 = torch.zeros(768, device=dev)
z = torch.empty(M, 3072, device=dev, dtype=dt)
dy = torch.randn(M, 768, device=dev).to(dt)
dqkv = torch.randn(M, 2304, device=dev).to(dt)
dz = torch.randn(M, 3072, device=dev).to(dt)
g768 = torch.zeros(768, 768, device=dev); g2304 = torch.zeros(2304, 768, device=dev)
g3072 = torch.zeros(3072, 768, device=dev); g768x = torch.zeros(768, 3072, device=dev)

which = os.environ.get("WHICH", "all")
if which in ("all", "gemm"):
    timeit("fwd qkv   5120x2304x768 +bias", lambda: ops.gemm(x, Wqkv, bias=b3), 2 * M * 2304 * 768)
    timeit("fwd out   5120x768x768 +bias+res f32out", lambda: ops.gemm(x, Wo, bias=b0, res=x32, out_dtype=torch.float32), 2 * M * 768 * 768)
    timeit("fwd ffn1  5120x3072x768 +bias+gelu+aux", lambda: ops.gemm(x, W1, bias=b1, act=ops.ACT_GELU, aux_out=z), 2 * M * 3072 * 768)
    timeit("fwd ffn2  5120x768x3072 +bias+res+drop f32out", lambda: ops.gemm(h, W2, bias=b0, res=x32, out_dtype=torch.float32, drop_p=0.1, drop_seed=5), 2 * M * 768 * 3072)
    timeit("dgrad ffn2 5120x3072x768 dgelu", lambda: ops.gemm(dy, W2, b_mn=True, act=ops.ACT_DGELU, aux_in=z), 2 * M * 3072 * 768)
    timeit("dgrad ffn1 5120x768x3072 +res f32out", lambda: ops.gemm(dz, W1, b_mn=True, res=x32, out_dtype=torch.float32), 2 * M * 768 * 3072)
    timeit("dgrad qkv 5120x768x2304 +res f32out", lambda: ops.gemm(dqkv, Wqkv, b_mn=True, res=x32, out_dtype=torch.float32), 2 * M * 768 * 2304)
    timeit("dgrad out 5120x768x768", lambda: ops.gemm(dy, Wo, b_mn=True), 2 * M * 768 * 768)
    timeit("wgrad out 768x768x5120 acc", lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, out=g768, accumulate=True), 2 * M * 768 * 768)
    timeit("wgrad qkv 2304x768x5120 acc", lambda: ops.gemm(dqkv, x, a_mn=True, b_mn=True, out=g2304, accumulate=True), 2 * M * 2304 * 768)
    timeit("wgrad ffn1 3072x768x5120 acc", lambda: ops.gemm(dz, x, a_mn=True, b_mn=True, out=g3072, accumulate=True), 2 * M * 3072 * 768)
    timeit("wgrad ffn2 768x3072x5120 acc", lambda: ops.gemm(dy, h, a_mn=True, b_mn=True, out=g768x, accumulate=True), 2 * M * 3072 * 768)
if which in ("all", "gemm", "gemmvp"):
    Mv = 2368
    xv = torch.randn(Mv, 768, device=dev).to(dt); xv32 = torch.randn(Mv, 768, device=dev)
    hv = torch.randn(Mv, 3072, device=dev).to(dt); zv = torch.empty(Mv, 3072, device=dev, dtype=dt)
    dqkvv = torch.randn(Mv, 2304, device=dev).to(dt); dzv = torch.randn(Mv, 3072, device=dev).to(dt)
    Wkv = (torch.randn(1536, 768, device=dev) * 0.02).to(dt); b15 = torch.zeros(1536, device=dev)
    dkv = torch.randn(M, 1536, device=dev).to(dt); g1536 = torch.zeros(1536, 768, device=dev)
    timeit("fwd kv    5120x1536x768 +bias", lambda: ops.gemm(x, Wkv, bias=b15), 2 * M * 1536 * 768)
    timeit("dgrad kv  5120x768x1536 f32out", lambda: ops.gemm(dkv, Wkv, b_mn=True, out_dtype=torch.float32), 2 * M * 1536 * 768)
    timeit("wgrad kv  1536x768x5120 acc", lambda: ops.gemm(dkv, x, a_mn=True, b_mn=True, out=g1536, accumulate=True), 2 * M * 1536 * 768)
    timeit("vp fwd qkv 2368x2304x768 +bias", lambda: ops.gemm(xv, Wqkv, bias=b3), 2 * Mv * 2304 * 768)
    timeit("vp fwd out 2368x768x768 +bias+res+drop f32", lambda: ops.gemm(xv, Wo, bias=b0, res=xv32, out_dtype=torch.float32, drop_p=0.1, drop_seed=5), 2 * Mv * 768 * 768)
    timeit("vp fwd ffn1 2368x3072x768 gelu", lambda: ops.gemm(xv, W1, bias=b1, act=ops.ACT_GELU, aux_out=zv), 2 * Mv * 3072 * 768)
    timeit("vp fwd ffn2 2368x768x3072 +res+drop f32", lambda: ops.gemm(hv, W2, bias=b0, res=xv32, out_dtype=torch.float32, drop_p=0.1, drop_seed=5), 2 * Mv * 768 * 3072)
    timeit("vp dgrad ffn2 2368x3072x768 dgelu", lambda: ops.gemm(xv, W2, b_mn=True, act=ops.ACT_DGELU, aux_in=zv), 2 * Mv * 3072 * 768)
    timeit("vp dgrad ffn1 2368x768x3072 +res f32", lambda: ops.gemm(dzv, W1, b_mn=True, res=xv32, out_dtype=torch.float32), 2 * Mv * 768 * 3072)
    timeit("vp dgrad out 2368x768x768", lambda: ops.gemm(xv, Wo, b_mn=True), 2 * Mv * 768 * 768)
    timeit("vp wgrad out 768x768x2368 acc", lambda: ops.gemm(xv, xv, a_mn=True, b_mn=True, out=g768, accumulate=True), 2 * Mv * 768 * 768)
    timeit("vp wgrad ffn1 3072x768x2368 acc", lambda: ops.gemm(dzv, xv, a_mn=True, b_mn=True, out=g3072, accumulate=True), 2 * Mv * 3072 * 768)
if which in ("all", "attn"):
    B, L, Nq = 64, 80, 37
    qkv = torch.randn(B, L, 2304, device=dev).to(dt)
    q, k, v = qkv[:, :, :768], qkv[:, :, 768:1536], qkv[:, :, 1536:]
    km = torch.zeros(B, L, device=dev)
    w = torch.randn(B, L, 768, device=dev).to(dt)
    dq = torch.empty_like(qkv)
    for p in (0.0, 0.1):
        o, lse = ops.attn_fwd(q, k, v, 12, km, drop_p=p, drop_seed=3)
        nbytes = 4 * B * L * 768 * 2
        timeit("attn fwd self B64 L80 p=%.1f" % p, lambda: ops.attn_fwd(q, k, v, 12, km, drop_p=p, drop_seed=3), 4 * B * 12 * L * L * 64, nbytes)
        timeit("attn bwd self B64 L80 p=%.1f" % p, lambda: ops.attn_bwd(w, q, k, v, o, lse, 12, dq[:, :, :768], dq[:, :, 768:1536], dq[:, :, 1536:], km, drop_p=p, drop_seed=3), 10 * B * 12 * L * L * 64, 2 * nbytes)
    qc = torch.randn(B, Nq, 768, device=dev).to(dt)
    kv = torch.randn(B, L, 1536, device=dev).to(dt)
    timeit("attn fwd cross B64 Nq37 Nk80", lambda: ops.attn_fwd(qc, kv[:, :, :768], kv[:, :, 768:], 12, km))
if which in ("all", "attnlong"):
    # BASELINE.json configs[4]: text self-attention of the 512-token stress (query-tiled kernels of attention_tc.cu)
    B, L = 32, 512
    qkv = torch.randn(B, L, 2304, device=dev).to(dt)
    q, k, v = qkv[:, :, :768], qkv[:, :, 768:1536], qkv[:, :, 1536:]
    km = torch.zeros(B, L, device=dev)
    w = torch.randn(B, L, 768, device=dev).to(dt)
    dq = torch.empty_like(qkv)
    o, lse = ops.attn_fwd(q, k, v, 12, km, drop_p=0.1, drop_seed=3)
    nbytes = 4 * B * L * 768 * 2
    timeit("attn fwd self B32 L512 p=0.1", lambda: ops.attn_fwd(q, k, v, 12, km, drop_p=0.1, drop_seed=3), 4 * B * 12 * L * L * 64, nbytes)
    timeit("attn bwd self B32 L512 p=0.1", lambda: ops.attn_bwd(w, q, k, v, o, lse, 12, dq[:, :, :768], dq[:, :, 768:1536], dq[:, :, 1536:], km, drop_p=0.1, drop_seed=3), 10 * B * 12 * L * L * 64, 2 * nbytes)
if which in ("all", "misc"):
    g = torch.ones(768, device=dev); bt = torch.zeros(768, device=dev)
    y32, y16, mean, rstd = ops.layernorm_fwd(x32, g, bt, 1e-12, True, dt)
    timeit("ln fwd 5120x768", lambda: ops.layernorm_fwd(x32, g, bt, 1e-12, True, dt), bytes_=M * 768 * 10)
    timeit("ln bwd 5120x768", lambda: ops.layernorm_bwd(x32, x32, g, mean, rstd, None, True, dt, 0.1, 3, None, want_colsum=True), bytes_=M * 768 * 14)
    gacc = torch.zeros(768, device=dev); bacc = torch.zeros(768, device=dev); cacc = torch.zeros(768, device=dev)
    timeit("ln bwd acc 5120x768", lambda: ops.layernorm_bwd(x32, x32, g, mean, rstd, None, True, dt, 0.1, 3, None, want_colsum=True, dgamma_out=gacc, dbeta_out=bacc, dcol_out=cacc, accumulate=True), bytes_=M * 768 * 14)
    timeit("colsum 5120x2304", lambda: ops.colsum(dqkv), bytes_=M * 2304 * 2)
    timeit("colsum 5120x3072", lambda: ops.colsum(dz), bytes_=M * 3072 * 2)
    c23 = torch.zeros(2304, device=dev); c30 = torch.zeros(3072, device=dev)
    timeit("colsum acc 5120x2304", lambda: ops.colsum(dqkv, out=c23, accumulate=True), bytes_=M * 2304 * 2)
    timeit("colsum acc 5120x3072", lambda: ops.colsum(dz, out=c30, accumulate=True), bytes_=M * 3072 * 2)
