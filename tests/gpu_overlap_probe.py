"""Measurement script (not a test): can the gradient exchange hide under the backward GEMMs on this box?

Run under torchrun with >= 2 ranks:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        tests/gpu_overlap_probe.py

Times, with CUDA events on rank 0 (max over ranks where it matters):
  1. the collectives of the sharded optimizer step alone on a gradient-sized buffer (reduce-scatter, all-gather,
     all-reduce, in 1 and in 8 pieces),
  2. a loop of the step's dominant GEMM alone,
  3. the same loop while an all-reduce / reduce-scatter of the buffer runs on NCCL's stream (eager, two streams),
  4. 3 captured as ONE CUDA graph (the way a captured training step would carry it) and replayed.
The persistent tcgen05 GEMM schedules its tiles statically over 74 CTA pairs: if the NCCL kernel's CTAs cannot share an
SM with a GEMM CTA (200 KB of shared memory), every GEMM launched meanwhile takes a second wave.  This script measures it.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from vln_goat_b200 import ops
    say = (lambda *a: print(*a, flush=True)) if rank == 0 else (lambda *a: None)

    n = 162631312 // (8 * world) * (8 * world)
    g = torch.randn(n, device=dev) * 1e-3
    shard = torch.empty(n // world, device=dev)
    x16 = torch.empty(n, device=dev, dtype=torch.float16)
    x16s = torch.empty(n // world, device=dev, dtype=torch.float16)

    say("== world %d, %d fp32 elements (%.0f MB)" % (world, n, n * 4 / 1e6))
    t_rs = timed(lambda: dist.reduce_scatter_tensor(shard, g))
    t_ag32 = timed(lambda: dist.all_gather_into_tensor(g, g[rank * (n // world):(rank + 1) * (n // world)]))
    t_ag16 = timed(lambda: dist.all_gather_into_tensor(x16, x16s))
    g.normal_().mul_(1e-3)
    t_ar = timed(lambda: dist.all_reduce(g))
    g.normal_().mul_(1e-3)
    pieces = [g[i * (n // 8):(i + 1) * (n // 8)] for i in range(8)]
    t_ar8 = timed(lambda: [dist.all_reduce(p) for p in pieces])
    say("reduce-scatter fp32 %.3f ms | all-gather fp32 in place %.3f ms | all-gather 16-bit %.3f ms (x2 for hi+lo) | "
        "all-reduce fp32 %.3f ms | all-reduce in 8 pieces %.3f ms" % (t_rs, t_ag32, t_ag16, t_ar, t_ar8))

    # the step's dominant backward GEMMs: dgrad (b_mn) and wgrad (a_mn, b_mn, fp32 accumulate)
    M, N, K = 5120, 3072, 768
    A = (torch.randn(M, K, device=dev) * 0.05).half()
    Wt = (torch.randn(K, N, device=dev) * 0.05).half()
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    dY = (torch.randn(M, N, device=dev) * 0.05).half()
    dW = torch.zeros(N, K, device=dev)
    REP = 60

    def gemms():
        for _ in range(REP):
            ops.gemm(A, Wt, b_mn=True, out=out)
            ops.gemm(dY, A, a_mn=True, b_mn=True, out=dW, accumulate=True)
    t_g = timed(gemms, iters=5)
    say("GEMM loop alone (%d dgrad + %d wgrad): %.3f ms" % (REP, REP, t_g))

    for name, coll in (("all-reduce (8 pieces)", lambda: [dist.all_reduce(p, async_op=True) for p in pieces]),
                       ("reduce-scatter", lambda: [dist.reduce_scatter_tensor(shard, g, async_op=True)])):
        def both():
            works = coll()
            gemms()
            for w in works:
                w.wait()
        g.normal_().mul_(1e-3)
        t_b = timed(both, iters=5)
        say("GEMM loop + %s on NCCL's stream, eager: %.3f ms  (serial would be %.3f)" %
            (name, t_b, t_g + (t_ar8 if "all-reduce" in name else t_rs)))

    # one CUDA graph holding both branches
    g.normal_().mul_(1e-3)
    ref = g.clone()
    dist.all_reduce(ref)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    ok = True
    try:
        with torch.cuda.stream(s):
            graph = torch.cuda.CUDAGraph()
            keep = g.clone()
            with torch.cuda.graph(graph, stream=s):
                works = [dist.all_reduce(p, async_op=True) for p in pieces]
                gemms()
                for w in works:
                    w.wait()
        torch.cuda.current_stream().wait_stream(s)
        g.copy_(keep)
        graph.replay()
        torch.cuda.synchronize()
        err = (g - ref).abs().max().item()
        say("captured graph (GEMM loop + all-reduce): replay result max |diff| vs eager all-reduce = %.3e" % err)

        def replay():
            graph.replay()
        t_cap = timed(replay, iters=5)
        say("captured graph replay: %.3f ms (GEMM alone %.3f, all-reduce alone %.3f)" % (t_cap, t_g, t_ar8))
    except Exception as e:      # report, do not hang the other rank
        ok = False
        say("capturing NCCL inside a CUDA graph failed: %r" % (e,))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
