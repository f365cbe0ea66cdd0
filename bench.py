#!/usr/bin/env python
"""bench.py -- R2R pretrain steps/sec on the GOAT cross-modal hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl goat|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload = "C2"): BASELINE.json configs[1], "full 9-layer GOAT cross-encoder fwd+bwd,
batch=64": LanguageEncoder (6 RobertaLayers, 80 tokens) + CrossmodalEncoder (3 BertCrossLayers, [stop]+36 view
tokens x 80 text tokens), hidden 768, dropout 0.1 as shipped, random-init weights, synthetic N(0,1) features.
One step = forward + backward + (N>1) NCCL gradient all-reduce + global-norm clip + AdamW, i.e. what
P/train_r2r_goat.py:301-366 does per batch.  Weak scaling: every rank runs its own batch of 64.

One JSON line on rank 0:
  value     device-resident steps/s (inputs already in HBM), N ranks x K steps / max-over-ranks CUDA-event time
  e2e       the same step driven from pinned HOST buffers: per step H2D of the batch + D2H of the loss
  roofline  the dominant kernel (the tcgen05 GEMM): algorithmic FLOPs of the step's GEMM launches / their
            CUDA-event time, re-timed live launch by launch after the timed region, vs MEASURED_PEAKS.json
  cpu_baseline  the CPU restatement of the reference path (oracle/, torch fp32, all host cores) on the same step
--impl reference times only that CPU path (the reference is pure PyTorch; /root/reference does not travel to the
GPU box, so the committed oracle port, pinned to the reference by tests/golden, stands in for it).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "R2R pretrain steps/sec (batch=64, 36 views×768, 80 tok) at 1/2/4/8 B200"
B, L, NQ, H = 64, 80, 37, 768
OPT = dict(lr=5e-5, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01, max_grad_norm=5.0)  # P/config/r2r_GOAT_pretrain.json


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="goat", choices=["goat", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic batch (same recipe on CPU and GPU arms)
# ------------------------------------------------------------------------------------------------
def make_batches(n, batch, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        txt = torch.randn(batch, L, H, generator=g)
        vp = torch.randn(batch, NQ, H, generator=g)
        vp[:, 0] = 0.0  # the [stop] token is a zero embedding, P/model/vilmodel_goat.py:379-388
        lens = torch.randint(L // 2, L + 1, (batch,), generator=g)
        lens[0] = L
        tm = torch.arange(L)[None, :] < lens[:, None]
        vm = torch.ones(batch, NQ, dtype=torch.bool)
        out.append((txt, tm, vp, vm))
    return out


# ------------------------------------------------------------------------------------------------
# CPU arm: oracle restatement of the reference path (checker code used here only as the timed baseline)
# ------------------------------------------------------------------------------------------------
def cpu_step_fn(batch, seed=0):
    import torch
    from oracle import goat_oracle as O
    torch.manual_seed(seed)
    P = {k: v.clone().requires_grad_(True) for k, v in O.seeded_params(O.c2_shapes(), seed=0).items()}
    M_ = {k: torch.zeros_like(v) for k, v in P.items()}
    V_ = {k: torch.zeros_like(v) for k, v in P.items()}
    state = {"t": 0}
    nodecay = ("bias", "LayerNorm.bias", "LayerNorm.weight")

    def step(txt, tm, vp, vm):
        for v in P.values():
            v.grad = None
        t, o = O.c2_forward(P, txt, tm, vp, vm)
        loss = O.c2_loss(t, o)
        loss.backward()
        state["t"] += 1
        with torch.no_grad():
            _, coef = O.clip_grad_norm([v.grad for v in P.values()], OPT["max_grad_norm"])
            for k, v in P.items():
                wd = 0.0 if any(nd in k for nd in nodecay) else OPT["weight_decay"]
                O.adamw_step(v, v.grad * coef, M_[k], V_[k], state["t"], OPT["lr"], OPT["betas"], OPT["eps"], wd)
        return float(loss)
    return step


def time_cpu(steps, warmup, sample_batch):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_step_fn(sample_batch)
    batches = make_batches(2, sample_batch, seed=123)
    for i in range(warmup):
        step(*batches[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        step(*batches[i % 2])
    dt = (time.perf_counter() - t0) / max(steps, 1)
    # one batch-64 step costs 64/sample_batch sample steps (every op is linear in the batch; the optimizer part
    # is batch independent and therefore over-counted in the reference's favour when sample_batch < 64)
    value = (sample_batch / float(B)) / dt
    return value, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 8
    steps = max(1, min(args.steps, 10))
    warm = max(1, min(args.warmup, 2))
    value, dt, cores = time_cpu(steps, warm, sample)
    n = max(1, args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": n, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3 * (B / sample), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: 6 RobertaLayer + 3 BertCrossLayer fwd+bwd+clip+AdamW, hidden 768, 80 tok x 37 view tokens",
                   "global_batch": B, "per_gpu_batch": B, "device": "host CPU (the reference path is pure PyTorch)"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": "%d timed steps on a %d-sample slice of the batch-64 step, scaled by %d/64; oracle/ "
                                   "restatement of the reference modules (pinned by tests/golden), torch fp32, %d threads"
                                   % (steps, sample, sample, cores)},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class Clocks(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        # "under load": samples above 60 % of the max draw seen
        if pw:
            thr = 0.6 * max(pw)
            sm_l = [s for s, p in zip(sm, pw) if p >= thr] or sm
        else:
            sm_l = sm
        sm_l.sort()
        return {"sm_mhz": sm_l[len(sm_l) // 2] if sm_l else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_goat(args):
    import torch
    import torch.distributed as dist
    from vln_goat_b200 import engine, ops, runtime, workloads
    from vln_goat_b200.config import GoatConfig

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl goat needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    cdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype]
    runtime.set_compute_dtype(cdt)

    torch.manual_seed(0)  # same initial weights on every rank, as DDP's broadcast would give
    model = workloads.C2CrossEncoder(GoatConfig()).to(dev).train()

    def loss_fn(txt, tm, vp, vm):
        t, v = model(txt, tm, vp, vm)
        return workloads.c2_loss(t, v)

    host = [tuple(t.pin_memory() for t in b) for b in make_batches(4, B, seed=1000 + rank)]
    devb = [tuple(t.to(dev) for t in b) for b in host]
    active = engine.active_parameters(model, loss_fn, devb[0])
    flat = engine.FlatParams(model, shadow_dtype=cdt if cdt != torch.float32 else None, only=active)
    ts = engine.TrainStep(flat, loss_fn, devb[0], use_graph=not args.no_graph, **OPT)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- device-resident: inputs already in HBM (rotating through 4 resident batches, D2D into the graph's inputs)
    def step_resident(i):
        ts.step(devb[i % len(devb)])

    for i in range(args.warmup):
        step_resident(i)
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    ms = timed(step_resident, args.steps)
    value = world * args.steps / (ms / 1e3)

    # ---- end to end: batch in pinned host memory -> H2D on a copy stream (overlapping the previous step) -> step
    #      -> loss D2H into pinned memory every step
    copy_stream = torch.cuda.Stream()
    stage = [tuple(torch.empty_like(t) for t in devb[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    losses = torch.zeros(args.steps + args.warmup + 1, dtype=torch.float32).pin_memory()
    state = {"n": 0}

    def prefetch(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            for d, h in zip(stage[s], host[i % len(host)]):
                d.copy_(h, non_blocking=True)
            ready[s].record(copy_stream)

    def step_e2e(i):
        s = i % 2
        cur = torch.cuda.current_stream()
        cur.wait_event(ready[s])
        ts.load_inputs(stage[s])
        consumed[s].record(cur)
        prefetch(i + 1)
        loss = ts.step()
        losses[state["n"]:state["n"] + 1].copy_(loss.reshape(1), non_blocking=True)
        state["n"] += 1

    for s in range(2):
        consumed[s].record(torch.cuda.current_stream())
    prefetch(0)
    for i in range(args.warmup):
        step_e2e(i)
    off = args.warmup
    ms_e2e = timed(lambda i: step_e2e(i + off), args.steps)
    e2e = world * args.steps / (ms_e2e / 1e3)
    clk = clocks.stop() if rank == 0 else None
    torch.cuda.synchronize()
    lv = losses[:state["n"]]
    if not bool(torch.isfinite(lv).all()):
        raise RuntimeError("non-finite loss in the timed run: %s" % lv.tolist())

    if rank == 0:
        roof = gemm_roofline(torch, ops, ts, cdt, ms / args.steps)
        roof_attn = attention_roofline(torch, ops, cdt, ms / args.steps)
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "C2: 6 RobertaLayer + 3 BertCrossLayer fwd+bwd+clip+AdamW, hidden 768, 80 tok x 37 view tokens",
                       "global_batch": B * world, "per_gpu_batch": B, "parallelism": "dp%d" % world,
                       "dropout": 0.1, "cuda_graph": not args.no_graph,
                       "l2": "per-step working set (%.0f MB params/grads/moments + activations) exceeds the 126 MB L2; "
                             "4 rotating input batches" % (flat.numel * 4 * 4 / 1e6),
                       "loss_first_last": [float(lv[0]), float(lv[-1])]},
            "clocks": clk,
            "e2e": {"value": e2e, "unit": "steps/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": ts.launches_per_step * args.steps,
            "gpu_launches_per_step": ts.launches_per_step,
            "roofline": roof,
            "roofline_attention": roof_attn,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, dt, cores = time_cpu(2, 1, 16)
            line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
                                    "sample": "2 timed steps on a 16-sample slice of the batch-64 step, scaled by 16/64; "
                                              "oracle/ restatement of the reference modules, torch fp32, %d threads" % cores}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def gemm_roofline(torch, ops, ts, cdt, step_ms):
    """Re-time every GEMM launch of one step (shapes recorded from the real step) with CUDA events."""
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained")
    which = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
    if peak is None:
        peak, which = 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"
    rec = []
    ops.GEMM_LOG = rec
    ts._fwd_bwd()  # eager pass: records (M,N,K,a_mn,b_mn) per launch and warms every shape
    ops.GEMM_LOG = None
    torch.cuda.synchronize()
    ts.flat.g.zero_()
    uniq = {}
    for r in rec:
        uniq[r] = uniq.get(r, 0) + 1
    dev = torch.device("cuda", torch.cuda.current_device())
    tot_ms, tot_flop = 0.0, 0.0
    umma_ms, umma_flop, n_umma = 0.0, 0.0, 0
    for (M, N, K, a_mn, b_mn, dt_, simt, acc_, act, has_bias, has_res, out32, has_drop, has_out2), cnt in uniq.items():
        dt_t = {0: torch.float32, 1: torch.float16, 2: torch.bfloat16}[dt_]
        A = torch.randn((K, M) if a_mn else (M, K), device=dev).to(dt_t)
        Bm = torch.randn((K, N) if b_mn else (N, K), device=dev).to(dt_t)
        out = (torch.zeros if acc_ else torch.empty)((M, N), device=dev, dtype=torch.float32 if out32 else dt_t)
        kw = dict(a_mn=bool(a_mn), b_mn=bool(b_mn), out=out, accumulate=bool(acc_), act=act)
        if has_bias:
            kw["bias"] = torch.zeros(N, device=dev)
        if has_res:
            kw["res"] = torch.randn(M, N, device=dev)
        if has_drop:
            kw["drop_p"], kw["drop_seed"] = 0.1, 11
        if has_out2:
            kw["out2"] = torch.empty((M, N), device=dev, dtype=dt_t)
        if act == ops.ACT_GELU:
            kw["aux_out"] = torch.empty((M, N), device=dev, dtype=dt_t)
        if act in (ops.ACT_DGELU, ops.ACT_DRELU):
            kw["aux_in"] = torch.randn(M, N, device=dev).to(dt_t)

        def launch():
            ops.gemm(A, Bm, **kw)
        for _ in range(3):
            launch()
        reps = 20
        # `reps` launches captured in one CUDA graph: the event pair sees device time, not the ~20 us of Python /
        # ctypes / tensor-map encoding per call (the step itself is replayed as a graph, too)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                launch()
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        del g
        fl = 2.0 * M * N * K
        tot_ms += ms * cnt
        tot_flop += fl * cnt
        if not simt:
            umma_ms += ms * cnt
            umma_flop += fl * cnt
            n_umma += cnt
    achieved = umma_flop / (umma_ms * 1e-3) / 1e12 if umma_ms > 0 else 0.0
    # DRAM bytes moved by the same launches, from the committed ncu pass over one step (profiles/r01d_*)
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r01d_gemm_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("gemm_launches") == n_umma:
            traffic = tj["gemm_dram_bytes_per_step"]
            traffic_src = ("profiles/r01d_launches_one_step.csv: dram__bytes_read.sum + dram__bytes_write.sum summed over the "
                           "%d GEMM launches of one step (ncu, cold cache)" % n_umma)
    except Exception:
        pass
    return {"bound": "tensor", "kernel": "gemm_umma2_kernel / gemm_umma_kernel (tcgen05.mma cta_group::2 + TMA; all %d tensor-core GEMM launches of one step, "
                      "each shape re-timed as 20 graph-captured launches)" % n_umma,
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_unit": "bytes per step over the same launches", "traffic_source": traffic_src,
            "peak_source": which, "flop_per_step": umma_flop, "gemm_ms_per_step": umma_ms,
            "gemm_share_of_step": umma_ms / step_ms if step_ms else None,
            "step_tflops": umma_flop / (step_ms * 1e-3) / 1e12 if step_ms else None}


def attention_roofline(torch, ops, cdt, step_ms):
    """The other regime of SURVEY.md 8d: the attention core (QK^T, softmax, PV and its backward) is HBM / latency
    bound.  Algorithmic bytes of the step's 24 attention launches / their device time (20 graph-captured launches each)."""
    if cdt == torch.float32:
        return None
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = peaks.get("hbm_gbs")
    which = "MEASURED_PEAKS.json hbm_gbs"
    if peak is None:
        peak, which = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    dev = torch.device("cuda", torch.cuda.current_device())
    heads, p_drop = 12, 0.1

    def graph_ms(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    tot_ms, tot_bytes, n = 0.0, 0.0, 0
    # (count per step, Nq, Nk): 6 text self-attentions, 3 view self-attentions, 3 view -> text cross-attentions
    for cnt, Nq, Nk in ((6, L, L), (3, NQ, NQ), (3, NQ, L)):
        q = torch.randn(B, Nq, H, device=dev).to(cdt)
        k = torch.randn(B, Nk, H, device=dev).to(cdt)
        v = torch.randn(B, Nk, H, device=dev).to(cdt)
        km = torch.zeros(B, Nk, device=dev)
        w = torch.randn(B, Nq, H, device=dev).to(cdt)
        o, lse = ops.attn_fwd(q, k, v, heads, km, drop_p=p_drop, drop_seed=3)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        f_ms = graph_ms(lambda: ops.attn_fwd(q, k, v, heads, km, drop_p=p_drop, drop_seed=3))
        b_ms = graph_ms(lambda: ops.attn_bwd(w, q, k, v, o, lse, heads, dq, dk, dv, km, drop_p=p_drop, drop_seed=3))
        f_bytes = 2.0 * B * H * (2 * Nq + 2 * Nk) + 4.0 * B * Nk          # q, o + k, v (16-bit) + key mask
        b_bytes = 2.0 * B * H * (4 * Nq + 4 * Nk) + 4.0 * B * Nk          # q, o, dO, dQ + k, v, dK, dV
        tot_ms += cnt * (f_ms + b_ms)
        tot_bytes += cnt * (f_bytes + b_bytes)
        n += 2 * cnt
    achieved = tot_bytes / (tot_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "attn_fwd_pipe_kernel + attn_bwd_pipe_kernel (all %d attention launches of one step, "
                                      "dropout 0.1, each shape re-timed as 20 graph-captured launches)" % n,
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "peak_source": which, "bytes_per_step": tot_bytes, "attention_ms_per_step": tot_ms,
            "attention_share_of_step": tot_ms / step_ms if step_ms else None}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_goat(args)


if __name__ == "__main__":
    main()
