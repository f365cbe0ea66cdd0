#!/usr/bin/env python
"""bench.py -- R2R pretrain steps/sec on the GOAT cross-modal hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl goat|reference] [--dtype fp16|bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload = "C3-step"): ONE optimizer step of the full pretraining model -- what the reference's loop
does per batch (P/train_r2r_goat.py:301-366): ``GlocalTextPathCMTPreTraining(batch, task, compute_loss=True)`` (208 M
parameters: RoBERTa embeddings, 6-layer language encoder, panorama embeddings + 2-layer pano encoder + adaptive fusion,
global-map and local 3-layer cross-modal encoders, MLM / SAP / CFP heads), ``loss.mean().backward()``, (N>1) gradient
exchange, global-norm clip, AdamW.  Tasks alternate MLM, SAP, CFP (round-robin over steps); batch = 64 samples per
GPU (weak scaling), 36 views x 768, 80 tokens, trajectories of 1-5 panoramas, dropout 0.1 as shipped, random-init
weights, synthetic features with the collate layout of P/data/tasks.py.

One JSON line on rank 0:
  value       device-resident steps/s (prepared batches already in HBM): N ranks x K steps / max-over-ranks CUDA-event time
  e2e         the same steps driven from HOST batch dicts: per step the host index builders + padding
              (batching.prepare_pretrain), H2D of the step's inputs from pinned memory, the step, D2H of the loss
  parity_max_err  max error of the benchmarked model / dtype / batch size against the CPU oracle (logits, CFP embeddings,
              MLM scores, losses), relative to max(1, max|ref|) -- the north-star tolerance is 1e-3 (16-bit) / 1e-5 (fp32)
  roofline    the dominant kernel (the tcgen05 GEMM): algorithmic FLOPs of the GEMM launches of one MLM+SAP+CFP round /
              their CUDA-event time, re-timed live launch by launch, vs MEASURED_PEAKS.json (burst: timed alone)
  roofline_attention  the attention core launches of the same round vs the HBM copy roofline
  sustained   the same device-resident loop run for >= 2 s (clocks settle well below the burst the K steps see)
  e2e_feature_bank  the e2e loop with the view features in a GPU-resident 16-bit bank (batches carry one int per panorama)
  cpu_baseline        the CPU restatement of the same step (oracle/, torch fp32, all host cores), one step per task
  eager_b200_baseline the same restatement run eagerly on the B200 (fp32 and autocast fp16): the honest denominator
  c2          second line: BASELINE.json configs[1] (9-layer cross-encoder slice, batch 64) on the same kernels
  c4          BASELINE.json configs[3]: 16-episode x 15-step fine-tune rollout (BACL + FACL), eager vs one captured graph
  c5          BASELINE.json configs[4]: the same slice with 512-token instructions, batch 32, + the long-text attention core
--impl reference times only the CPU path (the reference is pure PyTorch; /root/reference does not travel to the GPU
box, so the committed oracle port, pinned to the reference by tests/golden, stands in for it).
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
TASKS = ("mlm", "sap", "cfp")
N_BATCHES = 4            # distinct host batches per task and rank, cycled
OPT = dict(lr=5e-5, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01, max_grad_norm=5.0)  # P/config/r2r_GOAT_pretrain.json
WORKLOAD = ("C3-step: full GlocalTextPathCMTPreTraining (208 M params) fwd+bwd+clip+AdamW, tasks MLM/SAP/CFP round-robin, "
            "batch 64 per GPU, 36 views x 768, 80 tokens, 1-5 panoramas per trajectory")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=180)
    ap.add_argument("--warmup", type=int, default=9)
    ap.add_argument("--impl", default="goat", choices=["goat", "reference"])
    ap.add_argument("--dtype", default="fp16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-weight-split", action="store_true",
                    help="plain 16-bit weight operands in the forward GEMMs (faster, ~1.6x the forward error)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip parity / rooflines / baselines / C2 (profiling runs)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="STRONG scaling: fix the global batch (SURVEY C3: 256) and give each rank global/N samples; a step is "
                         "then one optimizer step over the global batch and `value` counts those (default: 64 per GPU, weak)")
    ap.add_argument("--selfcheck", action="store_true", help="(default at N>1; kept for old command lines)")
    ap.add_argument("--no-selfcheck", action="store_true",
                    help="N>1: skip the check of bit-identical parameters / equal forward outputs across ranks after the "
                         "timed data-parallel steps (runs outside the timed region)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------------
# CPU arm: oracle restatement of the reference step (checker code used here only as the timed baseline)
# ------------------------------------------------------------------------------------------------
def oracle_params(device="cpu"):
    """seeded parameters of the full pretraining model under the reference's state_dict names"""
    import torch
    from oracle import goat_oracle as O
    from vln_goat_b200 import pretrain_model
    from vln_goat_b200.config import GoatConfig
    with torch.device("meta"):
        m = pretrain_model.GlocalTextPathCMTPreTraining(GoatConfig(pretrain_tasks=TASKS))
    shapes = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    P = O.seeded_params(shapes, seed=0)
    P["bert.embeddings.word_embeddings.weight"] = P["mlm_head.predictions.decoder.weight"]   # tied
    return {k: v.to(device) for k, v in P.items()}


def oracle_step_fn(device="cpu", autocast=False):
    """-> step(batch, task): forward + backward + clip + per-tensor AdamW of the oracle model (what the reference loop
    does with its own modules, P/train_r2r_goat.py:301-366 + P/optim/adamw.py)."""
    import torch
    from oracle import goat_oracle as O
    from oracle import goat_pretrain_oracle as PO
    P = {k: v.clone().requires_grad_(True) for k, v in oracle_params(device).items() if "decoder.weight" not in k}
    P["mlm_head.predictions.decoder.weight"] = P["bert.embeddings.word_embeddings.weight"]
    leaves = {k: v for k, v in P.items() if "decoder.weight" not in k}
    M_ = {k: torch.zeros_like(v) for k, v in leaves.items()}
    V_ = {k: torch.zeros_like(v) for k, v in leaves.items()}
    state = {"t": 0}
    nodecay = ("bias", "LayerNorm.bias", "LayerNorm.weight")
    scaler = torch.amp.GradScaler("cuda") if autocast else None

    def step(batch, task):
        for v in leaves.values():
            v.grad = None
        if autocast:
            with torch.autocast("cuda", dtype=torch.float16):
                loss = PO.scalar_loss(P, batch, task)
            scaler.scale(loss).backward()
            inv = 1.0 / scaler.get_scale()
        else:
            loss = PO.scalar_loss(P, batch, task)
            loss.backward()
            inv = 1.0
        state["t"] += 1
        with torch.no_grad():
            used = {k: v for k, v in leaves.items() if v.grad is not None}
            _, coef = O.clip_grad_norm([v.grad * inv for v in used.values()], OPT["max_grad_norm"])
            for k, v in used.items():
                wd = 0.0 if any(nd in k for nd in nodecay) else OPT["weight_decay"]
                O.adamw_step(v, v.grad * (inv * coef), M_[k], V_[k], state["t"], OPT["lr"], OPT["betas"], OPT["eps"], wd)
        return float(loss.detach())
    return step


def time_cpu(steps, warmup, budget_s=240.0):
    """K timed steps (round-robin tasks) of the oracle step at the full batch 64 on all host cores."""
    import torch
    from vln_goat_b200 import workloads
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = oracle_step_fn("cpu")
    batches = [workloads.synthetic_pretrain_batch(B, L, seed=123 + i) for i in range(2)]
    t_w = time.perf_counter()
    for i in range(warmup):
        step(batches[i % 2], TASKS[i % 3])
    est = (time.perf_counter() - t_w) / max(warmup, 1)
    if warmup and est * steps > budget_s:
        steps = max(3, int(budget_s / est) // 3 * 3)
    t0 = time.perf_counter()
    for i in range(steps):
        step(batches[i % 2], TASKS[i % 3])
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return 1.0 / dt, dt, cores, steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    value, dt, cores, steps = time_cpu(steps, warm)
    n = max(1, args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": n, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": B, "per_gpu_batch": B,
                   "device": "host CPU (the reference path is pure PyTorch)"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": "%d timed optimizer steps (tasks MLM/SAP/CFP round-robin) at the FULL batch 64; oracle/ "
                                   "restatement of the reference model (pinned by tests/golden), torch fp32, %d threads"
                                   % (steps, cores)},
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
        if pw:
            thr = 0.6 * max(pw)        # "under load": samples above 60 % of the max draw seen
            sm_l = [s for s, p in zip(sm, pw) if p >= thr] or sm
        else:
            sm_l = sm
        sm_l.sort()
        return {"sm_mhz": sm_l[len(sm_l) // 2] if sm_l else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Prefetcher(object):
    """Background host workers: batch dict -> prepare_pretrain (index builders, padding straight into pinned memory) a
    few steps ahead, results handed out in step order (the reference overlaps collate with compute the same way:
    DataLoader workers + PrefetchLoader, P/data/loader.py:62-130)."""

    def __init__(self, batches, pad, depth=4, workers=2):
        self.batches, self.pad, self.depth, self.workers = batches, pad, depth, workers
        self.cv = threading.Condition()
        self.ready = {}
        self.next_get = 0
        self.stop = False
        self.threads = [threading.Thread(target=self._run, args=(w,), daemon=True) for w in range(workers)]
        for t in self.threads:
            t.start()

    def _run(self, w):
        from vln_goat_b200 import batching
        i = w
        while not self.stop:
            with self.cv:
                while not self.stop and i >= self.next_get + self.depth:
                    self.cv.wait(0.05)
            if self.stop:
                return
            task = TASKS[i % 3]
            b = self.batches[(i // 3) % len(self.batches)]
            P = batching.pin(batching.prepare_pretrain(b, task, pad=self.pad, pinned=True))
            with self.cv:
                self.ready[i] = (task, P)
                self.cv.notify_all()
            i += self.workers

    def get(self):
        with self.cv:
            while self.next_get not in self.ready:
                self.cv.wait(0.05)
            item = self.ready.pop(self.next_get)
            self.next_get += 1
            self.cv.notify_all()
        return item

    def close(self):
        self.stop = True
        with self.cv:
            self.cv.notify_all()


def run_goat(args):
    import torch
    import torch.distributed as dist
    from vln_goat_b200 import batching, engine, ops, pretrain_model, runtime, workloads
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
    runtime.set_weight_split(not args.no_weight_split)

    model = pretrain_model.GlocalTextPathCMTPreTraining(GoatConfig(pretrain_tasks=TASKS))
    model.load_state_dict(oracle_params(), strict=True)     # same seeded weights on every rank (and in the oracle)
    model.tie_weights()
    model = model.to(dev).train()
    pad = batching.PadSpec()

    host = [workloads.synthetic_pretrain_batch(B, L, seed=1000 + 17 * rank + i) for i in range(N_BATCHES)]

    def loss_fn_of(task):
        return lambda P: model.scalar_loss(P, task)

    def to_dev(P):
        return {k: v.to(dev, non_blocking=True) for k, v in P.items()}

    # parameters any task trains (the reference's AdamW never touches the rest: their grad stays None)
    active, seen = [], set()
    for task in TASKS:
        P0 = to_dev(batching.prepare_pretrain(host[0], task, pad=pad))
        for p in engine.active_parameters(model, loss_fn_of(task), (P0,)):
            if id(p) not in seen:
                seen.add(id(p))
                active.append(p)
    flat = engine.FlatParams(model, shadow_dtype=cdt if cdt != torch.float32 else None, only=active)
    if cdt == torch.float16:
        flat.enable_loss_scale()
    ts = engine.TrainStep(flat, use_graph=not args.no_graph, check_unwritten=False, **OPT)

    # parity of the benchmarked model / dtype / batch size against the CPU oracle, on the seeded parameters (before any
    # optimizer step: the timed steps train on random labels at a constant 5e-5, which is not a state worth certifying)
    parity_before = None
    extras = (not args.no_extras) and world == 1      # parity / rooflines / baselines / other configs: the 1-GPU line has them
    if extras:
        parity_before = parity(torch, model, host[0], dev, cdt, "seeded parameters, before the timed steps")

    # device-resident prepared batches; one captured graph per (task, padded-shape signature)
    resident = []
    for i in range(N_BATCHES):
        for task in TASKS:
            P = to_dev(batching.prepare_pretrain(host[i], task, pad=pad))
            key = (task, engine.input_signature(P))
            if not ts.has(key):
                ts.capture(key, loss_fn_of(task), P)
            resident.append((task, key, P))
    torch.cuda.synchronize()
    n_graphs = len(ts.entries)

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

    launches = [0]

    def step_resident(i):
        task, key, P = resident[i % len(resident)]
        ts.step(P, key)
        launches[0] += ts.launches_per_step

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    launches[0] = 0
    ms = timed(step_resident, args.steps)
    units = 1 if args.global_batch else world        # strong scaling counts global-batch steps, weak counts per-rank batches
    value = units * args.steps / (ms / 1e3)
    gpu_launches = launches[0]

    # ---- end to end: host batch dict -> index builders + padding -> pinned -> H2D on a copy stream -> step -> loss D2H
    def run_e2e(host_batches, steps):
        pf = Prefetcher(host_batches, pad)
        copy_stream = torch.cuda.Stream()
        wu = max(args.warmup, 3)
        losses = torch.zeros(steps + wu + 4, dtype=torch.float32).pin_memory()
        state = {"n": 0, "h2d": 0}
        slots = [None, None]
        ready = [torch.cuda.Event() for _ in range(2)]

        def stage(i):
            task, P = pf.get()
            s = i % 2
            with torch.cuda.stream(copy_stream):
                d = {k: v.to(dev, non_blocking=True) for k, v in P.items()}
                ready[s].record(copy_stream)
            slots[s] = (task, P, d)
            state["h2d"] = batching.h2d_bytes(P)

        def step_e2e(i):
            s = i % 2
            task, P, d = slots[s]
            cur = torch.cuda.current_stream()
            cur.wait_event(ready[s])
            key = (task, engine.input_signature(d))
            if not ts.has(key):
                ts.capture(key, loss_fn_of(task), d)     # an unseen padded shape: captured once, during the warm-up steps
            ts.load_inputs(d, key)
            for v in d.values():
                v.record_stream(cur)
            stage(i + 1)
            loss = ts.step(None, key)
            losses[state["n"]:state["n"] + 1].copy_(loss.reshape(1), non_blocking=True)
            state["n"] += 1

        stage(0)
        for i in range(max(wu, 3 * len(host_batches))):       # every (task, batch) pair once: all graphs exist before timing
            step_e2e(i)
        off = max(wu, 3 * len(host_batches))
        ms_ = timed(lambda i: step_e2e(i + off), steps)
        pf.close()
        torch.cuda.synchronize()
        lv_ = losses[:state["n"]].clone()
        if not bool(torch.isfinite(lv_).all()):
            raise RuntimeError("non-finite loss in the timed run: %s" % lv_.tolist())
        return (1 if args.global_batch else world) * steps / (ms_ / 1e3), ms_, state["h2d"], lv_

    e2e, ms_e2e, h2d_bytes, lv = run_e2e(host, args.steps)

    # ---- the same end-to-end loop over a GPU-resident 16-bit feature bank: a batch names its panoramas by
    #      bank row, the [S,36,768] features never cross PCIe
    bank_line = None
    if extras and cdt != torch.float32:
        feats = torch.cat([b["traj_view_img_fts"] for b in host], 0)
        model.bert.feature_bank = workloads.FeatureBank(feats, dtype=cdt, device=dev)
        host_bank, off = [], 0
        for b in host:
            nb = {k: v for k, v in b.items() if k != "traj_view_img_fts"}
            n = b["traj_view_img_fts"].shape[0]
            nb["traj_view_ids"] = torch.arange(off, off + n, dtype=torch.int32)
            off += n
            host_bank.append(nb)
        e2e_b, ms_b, h2d_b, _ = run_e2e(host_bank, args.steps)
        bank_line = {"value": e2e_b, "unit": "steps/s", "ms_per_step": ms_b / args.steps, "h2d_bytes_per_step": h2d_b,
                     "bank_bytes": int(feats.numel() * 2),
                     "note": "same host-driven loop; batches carry traj_view_ids (rows of workloads.FeatureBank, 16 bit, resident in "
                             "HBM) instead of [S,36,768] fp32 features; goat_gather_rows feeds img_linear's GEMM operand directly"}

    # ---- sustained: the device-resident loop for >= 2 s
    sustained = None
    if not args.no_extras:
        n_s = max(args.steps, int(2.2 * value / units) // 3 * 3 + 3)
        ms_s = timed(step_resident, n_s)
        sustained = {"steps": n_s, "seconds": ms_s / 1e3, "value": units * n_s / (ms_s / 1e3), "unit": "steps/s"}
    clk = clocks.stop() if rank == 0 else None

    if os.environ.get("GOAT_SHARD_TIMING") and world > 1:      # diagnostic, outside every timed region
        flat.profile_phases(True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(30):
            step_resident(i)
        ev1.record()
        pt = flat.phase_times()
        flat.profile_phases(False)
        if rank == 0 or os.environ["GOAT_SHARD_TIMING"] == "all":
            sys.stderr.write("sharded_step phases (ms, mean of 30 steps, rank %d): %s | step %.3f ms\n"
                             % (rank, json.dumps({k: round(v, 3) for k, v in pt.items()}), ev0.elapsed_time(ev1) / 30))

    check_line = None
    if world > 1 and not args.no_selfcheck:
        check_line = selfcheck(torch, dist, model, flat, resident, dev, rank)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.global_batch else "weak", "samples_per_s": B * world * args.steps / (ms / 1e3),
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B * world, "per_gpu_batch": B, "parallelism": "dp%d" % world,
                       "dropout": 0.1, "cuda_graph": not args.no_graph, "captured_graphs": n_graphs,
                       "padding": "S (trajectory steps) to x32, G (map nodes) to x8, masked tokens to x128",
                       "loss_scale": "dynamic (device-resident GradScaler semantics)" if cdt == torch.float16 else None,
                       "weight_split": bool(flat.split),
                       "params_trained": int(flat.numel),
                       "l2": "per-step working set (%.0f MB params/grads/moments + activations) exceeds the 126 MB L2; "
                             "%d rotating batches per task" % (flat.numel * 4 * 4 / 1e6, N_BATCHES),
                       "loss_first_last": [float(lv[0]), float(lv[-1])]},
            "clocks": clk,
            "e2e": {"value": e2e, "unit": "steps/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "host_work": "batching.prepare_pretrain (index builders + padding straight into pinned memory) on two worker "
                                 "threads, H2D on a copy stream one step ahead"},
            "gpu_launches": gpu_launches,
            "gpu_launches_per_step": gpu_launches / float(args.steps),
            "sustained": sustained,
            "e2e_feature_bank": bank_line,
        }
        if flat.scaler is not None:
            sc = flat.scaler.cpu().tolist()
            line["config"]["loss_scale_state"] = {"scale": sc[0], "skipped_steps": sc[3], "steps_taken": sc[4]}
        if extras:
            line["parity_max_err"] = parity_before
            after = parity(torch, model, host[0], dev, cdt, "the same parameters after all optimizer steps of this run")
            line["parity_after_training_steps"] = {"max": after["max"], "steps_taken": flat.step_count, "worst": max(
                (k for k in after if isinstance(after[k], float) and k not in ("max", "tolerance")), key=lambda k: after[k])}
            line["roofline"] = gemm_roofline(torch, ops, ts, resident, ms / args.steps)
            line["roofline_attention"] = attention_roofline(torch, ops, ts, resident, cdt, ms / args.steps)
            if world == 1:
                line["eager_b200_baseline"] = eager_baseline(torch, host, dev)
                line["c2"] = run_c2(torch, cdt, dev)
                line["c5"] = run_c2(torch, cdt, dev, steps=20, L=512, B=32, name="C5 (RxR long-instruction stress)")
                line["c4"] = run_c4(torch, cdt, dev)
                if not args.no_cpu_baseline:
                    v, dt, cores, n = time_cpu(3, 0)
                    line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
                                            "sample": "3 timed optimizer steps (one per task MLM/SAP/CFP) at the FULL batch 64; "
                                                      "oracle/ restatement of the reference model, torch fp32, %d threads" % cores}
        if world > 1:
            line["selfcheck"] = check_line
            line["config"]["gradient_exchange"] = (
                "fused over NVLink peer memory (csrc/exchange.cu): peer loads of the gradient shards + sum of squares, clip + "
                "AdamW on 1/world storing the new operands into every rank" if isinstance(flat._peer, dict) else
                "NCCL reduce-scatter + AdamW on 1/world + all-gather")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        flat.release_peers()
        dist.destroy_process_group()


def selfcheck(torch, dist, model, flat, resident, dev, rank):
    """After the sharded optimizer steps of the run every rank must hold bit-identical parameters (fp32 masters of what
    the kernels read in fp32, 16-bit shadow of the rest) and compute bit-identical forward outputs."""
    import hashlib
    flat.sync_master()
    torch.cuda.synchronize()
    h = hashlib.sha256(flat.p[:flat.numel].cpu().numpy().tobytes())
    if flat.shadow is not None:
        h.update(flat.shadow[:flat.numel].view(torch.int16).cpu().numpy().tobytes())
    from vln_goat_b200 import batching, workloads
    probe = workloads.synthetic_pretrain_batch(8, L, seed=4242)        # the SAME batch on every rank (built from the seed)
    worst = 0.0
    model.eval()
    with torch.no_grad():
        for task in TASKS:
            src = {k: v.to(dev) for k, v in batching.prepare_pretrain(probe, task, pad=None).items()}
            out = model.forward_prepared(src, task, compute_loss=False)
            for o in (out if isinstance(out, tuple) else (out,)):
                if torch.is_floating_point(o):
                    # same shapes on every rank (same probe batch); the 1-wide heads add split-K partials with atomics,
                    # so outputs are compared to rounding noise, parameters bit for bit
                    ref = o.float().contiguous().clone()
                    dist.broadcast(ref, src=0)
                    fin = torch.isfinite(ref)
                    if not torch.equal(fin, torch.isfinite(o)):
                        worst = float("inf")
                    if fin.any():
                        worst = max(worst, float((o.float()[fin] - ref[fin]).abs().max()))
    model.train()
    wt = torch.tensor([worst], device=dev, dtype=torch.float64)
    dist.all_reduce(wt, op=dist.ReduceOp.MAX)
    if float(wt) > 1e-5:
        raise RuntimeError("selfcheck: forward outputs differ across ranks by %g after the data-parallel steps" % float(wt))
    digest = int(h.hexdigest()[:15], 16)
    t = torch.tensor([digest], device=dev, dtype=torch.int64)
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if int(lo) != int(hi):
        raise RuntimeError("selfcheck: ranks hold different parameters / outputs after the data-parallel steps")
    msg = "ok: bit-identical parameters on all %d ranks (digest %x), forward outputs equal to %.1e" % (
        dist.get_world_size(), digest, float(wt))
    if rank == 0:
        print("selfcheck " + msg, file=sys.stderr)
    return msg


def parity(torch, model, batch, dev, cdt, state):
    """Eval-mode forward of the benchmarked model (same dtype, same batch size 64) against the CPU oracle."""
    from oracle import goat_pretrain_oracle as PO
    from vln_goat_b200 import batching
    torch.set_num_threads(os.cpu_count() or 1)
    Pr = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    model.eval()
    errs = {}

    def rel(got, ref):
        got, ref = got.detach().float().cpu(), ref.float()
        fin = torch.isfinite(ref)
        if not torch.equal(torch.isfinite(got), fin):
            return float("inf")
        return float((got[fin] - ref[fin]).abs().max() / max(1.0, float(ref[fin].abs().max()))) if fin.any() else 0.0
    with torch.no_grad():
        for task in TASKS:
            P = {k: v.to(dev) for k, v in batching.prepare_pretrain(batch, task, pad=batching.PadSpec()).items()}
            out = model.forward_prepared(P, task, compute_loss=False)
            loss = model.forward_prepared(P, task, compute_loss=True)
            if task == "mlm":
                scores, rloss = PO.forward_mlm(Pr, batch)
                n = scores.shape[0]
                errs["mlm_scores"] = rel(out[:n], scores)
                errs["mlm_loss"] = rel(loss[:n], rloss)
            elif task == "sap":
                gl, ll, fl, rloss = PO.forward_sap(Pr, batch)
                G_ = gl.shape[1]
                errs["sap_global_logits"] = rel(out[0][:, :G_], gl)
                errs["sap_local_logits"] = rel(out[1], ll)
                errs["sap_fused_logits"] = rel(out[2][:, :G_], fl)
                errs["sap_loss"] = rel(loss, rloss)
            else:
                go, vo, fo, to, rloss = PO.forward_cfp(Pr, batch)
                for nm, a, b in (("cfp_gmap", out[0], go), ("cfp_vp", out[1], vo), ("cfp_fused", out[2], fo), ("cfp_txt", out[3], to)):
                    errs[nm] = rel(a, b)
                errs["cfp_loss"] = rel(loss, rloss)
    model.train()
    errs["max"] = max(errs.values())
    errs["tolerance"] = 1e-5 if cdt == torch.float32 else 1e-3
    errs["reference"] = "oracle/goat_pretrain_oracle.py (fp32, CPU), batch 64, eval mode; " + state
    return errs


def eager_baseline(torch, host, dev):
    """The reference-style eager path ON the B200: the oracle restatement (plain torch ops, Python aggregation loops,
    per-tensor AdamW) with .cuda() tensors, fp32 and autocast fp16."""
    out = {}
    dev_batches = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()} for b in host[:2]]
    for name, ac in (("fp32", False), ("autocast_fp16", True)):
        try:
            step = oracle_step_fn(dev, autocast=ac)
            for i in range(3):
                step(dev_batches[i % 2], TASKS[i % 3])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 6
            for i in range(n):
                step(dev_batches[i % 2], TASKS[i % 3])
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
            out[name] = {"value": 1.0 / dt, "unit": "steps/s", "ms_per_step": dt * 1e3}
        except Exception as e:       # noqa: BLE001 -- a baseline that cannot run is reported, not fatal
            out[name] = {"error": repr(e)[:200]}
        torch.cuda.empty_cache()
    out["what"] = ("oracle/ restatement of the reference model run eagerly on the same B200 (torch ops, Python loops of the "
                   "reference kept), 6 timed steps, batch 64, wall clock with synchronize")
    return out


def _time_graph(torch, fn, reps=20):
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
    ms = e0.elapsed_time(e1) / reps
    del g
    return ms


def _record_round(torch, ops, ts, resident, what):
    """eager fwd+bwd of one MLM + SAP + CFP round with the launch log on -> list of records"""
    rec = []
    setattr(ops, what, rec)
    done = set()
    for task, key, P in resident:
        if task in done:
            continue
        done.add(task)
        ts.load_inputs(P, key)
        ts._fwd_bwd(ts.entries[key])
    setattr(ops, what, None)
    torch.cuda.synchronize()
    ts.flat.zero_grad()
    return rec


def gemm_roofline(torch, ops, ts, resident, step_ms):
    """Re-time every GEMM launch of one MLM + SAP + CFP round (shapes recorded from the real steps) with CUDA events."""
    pk = peaks()
    peak = pk.get("bf16_tflops")
    which = "MEASURED_PEAKS.json bf16_tflops (burst: each launch group is timed alone)"
    if peak is None:
        peak, which = 1650.0, "fallback (B200_PROFILING.md burst ~1.65 PFLOP/s)"
    rec = _record_round(torch, ops, ts, resident, "GEMM_LOG")
    uniq = {}
    for r in rec:
        uniq[r] = uniq.get(r, 0) + 1
    dev = torch.device("cuda", torch.cuda.current_device())
    umma_ms, umma_flop, n_umma, simt_ms, exec_flop = 0.0, 0.0, 0, 0.0, 0.0
    table = []
    from vln_goat_b200 import runtime
    n_split = 0
    for (M, N, K, a_mn, b_mn, dt_, simt, acc_, act, has_bias, has_res, out32, has_drop, has_out2, has_lo), cnt in uniq.items():
        dt_t = {0: torch.float32, 1: torch.float16, 2: torch.bfloat16}[dt_]
        A = (torch.randn((K, M) if a_mn else (M, K), device=dev) * 0.05).to(dt_t)
        if has_lo:       # split weight operand: [hi | lo] buffer, the launch runs the K loop twice (FLOPs counted once)
            Bm = runtime._cast_split(torch.randn(N, K, device=dev) * 0.05, dt_t)
            n_split += cnt
        else:
            Bm = (torch.randn((K, N) if b_mn else (N, K), device=dev) * 0.05).to(dt_t)
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
        ms = _time_graph(torch, lambda: ops.gemm(A, Bm, **kw))
        table.append((ms * cnt, "%s M=%d N=%d K=%d a_mn=%d b_mn=%d acc=%d act=%d bias=%d res=%d out32=%d drop=%d split_w=%d x%d  %.1f us  %.0f TFLOP/s"
                      % ("simt" if simt else "umma", M, N, K, a_mn, b_mn, acc_, act, has_bias, has_res, out32, has_drop, has_lo, cnt,
                         ms * 1e3, 2.0 * M * N * K / (ms * 1e-3) / 1e12)))
        if simt:
            simt_ms += ms * cnt
        else:
            umma_ms += ms * cnt
            umma_flop += 2.0 * M * N * K * cnt
            exec_flop += 2.0 * M * N * K * cnt * (2 if has_lo else 1)
            n_umma += cnt
    achieved = umma_flop / (umma_ms * 1e-3) / 1e12 if umma_ms > 0 else 0.0
    round_ms = 3.0 * step_ms
    if os.environ.get("GOAT_BENCH_SHAPES"):       # per-shape table, most expensive first (profiles/)
        with open(os.environ["GOAT_BENCH_SHAPES"], "w") as f:
            f.write("# every GEMM shape of one MLM+SAP+CFP round: total ms per round, then the launch\n")
            for tot, line in sorted(table, reverse=True):
                f.write("%8.3f ms  %s\n" % (tot, line))
    traffic, traffic_src = None, None
    try:       # DRAM bytes of the same launches, from the committed ncu pass over one round (profiles/)
        with open(os.path.join(ROOT, "profiles", "r02m_gemm_traffic.json")) as f:
            tj = json.load(f)
        if abs(tj.get("gemm_launches", 0) - n_umma) <= 0.05 * n_umma:
            traffic, traffic_src = tj["gemm_dram_bytes_per_round"], tj["source"]
    except Exception:
        pass
    algo_bytes = sum(c * 2.0 * (m_ * k_ + n_ * k_ + m_ * n_) for (m_, n_, k_, *_r), c in uniq.items() if not _r[3])
    return {"bound": "tensor", "kernel": "gemm_umma2_kernel / gemm_umma_kernel (tcgen05.mma cta_group::2 + TMA; all %d tensor-core GEMM "
                                          "launches of one MLM+SAP+CFP round, each shape re-timed as 20 graph-captured launches)" % n_umma,
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_unit": "DRAM bytes per round over the same launches", "traffic_source": traffic_src,
            "operand_bytes_per_round": algo_bytes,
            "peak_source": which, "flop_per_round": umma_flop, "gemm_ms_per_round": umma_ms, "simt_gemm_ms_per_round": simt_ms,
            "split_weight_launches": n_split,
            "executed_tflops": exec_flop / (umma_ms * 1e-3) / 1e12 if umma_ms > 0 else 0.0,
            "executed_frac": (exec_flop / (umma_ms * 1e-3) / 1e12 / peak) if umma_ms > 0 else 0.0,
            "split_weight_note": ("forward GEMMs run the K loop twice (16-bit weight hi + lo terms, ~22-bit weights) for parity; "
                                  "achieved / frac count the ALGORITHMIC 2MNK once; executed_* count what the tensor pipe actually ran") if n_split else None,
            "gemm_share_of_step": umma_ms / round_ms if round_ms else None,
            "step_tflops": umma_flop / (round_ms * 1e-3) / 1e12 if round_ms else None}


def attention_roofline(torch, ops, ts, resident, cdt, step_ms):
    """The other regime of SURVEY.md 8d: the attention core (QK^T, softmax, PV and its backward) is HBM / latency
    bound.  Algorithmic bytes of the attention launches of one round / their device time."""
    if cdt == torch.float32:
        return None
    pk = peaks()
    peak = pk.get("hbm_gbs")
    which = "MEASURED_PEAKS.json hbm_gbs"
    if peak is None:
        peak, which = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    rec = _record_round(torch, ops, ts, resident, "ATTN_LOG")
    uniq = {}
    for r in rec:
        uniq[r] = uniq.get(r, 0) + 1
    dev = torch.device("cuda", torch.cuda.current_device())
    heads = 12
    tot_ms, tot_bytes, n = 0.0, 0.0, 0
    for (kind, Bn, Nq, Nk, has_bias, p_drop), cnt in uniq.items():
        q = torch.randn(Bn, Nq, H, device=dev).to(cdt)
        k = torch.randn(Bn, Nk, H, device=dev).to(cdt)
        v = torch.randn(Bn, Nk, H, device=dev).to(cdt)
        km = torch.zeros(Bn, Nk, device=dev)
        bias = torch.zeros(Bn, Nq, Nk, device=dev) if has_bias else None
        o, lse = ops.attn_fwd(q, k, v, heads, km, bias, drop_p=p_drop, drop_seed=3)
        if kind == "fwd":
            ms = _time_graph(torch, lambda: ops.attn_fwd(q, k, v, heads, km, bias, drop_p=p_drop, drop_seed=3))
            nbytes = 2.0 * Bn * H * (2 * Nq + 2 * Nk) + 4.0 * Bn * Nk + (4.0 * Bn * Nq * Nk if has_bias else 0.0)
        else:
            w = torch.randn(Bn, Nq, H, device=dev).to(cdt)
            dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
            ms = _time_graph(torch, lambda: ops.attn_bwd(w, q, k, v, o, lse, heads, dq, dk, dv, km, bias, drop_p=p_drop,
                                                          drop_seed=3, want_dbias=bool(has_bias)))
            nbytes = 2.0 * Bn * H * (4 * Nq + 4 * Nk) + 4.0 * Bn * Nk + (8.0 * Bn * Nq * Nk if has_bias else 0.0)
        tot_ms += cnt * ms
        tot_bytes += cnt * nbytes
        n += cnt
    achieved = tot_bytes / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0
    round_ms = 3.0 * step_ms
    return {"bound": "hbm", "kernel": "attention core kernels (all %d launches of one MLM+SAP+CFP round, dropout as run, each shape "
                                      "re-timed as 20 graph-captured launches)" % n,
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "peak_source": which, "bytes_per_round": tot_bytes, "attention_ms_per_round": tot_ms,
            "attention_share_of_step": tot_ms / round_ms if round_ms else None}


# ------------------------------------------------------------------------------------------------
# second line: BASELINE.json configs[1] (the round-1 headline): 6 RobertaLayer + 3 BertCrossLayer, batch 64
# ------------------------------------------------------------------------------------------------
def run_c2(torch, cdt, dev, steps=60, L=L, B=B, name="C2"):
    from vln_goat_b200 import engine, ops, workloads
    from vln_goat_b200.config import GoatConfig
    torch.manual_seed(0)
    model = workloads.C2CrossEncoder(GoatConfig()).to(dev).train()

    def loss_fn(txt, tm, vp, vm):
        t, v = model(txt, tm, vp, vm)
        return workloads.c2_loss(t, v)
    g = torch.Generator().manual_seed(5)
    batches = []
    for _ in range(2 if L > 128 else 4):
        txt = torch.randn(B, L, H, generator=g)
        vp = torch.randn(B, NQ, H, generator=g)
        vp[:, 0] = 0.0
        lens = torch.randint(L // 2, L + 1, (B,), generator=g)
        lens[0] = L
        tm = torch.arange(L)[None, :] < lens[:, None]
        batches.append(tuple(t.to(dev) for t in (txt, tm, vp, torch.ones(B, NQ, dtype=torch.bool))))
    active = engine.active_parameters(model, loss_fn, batches[0])
    flat = engine.FlatParams(model, shadow_dtype=cdt if cdt != torch.float32 else None, only=active)
    if cdt == torch.float16:
        flat.enable_loss_scale()
    ts = engine.TrainStep(flat, loss_fn, batches[0], **OPT)
    nb = len(batches)
    for i in range(5):
        ts.step(batches[i % nb])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        ts.step(batches[i % nb])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"workload": "%s: 6 RobertaLayer + 3 BertCrossLayer fwd+bwd+clip+AdamW (70.9 M params), batch %d, %d text tokens x %d "
                       "view tokens, device-resident" % (name, B, L, NQ),
           "value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms, "steps": steps, "gpu_launches_per_step": ts.launches_per_step}
    if L > 128 and cdt != torch.float32:
        # the long-instruction attention core alone (query-tiled tcgen05 kernels, attention_tc.cu): achieved HBM GB/s
        heads = 12
        q = torch.randn(B, L, H, device=dev).to(cdt)
        km = torch.zeros(B, L, device=dev)
        o, lse = ops.attn_fwd(q, q, q, heads, km, drop_p=0.1, drop_seed=3)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
        f_ms = _time_graph(torch, lambda: ops.attn_fwd(q, q, q, heads, km, drop_p=0.1, drop_seed=3))
        b_ms = _time_graph(torch, lambda: ops.attn_bwd(q, q, q, q, o, lse, heads, dq, dk, dv, km, drop_p=0.1, drop_seed=3))
        fb = 2.0 * B * H * 4 * L + 4.0 * B * L
        bb = 2.0 * B * H * 8 * L + 4.0 * B * L
        pk = peaks().get("hbm_gbs") or 6650.0
        fl = 4.0 * B * L * L * H
        out["attention_self_%d" % L] = {
            "fwd_us": f_ms * 1e3, "bwd_us": b_ms * 1e3, "fwd_gbs": fb / (f_ms * 1e-3) / 1e9, "bwd_gbs": bb / (b_ms * 1e-3) / 1e9,
            "fwd_frac_of_hbm": fb / (f_ms * 1e-3) / 1e9 / pk, "bwd_frac_of_hbm": bb / (b_ms * 1e-3) / 1e9 / pk,
            "fwd_tflops": fl / (f_ms * 1e-3) / 1e12, "bwd_tflops": 2.5 * fl / (b_ms * 1e-3) / 1e12,
            "note": "text self-attention B%d x 12 heads x %d x %d, dropout 0.1; bytes = Q,K,V,O (+ dO,dQ,dK,dV) in 16 bit + key mask" % (B, L, L)}
    return out


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: fine-tune rollout, 16 parallel episodes x 15 steps, BACL + FACL on (per GPU; episodes are
# independent, so N GPUs run N such rollouts with the same gradient exchange as the pretraining step)
# ------------------------------------------------------------------------------------------------
def run_c4(torch, cdt, dev, episodes=16, T=15, reps=4):
    from vln_goat_b200 import engine, nav_model, workloads
    from vln_goat_b200.config import GoatConfig
    cfg = GoatConfig(layer_norm_eps=1e-5, pad_token_id=1, dataset="r2r", mode="train", obj_feat_size=0, feat_dropout=0.4,
                     do_back_img=True, do_back_txt=True, do_front_img=True, do_front_his=True, do_front_txt=True,
                     do_back_txt_type="type_2", do_back_img_type="type_1", do_add_method="door", use_lang2visn_attn=False,
                     fix_lang_embedding=False, fix_pano_embedding=False, fix_local_branch=False)
    torch.manual_seed(0)
    model = nav_model.GlocalTextPathNavCMT(cfg).to(dev).train()
    lang, steps, targets, mem0 = workloads.synthetic_nav_rollout(episodes, L, T, seed=7)
    flat_in = {"mem0": mem0}
    for k, v in lang.items():
        flat_in["lang." + k] = v
    for t, ((pano, nav), tgt) in enumerate(zip(steps, targets)):
        for k, v in pano.items():
            if torch.is_tensor(v):
                flat_in["s%02d.p.%s" % (t, k)] = v
        for k, v in nav.items():
            flat_in["s%02d.n.%s" % (t, k)] = v
        flat_in["s%02d.t" % t] = tgt
    h2d = sum(v.numel() * v.element_size() for v in flat_in.values())
    dev_in = {k: v.to(dev) for k, v in flat_in.items()}

    def loss_fn(D):
        lg = {k[5:]: v for k, v in D.items() if k.startswith("lang.")}
        st, tg = [], []
        for t in range(T):
            pre = "s%02d." % t
            pano = {k[len(pre) + 2:]: v for k, v in D.items() if k.startswith(pre + "p.")}
            pano["already_dropout"] = True
            nav = {k[len(pre) + 2:]: v for k, v in D.items() if k.startswith(pre + "n.")}
            st.append((pano, nav))
            tg.append(D[pre + "t"])
        return workloads.nav_rollout_loss(model, lg, st, tg, D["mem0"])

    # (a) eager autograd (how the reference's agent drives the model: one Python call per op)
    def eager():
        model.zero_grad(set_to_none=True)
        loss = loss_fn(dev_in)
        loss.backward()
        return loss
    for _ in range(2):
        eager()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        eager()
    torch.cuda.synchronize()
    eager_ms = (time.perf_counter() - t0) / reps * 1e3
    # (b) the same rollout as ONE captured CUDA graph (forward + backward of all 15 steps) + fused clip / AdamW
    active = engine.active_parameters(model, loss_fn, (dev_in,))
    flat = engine.FlatParams(model, shadow_dtype=cdt if cdt != torch.float32 else None, only=active)
    if cdt == torch.float16:
        flat.enable_loss_scale()
    ts = engine.TrainStep(flat, loss_fn, dev_in, lr=1e-5, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01, max_grad_norm=40.0,
                          check_unwritten=False)
    for _ in range(3):
        ts.step(dev_in)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3 * reps
    e0.record()
    for _ in range(n):
        loss = ts.step(dev_in)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    if not bool(torch.isfinite(loss)):
        raise RuntimeError("non-finite rollout loss")
    return {"workload": "C4: teacher-forced fine-tune rollout, %d episodes x %d steps (language once, panorama + navigation per "
                        "step, BACL + FACL on, global map growing 8 -> 56 nodes), fwd + bwd + clip + AdamW, 188 M params" % (episodes, T),
            "value": 1e3 / ms, "unit": "rollouts/s", "ms_per_rollout": ms, "ms_per_nav_step": ms / T,
            "eager_ms_per_rollout": eager_ms, "graph_speedup_vs_eager": eager_ms / ms,
            "gpu_launches_per_rollout": ts.launches_per_step, "h2d_bytes_per_rollout": h2d, "loss": float(loss),
            "note": "the whole rollout (all steps, one backward) is ONE CUDA graph: with teacher forcing the observations do "
                    "not depend on the model's outputs, so every step's inputs and logit-fusion indices are built up front"}


def main():
    global B, WORKLOAD
    args = parse()
    if args.global_batch:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.global_batch % world:
            raise SystemExit("--global-batch must be a multiple of the number of ranks")
        B = args.global_batch // world
        WORKLOAD = WORKLOAD.replace("batch 64 per GPU", "GLOBAL batch %d = %d per GPU (strong scaling)" % (args.global_batch, B))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_goat(args)


if __name__ == "__main__":
    main()
