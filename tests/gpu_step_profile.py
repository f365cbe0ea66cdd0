"""ncu target (not a test): one eager MLM + SAP + CFP round of the full pretraining step between cudaProfilerStart/Stop.

    GOAT_PDL=0 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tests/gpu_step_profile.py [fp16|bf16]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from vln_goat_b200 import batching, engine, pretrain_model, runtime, workloads
from vln_goat_b200.config import GoatConfig

cdt = {"fp16": torch.float16, "bf16": torch.bfloat16}[sys.argv[1] if len(sys.argv) > 1 else "fp16"]
only = sys.argv[2] if len(sys.argv) > 2 else None
runtime.set_compute_dtype(cdt)
dev = torch.device("cuda", 0)
model = pretrain_model.GlocalTextPathCMTPreTraining(GoatConfig(pretrain_tasks=bench.TASKS))
model.load_state_dict(bench.oracle_params(), strict=True)
model.tie_weights()
model = model.to(dev).train()
host = workloads.synthetic_pretrain_batch(bench.B, bench.L, seed=1000)
pad = batching.PadSpec()
preps = {t: {k: v.to(dev) for k, v in batching.prepare_pretrain(host, t, pad=pad).items()} for t in bench.TASKS}
fns = {t: (lambda P, t=t: model.scalar_loss(P, t)) for t in bench.TASKS}
active, seen = [], set()
for t in bench.TASKS:
    for p in engine.active_parameters(model, fns[t], (preps[t],)):
        if id(p) not in seen:
            seen.add(id(p))
            active.append(p)
flat = engine.FlatParams(model, shadow_dtype=cdt, only=active)
if cdt == torch.float16:
    flat.enable_loss_scale()
ts = engine.TrainStep(flat, use_graph=False, check_unwritten=False, **bench.OPT)
for t in bench.TASKS:
    ts.capture(t, fns[t], preps[t])
    ts.step(preps[t], t)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for t in bench.TASKS:
    if only is None or t == only:
        ts.step(preps[t], t)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one round:", [t for t in bench.TASKS if only is None or t == only])
