"""Scratch probe (not a test): per-kernel time breakdown of the C2 training step (engine.TrainStep, bf16)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vln_goat_b200 import engine, runtime, workloads
from vln_goat_b200.config import GoatConfig

drop = float(os.environ.get("PROBE_DROPOUT", "0.1"))
dev = torch.device("cuda", 0)
runtime.set_compute_dtype(torch.bfloat16)
torch.manual_seed(0)
model = workloads.C2CrossEncoder(GoatConfig(hidden_dropout_prob=drop, attention_probs_dropout_prob=drop)).to(dev).train()

def loss_fn(txt, tm, vp, vm):
    t, v = model(txt, tm, vp, vm)
    return 0.5 * (t * t).mean() + 0.5 * (v * v).mean()

batch = tuple(t.to(dev) for t in bench.make_batches(1, 64, seed=1)[0])
active = engine.active_parameters(model, loss_fn, batch)
flat = engine.FlatParams(model, shadow_dtype=torch.bfloat16, only=active)
ts = engine.TrainStep(flat, loss_fn, batch, **bench.OPT)
for _ in range(3):
    ts.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ts.step()
e1.record()
torch.cuda.synchronize()
print("dropout %.2f: %.3f ms/step, %d launches/step" % (drop, e0.elapsed_time(e1) / 10, ts.launches_per_step))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ts.step(); ts.step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
