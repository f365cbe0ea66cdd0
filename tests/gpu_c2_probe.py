"""Scratch probe (not a test): C2 fwd+bwd timing, eager vs CUDA graph, bf16."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vln_goat_b200 import runtime, workloads
from vln_goat_b200.config import GoatConfig
from oracle import goat_oracle as O

B, L, Nq = 64, 80, 37
torch.manual_seed(0)
cfg = GoatConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
m = workloads.C2CrossEncoder(cfg)
m.load_state_dict(O.seeded_params(O.c2_shapes(), seed=0), strict=False)
m = m.cuda().train()
txt = torch.randn(B, L, 768, device="cuda")
vp = torch.randn(B, Nq, 768, device="cuda")
tm = torch.ones(B, L, dtype=torch.bool, device="cuda")
vm = torch.ones(B, Nq, dtype=torch.bool, device="cuda")
runtime.set_compute_dtype(torch.bfloat16)

def step():
    for p in m.parameters():
        p.grad = None
    t, v = m(txt, tm, vp, vm)
    (v.sum() + t.sum()).backward()

def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.time() - t0) / n * 1e3

print("eager ms/step (gpu, wall):", timeit(step))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    step()
print("graph ms/step (gpu, wall):", timeit(g.replay))
# per-kernel breakdown through torch profiler
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay(); g.replay()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=80))
