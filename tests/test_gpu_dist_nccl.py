"""2-rank NCCL check of the data-parallel step on real GPUs (skipped with fewer than 2 devices): after one
engine.TrainStep on DIFFERENT per-rank batches every rank holds identical parameters, and they equal a 1-rank step whose
gradient is the average of the two per-rank gradients (DDP semantics, P/utils/misc.py:52-58)."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup(dev):
    from vln_goat_b200 import engine, runtime, workloads
    from vln_goat_b200.config import GoatConfig
    runtime.set_compute_dtype(torch.bfloat16)
    torch.manual_seed(0)
    model = workloads.C2CrossEncoder(GoatConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)).to(dev).train()

    def loss_fn(txt, tm, vp, vm):
        t, v = model(txt, tm, vp, vm)
        return workloads.c2_loss(t, v)
    return engine, model, loss_fn


def _batch(seed, dev, B=4, L=80, Nq=37, H=768):
    g = torch.Generator().manual_seed(seed)
    txt = torch.randn(B, L, H, generator=g)
    vp = torch.randn(B, Nq, H, generator=g)
    tm = torch.ones(B, L, dtype=torch.bool)
    vm = torch.ones(B, Nq, dtype=torch.bool)
    return tuple(t.to(dev) for t in (txt, tm, vp, vm))


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=dev)
    try:
        engine, model, loss_fn = _setup(dev)
        b = _batch(100 + rank, dev)
        active = engine.active_parameters(model, loss_fn, b)
        flat = engine.FlatParams(model, shadow_dtype=torch.bfloat16, only=active)
        ts = engine.TrainStep(flat, loss_fn, b, use_graph=False, lr=1e-3, max_grad_norm=-1.0)
        ts.step(b)
        torch.cuda.synchronize()
        ret.put((rank, flat.p.detach().cpu()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_step_matches_averaged_gradient_step():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(ret.get(timeout=400) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert torch.equal(got[0], got[1])          # identical replicas after the all-reduced step
    # 1-rank reference: accumulate both batches' gradients, scale by 1/2
    dev = torch.device("cuda", 0)
    engine, model, loss_fn = _setup(dev)
    b0, b1 = _batch(100, dev), _batch(101, dev)
    active = engine.active_parameters(model, loss_fn, b0)
    flat = engine.FlatParams(model, shadow_dtype=torch.bfloat16, only=active)
    flat.g.zero_()
    for b in (b0, b1):
        flat.begin_step()
        loss_fn(*b).backward()
    flat.adamw_step(lr=1e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01, max_grad_norm=-1.0, grad_scale=0.5)
    torch.cuda.synchronize()
    ref = flat.p.detach().cpu()
    # AdamW's first step moves every weight by ~lr * sign(g): compare the updates, tolerance for atomic-order noise
    err = (got[0] - ref).abs().max().item()
    assert err < 2e-4, err
