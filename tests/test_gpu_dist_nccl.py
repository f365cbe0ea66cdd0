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


def _worker(rank, world, port, ret, shard):
    import torch.distributed as dist
    os.environ["GOAT_PEER_EXCHANGE"] = "0" if shard == "nccl" else "1"
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=dev)
    try:
        engine, model, loss_fn = _setup(dev)
        b = _batch(100 + rank, dev)
        active = engine.active_parameters(model, loss_fn, b)
        flat = engine.FlatParams(model, shadow_dtype=torch.bfloat16, only=active)
        ts = engine.TrainStep(flat, loss_fn, b, use_graph=False, lr=1e-3, max_grad_norm=0.05, shard_optimizer=bool(shard))
        ts.step(b)
        ts.step(b)
        if shard == "peer":       # the NVLink peer-memory exchange must be what ran, not the NCCL fallback
            assert isinstance(flat._peer, dict), "peer-memory exchange was not set up"
        flat.sync_master()          # sharded steps leave the fp32 master weights current on their owner rank only
        torch.cuda.synchronize()
        ret.put((rank, flat.p[:flat.numel].detach().cpu().numpy(), flat.shadow[:flat.numel].float().cpu().numpy(),
                 flat.grad_norm.item()))     # numpy: pickled by value (tensors travel as handles that die with the worker)
        dist.barrier()
        flat.release_peers()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("shard", [False, "nccl", "peer"])
def test_two_rank_step_matches_averaged_gradient_step(shard):
    """shard=False: all-reduce + full AdamW on every rank; "nccl": reduce-scatter + AdamW on 1/world + all-gather;
    "peer": the same step through csrc/exchange.cu (peer loads of the gradient shards, AdamW storing into every rank)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret, shard)) for r in range(2)]
    for p in procs:
        p.start()
    import queue
    res = []
    for _ in range(120):                      # poll: a worker that died must fail the test, not hang it
        try:
            res.append(ret.get(timeout=2))
        except queue.Empty:
            pass
        if len(res) == 2:
            break
        assert all(p.is_alive() or p.exitcode == 0 for p in procs), "a worker exited with %s" % [p.exitcode for p in procs]
    assert len(res) == 2, "workers timed out"
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = {r[0]: torch.from_numpy(r[1]) for r in res}
    shadows = {r[0]: torch.from_numpy(r[2]) for r in res}
    norms = {r[0]: r[3] for r in res}
    assert torch.equal(got[0], got[1])          # identical replicas after the data-parallel steps
    assert torch.equal(shadows[0], shadows[1])
    assert torch.equal(shadows[0], got[0].bfloat16().float())      # the operand shadow is the rounded master copy
    assert abs(norms[0] - norms[1]) < 1e-6 * max(1.0, norms[0])
    # 1-rank reference: accumulate both batches' gradients, scale by 1/2
    dev = torch.device("cuda", 0)
    engine, model, loss_fn = _setup(dev)
    b0, b1 = _batch(100, dev), _batch(101, dev)
    active = engine.active_parameters(model, loss_fn, b0)
    flat = engine.FlatParams(model, shadow_dtype=torch.bfloat16, only=active)
    flat.g.zero_()
    for _ in range(2):              # two optimizer steps, each on the average of the two ranks' gradients
        for b in (b0, b1):
            flat.begin_step()
            loss_fn(*b).backward()
        flat.adamw_step(lr=1e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01, max_grad_norm=0.05, grad_scale=0.5)
    torch.cuda.synchronize()
    ref = flat.p[:flat.numel].detach().cpu()
    assert abs(flat.grad_norm.item() - norms[0]) < 1e-3 * max(1.0, norms[0])
    # AdamW's first step moves every weight by ~lr * sign(g): compare the updates, tolerance for atomic-order noise
    err = (got[0] - ref).abs().max().item()
    assert err < 2e-4, err
