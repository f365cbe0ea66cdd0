"""world_size-2 data-parallel logic on CPU (gloo): shard the batch, all-reduce the flat gradient, average.

Mirrors SURVEY.md Appendix B item 5 with the CPU oracle as the model: the averaged 2-rank gradient of per-rank mean
losses equals the 1-rank gradient of the global-batch mean loss when the shards are equal (what DDP guarantees and
what engine.FlatParams.all_reduce + adamw_step(grad_scale=1/world) reproduce on the GPUs)."""
import os
import socket
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    from oracle import goat_oracle as O
    torch.manual_seed(3)
    B, Nq, Nk, H = 4, 5, 7, 768
    params = O.seeded_params(O.cross_layer_shapes("l."), seed=1)
    x = torch.randn(B, Nq, H)
    kv = torch.randn(B, Nk, H)
    lens = torch.tensor([7, 4, 6, 7])
    return O, params, x, kv, O.gen_seq_masks(lens, Nk)


def _loss(O, P, x, kv, kvm):
    qm = torch.ones(x.shape[0], x.shape[1], dtype=torch.bool)
    out = O.cross_layer(P, "l.", x, kv, O.extend_neg_masks(qm), O.extend_neg_masks(kvm))
    return 0.5 * (out ** 2).mean()


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    from vln_goat_b200 import dist_utils as D
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        O, params, x, kv, kvm = _problem()
        P = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        xs, kvs, ms = D.shard_batch((x, kv, kvm), rank, world)
        assert xs.shape[0] == x.shape[0] // world
        _loss(O, P, xs, kvs, ms).backward()
        flat, sizes = D.flatten_grads(list(P.values()))
        w = D.all_reduce_sum_(flat)
        assert w == world
        flat *= 1.0 / w                       # the optimizer kernel's grad_scale
        if rank == 0:
            ret.put(flat.numpy().copy())     # by value: a torch tensor would travel as a shared-memory handle that dies with the worker
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    from vln_goat_b200 import dist_utils as D
    for n in (1, 7, 64, 256):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    with pytest.raises(ValueError):
        D.shard_range(8, 2, 2)
    assert D.all_reduce_sum_(torch.ones(3)) == 1      # no process group: no-op, world 1


@pytest.mark.timeout(300)
def test_two_rank_gradient_average_equals_global_batch_gradient():
    import torch.multiprocessing as mp
    from vln_goat_b200 import dist_utils as D
    O, params, x, kv, kvm = _problem()
    P = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    _loss(O, P, x, kv, kvm).backward()
    ref, _ = D.flatten_grads(list(P.values()))
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = torch.from_numpy(ret.get(timeout=240))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    err = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)
    assert err < 1e-5, err
