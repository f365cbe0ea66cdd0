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


# ----------------------------------------------------------------------------------------------
# sharded optimizer step: layout arithmetic + the data flow of engine.FlatParams.sharded_step, emulated on CPU
# ----------------------------------------------------------------------------------------------
def test_shard_layout_and_tail_pieces():
    from vln_goat_b200 import dist_utils as D
    for world in (1, 2, 4, 8):
        numel = 1000 * world + 123
        padded = (numel + 8 * world - 1) // (8 * world) * (8 * world)
        n_decay = 700 * world + 16
        covered, decayed = 0, 0
        for r in range(world):
            S, lo, ndl = D.shard_layout(padded, n_decay, world, r)
            assert S % 8 == 0 and lo == r * S and 0 <= ndl <= S
            covered += S
            decayed += ndl
        assert covered == padded and decayed == n_decay        # every element owned once, decay region preserved
        pieces = D.tail_pieces(n_decay, numel, padded // world, world)
        assert pieces[0][1] == n_decay and pieces[-1][2] == numel
        assert all(a[2] == b[1] for a, b in zip(pieces, pieces[1:]))
        assert all(r * (padded // world) <= a and b <= (r + 1) * (padded // world) for r, a, b in pieces)
    with pytest.raises(ValueError):
        D.shard_layout(100, 10, 3, 0)


def _sharded_worker(rank, world, port, ret):
    """engine.FlatParams.sharded_step with torch CPU arithmetic in place of the two kernels (oracle AdamW = the numerics
    goat_adamw_step implements): reduce-scatter, global norm from shard sums, AdamW on the shard, all-gather, tail."""
    import torch.distributed as dist
    from oracle import goat_oracle as O
    from vln_goat_b200 import dist_utils as D
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        numel, n_decay, padded, p0, grads = _sharded_problem(world)
        p = torch.zeros(padded); p[:numel] = p0
        m, v = torch.zeros(padded), torch.zeros(padded)
        S, lo, ndl = D.shard_layout(padded, n_decay, world, rank)
        for step in (1, 2):
            g = torch.zeros(padded); g[:numel] = grads[rank] * step          # this rank's local gradient
            gs = torch.empty(S)
            D.reduce_scatter_sum(gs, g)
            tot = (gs.double() ** 2).sum().float().reshape(1)
            dist.all_reduce(tot)
            pre = 1.0 / world
            norm = tot.sqrt() * pre
            coef = pre * min(1.0, (MAXN / (norm + 1e-6)).item())
            sl = slice(lo, lo + S)
            gg = gs * coef
            if ndl > 0:
                O.adamw_step(p[lo:lo + ndl], gg[:ndl], m[lo:lo + ndl], v[lo:lo + ndl], step, LR, BETAS, 1e-6, WD)
            if ndl < S:
                O.adamw_step(p[lo + ndl:lo + S], gg[ndl:], m[lo + ndl:lo + S], v[lo + ndl:lo + S], step, LR, BETAS, 1e-6, 0.0)
            D.all_gather_flat(p, p[sl].clone())        # (the product gathers the 16-bit shadow + the fp32 tail; same layout)
        if rank == 0:
            ret.put(p[:numel].numpy().copy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


LR, BETAS, WD, MAXN = 1e-2, (0.9, 0.98), 0.01, 0.5


def _sharded_problem(world):
    g = torch.Generator().manual_seed(17)
    numel, n_decay = 1000 + 37, 800
    padded = (numel + 8 * world - 1) // (8 * world) * (8 * world)
    p0 = torch.randn(numel, generator=g)
    grads = [torch.randn(numel, generator=g) for _ in range(world)]
    return numel, n_decay, padded, p0, grads


@pytest.mark.timeout(300)
def test_sharded_step_data_flow_matches_unsharded_reference():
    import torch.multiprocessing as mp
    from oracle import goat_oracle as O
    world = 2
    numel, n_decay, padded, p0, grads = _sharded_problem(world)
    # 1-process reference: average the ranks' gradients, clip by the global norm, AdamW with decay on [0, n_decay)
    p, m, v = p0.clone(), torch.zeros(numel), torch.zeros(numel)
    for step in (1, 2):
        g = sum(gr * step for gr in grads) / world
        _, coef = O.clip_grad_norm([g], MAXN)
        g = g * coef
        O.adamw_step(p[:n_decay], g[:n_decay], m[:n_decay], v[:n_decay], step, LR, BETAS, 1e-6, WD)
        O.adamw_step(p[n_decay:], g[n_decay:], m[n_decay:], v[n_decay:], step, LR, BETAS, 1e-6, 0.0)
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, ret)) for r in range(world)]
    for q in procs:
        q.start()
    got = torch.from_numpy(ret.get(timeout=240))
    for q in procs:
        q.join(timeout=60)
        assert q.exitcode == 0
    assert (got - p).abs().max().item() < 1e-6
