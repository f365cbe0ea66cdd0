"""Training-step plumbing around the kernels: flat parameter storage, fused optimizer, CUDA-graph replay.

What the reference does per optimizer step (P/train_r2r_goat.py:301-366) and what replaces it here:

  reference                                                     here
  ------------------------------------------------------------  ------------------------------------------------
  ~600 separate parameter tensors, fp32                         FlatParams: ONE fp32 buffer (params are views), one
                                                                fp32 gradient buffer, one 16-bit operand shadow;
                                                                q/k/v weights of each attention sit back to back so
                                                                the fused-QKV GEMM reads them with no concat
  autograd accumulates every weight grad into p.grad            wgrad GEMMs / LN-backward write straight into the flat
                                                                gradient views (functional.py ``_goat_grad``)
  DDP bucketed all-reduce (P/utils/misc.py:52-58)               one NCCL all-reduce (sum) over the flat buffer; the
                                                                1/world average is folded into the optimizer kernel
  clip_grad_norm_ + per-tensor AdamW loop (P/optim/adamw.py)    goat_sumsq + goat_adamw_step (2 launches)
  ~500 eager launches per fwd+bwd                               the whole fwd+bwd is captured once in a CUDA graph
"""
import ctypes as C

import torch

from . import _lib, functional as Fn, ops

NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight")  # P/optim/misc.py:13


def _pad8(n):
    return (n + 7) // 8 * 8


def _p_or_none(t):
    return None if t is None else t.data_ptr()


def _fold_autograd_grad(p):
    if p.grad is not None:
        p._goat_grad.add_(p.grad.to(torch.float32).view_as(p._goat_grad))
        p.grad = None
        p._goat_fresh = False


class FlatParams(object):
    """Re-homes every trainable parameter of ``model`` (already on its CUDA device) into one flat fp32 buffer.

    Layout: [weight-decayed tensors | no-decay tensors], each padded to 8 elements (16-byte aligned fp32 and
    16-bit views).  ``named_parameters`` order is kept inside each group, which puts query/key/value weights
    (and, in the other group, their biases) back to back.
    """

    def __init__(self, model, shadow_dtype=None, no_decay=NO_DECAY, only=None, group=None):
        """``only``: optional collection of parameters to flatten (see ``active_parameters``); the rest of the
        model is left untouched and never updated -- the reference's AdamW likewise skips parameters whose
        ``grad`` is None (P/optim/adamw.py:66-67), which DDP's find_unused_parameters=True relies on.
        ``group``: the data-parallel process group (default: the world); the buffers are padded so that they split
        into equal 16-byte aligned shards, one per rank (``sharded_step``).

        Layout: [decayed matrices the GEMMs read through the 16-bit shadow | decayed tensors the kernels read in fp32
        (embedding tables, LayerNorm gains not called ``LayerNorm``, pooling / gate vectors, 7- and 14-wide position
        projections) | no-decay vectors (biases, ``LayerNorm.*``)].  ``n_decay`` ends the second region (weight decay);
        ``n_shadow_only`` ends the first: everything after it must be current in FP32 on every rank after a sharded
        optimizer step, everything before it only in the shadow."""
        named = []
        seen = set()
        keep = None if only is None else set(id(p) for p in only)
        for n, p in model.named_parameters():
            if p.requires_grad and id(p) not in seen and (keep is None or id(p) in keep):
                seen.add(id(p))
                named.append((n, p))
        if not named:
            raise ValueError("model has no trainable parameters")
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("FlatParams needs the model on a CUDA device (no CPU path)")
        emb = set(id(m.weight) for m in model.modules() if isinstance(m, torch.nn.Embedding))

        def fp32_read(p):
            # what functional.LinearFn / EmbedFn / the head kernels read straight from the fp32 master
            return (id(p) in emb or p.dim() != 2 or p.shape[1] % 8 != 0 or p.shape[1] < 16 or p.shape[0] < 8)
        is_nd = lambda n: any(nd in n for nd in no_decay)
        decay_w = [(n, p) for n, p in named if not is_nd(n) and not fp32_read(p)]
        decay_f = [(n, p) for n, p in named if not is_nd(n) and fp32_read(p)]
        nodecay = [(n, p) for n, p in named if is_nd(n)]
        self.names, self.params, self.offsets = [], [], []
        off = 0
        self.n_shadow_only = self.n_decay = 0
        for gi, grp in enumerate((decay_w, decay_f, nodecay)):
            for n, p in grp:
                self.names.append(n)
                self.params.append(p)
                self.offsets.append(off)
                off += _pad8(p.numel())
            if gi == 0:
                self.n_shadow_only = off
            if gi == 1:
                self.n_decay = off
        from .dist_utils import world_size
        self.group = group
        self.world = world_size(group)
        self.numel = off
        unit = 8 * self.world
        self.padded = (off + unit - 1) // unit * unit        # equal shards of a multiple of 8 elements per rank
        tot = self.padded
        self.p = torch.zeros(tot, device=dev, dtype=torch.float32)
        self.g = torch.zeros(tot, device=dev, dtype=torch.float32)
        self.m = torch.zeros(tot, device=dev, dtype=torch.float32)
        self.v = torch.zeros(tot, device=dev, dtype=torch.float32)
        self.shadow_dtype = shadow_dtype if shadow_dtype in (torch.float16, torch.bfloat16) else None
        from . import runtime
        self.split = bool(self.shadow_dtype) and runtime.weight_split()
        self.shadow = self.shadow_lo = None
        if self.shadow_dtype:
            # [hi | lo] when weights are split (runtime.set_weight_split): lo = round(p - hi), written by the optimizer kernel
            buf = torch.zeros(tot * (2 if self.split else 1), device=dev, dtype=self.shadow_dtype)
            self.shadow = buf[:tot]
            if self.split:
                self.shadow_lo = buf[tot:]
                runtime.register_split_buffer(buf, tot)
            self._shadow_buf = buf
        self._g_shard = self._x_shard = None
        self._phase_events = None          # profile_phases(True): CUDA events around the phases of sharded_step
        self._peer = None                  # None: not set up yet; False: NCCL collectives; dict: peer pointer tables
        self.master_synced = True
        self.scaler = None
        self._scaler_cfg = None
        self._hooks = []
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            self.p[o:o + n].copy_(p.detach().reshape(-1))
            p.data = self.p[o:o + n].view(p.shape)
            p.grad = None
            p._goat_grad = self.g[o:o + n].view(p.shape)
            p._goat_fresh = True
            if self.shadow is not None:
                p._goat_shadow = self.shadow[o:o + n].view(p.shape)
            # a parameter that some torch-native op consumed gets its gradient through autograd: fold it into the flat
            # buffer as soon as it is accumulated (nothing on the GOAT path should need this; it keeps foreign ops
            # correct instead of silently untrained)
            self._hooks.append(p.register_post_accumulate_grad_hook(_fold_autograd_grad))
        self.refresh_shadow()
        ws = _lib.lib().goat_sumsq_workspace_bytes()
        self._partial = torch.zeros(ws // 4, device=dev, dtype=torch.float32)
        self.grad_norm = torch.zeros(1, device=dev, dtype=torch.float32)
        self._hp = torch.zeros(9, device=dev, dtype=torch.float32)
        self.step_count = 0

    # ------------------------------------------------------------------------------------------
    # fp16 loss scaling (torch.cuda.amp.GradScaler of the reference's 16-bit path, P/train_r2r_goat.py:279,325,351-363)
    # ------------------------------------------------------------------------------------------
    def enable_loss_scale(self, init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        """Dynamic loss scale held ON THE DEVICE (``scaler``: [scale, clean steps, overflowed, skipped, steps taken]):
        multiply the loss by ``loss_scale()`` before backward; the optimizer kernel unscales, skips the update when the
        gradient norm is not finite, and ``goat_scaler_update`` moves the scale -- no host synchronisation, so the
        scaled backward can sit inside a replayed CUDA graph."""
        self.scaler = torch.tensor([init_scale, 0.0, 0.0, 0.0, 0.0], device=self.p.device, dtype=torch.float32)
        self._scaler_cfg = (float(growth_factor), float(backoff_factor), int(growth_interval))
        return self.scaler

    def loss_scale(self):
        """device scalar (a view of the scaler state) to multiply the loss with, or None when scaling is off"""
        return None if self.scaler is None else self.scaler[0]

    def _scaler_update(self, nparts, st):
        if self.scaler is not None:
            g, b, i = self._scaler_cfg
            _lib.check(_lib.lib().goat_scaler_update(self.scaler.data_ptr(), self._partial.data_ptr(), nparts, g, b, i, st),
                       "goat_scaler_update")
            ops.LAUNCHES[0] += 1

    def refresh_shadow(self):
        """Re-derive the 16-bit operand copies from the fp32 masters (after load_state_dict etc.)."""
        from . import runtime
        if self.shadow is not None:
            if self.shadow_lo is not None:
                ops.split_cast(self.p, self.shadow, self.shadow_lo)
            else:
                ops.cast(self.p, self.shadow_dtype, out=self.shadow)
        runtime.bump_generation()

    def begin_step(self):
        """Reset the written-this-step diagnostic (``unwritten()``).  It does NOT touch the gradients: every backward
        ACCUMULATES into the flat gradient buffer (functional._grad_into), the optimizer step clears it; call
        ``zero_grad()`` to drop gradients without stepping.  Several backward passes between two optimizer steps sum,
        like ``.grad`` under the reference's gradient_accumulation_steps (P/train_r2r_goat.py:322-327)."""
        for p in self.params:
            p._goat_fresh = True

    def zero_grad(self):
        self.g.zero_()
        self.begin_step()

    def unwritten(self):
        return [n for n, p in zip(self.names, self.params) if p._goat_fresh]

    def grads_to_autograd(self):
        """Expose the flat gradient views as ``param.grad`` (for code that inspects grads the torch way)."""
        for p in self.params:
            p.grad = p._goat_grad

    def all_reduce(self, group=None):
        """Sum the flat gradient over ranks (the average is applied by the optimizer kernel's pre-scale)."""
        from .dist_utils import all_reduce_sum_
        return all_reduce_sum_(self.g, group)

    # ------------------------------------------------------------------------------------------
    # sharded optimizer step (world > 1): reduce-scatter -> clip + AdamW on this rank's 1/world of the buffer ->
    # all-gather of what the next forward reads.  Same arithmetic as all_reduce() + adamw_step(grad_scale=1/world)
    # -- every element is updated by exactly one rank from the same summed gradient and the same global norm -- at
    # 3/4 of the all-reduce's bytes (fp32 reduce-scatter + 16-bit all-gather) and 1/world of the optimizer's HBM pass.
    # ------------------------------------------------------------------------------------------
    def _hp_upload(self, lr, betas, eps, weight_decay, max_grad_norm, grad_scale, correct_bias):
        self.step_count += 1
        t = self.step_count
        # pageable source on purpose: the runtime stages a small pageable H2D copy before returning, so the
        # host tensor can be dropped / rebuilt next step without racing the DMA (a reused pinned buffer could)
        h = torch.tensor([lr, betas[0], betas[1], eps, weight_decay,
                          (1.0 - betas[0] ** t) if correct_bias else 1.0,
                          (1.0 - betas[1] ** t) if correct_bias else 1.0, max_grad_norm, grad_scale], dtype=torch.float32)
        self._hp.copy_(h, non_blocking=True)

    def profile_phases(self, on=True):
        """Record CUDA events around the phases of ``sharded_step`` (reduce-scatter, norm, AdamW, all-gather, fp32 tail);
        ``phase_times()`` synchronises and returns the mean milliseconds per phase."""
        self._phase_events = [] if on else None

    def _phase(self, name):
        if self._phase_events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._phase_events.append((name, e))

    def phase_times(self):
        ev = self._phase_events or []
        torch.cuda.synchronize()
        tot, cnt = {}, {}
        for (n0, e0), (n1, e1) in zip(ev[:-1], ev[1:]):
            if n1 == "begin":
                continue
            tot[n1] = tot.get(n1, 0.0) + e0.elapsed_time(e1)
            cnt[n1] = cnt.get(n1, 0) + 1
        return {k: tot[k] / cnt[k] for k in tot}

    def sharded_step(self, lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, max_grad_norm=-1.0, correct_bias=True):
        """One data-parallel optimizer step with the optimizer state's work split over the ranks (see above).
        Afterwards every rank holds the new 16-bit operand shadow and the new fp32 no-decay vectors (biases,
        LayerNorm, embedding tables, every tensor past ``n_shadow_only``); the fp32 master MATRICES that are only ever
        read through the shadow are current only on their owner rank until ``sync_master()``
        (call it before ``state_dict()`` / checkpointing).  Without a 16-bit shadow (fp32 compute) the whole fp32
        buffer is all-gathered every step instead."""
        import torch.distributed as dist
        W = self.world
        if W <= 1:
            return self.adamw_step(lr, betas, eps, weight_decay, max_grad_norm, 1.0, correct_bias)
        from . import dist_utils as D
        rank = dist.get_rank(self.group)
        S, lo, n_decay_local = D.shard_layout(self.padded, self.n_decay, W, rank)
        if self._g_shard is None:
            self._g_shard = torch.zeros(S, device=self.p.device, dtype=torch.float32)
        if self._peer is None:
            self._peer_setup()
        if self._peer:
            return self._sharded_step_peers(S, lo, rank, lr, betas, eps, weight_decay, max_grad_norm, correct_bias)
        if self._x_shard is None:
            self._x_shard = torch.zeros(S, device=self.p.device, dtype=self.shadow_dtype or torch.float32)
        self._phase("begin")
        D.reduce_scatter_sum(self._g_shard, self.g, self.group)
        self._phase("reduce_scatter")
        self.g.zero_()                      # the next step's gradients accumulate into a cleared buffer
        self._hp_upload(lr, betas, eps, weight_decay, max_grad_norm, 1.0 / W, correct_bias)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        nparts = C.c_int(0)
        L = _lib.lib()
        _lib.check(L.goat_sumsq(self._g_shard.data_ptr(), S, self._partial.data_ptr(), C.byref(nparts), st), "goat_sumsq")
        tot = self._partial[:nparts.value].sum().reshape(1)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)     # global squared norm of the summed gradient
        self._partial[:1].copy_(tot)
        self._phase("zero_and_norm")
        sd = ops.dt(self.shadow_dtype) if self.shadow is not None else 0
        _lib.check(L.goat_adamw_step(self.p[lo:lo + S].data_ptr(), self._g_shard.data_ptr(), self.m[lo:lo + S].data_ptr(),
                                     self.v[lo:lo + S].data_ptr(),
                                     self.shadow[lo:lo + S].data_ptr() if self.shadow is not None else None, sd, S,
                                     n_decay_local, self._hp.data_ptr(), self._partial.data_ptr(), 1,
                                     self.grad_norm.data_ptr(), 0, _p_or_none(self.scaler),
                                     self.shadow_lo[lo:lo + S].data_ptr() if self.shadow_lo is not None else None, st),
                   "goat_adamw_step")
        ops.LAUNCHES[0] += 2
        self._scaler_update(1, st)
        self._phase("adamw_shard")
        if self.shadow is not None:
            self._x_shard.copy_(self.shadow[lo:lo + S])
            D.all_gather_flat(self.shadow, self._x_shard, self.group)
            if self.shadow_lo is not None:
                self._x_shard.copy_(self.shadow_lo[lo:lo + S])
                D.all_gather_flat(self.shadow_lo, self._x_shard, self.group)
            self._phase("all_gather_shadow")
            # everything the kernels read in FP32 (embedding tables, LayerNorm gains, pooling / gate vectors, biases):
            # each owner broadcasts its piece of [n_shadow_only, numel)
            for r, a, b in D.tail_pieces(self.n_shadow_only, self.numel, S, W):
                dist.broadcast(self.p[a:b], src=D.group_src(r, self.group), group=self.group)
            self._phase("broadcast_fp32_tail")
            self.master_synced = False
        else:
            self._x_shard.copy_(self.p[lo:lo + S])
            D.all_gather_flat(self.p, self._x_shard, self.group)
        from . import runtime
        runtime.bump_generation()

    # ------------------------------------------------------------------------------------------
    # the same step over NVLink peer memory (csrc/exchange.cu): every rank maps the other ranks' p / g / shadow buffers
    # (CUDA IPC), reads its shard of everybody's gradients in the reduce kernel and stores the updated values straight
    # into everybody's buffers from the AdamW kernel.  Two kernels + three scalar all-reduces (norm / barriers) replace
    # reduce-scatter, staging copies, two all-gathers and the tail broadcasts.
    # ------------------------------------------------------------------------------------------
    def _peer_setup(self):
        """Collective (first sharded step).  Falls back to the NCCL collectives -- on every rank -- when the exchange
        is switched off (GOAT_PEER_EXCHANGE=0), the backend is not NCCL, the world exceeds GOAT_MAX_PEERS or a buffer
        cannot be exported / mapped."""
        import os
        import sys
        import torch.distributed as dist
        L = _lib.lib()
        W = self.world
        rank = dist.get_rank(self.group)
        sig = torch.zeros(L.goat_peer_signal_bytes() // 4, device=self.p.device, dtype=torch.float32)
        torch.cuda.synchronize()             # the signal block is zero before any rank can learn its address
        bufs = {"p": self.p, "g": self.g, "sig": sig}
        if self.shadow is not None:
            bufs["s"] = self._shadow_buf
        mine, why = {}, None
        if os.environ.get("GOAT_PEER_EXCHANGE", "1") == "0":
            mine, why = None, "GOAT_PEER_EXCHANGE=0"
        elif dist.get_backend(self.group) != "nccl" or W > _lib.MAX_PEERS:
            mine, why = None, "backend %s, world %d" % (dist.get_backend(self.group), W)
        else:
            for k, t in bufs.items():
                h = (C.c_ubyte * _lib.PEER_HANDLE_BYTES)()
                off = C.c_ulonglong(0)
                if L.goat_peer_export(C.c_void_p(t.data_ptr()), h, C.byref(off)) != 0:
                    mine, why = None, L.goat_last_error().decode()
                    break
                mine[k] = (bytes(h), int(off.value))
        every = [None] * W
        dist.all_gather_object(every, mine, group=self.group)
        ptrs, opened, ok = {k: [] for k in bufs}, {}, all(e is not None for e in every)
        if ok:
            for q in range(W):
                for k, t in bufs.items():
                    if q == rank:
                        ptrs[k].append(t.data_ptr())
                        continue
                    hb, off = every[q][k]
                    if hb not in opened:
                        base = C.c_void_p(0)
                        if L.goat_peer_open(hb, C.byref(base)) != 0:
                            ok, why = False, L.goat_last_error().decode()
                            break
                        opened[hb] = base.value
                    ptrs[k].append(opened[hb] + off)
                if not ok:
                    break
        flags = [None] * W
        dist.all_gather_object(flags, bool(ok), group=self.group)
        if not all(flags):
            for base in opened.values():
                L.goat_peer_close(C.c_void_p(base))
            if rank == 0 and os.environ.get("GOAT_PEER_EXCHANGE", "1") != "0":
                sys.stderr.write("vln_goat_b200: peer-memory gradient exchange unavailable (%s); using NCCL collectives\n"
                                 % (why or "another rank could not map the buffers"))
            self._peer = False
            return
        arr = lambda xs: (C.c_void_p * W)(*xs)
        esz = self.shadow.element_size() if self.shadow is not None else 0
        self._peer = {"p": arr(ptrs["p"]), "g": arr(ptrs["g"]), "opened": opened, "sig": arr(ptrs["sig"]), "sig_t": sig,
                      "epoch": 0, "side": torch.cuda.Stream(device=self.p.device),
                      "nccl_sync": os.environ.get("GOAT_PEER_SYNC", "flags") == "nccl",
                      "s": arr(ptrs["s"]) if self.shadow is not None else None,
                      "slo": arr([b + self.padded * esz for b in ptrs["s"]]) if self.shadow_lo is not None else None}
        self._sync = torch.zeros(1, device=self.p.device, dtype=torch.float32)

    def release_peers(self):
        """Unmap the other ranks' buffers (collective-free; call before the process group is destroyed)."""
        if self._peer:
            torch.cuda.synchronize()
            for base in self._peer["opened"].values():
                _lib.lib().goat_peer_close(C.c_void_p(base))
        self._peer = None

    def _peer_barrier(self, st, rank):
        P = self._peer
        if P["nccl_sync"]:
            import torch.distributed as dist
            dist.all_reduce(self._sync, group=self.group)
            return
        P["epoch"] += 1
        _lib.check(_lib.lib().goat_peer_barrier(P["sig"], self.world, rank, P["epoch"] & 0xFFFFFFFF, st), "goat_peer_barrier")
        ops.LAUNCHES[0] += 1

    def _sharded_step_peers(self, S, lo, rank, lr, betas, eps, weight_decay, max_grad_norm, correct_bias):
        L, P, W = _lib.lib(), self._peer, self.world
        cur = torch.cuda.current_stream()
        st = C.c_void_p(cur.cuda_stream)
        self._hp_upload(lr, betas, eps, weight_decay, max_grad_norm, 1.0 / W, correct_bias)
        self._phase("begin")
        self._peer_barrier(st, rank)                 # every rank's backward is done: its gradients may be read
        self._phase("barrier_backward_done")
        nparts = C.c_int(0)
        _lib.check(L.goat_peer_reduce_sumsq(P["g"], W, lo, S, self._g_shard.data_ptr(), self._partial.data_ptr(),
                                            C.byref(nparts), st), "goat_peer_reduce_sumsq")
        self._phase("peer_reduce_sumsq")
        # global squared norm, the same bits on every rank; also the barrier "everybody has read my gradients"
        if P["nccl_sync"]:
            import torch.distributed as dist
            tot = self._partial[:nparts.value].sum().reshape(1)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)
            self._partial[:1].copy_(tot)
        else:
            P["epoch"] += 1
            _lib.check(L.goat_peer_sum_scalar(P["sig"], W, rank, P["epoch"] & 0xFFFFFFFF, self._partial.data_ptr(),
                                              nparts.value, self._partial.data_ptr(), st), "goat_peer_sum_scalar")
        self._phase("norm_exchange")
        P["side"].wait_stream(cur)                   # clearing the gradient buffer (HBM) runs beside the AdamW kernel,
        with torch.cuda.stream(P["side"]):           # whose time at larger worlds is the NVLink stores
            self.g.zero_()
        sd = ops.dt(self.shadow_dtype) if self.shadow is not None else 0
        _lib.check(L.goat_adamw_step_peers(P["p"], P["s"], P["slo"], sd, W, rank, self._g_shard.data_ptr(),
                                           self.m.data_ptr(), self.v.data_ptr(), lo, S, self.n_decay,
                                           self.n_shadow_only if self.shadow is not None else 0, self._hp.data_ptr(),
                                           self._partial.data_ptr(), 1, self.grad_norm.data_ptr(), _p_or_none(self.scaler), st),
                   "goat_adamw_step_peers")
        ops.LAUNCHES[0] += 3
        self._scaler_update(1, st)
        cur.wait_stream(P["side"])
        self._phase("adamw_peers_and_zero")
        self._peer_barrier(st, rank)                 # every rank's stores have landed before anybody's next forward
        self._phase("barrier_stores_done")
        if self.shadow is not None and self.numel > self.n_shadow_only:
            # the fp32-read region arrived as fp32 (4 B per element like hi + lo elsewhere): its operand copies are local work
            a, n = self.n_shadow_only, self.numel - self.n_shadow_only
            _lib.check(L.goat_split_cast(self.p[a:].data_ptr(), self.shadow[a:].data_ptr(),
                                         self.shadow_lo[a:].data_ptr() if self.shadow_lo is not None else None, sd, n, st),
                       "goat_split_cast")
            ops.LAUNCHES[0] += 1
            self._phase("split_cast_fp32_region")
        if self.shadow is not None:
            self.master_synced = False
        from . import runtime
        runtime.bump_generation()

    def sync_master(self):
        """All-gather the fp32 master weights after sharded steps (every rank then holds the full, identical fp32
        parameters: call before state_dict() / checkpointing)."""
        import torch.distributed as dist
        if self.world > 1 and not self.master_synced:
            from . import dist_utils as D
            S, lo, _ = D.shard_layout(self.padded, self.n_decay, self.world, dist.get_rank(self.group))
            D.all_gather_flat(self.p, self.p[lo:lo + S].clone(), self.group)
            self.master_synced = True

    def adamw_step(self, lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, max_grad_norm=-1.0, grad_scale=1.0,
                   correct_bias=True):
        """clip_grad_norm_(max_grad_norm) + AdamW (P/optim/adamw.py:85-110 numerics) on the flat buffer; the gradient
        buffer is cleared in the same pass (weight-gradient GEMMs accumulate into it, see functional._wgrad)."""
        self._hp_upload(lr, betas, eps, weight_decay, max_grad_norm, grad_scale, correct_bias)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        nparts = C.c_int(0)
        L = _lib.lib()
        _lib.check(L.goat_sumsq(self.g.data_ptr(), self.numel, self._partial.data_ptr(), C.byref(nparts), st), "goat_sumsq")
        sd = ops.dt(self.shadow_dtype) if self.shadow is not None else 0
        _lib.check(L.goat_adamw_step(self.p.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                     self.shadow.data_ptr() if self.shadow is not None else None, sd, self.numel,
                                     self.n_decay, self._hp.data_ptr(), self._partial.data_ptr(), nparts.value,
                                     self.grad_norm.data_ptr(), 1, _p_or_none(self.scaler), _p_or_none(self.shadow_lo), st),
                   "goat_adamw_step")
        ops.LAUNCHES[0] += 2
        self._scaler_update(nparts.value, st)
        from . import runtime
        runtime.bump_generation()


def active_parameters(model, loss_fn, inputs):
    """One eager forward+backward through plain autograd; returns the parameters that received a gradient
    (the set a given task / mode actually trains -- e.g. the pretrain-only ``lang_*`` blocks of BertCrossLayer,
    P/model/Bert_backbone.py:673-676, take no part in the SAP / CFP / navigation forward)."""
    for p in model.parameters():
        p.grad = None
    loss = loss_fn(*inputs)
    loss.backward()
    torch.cuda.synchronize()
    act = [p for p in model.parameters() if p.grad is not None]
    for p in model.parameters():
        p.grad = None
    return act


def warmup_linear(step, warmup_step, tot_step):
    """BERT schedule, P/optim/sched.py:17-21"""
    if step < warmup_step:
        return step / warmup_step
    return max(0, (tot_step - step) / (tot_step - warmup_step))


def _map_inputs(inputs, fn):
    if isinstance(inputs, dict):
        return {k: fn(v) for k, v in inputs.items()}
    return [fn(v) for v in inputs]


def _input_items(inputs):
    return list(inputs.items()) if isinstance(inputs, dict) else list(enumerate(inputs))


def input_signature(inputs):
    """Shape / dtype signature of a set of step inputs: the key of the captured-graph cache."""
    return tuple((k, tuple(v.shape), str(v.dtype)) for k, v in _input_items(inputs))


class _Captured(object):
    """forward + backward of one (loss_fn, input signature) pair, captured in a CUDA graph over static input buffers"""
    __slots__ = ("loss_fn", "static_inputs", "graph", "loss", "launches")


class TrainStep(object):
    """One optimizer step = forward + backward (captured in a CUDA graph) + gradient exchange + fused AdamW.

    ``loss_fn(*static_inputs)`` (list inputs) or ``loss_fn(static_inputs)`` (dict inputs) -> scalar loss tensor; it must
    be shape-static.  ``step(new_inputs)`` takes tensors of the same shapes (device or pinned host; they are copied into
    the captured input buffers).  Several (loss_fn, shapes) pairs can share one TrainStep -- the pretraining loop
    alternates MLM / SAP / CFP batches whose padded shapes fall into a few buckets (P/train_r2r_goat.py:301-314):
    ``capture(key, loss_fn, example_inputs)`` adds a graph, ``step(inputs, key)`` replays it.
    """

    SEED_STRIDE = 16     # a block uses host seeds seed .. seed+3 for its dropout sites: advance past all of them per step

    def __init__(self, flat, loss_fn=None, example_inputs=None, lr=5e-5, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01,
                 max_grad_norm=5.0, use_graph=True, warmup_iters=2, shard_optimizer=True, check_unwritten=True):
        """``shard_optimizer``: at world size > 1 use FlatParams.sharded_step (reduce-scatter, 1/world of the AdamW
        pass per rank, all-gather of the operand shadow) instead of all-reduce + a full AdamW pass on every rank.
        If ``flat.enable_loss_scale()`` was called, the loss is multiplied by the device-resident scale before
        backward and the optimizer kernel unscales / skips on overflow."""
        self.flat = flat
        self.shard_optimizer = shard_optimizer
        self.use_graph = use_graph
        self.warmup_iters = warmup_iters
        self.check_unwritten = check_unwritten
        self.opt = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        dev = flat.p.device
        self.seed = torch.zeros(1, device=dev, dtype=torch.int64)
        Fn.set_seed_ptr(self.seed)
        self.entries = {}
        self.loss = None
        self.launches_per_step = None
        if loss_fn is not None:
            self.capture(None, loss_fn, example_inputs)

    # compatibility with the single-graph form
    @property
    def static_inputs(self):
        return self.entries[None].static_inputs

    @property
    def graph(self):
        return self.entries[None].graph

    def has(self, key):
        return key in self.entries

    def capture(self, key, loss_fn, example_inputs):
        from . import runtime
        flat = self.flat
        e = _Captured()
        e.loss_fn = loss_fn
        dev = flat.p.device
        e.static_inputs = _map_inputs(example_inputs, lambda t: t.clone() if t.is_cuda else t.to(dev))
        e.graph = None
        e.loss = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup_iters):
                self._fwd_bwd(e)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if self.check_unwritten:
            missing = flat.unwritten()
            if missing:
                raise RuntimeError("parameters without a gradient in the captured step: %s" % missing[:8])
        # operand casts made by the warm-up passes must not be reused by the capture (they would be baked in as
        # constants); a new generation makes every cast part of the graph
        runtime.bump_generation()
        n0 = ops.LAUNCHES[0]
        if self.use_graph:
            e.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(e.graph):
                self._fwd_bwd(e)
        else:
            self._fwd_bwd(e)
        e.launches = ops.LAUNCHES[0] - n0 + 2 + (1 if flat.scaler is not None else 0)  # + sumsq + adamw (+ scaler)
        self.launches_per_step = e.launches
        torch.cuda.synchronize()
        flat.zero_grad()   # drop what the warm-up / capture passes accumulated
        self.entries[key] = e
        return e

    def _fwd_bwd(self, e):
        self.flat.begin_step()
        if isinstance(e.static_inputs, dict):
            loss = e.loss_fn(e.static_inputs)
        else:
            loss = e.loss_fn(*e.static_inputs)
        scale = self.flat.loss_scale()
        (loss if scale is None else loss * scale).backward()
        self.seed.add_(self.SEED_STRIDE)
        e.loss = loss.detach()

    def load_inputs(self, inputs, key=None):
        e = self.entries[key]
        if isinstance(e.static_inputs, dict):
            for k, dst in e.static_inputs.items():
                dst.copy_(inputs[k], non_blocking=True)
        else:
            for dst, src in zip(e.static_inputs, inputs):
                dst.copy_(src, non_blocking=True)

    def step(self, inputs=None, key=None):
        e = self.entries[key]
        if inputs is not None:
            self.load_inputs(inputs, key)
        if e.graph is not None:
            e.graph.replay()
        else:
            self._fwd_bwd(e)
        self.launches_per_step = e.launches
        self.optimizer_step()
        self.loss = e.loss
        return e.loss

    def optimizer_step(self):
        if self.shard_optimizer and self.flat.world > 1:
            self.flat.sharded_step(**self.opt)
        else:
            world = self.flat.all_reduce(self.flat.group)
            self.flat.adamw_step(grad_scale=1.0 / world, **self.opt)


__all__ = ["FlatParams", "TrainStep", "active_parameters", "warmup_linear", "input_signature", "NO_DECAY"]
