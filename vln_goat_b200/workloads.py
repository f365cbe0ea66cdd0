"""The BASELINE.json workloads as small nn.Modules built from the drop-in blocks.

C2 ("full 9-layer GOAT cross-encoder fwd+bwd, batch=64"): the local branch of
GlocalTextPathCMT.forward -- LanguageEncoder (6 RobertaLayers) over the instruction tokens, then
LocalVPEncoder.encoder = CrossmodalEncoder (3 BertCrossLayers) with the [stop]+36 view tokens as
queries and the text as keys/values (P/model/vilmodel_goat.py:563-564 and :399).  Parameter names
follow the reference model (``lang_encoder.layer.N.*``, ``local_encoder.encoder.crossattention.N.*``).
"""
import torch
from torch import nn

from . import modules as M


class _HalfMeanSquare(torch.autograd.Function):
    """0.5 * mean(x^2) as one dot product forward and one scaled copy backward (the plain torch expression costs eight
    elementwise / reduction launches over the output streams per step)."""

    @staticmethod
    def forward(ctx, x):
        xf = x.reshape(-1)
        ctx.save_for_backward(x)
        return torch.dot(xf, xf) * (0.5 / xf.numel())

    @staticmethod
    def backward(ctx, go):
        x, = ctx.saved_tensors
        return x * (go * (1.0 / x.numel()))


def c2_loss(txt_out, vp_out):
    """The synthetic objective of the C2 workload (oracle.goat_oracle.c2_loss): mean square of both output streams."""
    return _HalfMeanSquare.apply(txt_out) + _HalfMeanSquare.apply(vp_out)


class _LocalBranch(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.encoder = M.CrossmodalEncoder(config)


class C2CrossEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.lang_encoder = M.LanguageEncoder(config)
        self.local_encoder = _LocalBranch(config)

    def forward(self, txt_embeds, txt_masks, vp_embeds, vp_masks):
        """txt_embeds [B,L,H], txt_masks bool [B,L], vp_embeds [B,Nq,H], vp_masks bool [B,Nq]
        -> (txt_out [B,L,H], vp_out [B,Nq,H])"""
        t = self.lang_encoder.run(txt_embeds, txt_masks)
        v = self.local_encoder.encoder.run(vp_embeds, vp_masks, t, txt_masks)
        return t.tensor(), v.tensor()


# --------------------------------------------------------------------------------------
# C3: the full pretraining step (GlocalTextPathCMTPreTraining, tasks MLM / SAP / CFP) on synthetic batches with the
# layout of the reference's collates (P/data/tasks.py:110-166, :392-451, :618-677)
# --------------------------------------------------------------------------------------
def _loc_fts(n, g):
    import math
    h = (torch.rand(n, generator=g) * 2 - 1) * math.pi
    e = torch.rand(n, generator=g) - 0.5
    one = torch.ones(n)
    return torch.stack([torch.sin(h), torch.cos(h), torch.sin(e), torch.cos(e), one, one, one], 1)


def synthetic_pretrain_batch(B=64, L=80, seed=0, views=36, max_steps=5, H=768):
    """One collated pretraining batch (host tensors + the Python lists of viewpoint ids the reference carries):
    per sample a trajectory of 1..max_steps panoramas of ``views`` N(0,1) feature vectors (stand-in for CLIP ViT-B/16),
    3..6 candidate views per step (one leads back to the previous node, one to the next node of the path, the rest to
    unvisited nodes, one of which is seen from every step), the global map [None] + visited + unvisited with symmetric
    U(0,10) pair distances, an instruction of L/2..L tokens with 15 % masked (MLM labels), and SAP labels pointing at an
    unvisited candidate of the current panorama (every third sample: stop)."""
    g = torch.Generator().manual_seed(seed)
    txt_lens = torch.randint(L // 2, L + 1, (B,), generator=g)
    txt_lens[0] = L
    txt_ids = torch.zeros(B, L, dtype=torch.int64)
    txt_labels = torch.full((B, L), -1, dtype=torch.int64)
    for i in range(B):
        n = int(txt_lens[i])
        txt_ids[i, :n] = torch.randint(3, 50000, (n,), generator=g)
        txt_ids[i, 0] = 0
        txt_ids[i, n - 1] = 2
        k = max(1, int(0.15 * n))
        pos = torch.randperm(n - 2, generator=g)[:k] + 1
        txt_labels[i, pos] = txt_ids[i, pos]
        txt_ids[i, pos] = 50264
    step_lens = [int(x) for x in torch.randint(1, max_steps + 1, (B,), generator=g)]
    S = sum(step_lens)
    view_lens = torch.full((S,), views, dtype=torch.int64)
    feats = torch.randn(S, views, H, generator=g)
    loc = _loc_fts(S * views, g).view(S, views, 7)
    nav_types = torch.zeros(S, views, dtype=torch.int64)
    n_extra = torch.randint(1, 4, (S,), generator=g).tolist()
    traj_vpids, traj_cand_vpids, gmap_vpids, n_cands = [], [], [], []
    s = 0
    for i, n in enumerate(step_lens):
        path = ["s%d_v%d" % (i, t) for t in range(n)]
        shared = "s%d_u_shared" % i
        cands_i = []
        for t in range(n):
            c = []
            if t > 0:
                c.append(path[t - 1])
            if t < n - 1:
                c.append(path[t + 1])
            c.append(shared)
            c.extend("s%d_u%d_%d" % (i, t, k) for k in range(n_extra[s + t] + (2 if n == 1 else 0)))
            c = c[:6]
            cands_i.append(c)
            nav_types[s + t, :len(c)] = 1
            n_cands.append(len(c))
        s += n
        seen = set(path)
        unvisited = []
        for c in cands_i:
            for vp in c:
                if vp not in seen:
                    seen.add(vp)
                    unvisited.append(vp)
        traj_vpids.append(path)
        traj_cand_vpids.append(cands_i)
        gmap_vpids.append([None] + path + unvisited)
    gmap_lens = torch.tensor([len(v) for v in gmap_vpids])
    Gm = int(gmap_lens.max())
    gmap_step_ids = torch.zeros(B, Gm, dtype=torch.int64)
    gmap_visited = torch.zeros(B, Gm, dtype=torch.bool)
    gmap_pos = torch.zeros(B, Gm, 7)
    pair = torch.zeros(B, Gm, Gm)
    global_lab = torch.zeros(B, dtype=torch.int64)
    local_lab = torch.zeros(B, dtype=torch.int64)
    for i, n in enumerate(step_lens):
        gl = int(gmap_lens[i])
        gmap_step_ids[i, 1:1 + n] = torch.arange(1, n + 1)
        gmap_visited[i, 1:1 + n] = True
        p = _loc_fts(gl, g)
        p[:, 4:] = torch.rand(gl, 3, generator=g)
        gmap_pos[i, :gl] = p
        d = torch.rand(gl, gl, generator=g) * 10
        d = (d + d.t()) / 2
        d.fill_diagonal_(0)
        d[0, :] = 0
        d[:, 0] = 0
        pair[i, :gl, :gl] = d
        if i % 3 != 2:
            tgt = traj_cand_vpids[i][-1][-1]
            global_lab[i] = gmap_vpids[i].index(tgt)
            local_lab[i] = 1 + traj_cand_vpids[i][-1].index(tgt)
    vp_pos = torch.zeros(B, views + 1, 14)
    s = 0
    for i, n in enumerate(step_lens):
        s += n
        vp_pos[i, :, :7] = _loc_fts(1, g)
        nc = n_cands[s - 1]
        vp_pos[i, 1:1 + nc, 7:] = _loc_fts(nc, g)
    return {
        "txt_ids": txt_ids, "txt_lens": txt_lens, "txt_labels": txt_labels,
        "traj_view_img_fts": feats, "traj_loc_fts": loc, "traj_nav_types": nav_types, "traj_step_lens": step_lens,
        "traj_vp_view_lens": view_lens, "traj_vpids": traj_vpids, "traj_cand_vpids": traj_cand_vpids,
        "gmap_vpids": gmap_vpids, "gmap_lens": gmap_lens, "gmap_step_ids": gmap_step_ids, "gmap_pos_fts": gmap_pos,
        "gmap_pair_dists": pair, "gmap_visited_masks": gmap_visited, "vp_pos_fts": vp_pos,
        "global_act_labels": global_lab, "local_act_labels": local_lab, "extra_heads": [True] * B,
    }


PRETRAIN_TASKS = ("mlm", "sap", "cfp")
