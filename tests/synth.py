"""Seeded synthetic inputs with the reference's batch layouts: used both by
tests/golden/make_golden.py (fed to the unmodified reference on CPU) and by the GPU parity tests (fed to the
CUDA modules), so the two sides see identical tensors.  Pure torch-CPU; nothing here imports the reference."""
import math

import torch


def _loc_fts(n, g):
    h = (torch.rand(n, generator=g) * 2 - 1) * math.pi
    e = (torch.rand(n, generator=g) - 0.5) * 1.0
    one = torch.ones(n)
    return torch.stack([torch.sin(h), torch.cos(h), torch.sin(e), torch.cos(e), one, one, one], 1)


def pretrain_batch(B=3, L=24, seed=0, views=36, ragged_views=True):
    """Batch dict of P/data/tasks.py collates (mlm / sap / cfp share the layout).  Trajectories of 1..3 steps on a
    small random graph; gmap = [None] + visited + unvisited; some unvisited nodes are seen from two steps (their
    feature is the mean of two candidate views) and later steps see already-visited neighbours (backtrack logit)."""
    g = torch.Generator().manual_seed(seed)
    H = 768
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
    step_lens = [int(x) for x in torch.randint(1, 4, (B,), generator=g)]
    step_lens[0] = 3
    S = sum(step_lens)
    view_lens = torch.full((S,), views, dtype=torch.int64)
    if ragged_views:
        view_lens[torch.randperm(S, generator=g)[: max(1, S // 3)]] = views - 5
        last = 0
        for n in step_lens:                       # keep one current panorama at full length (max_vp_len = views + 1)
            last += n
        view_lens[step_lens[0] - 1] = views
    feats = torch.randn(S, views, H, generator=g)
    loc = torch.stack([_loc_fts(views, g) for _ in range(S)], 0)
    nav_types = torch.zeros(S, views, dtype=torch.int64)
    traj_vpids, traj_cand_vpids, gmap_vpids = [], [], []
    s = 0
    n_cands = []
    for i, n in enumerate(step_lens):
        path = ["s%d_v%d" % (i, t) for t in range(n)]
        cands_i = []
        shared = "s%d_u_shared" % i
        for t in range(n):
            c = []
            if t > 0:
                c.append(path[t - 1])                      # visited neighbour -> backtrack logit
            if t < n - 1:
                c.append(path[t + 1])                      # next node (unvisited now, visited later)
            c.append(shared)                               # seen from every step -> mean of several views
            for k in range(int(torch.randint(1, 3, (1,), generator=g))):
                c.append("s%d_u%d_%d" % (i, t, k))
            cands_i.append(c)
            nav_types[s + t, :len(c)] = 1
            n_cands.append(len(c))
        s += n
        visited = list(path)
        unvisited = []
        for c in cands_i:
            for vp in c:
                if vp not in visited and vp not in unvisited:
                    unvisited.append(vp)
        traj_vpids.append(path)
        traj_cand_vpids.append(cands_i)
        gmap_vpids.append([None] + visited + unvisited)
    for k in range(S):
        feats[k, int(view_lens[k]):] = 0
        loc[k, int(view_lens[k]):] = 0
    gmap_lens = torch.tensor([len(v) for v in gmap_vpids])
    G = int(gmap_lens.max())
    gmap_step_ids = torch.zeros(B, G, dtype=torch.int64)
    gmap_visited = torch.zeros(B, G, dtype=torch.bool)
    gmap_pos = torch.zeros(B, G, 7)
    pair = torch.zeros(B, G, G)
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
        if i % 3 == 2:
            global_lab[i], local_lab[i] = 0, 0             # stop
        else:
            tgt = traj_cand_vpids[i][-1][-1]               # an unvisited candidate of the current panorama
            global_lab[i] = gmap_vpids[i].index(tgt)
            local_lab[i] = 1 + traj_cand_vpids[i][-1].index(tgt)
    last_rows = []
    s = 0
    for n in step_lens:
        s += n
        last_rows.append(s - 1)
    max_vp = int(view_lens[last_rows].max()) + 1
    vp_pos = torch.zeros(B, max_vp, 14)
    for i, r in enumerate(last_rows):
        vp_pos[i, :, :7] = _loc_fts(1, g)
        nc = n_cands[r]
        vp_pos[i, 1:1 + nc, 7:] = _loc_fts(nc, g)
    return {
        "txt_ids": txt_ids, "txt_lens": txt_lens, "txt_labels": txt_labels,
        "traj_view_img_fts": feats, "traj_loc_fts": loc, "traj_nav_types": nav_types, "traj_step_lens": step_lens,
        "traj_vp_view_lens": view_lens, "traj_vpids": traj_vpids, "traj_cand_vpids": traj_cand_vpids,
        "gmap_vpids": gmap_vpids, "gmap_lens": gmap_lens, "gmap_step_ids": gmap_step_ids, "gmap_pos_fts": gmap_pos,
        "gmap_pair_dists": pair, "gmap_visited_masks": gmap_visited, "vp_pos_fts": vp_pos,
        "global_act_labels": global_lab, "local_act_labels": local_lab, "extra_heads": [True] * B,
    }


def batch_to(batch, device):
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in batch.items()}


def nav_inputs(B=3, L=20, seed=0, views=36, G=9):
    """Per-step inputs of the fine-tune modes (M/r2r/agent.py:38-304, M/utils/efficiency_count.py:16-109):
    -> (language batch, panorama batch, navigation batch without txt_embeds / pano-derived tensors)."""
    g = torch.Generator().manual_seed(seed)
    H = 768
    txt_lens = torch.randint(L // 2, L + 1, (B,), generator=g)
    txt_lens[0] = L
    txt_ids = torch.ones(B, L, dtype=torch.int64)
    for i in range(B):
        n = int(txt_lens[i])
        txt_ids[i, :n] = torch.randint(3, 50000, (n,), generator=g)
    txt_masks = torch.arange(L)[None, :] < txt_lens[:, None]

    def pz(n):
        p = torch.rand(B, n, 1, generator=g, dtype=torch.float64)
        return p / p.sum(1, keepdim=True)
    lang = {"txt_ids": txt_ids, "txt_masks": txt_masks,
            "instr_z_direction_features": torch.randn(B, 35, H, generator=g), "instr_z_direction_pzs": pz(35),
            "instr_z_landmark_features": torch.randn(B, 39, H, generator=g), "instr_z_landmark_pzs": pz(39),
            "front_txt_feats": torch.tanh(torch.randn(B, 24, H, generator=g))}
    view_lens = torch.full((B,), views, dtype=torch.int64)
    if B > 1:
        view_lens[1] = views - 4
    feats = torch.randn(B, views, H, generator=g)
    loc = torch.stack([_loc_fts(views, g) for _ in range(B)], 0)
    nav_types = torch.zeros(B, views, dtype=torch.int64)
    cand_vpids = []
    for i in range(B):
        nc = 3 + i % 3
        nav_types[i, :nc] = 1
        feats[i, int(view_lens[i]):] = 0
        loc[i, int(view_lens[i]):] = 0
        cand_vpids.append(["n%d_c%d" % (i, k) for k in range(nc)])
    pano = {"view_img_fts": feats, "loc_fts": loc, "nav_types": nav_types, "view_lens": view_lens,
            "z_img_features": torch.randn(B, 50, H, generator=g), "z_img_pzs": pz(50), "already_dropout": True}
    # global map: [stop], [MEM], visited..., unvisited...
    gmap_lens = torch.tensor([G - (i % 3) for i in range(B)])
    gmap_vpids, visited_masks = [], torch.zeros(B, G, dtype=torch.bool)
    step_ids = torch.zeros(B, G, dtype=torch.int64)
    pos = torch.zeros(B, G, 7)
    pair = torch.zeros(B, G, G)
    for i in range(B):
        gl = int(gmap_lens[i])
        nvis = 2
        unv = gl - 2 - nvis
        vis = ["n%d_p%d" % (i, k) for k in range(nvis)]
        # first unvisited nodes are the current candidates (cand 0 leads back to a visited node instead)
        cand_vpids[i][0] = vis[0]
        unvis = cand_vpids[i][1:1 + unv]
        unvis = unvis + ["n%d_far%d" % (i, k) for k in range(unv - len(unvis))]
        gmap_vpids.append([None, None] + vis + unvis)
        visited_masks[i, 1:2 + nvis] = True
        step_ids[i, 2:2 + nvis] = torch.arange(1, nvis + 1)
        p = _loc_fts(gl, g)
        p[:, 4:] = torch.rand(gl, 3, generator=g)
        pos[i, :gl] = p
        d = torch.rand(gl, gl, generator=g) * 10
        d = (d + d.t()) / 2
        d.fill_diagonal_(0)
        d[:2, :] = 0
        d[:, :2] = 0
        pair[i, :gl, :gl] = d
    gmap_masks = torch.arange(G)[None, :] < gmap_lens[:, None]
    gmap_masks[:, 1] = False
    gmap_img = torch.randn(B, G, H, generator=g) * 0.5
    gmap_img[:, 0] = 0
    for i in range(B):
        gmap_img[i, int(gmap_lens[i]):] = 0
    vp_pos = torch.zeros(B, views + 2, 14)
    for i in range(B):
        vp_pos[i, :, :7] = _loc_fts(1, g)
        nc = len(cand_vpids[i])
        vp_pos[i, 2:2 + nc, 7:] = _loc_fts(nc, g)
    nav = {"txt_masks": txt_masks, "gmap_img_embeds": gmap_img, "gmap_step_ids": step_ids, "gmap_pos_fts": pos,
           "gmap_masks": gmap_masks, "gmap_pair_dists": pair, "gmap_visited_masks": visited_masks,
           "gmap_vpids": gmap_vpids, "vp_pos_fts": vp_pos, "vp_masks": torch.arange(views + 2)[None, :] < (view_lens + 2)[:, None],
           "vp_nav_masks": torch.cat([torch.ones(B, 1, dtype=torch.bool), torch.zeros(B, 1, dtype=torch.bool), nav_types == 1], 1),
           "vp_cand_vpids": [[None, None] + c for c in cand_vpids],
           "front_vp_feats": torch.tanh(torch.randn(B, 24, H, generator=g)),
           "front_gmap_feats": torch.tanh(torch.randn(B, 24, H, generator=g)),
           "mem_embeds": torch.randn(B, H, generator=g) * 0.5}
    return lang, pano, nav


def rollout_inputs(steps=3, B=3, L=20, seed=40):
    """Synthetic 3-step fine-tune rollout (shared by tests/golden/make_golden.py and tests/test_gpu_models.py): one instruction,
    per-step panorama / navigation inputs, the FACL prototypes constant over the rollout as in M/r2r/agent.py:567-583."""
    lang, _, nav0 = nav_inputs(B=B, L=L, seed=seed)
    per_step = []
    for t in range(steps):
        _, pano, nav = nav_inputs(B=B, L=L, seed=seed + 1 + t)
        nav["txt_masks"] = lang["txt_masks"]
        nav["front_vp_feats"], nav["front_gmap_feats"] = nav0["front_vp_feats"], nav0["front_gmap_feats"]
        per_step.append((pano, nav))
    targets = [torch.tensor([4, 0, 5]), torch.tensor([0, 4, 0]), torch.tensor([5, 4, 6])][:steps]   # [stop] or unvisited nodes
    return lang, per_step, targets


def run_rollout(model, lang, per_step, targets, to_dev=lambda d: d):
    """language once, then panorama + navigation per step; the [MEM] token of step t+1 is step t's cls_embeds (agent.py:592);
    teacher-forced cross-entropy summed over the steps.  -> (loss, [fused_logits], [cls_embeds], txt_embeds)"""
    from collections import defaultdict
    dd = lambda d: defaultdict(lambda: None, to_dev(d))
    txt = model("language", dd(lang))
    loss, logits, clss, mem = 0.0, [], [], None
    for (pano, nav), tgt in zip(per_step, targets):
        pe, pm, pf = model("panorama", dd(pano))
        navb = dict(to_dev(nav))
        m0 = navb.pop("mem_embeds")
        mem_t = m0 if mem is None else mem
        navb["txt_embeds"] = txt
        navb["vp_img_embeds"] = torch.cat([torch.zeros_like(pe[:, :1]), mem_t.unsqueeze(1), pe], 1)
        gi = navb["gmap_img_embeds"].clone()
        gi = torch.cat([gi[:, :1], mem_t.unsqueeze(1), gi[:, 2:]], 1)          # gmap row 1 = [MEM] as well (agent.py:175)
        navb["gmap_img_embeds"] = gi
        outs = model("navigation", defaultdict(lambda: None, navb))
        mem = outs["cls_embeds"]
        logits.append(outs["fused_logits"])
        clss.append(outs["cls_embeds"])
        loss = loss + torch.nn.functional.cross_entropy(outs["fused_logits"], tgt.to(outs["fused_logits"].device), reduction="sum")
    return loss, logits, clss, txt
