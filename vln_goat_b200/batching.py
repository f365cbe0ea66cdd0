"""Host side of a pretraining step: batch dict of the reference collates -> flat, index-addressed, optionally
statically padded tensors that ``pretrain_model.GlocalTextPathCMTPreTraining.forward_prepared`` consumes.

The reference walks Python lists of viewpoint-id strings INSIDE the model forward (global-map aggregation
P/model/vilmodel_goat.py:430-468, SAP logit fusion P/model/pretrain_goat.py:328-345, per-sample ``x[-1]`` selection of
the current panorama :379-381, boolean-mask selection of the masked tokens :200-206).  Here all of that happens once
per batch on the host and leaves only integer index tensors, so the device side is shape-static: the whole forward +
backward of a (task, padded shape) pair can be captured in one CUDA graph and replayed (engine.TrainStep).

Padding (``pad`` given) rounds the data-dependent extents up to bucket multiples:
  S   total trajectory steps of the batch          (padded steps: one zero view, never referenced by an index)
  G   global-map length                            (padded nodes: masked by gmap_lens, no label points at them)
  NM  number of masked tokens (MLM)                (padded rows: label -1 = ignore_index, zero loss / gradient)
  K   index-list widths                            (-1 = empty slot)
and records the batch's own padded extents (``n_gmap``, ``n_vp``, ``n_txt``) for the CFP pooling, which the reference
runs over padding too (so extra padding must not add tokens).
"""
import torch

from . import goat_blocks as G


class PadSpec(object):
    """Bucket multiples for the data-dependent extents (see module docstring)."""

    def __init__(self, S=32, G=8, NM=128, K=8, KF=8, L=None):
        """L: pad the token axis to exactly L (None: keep the batch's own padded length)"""
        self.S, self.G, self.NM, self.K, self.KF, self.L = S, G, NM, K, KF, L


def _up(n, m):
    return (n + m - 1) // m * m


def _cpu(t):
    return t.detach().cpu() if torch.is_tensor(t) and t.is_cuda else t


def _pad_dim0(t, n, value=0, pinned=False):
    """pad the leading dimension to n; pinned=True: the result lives in page-locked host memory (ONE copy does the
    padding and the staging for the H2D transfer)"""
    pinned = pinned and not t.is_cuda
    if t.shape[0] == n and not pinned:
        return t
    out = torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device, pin_memory=pinned)
    out[:t.shape[0]] = t
    if n > t.shape[0]:
        out[t.shape[0]:] = value
    return out


def _pad_dim1(t, n, value=0, pinned=False):
    pinned = pinned and not t.is_cuda
    if t.shape[1] == n and not pinned:
        return t
    out = torch.empty((t.shape[0], n) + tuple(t.shape[2:]), dtype=t.dtype, device=t.device, pin_memory=pinned)
    out[:, :t.shape[1]] = t
    if n > t.shape[1]:
        out[:, t.shape[1]:] = value
    return out


_Z_KEYS = ("instr_z_direction_features", "instr_z_direction_pzs", "instr_z_landmark_features", "instr_z_landmark_pzs",
           "img_z_features", "img_z_pzs")


def prepare_pretrain(batch, task, pad=None, pano_fusion=True, pinned=False):
    """batch: dict with the keys of P/data/tasks.py's collates, tensors on the host or on the
    device.  -> dict of tensors (same placement as the inputs; integer index tensors are built on the host).
    pano_fusion: config.adaptive_pano_fusion (a visited node is its fused panorama, else the mean of its views).
    pinned: write the large host tensors (view / location features) straight into page-locked memory."""
    task = task.split("_")[0]
    if task not in ("mlm", "sap", "cfp"):
        raise NotImplementedError("task %r is outside the hot-path scope (MRC / OG are REVERIE-only)" % task)
    if batch.get("traj_obj_img_fts") is not None:
        raise NotImplementedError("object features (REVERIE / SOON) are outside the hot-path scope")
    view_ids = batch.get("traj_view_ids")          # rows of a GPU-resident workloads.FeatureBank instead of the features
    feats = batch.get("traj_view_img_fts")
    dev = (feats if feats is not None else batch["traj_loc_fts"]).device
    step_lens = [int(x) for x in batch["traj_step_lens"]]
    B = len(step_lens)
    S = sum(step_lens)
    V = feats.shape[1] if feats is not None else batch["traj_loc_fts"].shape[1]
    view_lens = batch["traj_vp_view_lens"]
    view_lens_h = _cpu(view_lens).tolist()
    gmap_step_ids = batch["gmap_step_ids"]
    Gn = gmap_step_ids.shape[1]
    vp_pos = batch["vp_pos_fts"]
    Nq = vp_pos.shape[1]
    L = batch["txt_ids"].shape[1]

    Sp = S if pad is None else _up(S, pad.S)
    Gp = Gn if pad is None else _up(Gn, pad.G)
    Nqp = Nq if pad is None else V + 1

    # ---- index lists (host) ----
    idx = G.build_gmap_index(step_lens, view_lens_h, batch["traj_vpids"], batch["traj_cand_vpids"], batch["gmap_vpids"],
                             pano_fusion, V)
    if pano_fusion:
        idx_f, idx_v = G.split_gmap_index(idx, S)
    else:
        idx_f, idx_v = torch.full_like(idx[..., :1], -1), idx - S      # every entry is a view row; -1 - S stays negative
    K = idx_v.shape[2] if pad is None else _up(idx_v.shape[2], pad.K)
    idx_f = _pad_dim1(idx_f, Gp - 1, -1)
    idx_v = _pad_dim1(idx_v, Gp - 1, -1)
    if idx_v.shape[2] < K:
        idx_v = torch.cat([idx_v, idx_v.new_full(idx_v.shape[:2] + (K - idx_v.shape[2],), -1)], 2)
    ends = torch.tensor(step_lens, dtype=torch.int64).cumsum(0)
    last_rows = (ends - 1).to(torch.int32).view(B, 1)

    def put(t):
        return t.to(dev) if t.device != dev else t

    txt_ids = batch["txt_ids"]
    Lp = L if (pad is None or pad.L is None) else pad.L
    if Lp < L:
        raise ValueError("the batch has %d tokens per instruction, more than the padded length %d" % (L, Lp))
    out = {
        "txt_ids": _pad_dim1(txt_ids, Lp), "txt_lens": batch["txt_lens"],
        "loc_fts": _pad_dim0(batch["traj_loc_fts"], Sp, 0, pinned),
        "view_lens": _pad_dim0(view_lens, Sp, 1), "last_rows": put(last_rows),
        "gmap_idx_f": put(idx_f.contiguous()), "gmap_idx_v": put(idx_v.contiguous()),
        "gmap_step_ids": _pad_dim1(gmap_step_ids, Gp), "gmap_pos_fts": _pad_dim1(batch["gmap_pos_fts"], Gp),
        "gmap_lens": batch["gmap_lens"], "vp_pos_fts": _pad_dim1(vp_pos, Nqp),
        "n_gmap": put(torch.tensor([Gn], dtype=torch.int32)), "n_vp": put(torch.tensor([Nq], dtype=torch.int32)),
        "n_txt": put(torch.tensor([L], dtype=torch.int32)),
    }
    if view_ids is not None:
        out["view_idx"] = _pad_dim0(view_ids.to(torch.int32), Sp, -1)
    else:
        out["view_fts"] = _pad_dim0(feats, Sp, 0, pinned)
    for k in _Z_KEYS:
        if batch.get(k) is not None:
            out[k] = batch[k]
    if task == "sap" or task == "fwd":
        pair = batch["gmap_pair_dists"]
        if pair.shape[1] != Gp:
            p2 = pair.new_zeros(B, Gp, Gp)
            p2[:, :Gn, :Gn] = pair
            pair = p2
        vis = batch["gmap_visited_masks"]
        fidx = G.build_fusion_index(batch["gmap_vpids"], _cpu(vis), [c[-1] for c in batch["traj_cand_vpids"]], Nqp, 1, 1)
        KF = fidx.shape[2] if pad is None else _up(fidx.shape[2], pad.KF)
        fidx = _pad_dim1(fidx, Gp, -1)
        if fidx.shape[2] < KF:
            fidx = torch.cat([fidx, fidx.new_full(fidx.shape[:2] + (KF - fidx.shape[2],), -1)], 2)
        out.update({
            "gmap_pair_dists": pair, "gmap_visited_masks": _pad_dim1(vis, Gp, True),
            "nav_types": _pad_dim0(batch["traj_nav_types"], Sp), "fuse_idx": put(fidx.contiguous()),
            "global_act_labels": batch["global_act_labels"], "local_act_labels": batch["local_act_labels"],
            "loss_inv": put(torch.tensor([1.0 / B], dtype=torch.float32)),
        })
    elif task == "mlm":
        labels = _pad_dim1(_cpu(batch["txt_labels"]), Lp, -1)
        rows = torch.nonzero(labels.reshape(-1) != -1).view(-1)                 # row-major, as boolean indexing orders them
        nm = int(rows.numel())
        NMp = nm if pad is None else max(_up(nm, pad.NM), pad.NM)
        mrows = torch.zeros(NMp, 1, dtype=torch.int32)
        mrows[:nm, 0] = rows.to(torch.int32)
        mlab = torch.full((NMp,), -1, dtype=torch.int64)
        mlab[:nm] = labels.reshape(-1)[rows]
        out.update({"mlm_rows": put(mrows), "mlm_labels": put(mlab),
                    "loss_inv": put(torch.tensor([1.0 / max(nm, 1)], dtype=torch.float32))})
    else:
        out["loss_inv"] = put(torch.tensor([1.0 / B], dtype=torch.float32))
    return out


def pin(prepared):
    """The (host) tensors in pinned memory, ready for non-blocking H2D copies (already-pinned ones are kept)."""
    return {k: (v if (v.is_cuda or v.is_pinned()) else v.pin_memory()) for k, v in prepared.items()}


def h2d_bytes(prepared):
    return sum(v.numel() * v.element_size() for v in prepared.values() if not v.is_cuda)
