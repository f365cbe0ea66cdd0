"""CPU tests of the host-side logic (no GPU, no compute calls into the library):
  * the shared library loads and exports every symbol include/goat_sm100.h declares;
  * the drop-in models expose exactly the reference's state_dict keys (checked against the key lists stored in the
    fixtures that tests/golden/make_golden.py took from the unmodified reference classes);
  * the index lists that replace the reference's Python loops over viewpoint-id strings (global-map aggregation,
    logit fusion) reproduce those loops, restated here literally from the reference;
  * the flat-parameter engine's layout and its gradient all-reduce on a 2-rank gloo group.
"""
import os
import re

import numpy as np
import pytest
import torch

from tests import synth
from tests.helpers import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    from vln_goat_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "goat_sm100.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)      # drop comments (they mention goat_* names too)
    declared = set(re.findall(r"\b(goat_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    lib = _lib.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.goat_version() >= 100


def _nav_cfg():
    from vln_goat_b200.config import GoatConfig
    return GoatConfig(layer_norm_eps=1e-5, pad_token_id=1, dataset="r2r", mode="train", obj_feat_size=0, feat_dropout=0.4,
                      do_back_img=True, do_back_txt=True, do_front_img=True, do_front_his=True, do_front_txt=True,
                      do_back_txt_type="type_2", do_back_img_type="type_1", do_add_method="door",
                      use_lang2visn_attn=False, fix_lang_embedding=False, fix_pano_embedding=False, fix_local_branch=False)


def test_state_dict_keys_match_reference_models():
    from vln_goat_b200 import nav_model, pretrain_model
    from vln_goat_b200.config import GoatConfig
    ref = str(golden("pretrain_full")["state_dict_keys"]).split("\n")
    m = pretrain_model.GlocalTextPathCMTPreTraining(GoatConfig())
    assert list(m.state_dict().keys()) == ref
    assert abs(sum(p.numel() for p in m.parameters()) / 1e6 - 208.12) < 0.01
    assert m.mlm_head.predictions.decoder.weight is m.bert.embeddings.word_embeddings.weight
    ref = str(golden("nav_full")["state_dict_keys"]).split("\n")
    m = nav_model.GlocalTextPathNavCMT(_nav_cfg())
    assert list(m.state_dict().keys()) == ref


def _reference_aggregate(split_embeds, split_lens, traj_vpids, traj_cand_vpids, gmap_vpids, split_fused):
    """P/model/vilmodel_goat.py:430-468 restated on CPU tensors (dicts keyed by viewpoint id)."""
    out = []
    for i in range(len(split_embeds)):
        visited, unvisited = {}, {}
        lens = split_lens[i]
        max_len = int(max(lens))
        masks = (torch.arange(max_len)[None, :] < lens[:, None])
        emb = split_embeds[i][:, :max_len] * masks.unsqueeze(2)
        for t in range(len(split_embeds[i])):
            if split_fused is not None:
                visited[traj_vpids[i][t]] = split_fused[i][t]
            else:
                visited[traj_vpids[i][t]] = torch.sum(emb[t], 0) / lens[t]
            for j, vp in enumerate(traj_cand_vpids[i][t]):
                if vp not in visited:
                    unvisited.setdefault(vp, []).append(emb[t][j])
        fts = []
        for vp in gmap_vpids[i][1:]:
            fts.append(visited[vp] if vp in visited else torch.mean(torch.stack(unvisited[vp], 0), 0))
        out.append(torch.stack(fts, 0))
    return out


@pytest.mark.parametrize("use_fused", [True, False])
def test_gmap_index_reproduces_reference_aggregation(use_fused):
    from vln_goat_b200 import goat_blocks as G
    b = synth.pretrain_batch(B=4, L=8, seed=11)
    H, V = 16, 36
    g = torch.Generator().manual_seed(0)
    S = sum(b["traj_step_lens"])
    views = torch.randn(S, V, H, generator=g)
    fused = torch.randn(S, H, generator=g)
    lens = b["traj_vp_view_lens"]
    split_e = torch.split(views, b["traj_step_lens"], 0)
    split_l = torch.split(lens, b["traj_step_lens"], 0)
    split_f = torch.split(fused, b["traj_step_lens"], 0) if use_fused else None
    ref = _reference_aggregate(split_e, split_l, b["traj_vpids"], b["traj_cand_vpids"], b["gmap_vpids"], split_f)
    idx = G.build_gmap_index(b["traj_step_lens"], lens.tolist(), b["traj_vpids"], b["traj_cand_vpids"], b["gmap_vpids"],
                             use_fused, V)
    vmask = (torch.arange(V)[None, :] < lens[:, None]).unsqueeze(2).float()
    src = torch.cat([fused if use_fused else torch.zeros(S, H), (views * vmask).reshape(S * V, H)], 0)
    for i, r in enumerate(ref):
        for gi in range(r.shape[0]):
            sel = [int(k) for k in idx[i, gi] if k >= 0]
            got = src[sel].mean(0)
            assert torch.allclose(got, r[gi], atol=1e-6), (i, gi)
        assert (idx[i, r.shape[0]:] == -1).all()


def _reference_fuse(global_logits, local_logits, gmap_vpids, visited_masks, cand_lists, pretrain):
    """P/model/pretrain_goat.py:328-345 (pretrain) / M/models/vilmodel_GOAT.py:793-813 (fine-tune), literally."""
    fused = global_logits.clone()
    fused[:, 0] += local_logits[:, 0]
    for i in range(len(gmap_vpids)):
        visited = set([vp for vp, m in zip(gmap_vpids[i], visited_masks[i]) if m])
        tmp, bw = {}, 0
        for j, vp in enumerate(cand_lists[i]):
            if pretrain:
                if vp in visited:
                    bw += local_logits[i, j + 1]
                else:
                    tmp[vp] = local_logits[i, j + 1]
            elif j > 1:
                if vp in visited:
                    bw += local_logits[i, j]
                else:
                    tmp[vp] = local_logits[i, j]
        for j, vp in enumerate(gmap_vpids[i]):
            if j > (0 if pretrain else 1) and vp not in visited:
                fused[i, j] += tmp[vp] if vp in tmp else bw
    return fused


def _apply_fusion_index(global_logits, local_logits, idx):
    flat = local_logits.reshape(-1)
    add = torch.zeros_like(global_logits)
    for i in range(idx.shape[0]):
        for j in range(idx.shape[1]):
            sel = [int(k) for k in idx[i, j] if k >= 0]
            if sel:
                add[i, j] = flat[sel].sum()
    return global_logits + add


def test_fusion_index_reproduces_reference_loops():
    from vln_goat_b200 import goat_blocks as G
    g = torch.Generator().manual_seed(1)
    b = synth.pretrain_batch(B=5, L=8, seed=12)
    B, Gn = b["gmap_visited_masks"].shape
    gl, ll = torch.randn(B, Gn, generator=g), torch.randn(B, 37, generator=g)
    cands = [c[-1] for c in b["traj_cand_vpids"]]
    ref = _reference_fuse(gl, ll, b["gmap_vpids"], b["gmap_visited_masks"], cands, True)
    idx = G.build_fusion_index(b["gmap_vpids"], b["gmap_visited_masks"], cands, 37, 1, 1)
    assert torch.allclose(_apply_fusion_index(gl, ll, idx), ref, atol=1e-6)
    _, _, nav = synth.nav_inputs(B=4, L=8, seed=13)
    B, Gn = nav["gmap_visited_masks"].shape
    gl, ll = torch.randn(B, Gn, generator=g), torch.randn(B, 38, generator=g)
    ref = _reference_fuse(gl, ll, nav["gmap_vpids"], nav["gmap_visited_masks"], nav["vp_cand_vpids"], False)
    idx = G.build_fusion_index(nav["gmap_vpids"], nav["gmap_visited_masks"], nav["vp_cand_vpids"], 38, 2, 2)
    assert torch.allclose(_apply_fusion_index(gl, ll, idx), ref, atol=1e-6)
    assert (ref != gl).any()


def _remap_pretrain_to_finetune(ckpt):
    """The key remap of M/models/vlnbert_init.py:52-69 restated, followed by what HF's from_pretrained does when the
    target class IS the base model (base_model_prefix 'bert'): a leading 'bert.' is stripped."""
    out = {}
    for k, v in ckpt.items():
        if k.startswith("module"):
            k = k[7:]
        if k.startswith("vln_bert"):
            k = "bert" + k[8:]
        if "_head" in k or "sap_fuse" in k:
            out["bert." + k] = v
        elif "tim" in k or "temperature" in k:
            out[("bert." + k) if "self_encoder" not in k else k] = v
        else:
            out[k] = v
    return {(k[5:] if k.startswith("bert.") else k): v for k, v in out.items()}


def test_pretrain_checkpoint_remaps_into_the_finetune_model():
    """a checkpoint saved from the pretrain model (ModelSaver strips 'module.',
    P/utils/save.py:47-63) goes through the reference's own key remap and loads into the fine-tune model: every shared
    block is covered with identical shapes; only pretrain-only heads are left over, only fine-tune-only modules missing."""
    from vln_goat_b200 import nav_model, pretrain_model
    from vln_goat_b200.config import GoatConfig
    pre = pretrain_model.GlocalTextPathCMTPreTraining(GoatConfig())
    ckpt = {("module." + k): v for k, v in pre.state_dict().items()}            # as saved under DDP
    remapped = _remap_pretrain_to_finetune(ckpt)
    nav = nav_model.GlocalTextPathNavCMT(_nav_cfg())
    target = nav.state_dict()
    shared = [k for k in remapped if k in target]
    assert all(tuple(remapped[k].shape) == tuple(target[k].shape) for k in shared)
    res = nav.load_state_dict({k: remapped[k] for k in shared}, strict=False)
    assert not res.unexpected_keys
    # the whole cross-modal core arrives from the pretrain checkpoint
    for fam in ("embeddings.", "lang_encoder.layer.", "img_embeddings.img_linear", "img_embeddings.img_self_encoder.",
                "local_encoder.encoder.crossattention.", "global_encoder.encoder.crossattention.",
                "global_encoder.gmap_pos_embeddings.", "local_encoder.vp_pos_embeddings.", "global_sap_head.",
                "local_sap_head.", "sap_fuse_linear."):
        fam_keys = [k for k in target if k.startswith(fam)]
        assert fam_keys and all(k in remapped for k in fam_keys if "lang_" not in k[len(fam):]), fam
    # left over: pretrain-only heads / pretrain-only sub-blocks of BertCrossLayer (P/model/Bert_backbone.py:673-676)
    leftover = [k for k in remapped if k not in target]
    assert leftover and all(any(t in k for t in ("mlm_head", "lang_self_attn", "lang_inter", "lang_output", "tim_",
                                                  "img_self_attn")) for k in leftover), leftover[:5]
    # missing: modules that only exist in the fine-tune model (BACL dictionaries / projections, FACL encoders, poolers,
    # the per-step history token)
    missing = set(res.missing_keys)
    assert missing and all(any(t in k for t in ("z_", "do_img", "front_", "pooler", "local_his", "instr_", "concat_linear",
                                                 "img_after_linear")) for k in missing), sorted(missing)[:8]


def _segment_reduce_cpu(src, idx, mean):
    """what goat_segment_reduce_fwd computes (CPU restatement for the host-logic tests)"""
    out = torch.zeros(idx.shape[0], src.shape[1])
    for r in range(idx.shape[0]):
        sel = [int(i) for i in idx[r] if i >= 0]
        if sel:
            out[r] = src[sel].sum(0) / (len(sel) if (mean and len(sel) > 1) else 1)
    return out


@pytest.mark.parametrize("padded", [False, True])
def test_prepare_pretrain_indices_reproduce_the_reference_loops(padded):
    """batching.prepare_pretrain (index tensors, optional static padding) against the string-keyed loops of the reference
    as restated in oracle/goat_pretrain_oracle.py: global-map aggregation, current-panorama selection, masked-token
    selection and SAP logit fusion give the same numbers for the real rows, padded or not."""
    from oracle import goat_oracle as O
    from oracle import goat_pretrain_oracle as PO
    from vln_goat_b200 import batching
    batch = synth.pretrain_batch(B=5, L=24, seed=11)
    pad = batching.PadSpec(S=8, G=8, NM=16, K=4, KF=4) if padded else None
    H = 16
    g = torch.Generator().manual_seed(0)
    S, V = batch["traj_view_img_fts"].shape[:2]
    views, fused = torch.randn(S, V, H, generator=g), torch.randn(S, H, generator=g)
    ref = PO.aggregate_gmap(views, fused, batch)                                   # [B, G, H]
    for task in ("mlm", "sap", "cfp"):
        P = batching.prepare_pretrain(batch, task, pad=pad)
        Sp = P["view_fts"].shape[0]
        assert Sp % (8 if padded else 1) == 0 and Sp >= S
        vp = torch.cat([views, torch.zeros(Sp - S, V, H)], 0)
        fp = torch.cat([fused, torch.zeros(Sp - S, H)], 0)
        B, G1, K = P["gmap_idx_v"].shape
        got = _segment_reduce_cpu(vp.reshape(Sp * V, H), P["gmap_idx_v"].view(B * G1, K), True) + \
            _segment_reduce_cpu(fp, P["gmap_idx_f"].view(B * G1, 1), False)
        got = got.view(B, G1, H)
        Gn = ref.shape[1]
        assert torch.allclose(got[:, :Gn - 1], ref[:, 1:], atol=1e-6)
        assert float(got[:, Gn - 1:].abs().max()) == 0.0 if G1 > Gn - 1 else True
        assert int(P["n_gmap"]) == Gn and P["gmap_step_ids"].shape[1] == G1 + 1
        # current panorama of every sample
        last = (torch.tensor(batch["traj_step_lens"]).cumsum(0) - 1)
        assert torch.equal(P["last_rows"].view(-1).long(), last)
        assert torch.equal(P["view_lens"][:S], batch["traj_vp_view_lens"]) and bool((P["view_lens"][S:] == 1).all())
    # masked tokens (row-major order of boolean indexing) and their labels; padded rows carry the ignore label
    P = batching.prepare_pretrain(batch, "mlm", pad=pad)
    sel = batch["txt_labels"] != -1
    nm = int(sel.sum())
    x = torch.randn(batch["txt_ids"].numel(), H, generator=g)
    got = _segment_reduce_cpu(x, P["mlm_rows"], False)
    assert torch.equal(got[:nm], x.view(*batch["txt_ids"].shape, H)[sel])
    assert torch.equal(P["mlm_labels"][:nm], batch["txt_labels"][sel]) and bool((P["mlm_labels"][nm:] == -1).all())
    assert abs(float(P["loss_inv"]) - 1.0 / nm) < 1e-7
    # SAP logit fusion
    P = batching.prepare_pretrain(batch, "sap", pad=pad)
    B, Gp = P["gmap_visited_masks"].shape
    Nq = P["vp_pos_fts"].shape[1]
    gl, ll = torch.randn(B, Gp, generator=g), torch.randn(B, Nq, generator=g)
    Gn = batch["gmap_step_ids"].shape[1]
    ref = O.sap_fuse_logits(gl[:, :Gn], ll, batch["gmap_vpids"], batch["gmap_visited_masks"].tolist(),
                            [c[-1] for c in batch["traj_cand_vpids"]], skip=1)
    add = _segment_reduce_cpu(ll.reshape(-1, 1), P["fuse_idx"].view(B * Gp, -1), False).view(B, Gp)
    assert torch.allclose((gl + add)[:, :Gn], ref, atol=1e-6)
    assert float(add[:, Gn:].abs().max()) == 0.0 if Gp > Gn else True
    assert bool(P["gmap_visited_masks"][:, Gn:].all()) if Gp > Gn else True


def test_node_embed_bank_matches_reference_graphmap_semantics():
    """graph_map.NodeEmbedBank (index ops over (episode, node) pairs) against the reference's dict-of-[sum, count] GraphMap
    bookkeeping, restated literally from M/models/graph_utils.py:110-121: rewrite for the visited node, accumulation for
    candidate views (duplicates included), sum / count reads, gradients through both."""
    from vln_goat_b200.graph_map import NodeEmbedBank
    g = torch.Generator().manual_seed(2)
    E, N, H = 3, 7, 16
    bank = NodeEmbedBank(E, N, H, device="cpu")
    ref = [dict() for _ in range(E)]

    def ref_update(e, vp, emb, rewrite=False):
        if rewrite or vp not in ref[e]:
            ref[e][vp] = [emb, 1]
        else:
            ref[e][vp] = [ref[e][vp][0] + emb, ref[e][vp][1] + 1]
    leaves = []
    for step in range(4):
        # candidate views: several per episode, duplicates allowed
        ep = torch.tensor([0, 0, 1, 2, 2, 2, 0])
        nd = torch.tensor([1 + step % 3, 2, 3, 4, 4, 5, 1 + step % 3])
        emb = torch.randn(len(ep), H, generator=g).requires_grad_(True)
        leaves.append(emb)
        bank.update(ep, nd, emb)
        for i in range(len(ep)):
            ref_update(int(ep[i]), int(nd[i]), emb[i])
        # visited node of every episode: rewrite
        ve, vn = torch.arange(E), torch.tensor([step % N, (step + 1) % N, (step + 2) % N])
        vemb = torch.randn(E, H, generator=g).requires_grad_(True)
        leaves.append(vemb)
        bank.update(ve, vn, vemb, rewrite=True)
        for i in range(E):
            ref_update(i, int(vn[i]), vemb[i], rewrite=True)
    table = torch.full((E, N), -1, dtype=torch.int64)
    for e in range(E):
        for j, vp in enumerate(sorted(ref[e])):
            table[e, j] = vp
    got = bank.get_padded(table)
    w = torch.randn(E, N, H, generator=g)
    exp = torch.zeros(E, N, H)
    for e in range(E):
        for j, vp in enumerate(sorted(ref[e])):
            exp[e, j] = ref[e][vp][0] / ref[e][vp][1]
    assert torch.allclose(got, exp, atol=1e-6)
    g1 = torch.autograd.grad((got * w).sum(), leaves, allow_unused=True)
    g2 = torch.autograd.grad((exp * w).sum(), leaves, allow_unused=True)
    for a, b in zip(g1, g2):
        assert (a is None and (b is None or float(b.abs().max()) == 0)) or torch.allclose(a, b if b is not None else torch.zeros_like(a), atol=1e-6)
