"""CPU oracle of the FULL pretraining forward (tasks MLM / SAP / CFP): a plain restatement of
``GlocalTextPathCMTPreTraining.forward(batch, task, compute_loss)`` composed from the block oracles in
``goat_oracle.py``, over the reference's own ``state_dict`` key names and the reference's batch dict.

TEST INFRASTRUCTURE ONLY -- not product code (same rules as goat_oracle.py: imported by tests/, smoke() and bench.py's
cpu_baseline / --impl reference / parity legs; nothing under vln_goat_b200/ imports it).

Pinning: ``tests/test_oracle_golden.py::test_pretrain_oracle_matches_reference_fixture`` checks every output of this
file against ``tests/golden/pretrain_full.npz``, which ``tests/golden/make_golden.py --tree pretrain_full`` produced by
running the UNMODIFIED reference model (through oracle/ref_shim.py) on ``tests/synth.pretrain_batch``.

Reference lines (P/ = /root/reference/pretrain_src/):
  text / trajectory front end     P/model/vilmodel_goat.py:546-575 (forward), :597-648 (forward_mlm), :650-696 (forward_cfp)
  CausalImageEmbeddings           P/model/vilmodel_goat.py:288-364 (R2R branch: no objects, no BACL in the shipped config)
  LocalVPEncoder / GlobalMapEncoder  P/model/vilmodel_goat.py:366-527
  MLM / SAP / CFP heads + losses  P/model/pretrain_goat.py:188-224, :286-354, :467-541
The Python loops over viewpoint-id strings are kept AS LOOPS here (this is the checker, the product replaces them by
index tensors + one kernel).
"""
import torch

from . import goat_oracle as O

EPS = 1e-12     # P/config/r2r_GOAT_model_config.json: layer_norm_eps


def _pos_embed(p, pre, x):
    """nn.Sequential(Linear, BertLayerNorm(eps=1e-12))  -- P/model/vilmodel_goat.py:369-372, :415-418"""
    return O.layernorm(O.linear(x, p[pre + "0.weight"], p[pre + "0.bias"]), p[pre + "1.weight"], p[pre + "1.bias"], 1e-12)


def encode_text(p, txt_ids, txt_lens):
    """embeddings -> 6 x RobertaLayer (dropout is the identity in eval)  -- P/model/vilmodel_goat.py:557-565"""
    txt_masks = O.gen_seq_masks(txt_lens, txt_ids.shape[1])
    emb = O.roberta_embeddings(p, "bert.embeddings.", txt_ids, EPS)
    return O.lang_encoder(p, "bert.lang_encoder.", emb, txt_masks, 6, EPS), txt_masks


def encode_traj(p, batch):
    """-> (views [S,V,H], fused [S,H])  -- P/model/vilmodel_goat.py:294-312 (embeddings + pano encoder), :354-362 (fusion)"""
    pre = "bert.img_embeddings."
    x = O.layernorm(O.linear(batch["traj_view_img_fts"], p[pre + "img_linear.weight"], p[pre + "img_linear.bias"]),
                    p[pre + "img_layer_norm.weight"], p[pre + "img_layer_norm.bias"], 1e-12)
    x = x + O.layernorm(O.linear(batch["traj_loc_fts"], p[pre + "loc_linear.weight"], p[pre + "loc_linear.bias"]),
                        p[pre + "loc_layer_norm.weight"], p[pre + "loc_layer_norm.bias"], 1e-12)
    img_masks = O.gen_seq_masks(batch["traj_vp_view_lens"], x.shape[1])
    x = O.pano_encoder(p, pre + "img_self_encoder.", x, img_masks.logical_not(), 2)
    fused = O.pano_fuse(x, p[pre + "adaptive_pano_attn.weight"], p[pre + "adaptive_pano_attn.bias"])
    return x, fused


def aggregate_gmap(views, fused, batch):
    """the string-keyed aggregation loop, verbatim in structure  -- P/model/vilmodel_goat.py:430-468"""
    step_lens = [int(x) for x in batch["traj_step_lens"]]
    H = views.shape[-1]
    out, off = [], 0
    for i, n in enumerate(step_lens):
        visited, unvisited = {}, {}
        for t in range(n):
            visited[batch["traj_vpids"][i][t]] = fused[off + t]
            for j, vp in enumerate(batch["traj_cand_vpids"][i][t]):
                if vp not in visited:
                    unvisited.setdefault(vp, []).append(views[off + t, j])
        off += n
        fts = [visited[vp] if vp in visited else torch.stack(unvisited[vp], 0).mean(0) for vp in batch["gmap_vpids"][i][1:]]
        out.append(torch.stack(fts, 0))
    G1 = max(x.shape[0] for x in out)
    padded = torch.stack([torch.cat([x, x.new_zeros(G1 - x.shape[0], H)], 0) for x in out], 0)
    return torch.cat([padded.new_zeros(len(out), 1, H), padded], 1)          # [stop] first


def gmap_inputs(p, views, fused, batch):
    pre = "bert.global_encoder."
    img = aggregate_gmap(views, fused, batch)
    emb = img + p[pre + "gmap_step_embeddings.weight"][batch["gmap_step_ids"]] + \
        _pos_embed(p, pre + "gmap_pos_embeddings.", batch["gmap_pos_fts"])
    return emb, O.gen_seq_masks(batch["gmap_lens"], emb.shape[1])


def vp_inputs(p, views, batch):
    """current (last) panorama of every sample, [stop] token first  -- P/model/vilmodel_goat.py:377-392"""
    step_lens = [int(x) for x in batch["traj_step_lens"]]
    last, s = [], 0
    for n in step_lens:
        s += n
        last.append(s - 1)
    cur = views[last]
    vp_lens = batch["traj_vp_view_lens"][last] + 1
    Nq = batch["vp_pos_fts"].shape[1]
    vp_img = torch.cat([cur.new_zeros(len(last), 1, cur.shape[-1]), cur], 1)[:, :Nq]
    emb = vp_img + _pos_embed(p, "bert.local_encoder.vp_pos_embeddings.", batch["vp_pos_fts"])
    return emb, O.gen_seq_masks(vp_lens, Nq), last


def _fuse_weights(p, g, v):
    return torch.sigmoid(O.cls_prediction(p, "sap_fuse_linear.", torch.cat([g[:, 0], v[:, 0]], 1)))


def forward_mlm(p, batch):
    """-> (prediction_scores [n_masked, vocab], per-token loss)"""
    txt, txt_masks = encode_text(p, batch["txt_ids"], batch["txt_lens"])
    views, fused = encode_traj(p, batch)
    g_in, g_m = gmap_inputs(p, views, fused, batch)
    v_in, v_m, _ = vp_inputs(p, views, batch)
    g_txt = O.crossmodal_encoder(p, "bert.global_encoder.encoder.", txt, txt_masks, g_in, g_m, None, 3, EPS)
    v_txt = O.crossmodal_encoder(p, "bert.local_encoder.encoder.", txt, txt_masks, v_in, v_m, None, 3, EPS)
    out = g_txt + v_txt
    sel = batch["txt_labels"] != -1
    h = O.head_transform(p, "mlm_head.predictions.transform.", out[sel], EPS)
    scores = O.linear(h, p["bert.embeddings.word_embeddings.weight"], p["mlm_head.predictions.bias"])   # tied decoder
    return scores, O.cross_entropy_rows(scores, batch["txt_labels"][sel])


def forward_sap(p, batch):
    """-> (global_logits, local_logits, fused_logits, per-sample loss)"""
    txt, txt_masks = encode_text(p, batch["txt_ids"], batch["txt_lens"])
    views, fused = encode_traj(p, batch)
    g_in, g_m = gmap_inputs(p, views, fused, batch)
    v_in, v_m, last = vp_inputs(p, views, batch)
    w, b = p["bert.global_encoder.sprel_linear.weight"], p["bert.global_encoder.sprel_linear.bias"]
    sprels = (batch["gmap_pair_dists"] * w.view(()) + b.view(())).unsqueeze(1)
    g = O.crossmodal_encoder(p, "bert.global_encoder.encoder.", g_in, g_m, txt, txt_masks, sprels, 3, EPS)
    v = O.crossmodal_encoder(p, "bert.local_encoder.encoder.", v_in, v_m, txt, txt_masks, None, 3, EPS)
    fw = _fuse_weights(p, g, v)
    ninf = -float("inf")
    gl = O.cls_prediction(p, "global_sap_head.", g).squeeze(2) * fw
    gl = gl.masked_fill(batch["gmap_visited_masks"], ninf).masked_fill(g_m.logical_not(), ninf)
    ll = O.cls_prediction(p, "local_sap_head.", v).squeeze(2) * (1 - fw)
    Nq = ll.shape[1]
    cur_nav = (batch["traj_nav_types"][last] != 1)[:, :Nq - 1]
    ll = ll.masked_fill(torch.cat([cur_nav.new_zeros(len(last), 1), cur_nav], 1), ninf)
    fl = O.sap_fuse_logits(gl, ll, batch["gmap_vpids"], batch["gmap_visited_masks"].tolist(),
                           [c[-1] for c in batch["traj_cand_vpids"]], skip=1)
    ga, la = batch["global_act_labels"], batch["local_act_labels"]
    loss = O.cross_entropy_rows(gl, ga) + O.cross_entropy_rows(ll, la) + O.cross_entropy_rows(fl, ga)
    return gl, ll, fl, loss


def forward_cfp(p, batch, temperature=1.0):
    """-> (gmap, vp, fused, txt pooled embeddings, per-sample loss)"""
    txt, txt_masks = encode_text(p, batch["txt_ids"], batch["txt_lens"])
    views, fused = encode_traj(p, batch)
    g_in, g_m = gmap_inputs(p, views, fused, batch)
    v_in, v_m, _ = vp_inputs(p, views, batch)
    g = O.bert_attention(p, "bert.global_encoder.tim_self_encoder.", g_in, O.extend_neg_masks(g_m), eps=EPS)
    v = O.bert_attention(p, "bert.local_encoder.tim_self_encoder.", v_in, O.extend_neg_masks(v_m), eps=EPS)
    g = O.head_transform(p, "tim_global_head.", g, EPS)
    v = O.head_transform(p, "tim_local_head.", v, EPS)
    t = O.head_transform(p, "tim_txt_head.", txt, EPS)
    fw = _fuse_weights(p, g, v)
    go, vo, to = O.cfp_pool(g, p["tim_global_attn"]), O.cfp_pool(v, p["tim_local_attn"]), O.cfp_pool(t, p["tim_txt_attn"])
    fo = go * fw + vo * (1 - fw)
    loss = O.infonce_sym(go, to, temperature) + O.infonce_sym(vo, to, temperature) + O.infonce_sym(fo, to, temperature)
    return go, vo, fo, to, loss


def scalar_loss(p, batch, task):
    """``loss.mean()`` of P/train_r2r_goat.py:317"""
    if task == "mlm":
        return forward_mlm(p, batch)[1].mean()
    if task == "sap":
        return forward_sap(p, batch)[3].mean()
    if task == "cfp":
        return forward_cfp(p, batch)[4].mean()
    raise ValueError(task)
