"""Pretraining model of the GOAT path (R2R / RxR tasks MLM, SAP, CFP), backed by libgoat_sm100.

Drop-in for the reference's ``GlocalTextPathCMT`` (P/model/vilmodel_goat.py:529-696) and
``GlocalTextPathCMTPreTraining`` (P/model/pretrain_goat.py:40-541): same constructor (``config``), same
``forward(batch, task, compute_loss)`` contract, same batch-dict keys and the same
``state_dict`` keys, so ``P/train_r2r_goat.py`` can build it instead of the reference class.  The MRC / OG tasks
are REVERIE-only and raise.  P/ = pretrain_src/ of CrystalSixone/VLN-GOAT.
"""
from collections import defaultdict

import torch
from torch import nn

from . import batching
from . import functional as Fn
from . import goat_blocks as G
from . import modules as M
from .modules import extend_neg_masks, gen_seq_masks


class GlocalTextPathCMT(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = G.RobertaEmbeddings(config, tuple_output=True)
        if config.do_back_txt:
            self.lang_encoder = G.LanguageEncoderDo(config, pretrain_layout=True)
        else:
            self.lang_encoder = M.LanguageEncoder(config)
        self.img_embeddings = G.CausalImageEmbeddings(config, pretrain_layout=True)
        self.local_encoder = G.LocalVPEncoder(config)
        self.global_encoder = G.GlobalMapEncoder(config)
        self.apply(_init_bert_weights(config))

    # -- shared front end: text encoder + panorama embeddings --------------------------------
    def _encode_text(self, txt_ids, txt_lens, zd_f, zd_p, zl_f, zl_p):
        txt_masks = gen_seq_masks(txt_lens)
        if self.config.do_back_txt:
            emb, zd, zl = self.embeddings(txt_ids, instr_z_direction_features=zd_f, instr_z_landmark_features=zl_f)
            txt = self.lang_encoder(emb, txt_masks, z_direc_embeds=zd, z_direc_pzs=zd_p, z_landm_embeds=zl, z_landm_pzs=zl_p)
        else:
            txt = self.lang_encoder(self.embeddings(txt_ids)[0], txt_masks)
        return txt, txt_masks

    def _encode_traj(self, traj_view_img_fts, traj_obj_img_fts, traj_loc_fts, traj_nav_types, traj_step_lens,
                     traj_vp_view_lens, z_img_features, z_img_pzs):
        if traj_obj_img_fts is not None:
            raise NotImplementedError("object features (REVERIE / SOON) are outside the hot-path scope")
        return self.img_embeddings(traj_view_img_fts, traj_loc_fts, traj_nav_types, traj_step_lens, traj_vp_view_lens,
                                   self.embeddings.token_type_embeddings, z_img_features=z_img_features,
                                   z_img_pzs=z_img_pzs)

    # -- index-addressed, shape-static form (batching.prepare_pretrain) ------------------------
    def encode_prepared(self, P):
        """-> (txt_embeds, txt_masks, views [S,V,H], fused [S,H] or None)"""
        L = P["txt_ids"].shape[1]
        txt_masks = gen_seq_masks(P["txt_lens"], L)
        if self.config.do_back_txt:
            emb, zd, zl = self.embeddings(P["txt_ids"], instr_z_direction_features=P.get("instr_z_direction_features"),
                                          instr_z_landmark_features=P.get("instr_z_landmark_features"))
            txt = self.lang_encoder(emb, txt_masks, z_direc_embeds=zd, z_direc_pzs=P.get("instr_z_direction_pzs"),
                                    z_landm_embeds=zl, z_landm_pzs=P.get("instr_z_landmark_pzs"))
        else:
            txt = self.lang_encoder(self.embeddings(P["txt_ids"])[0], txt_masks)
        if "view_idx" in P:
            bank = getattr(self, "feature_bank", None)
            if bank is None:
                raise RuntimeError("the batch names panoramas by feature-bank row (traj_view_ids) but no bank is attached: "
                                   "set model.bert.feature_bank = workloads.FeatureBank(features)")
            view_fts = bank.gather(P["view_idx"])
        else:
            view_fts = P["view_fts"]
        views, _, fused = self.img_embeddings.encode(view_fts, P["loc_fts"], P["view_lens"], P.get("img_z_features"),
                                                     P.get("img_z_pzs"))
        return txt, txt_masks, views, fused

    def gmap_inputs_prepared(self, P, views, fused):
        ge = self.global_encoder
        img = ge.aggregate_flat(views, fused, P["gmap_idx_f"], P["gmap_idx_v"])
        return ge.embed_nodes(img, P["gmap_step_ids"], P["gmap_pos_fts"], P["gmap_lens"])

    def vp_inputs_prepared(self, P, views):
        return self.local_encoder.vp_input_embedding_flat(views, P["view_lens"], P["last_rows"], P["vp_pos_fts"])

    def forward(self, txt_ids, txt_lens, traj_view_img_fts, traj_obj_img_fts, traj_loc_fts, traj_nav_types,
                traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens, traj_vpids, traj_cand_vpids, gmap_lens,
                gmap_step_ids, gmap_pos_fts, gmap_pair_dists, gmap_vpids, vp_pos_fts, return_gmap_embeds=True,
                z_img_features=None, z_img_pzs=None, traj_reverie_loc_fts=None, return_txt_embeds=False,
                traj_reverie_obj_names=None, instr_z_landmark_features=None, instr_z_landmark_pzs=None,
                instr_z_direction_features=None, instr_z_direction_pzs=None):
        txt_embeds, txt_masks = self._encode_text(txt_ids, txt_lens, instr_z_direction_features, instr_z_direction_pzs,
                                                  instr_z_landmark_features, instr_z_landmark_pzs)
        split_embeds, split_lens, split_fused = self._encode_traj(
            traj_view_img_fts, traj_obj_img_fts, traj_loc_fts, traj_nav_types, traj_step_lens, traj_vp_view_lens,
            z_img_features, z_img_pzs)
        gmap_embeds = None
        if return_gmap_embeds:
            gmap_embeds = self.global_encoder(txt_embeds, txt_masks, split_embeds, split_lens, traj_vpids, traj_cand_vpids,
                                              gmap_vpids, gmap_step_ids, gmap_pos_fts, gmap_lens,
                                              graph_sprels=gmap_pair_dists, split_traj_fused_embeds=split_fused)
        vp_embeds = self.local_encoder(txt_embeds, txt_masks, split_embeds, split_lens, vp_pos_fts)
        if return_txt_embeds:
            return gmap_embeds, vp_embeds, txt_embeds
        return gmap_embeds, vp_embeds

    def forward_mlm(self, txt_ids, txt_lens, traj_view_img_fts, traj_obj_img_fts, traj_loc_fts, traj_nav_types,
                    traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens, traj_vpids, traj_cand_vpids, gmap_lens,
                    gmap_step_ids, gmap_pos_fts, gmap_pair_dists, gmap_vpids, vp_pos_fts, z_img_features=None,
                    z_img_pzs=None, traj_reverie_loc_fts=None, traj_reverie_obj_names=None,
                    instr_z_landmark_features=None, instr_z_landmark_pzs=None, instr_z_direction_features=None,
                    instr_z_direction_pzs=None):
        """text queries attend to the map / panorama tokens (roles swapped w.r.t. forward), P:597-648"""
        txt_embeds, txt_masks = self._encode_text(txt_ids, txt_lens, instr_z_direction_features, instr_z_direction_pzs,
                                                  instr_z_landmark_features, instr_z_landmark_pzs)
        ext_txt = extend_neg_masks(txt_masks)
        split_embeds, split_lens, split_fused = self._encode_traj(
            traj_view_img_fts, traj_obj_img_fts, traj_loc_fts, traj_nav_types, traj_step_lens, traj_vp_view_lens,
            z_img_features, z_img_pzs)
        gmap_in, gmap_masks = self.global_encoder.gmap_input_embedding(
            split_embeds, split_lens, traj_vpids, traj_cand_vpids, gmap_vpids, gmap_step_ids, gmap_pos_fts, gmap_lens,
            split_traj_fused_embeds=split_fused)
        g_txt = self.global_encoder.encoder(txt_embeds, ext_txt, gmap_in, extend_neg_masks(gmap_masks))
        vp_in, vp_masks = self.local_encoder.vp_input_embedding(split_embeds, split_lens, vp_pos_fts)
        v_txt = self.local_encoder.encoder(txt_embeds, ext_txt, vp_in, extend_neg_masks(vp_masks))
        return g_txt + v_txt

    def forward_cfp(self, txt_ids, txt_lens, traj_view_img_fts, traj_obj_img_fts, traj_loc_fts, traj_nav_types,
                    traj_step_lens, traj_vp_view_lens, traj_vp_obj_lens, traj_vpids, traj_cand_vpids, gmap_lens,
                    gmap_step_ids, gmap_pos_fts, gmap_pair_dists, gmap_vpids, vp_pos_fts, return_gmap_embeds=True,
                    z_img_features=None, z_img_pzs=None, traj_reverie_loc_fts=None, return_txt_embeds=False,
                    traj_reverie_obj_names=None, instr_z_landmark_features=None, instr_z_landmark_pzs=None,
                    instr_z_direction_features=None, instr_z_direction_pzs=None):
        txt_embeds, txt_masks = self._encode_text(txt_ids, txt_lens, instr_z_direction_features, instr_z_direction_pzs,
                                                  instr_z_landmark_features, instr_z_landmark_pzs)
        split_embeds, split_lens, split_fused = self._encode_traj(
            traj_view_img_fts, traj_obj_img_fts, traj_loc_fts, traj_nav_types, traj_step_lens, traj_vp_view_lens,
            z_img_features, z_img_pzs)
        gmap_embeds = None
        if return_gmap_embeds:
            gmap_embeds = self.global_encoder.forward_cfp(split_embeds, split_lens, traj_vpids, traj_cand_vpids, gmap_vpids,
                                                          gmap_step_ids, gmap_pos_fts, gmap_lens,
                                                          graph_sprels=gmap_pair_dists, split_traj_fused_embeds=split_fused)
        vp_embeds = self.local_encoder.forward_cfp(split_embeds, split_lens, vp_pos_fts)
        if return_txt_embeds:
            return gmap_embeds, vp_embeds, txt_embeds
        return gmap_embeds, vp_embeds


def _init_bert_weights(config):
    """v4 BertPreTrainedModel._init_weights: N(0, initializer_range) Linear / Embedding, zero bias, unit LayerNorm."""
    std = getattr(config, "initializer_range", 0.02)

    def fn(module):
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.padding_idx is not None:
                module.weight.data[module.padding_idx].zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
    return fn


_BERT_ARGS = ("txt_ids", "txt_lens", "traj_view_img_fts", "traj_obj_img_fts", "traj_loc_fts", "traj_nav_types",
              "traj_step_lens", "traj_vp_view_lens", "traj_vp_obj_lens", "traj_vpids", "traj_cand_vpids", "gmap_lens",
              "gmap_step_ids", "gmap_pos_fts", "gmap_pair_dists", "gmap_vpids", "vp_pos_fts")


class GlocalTextPathCMTPreTraining(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        tasks = config.pretrain_tasks
        if "mrc" in tasks or "og" in tasks:
            raise NotImplementedError("MRC / OG are REVERIE-only tasks, outside the hot-path scope (SURVEY.md 8a)")
        self.bert = GlocalTextPathCMT(config)
        H = config.hidden_size
        if "mlm" in tasks:
            self.mlm_head = G.BertOnlyMLMHead(config)
        if "sap" in tasks:
            self.global_sap_head = G.ClsPrediction(H)
            self.local_sap_head = G.ClsPrediction(H)
            self.sap_fuse_linear = G.ClsPrediction(H, input_size=H * 2) if config.glocal_fuse else None
        if "cfp" in tasks:
            self.tim_txt_head = G.BertPredictionHeadTransform(config)
            self.tim_global_head = G.BertPredictionHeadTransform(config)
            self.tim_local_head = G.BertPredictionHeadTransform(config)
            self.tim_fused_head = G.BertPredictionHeadTransform(config)
            for name in ("tim_txt_attn", "tim_global_attn", "tim_local_attn", "tim_fused_attn"):
                p = nn.Parameter(torch.empty(H, 1))
                nn.init.uniform_(p, -0.1, 0.1)
                setattr(self, name, p)
            self.temperature = config.cfp_temperature
        self.apply(_init_bert_weights(config))
        self.tie_weights()

    def tie_weights(self):
        if "mlm" in self.config.pretrain_tasks:
            self.mlm_head.predictions.decoder.weight = self.bert.embeddings.word_embeddings.weight

    # ------------------------------------------------------------------------------------------
    def forward(self, batch, task, compute_loss=True):
        """``batch``: the collate dict of P/data/tasks.py (tensors on the device).  The Python-list fields (viewpoint-id
        strings, step lengths) are turned into index tensors on the host first (batching.prepare_pretrain), then the
        shape-static device path runs (forward_prepared)."""
        batch = defaultdict(lambda: None, batch)
        task0 = task.split("_")[0]
        if task0 in ("mrc", "og", "valid"):
            raise NotImplementedError("task %r is REVERIE-only, outside the hot-path scope" % task)
        if task0 not in ("mlm", "sap", "cfp"):
            raise ValueError("invalid task")
        P = batching.prepare_pretrain(batch, task0, pad=None, pano_fusion=bool(self.config.adaptive_pano_fusion))
        if task0 == "cfp" and not batch["extra_heads"]:
            P["_no_extra_heads"] = True
        return self.forward_prepared(P, task0, compute_loss)

    def forward_prepared(self, P, task, compute_loss=True):
        """P: dict of device tensors from batching.prepare_pretrain (optionally padded to static bucket shapes).
        Returns what the reference returns: per-sample (per-masked-token) loss vector, or the logits / embeddings."""
        if task == "mlm":
            return self._mlm(P, compute_loss)
        if task == "sap":
            return self._sap(P, compute_loss)
        if task == "cfp":
            return self._cfp(P, compute_loss)
        raise ValueError("invalid task")

    def scalar_loss(self, P, task):
        """mean of the loss vector over the REAL rows (``loss.mean()`` of P/train_r2r_goat.py:317; padded rows are 0)"""
        return self.forward_prepared(P, task, True).sum() * P["loss_inv"][0]

    def _mlm(self, P, compute_loss):
        """text queries attend to the map / panorama tokens (P/model/vilmodel_goat.py:597-648), then the masked tokens
        go through the tied vocabulary projection (P/model/pretrain_goat.py:188-224)"""
        bert = self.bert
        txt, txt_masks, views, fused = bert.encode_prepared(P)
        ext_txt = extend_neg_masks(txt_masks)
        gmap_in, gmap_masks = bert.gmap_inputs_prepared(P, views, fused)
        g_txt = bert.global_encoder.encoder(txt, ext_txt, gmap_in, extend_neg_masks(gmap_masks))
        vp_in, vp_masks = bert.vp_inputs_prepared(P, views)
        v_txt = bert.local_encoder.encoder(txt, ext_txt, vp_in, extend_neg_masks(vp_masks))
        txt_embeds = g_txt + v_txt
        H = txt_embeds.shape[-1]
        masked_output = Fn.SegmentReduceFn.apply(txt_embeds.reshape(-1, H), P["mlm_rows"], False)
        if compute_loss:
            # chunked vocabulary projection + cross-entropy: the [n, 50265] logits never exist in memory
            return self.mlm_head.loss(masked_output, P["mlm_labels"], ignore_index=-1)
        return self.mlm_head(masked_output)

    def _fuse_weights(self, gmap_embeds, vp_embeds):
        if self.sap_fuse_linear is None:
            return 0.5
        return torch.sigmoid(self.sap_fuse_linear(torch.cat([gmap_embeds[:, 0], vp_embeds[:, 0]], 1)))

    def _sap(self, P, compute_loss):
        """P/model/pretrain_goat.py:286-354"""
        bert = self.bert
        txt, txt_masks, views, fused = bert.encode_prepared(P)
        gmap_in, gmap_masks = bert.gmap_inputs_prepared(P, views, fused)
        ge = bert.global_encoder
        gmap_embeds = ge.encoder(gmap_in, gmap_masks, txt, txt_masks, graph_sprels=ge.sprels(P["gmap_pair_dists"]))
        vp_in, vp_masks = bert.vp_inputs_prepared(P, views)
        vp_embeds = bert.local_encoder.encoder(vp_in, vp_masks, txt, txt_masks)
        fuse_weights = self._fuse_weights(gmap_embeds, vp_embeds)
        neg_inf = -float("inf")
        global_logits = self.global_sap_head(gmap_embeds).squeeze(2) * fuse_weights
        global_logits = global_logits.masked_fill(P["gmap_visited_masks"], neg_inf)
        global_logits = global_logits.masked_fill(gmap_masks.logical_not(), neg_inf)
        local_logits = self.local_sap_head(vp_embeds).squeeze(2) * (1 - fuse_weights)
        Nq = local_logits.size(1)
        cur_nav = P["nav_types"][P["last_rows"].view(-1).long()] != 1                     # current panorama: not navigable
        cur_nav = cur_nav[:, :Nq - 1]
        if cur_nav.shape[1] < Nq - 1:
            cur_nav = torch.cat([cur_nav, cur_nav.new_ones(len(cur_nav), Nq - 1 - cur_nav.shape[1])], 1)
        vp_nav_masks = torch.cat([cur_nav.new_zeros(len(cur_nav), 1), cur_nav], 1)        # [stop] is never masked
        local_logits = local_logits.masked_fill(vp_nav_masks, neg_inf)
        fused_logits = G.fuse_logits(global_logits, local_logits, P["fuse_idx"])
        gl, ll = P["global_act_labels"], P["local_act_labels"]
        if compute_loss:
            return G.cross_entropy(global_logits, gl) + G.cross_entropy(local_logits, ll) + G.cross_entropy(fused_logits, gl)
        return global_logits, local_logits, fused_logits, gl, ll

    def _cfp(self, P, compute_loss):
        """P/model/pretrain_goat.py:467-541; the poolings run over the batch's own padded lengths (n_gmap / n_vp / n_txt)"""
        bert = self.bert
        txt_embeds, txt_masks, views, fused = bert.encode_prepared(P)
        gmap_in, gmap_masks = bert.gmap_inputs_prepared(P, views, fused)
        gmap_embeds = bert.global_encoder.tim_self_encoder(gmap_in, extend_neg_masks(gmap_masks))[0]
        vp_in, vp_masks = bert.vp_inputs_prepared(P, views)
        vp_embeds = bert.local_encoder.tim_self_encoder(vp_in, extend_neg_masks(vp_masks))[0]
        if not P.get("_no_extra_heads"):        # a non-empty python list after collate: always true in the reference loop
            gmap_embeds = self.tim_global_head(gmap_embeds)
            vp_embeds = self.tim_local_head(vp_embeds)
            txt_embeds = self.tim_txt_head(txt_embeds)
        fuse_weights = self._fuse_weights(gmap_embeds, vp_embeds)
        gmap_outputs = G.attn_pool_cfp(gmap_embeds, self.tim_global_attn, P["n_gmap"])
        vp_outputs = G.attn_pool_cfp(vp_embeds, self.tim_local_attn, P["n_vp"])
        txt_outputs = G.attn_pool_cfp(txt_embeds, self.tim_txt_attn, P["n_txt"])
        fused_outputs = gmap_outputs * fuse_weights + vp_outputs * (1 - fuse_weights)
        if compute_loss:
            T = self.temperature
            return (G.infonce(gmap_outputs, txt_outputs, T) + G.infonce(vp_outputs, txt_outputs, T) +
                    G.infonce(fused_outputs, txt_outputs, T))
        return gmap_outputs, vp_outputs, fused_outputs, txt_outputs
