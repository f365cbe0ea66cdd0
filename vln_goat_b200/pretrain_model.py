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
        batch = defaultdict(lambda: None, batch)
        args = [batch[k] for k in _BERT_ARGS]
        kw = dict(traj_reverie_loc_fts=batch["traj_reverie_loc_fts"], traj_reverie_obj_names=batch["traj_reverie_obj_names"],
                  instr_z_landmark_features=batch["instr_z_landmark_features"], instr_z_landmark_pzs=batch["instr_z_landmark_pzs"],
                  instr_z_direction_features=batch["instr_z_direction_features"],
                  instr_z_direction_pzs=batch["instr_z_direction_pzs"], z_img_features=batch["img_z_features"],
                  z_img_pzs=batch["img_z_pzs"])
        if task.startswith("mlm"):
            return self.forward_mlm(args, kw, batch["txt_labels"], compute_loss)
        if task.startswith("sap"):
            return self.forward_sap(args, kw, batch["gmap_visited_masks"], batch["global_act_labels"],
                                    batch["local_act_labels"], compute_loss)
        if task.startswith("cfp"):
            return self.forward_cfp(args, kw, compute_loss, batch["extra_heads"])
        if task.startswith(("mrc", "og", "valid_sap_og")):
            raise NotImplementedError("task %r is REVERIE-only, outside the hot-path scope" % task)
        raise ValueError("invalid task")

    def forward_mlm(self, args, kw, txt_labels, compute_loss):
        txt_embeds = self.bert.forward_mlm(*args, **kw)
        sel = txt_labels != -1
        masked_output = txt_embeds[sel]                       # only the masked tokens go through the vocabulary GEMM
        prediction_scores = self.mlm_head(masked_output)
        if compute_loss:
            return G.cross_entropy(prediction_scores, txt_labels[sel])
        return prediction_scores

    def _fuse_weights(self, gmap_embeds, vp_embeds):
        if self.sap_fuse_linear is None:
            return 0.5
        return torch.sigmoid(self.sap_fuse_linear(torch.cat([gmap_embeds[:, 0], vp_embeds[:, 0]], 1)))

    def forward_sap(self, args, kw, gmap_visited_masks, global_act_labels, local_act_labels, compute_loss):
        (txt_ids, _, _, _, _, traj_nav_types, traj_step_lens, _, _, _, traj_cand_vpids, gmap_lens, _, _, _, gmap_vpids,
         _) = args
        gmap_embeds, vp_embeds = self.bert(*args, **kw)
        fuse_weights = self._fuse_weights(gmap_embeds, vp_embeds)
        neg_inf = -float("inf")
        global_logits = self.global_sap_head(gmap_embeds).squeeze(2) * fuse_weights
        global_logits = global_logits.masked_fill(gmap_visited_masks, neg_inf)
        global_logits = global_logits.masked_fill(gen_seq_masks(gmap_lens).logical_not(), neg_inf)
        local_logits = self.local_sap_head(vp_embeds).squeeze(2) * (1 - fuse_weights)
        Nq = local_logits.size(1)
        cur_nav = torch.stack([x[-1] != 1 for x in torch.split(traj_nav_types, traj_step_lens)], 0)[:, :Nq - 1]
        vp_nav_masks = torch.cat([cur_nav.new_zeros(len(cur_nav), 1), cur_nav], 1)      # [stop] is never masked
        local_logits = local_logits.masked_fill(vp_nav_masks, neg_inf)
        idx = G.build_fusion_index(gmap_vpids, gmap_visited_masks, [c[-1] for c in traj_cand_vpids], Nq, 1, 1)
        fused_logits = G.fuse_logits(global_logits, local_logits, idx.to(global_logits.device))
        if compute_loss:
            return (G.cross_entropy(global_logits, global_act_labels) + G.cross_entropy(local_logits, local_act_labels) +
                    G.cross_entropy(fused_logits, global_act_labels))
        return global_logits, local_logits, fused_logits, global_act_labels, local_act_labels

    def forward_cfp(self, args, kw, compute_loss, extra_heads=False):
        kw = dict(kw)
        kw["return_txt_embeds"] = True
        gmap_embeds, vp_embeds, txt_embeds = self.bert.forward_cfp(*args, **kw)
        if extra_heads:        # a non-empty python list after collate: always true in the reference loop
            gmap_embeds = self.tim_global_head(gmap_embeds)
            vp_embeds = self.tim_local_head(vp_embeds)
            txt_embeds = self.tim_txt_head(txt_embeds)
        fuse_weights = self._fuse_weights(gmap_embeds, vp_embeds)
        gmap_outputs = G.attn_pool_cfp(gmap_embeds, self.tim_global_attn)
        vp_outputs = G.attn_pool_cfp(vp_embeds, self.tim_local_attn)
        txt_outputs = G.attn_pool_cfp(txt_embeds, self.tim_txt_attn)
        fused_outputs = gmap_outputs * fuse_weights + vp_outputs * (1 - fuse_weights)
        if compute_loss:
            T = self.temperature
            return (G.infonce(gmap_outputs, txt_outputs, T) + G.infonce(vp_outputs, txt_outputs, T) +
                    G.infonce(fused_outputs, txt_outputs, T))
        return gmap_outputs, vp_outputs, fused_outputs, txt_outputs
