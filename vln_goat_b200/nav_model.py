"""Fine-tuning / navigation model of the GOAT path, backed by libgoat_sm100.

Drop-in for the reference's ``GlocalTextPathNavCMT`` (M/models/vilmodel_GOAT.py:556-927) and the ``VLNBert``
wrapper (M/models/model.py:12-38): ``forward(mode, batch)`` with modes ``language`` (once per rollout),
``panorama`` and ``navigation`` (every step), ``instr_zdict_update`` and ``extract_cfp_features``; the batch
dicts are ``defaultdict(lambda: None)`` (a missing key switches the feature off) with the keys of SURVEY.md
appendix A.2; BACL (back-door) and FACL (front-door) interventions as configured.  Same ``state_dict`` keys, so
``Seq2SeqAgent.load`` and the pretrain -> fine-tune remap of M/models/vlnbert_init.py:52-69 keep working.
M/ = map_nav_src/ of CrystalSixone/VLN-GOAT.  R2R / RxR only (no object branch).
"""
import collections

import torch
from torch import nn

from . import goat_blocks as G
from . import modules as M
from .modules import layer_norm, linear
from .pretrain_model import _init_bert_weights


class GlocalTextPathNavCMT(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        if getattr(config, "obj_feat_size", 0) > 0:
            raise NotImplementedError("object grounding (REVERIE / SOON) is outside the hot-path scope")
        H = config.hidden_size
        extract = getattr(config, "mode", None) == "extract_cfp_features"
        self.embeddings = G.RobertaEmbeddings(config, tuple_output=False)
        if config.do_back_txt or config.do_front_txt:
            self.lang_encoder = G.LanguageEncoderDo(config)
        else:
            self.lang_encoder = M.LanguageEncoder(config)
        self.img_embeddings = G.CausalImageEmbeddings(config)
        self.local_encoder = G.LocalVPEncoder(config, with_cfp=extract)
        self.global_encoder = G.GlobalMapEncoder(config, with_cfp=extract)
        self.global_sap_head = G.ClsPrediction(H)
        self.local_sap_head = G.ClsPrediction(H)
        self.sap_fuse_linear = G.ClsPrediction(H, input_size=H * 2) if config.glocal_fuse else None
        self.object_encoder = None
        self.extra_drop = nn.Dropout(0.2)
        self.gmap_pooler = G.BertPooler(config)
        self.vp_pooler = G.BertPooler(config)
        self.txt_pooler = G.BertPooler(config)
        self.local_his_map = nn.Linear(H * 3, H)
        self.local_his_ln = M.BertLayerNorm(H, eps=config.layer_norm_eps)
        self.drop_env = nn.Dropout(p=config.feat_dropout)
        if extract:
            self.tim_local_head = G.BertPredictionHeadTransform(config)
            self.tim_local_attn = nn.Parameter(torch.empty(H, 1).uniform_(-0.1, 0.1))
            self.temperature = config.cfp_temperature
        if config.do_front_img:
            self.front_local_encoder = G.FrontDoorEncoder(config)
        if extract:
            self.tim_global_head = G.BertPredictionHeadTransform(config)
            self.tim_global_attn = nn.Parameter(torch.empty(H, 1).uniform_(-0.1, 0.1))
        if config.do_front_his:
            self.front_global_encoder = G.FrontDoorEncoder(config)
        if extract:
            self.tim_txt_head = G.BertPredictionHeadTransform(config)
            self.tim_txt_attn = nn.Parameter(torch.empty(H, 1).uniform_(-0.1, 0.1))
        if config.do_front_txt:
            self.front_txt_encoder = G.FrontDoorEncoder(config)      # constructed, never called (reference :607-608)
        self.apply(_init_bert_weights(config))
        # rollout-level K|V projection cache of the cross-attentions; config.kv_cache=False turns it off
        self._kv_cache = M.KVCache() if getattr(config, "kv_cache", True) else None
        if getattr(config, "fix_lang_embedding", False) or getattr(config, "fix_local_branch", False):
            for mod in (self.embeddings, self.lang_encoder):
                for p in mod.parameters():
                    p.requires_grad = False
        if getattr(config, "fix_pano_embedding", False) or getattr(config, "fix_local_branch", False):
            for p in self.img_embeddings.parameters():
                p.requires_grad = False
        if getattr(config, "fix_local_branch", False):
            for mod in (self.local_encoder, self.local_sap_head):
                for p in mod.parameters():
                    p.requires_grad = False

    # ------------------------------------------------------------------------------------------
    def forward_text(self, txt_ids, txt_masks, instr_z_direction_features=None, instr_z_direction_pzs=None,
                     instr_z_landmark_features=None, instr_z_landmark_pzs=None, front_txt_embeds=None):
        txt_embeds = self.embeddings(txt_ids)
        if self.config.do_back_txt or self.config.do_front_txt:
            return self.lang_encoder(txt_embeds, txt_masks, instr_z_direction_features, instr_z_direction_pzs,
                                     instr_z_landmark_features, instr_z_landmark_pzs, front_txt_embeds)
        return self.lang_encoder(txt_embeds, txt_masks)

    def forward_panorama_do_per_step(self, view_img_fts, loc_fts, nav_types, view_lens, z_img_features=None,
                                     z_img_pzs=None, reverie_obj_fts=None, reverie_obj_lens=None, reverie_obj_names=None):
        if reverie_obj_fts is not None:
            raise NotImplementedError("object features (REVERIE / SOON) are outside the hot-path scope")
        return self.img_embeddings.encode(view_img_fts, loc_fts, view_lens, z_img_features, z_img_pzs, loc_after_do=True)

    def forward_navigation_per_step(self, txt_embeds, txt_masks, gmap_img_embeds, gmap_step_ids, gmap_pos_fts, gmap_masks,
                                    gmap_pair_dists, gmap_visited_masks, gmap_vpids, vp_img_embeds, vp_pos_fts, vp_masks,
                                    vp_nav_masks, vp_obj_masks, vp_cand_vpids, front_vp_feats=None, front_gmap_feats=None,
                                    flops_count=False, fuse_idx=None):
        """fuse_idx: optional int32 [B,G,K] device tensor from goat_blocks.build_fusion_index, built by the caller on the
        host (then ``gmap_vpids`` / ``vp_cand_vpids`` are not needed and the step has no host synchronisation: a whole
        teacher-forced rollout can be captured in one CUDA graph)."""
        with M.kv_cache_scope(self._kv_cache):
            return self._navigation_step(txt_embeds, txt_masks, gmap_img_embeds, gmap_step_ids, gmap_pos_fts, gmap_masks,
                                         gmap_pair_dists, gmap_visited_masks, gmap_vpids, vp_img_embeds, vp_pos_fts,
                                         vp_masks, vp_nav_masks, vp_obj_masks, vp_cand_vpids, front_vp_feats,
                                         front_gmap_feats, flops_count, fuse_idx)

    def _navigation_step(self, txt_embeds, txt_masks, gmap_img_embeds, gmap_step_ids, gmap_pos_fts, gmap_masks,
                         gmap_pair_dists, gmap_visited_masks, gmap_vpids, vp_img_embeds, vp_pos_fts, vp_masks,
                         vp_nav_masks, vp_obj_masks, vp_cand_vpids, front_vp_feats, front_gmap_feats, flops_count,
                         fuse_idx=None):
        ge, le = self.global_encoder, self.local_encoder
        # global branch
        gmap_embeds = gmap_img_embeds + ge.step_embed(gmap_step_ids) + G._pos_embed(ge.gmap_pos_embeddings, gmap_pos_fts)
        graph_sprels = ge.sprels(gmap_pair_dists)
        if front_gmap_feats is not None:
            gmap_embeds = self.front_global_encoder(gmap_embeds, front_gmap_feats, gmap_masks)
        gmap_embeds = ge.encoder(gmap_embeds, gmap_masks, txt_embeds, txt_masks, graph_sprels=graph_sprels)
        # local branch
        vp_embeds = vp_img_embeds + G._pos_embed(le.vp_pos_embeddings, vp_pos_fts)
        if front_vp_feats is not None:
            vp_embeds = self.front_local_encoder(vp_embeds, front_vp_feats, vp_masks)
        vp_embeds = le.encoder(vp_embeds, vp_masks, txt_embeds, txt_masks)
        # action logits
        if self.sap_fuse_linear is None:
            fuse_weights = 0.5
        else:
            fuse_weights = torch.sigmoid(self.sap_fuse_linear(torch.cat([gmap_embeds[:, 0], vp_embeds[:, 0]], 1)))
        neg_inf = -float("inf")
        global_logits = self.global_sap_head(gmap_embeds).squeeze(2) * fuse_weights
        local_logits = self.local_sap_head(vp_embeds).squeeze(2) * (1 - fuse_weights)
        global_logits = global_logits.masked_fill(gmap_visited_masks, neg_inf)
        global_logits = global_logits.masked_fill(gmap_masks.logical_not(), neg_inf)
        local_logits = local_logits.masked_fill(vp_nav_masks.logical_not(), neg_inf)
        if flops_count:
            fused_logits = global_logits.clone()
            fused_logits[:, 0] = fused_logits[:, 0] + local_logits[:, 0]
        else:
            if fuse_idx is None:
                fuse_idx = G.build_fusion_index(gmap_vpids, gmap_visited_masks, vp_cand_vpids, local_logits.size(1), 2, 2)
            fused_logits = G.fuse_logits(global_logits, local_logits, fuse_idx.to(global_logits.device))
        # per-step history token
        cls = torch.cat((self.gmap_pooler(gmap_embeds, location=0), self.vp_pooler(vp_embeds, location=0),
                         self.txt_pooler(txt_embeds, location=0)), dim=-1)
        cls_embeds = layer_norm(self.local_his_ln, linear(self.local_his_map, cls))
        return {"gmap_embeds": gmap_embeds, "vp_embeds": vp_embeds, "global_logits": global_logits,
                "local_logits": local_logits, "fused_logits": fused_logits, "obj_logits": None, "txt_embeds": txt_embeds,
                "cls_embeds": cls_embeds}

    def forward(self, mode, batch, **kwargs):
        if mode == "language":
            if self._kv_cache is not None:
                self._kv_cache.clear()      # a new rollout: new instruction embeddings
            return self.forward_text(batch["txt_ids"], batch["txt_masks"], batch["instr_z_direction_features"],
                                     batch["instr_z_direction_pzs"], batch["instr_z_landmark_features"],
                                     batch["instr_z_landmark_pzs"], batch["front_txt_feats"])
        if mode == "panorama":
            return self.forward_panorama_do_per_step(batch["view_img_fts"], batch["loc_fts"], batch["nav_types"],
                                                     batch["view_lens"], batch["z_img_features"], batch["z_img_pzs"],
                                                     batch["reverie_obj_img_fts"], batch["reverie_obj_lens"],
                                                     batch["reverie_obj_names"])
        if mode == "navigation":
            return self.forward_navigation_per_step(
                batch["txt_embeds"], batch["txt_masks"], batch["gmap_img_embeds"], batch["gmap_step_ids"],
                batch["gmap_pos_fts"], batch["gmap_masks"], batch["gmap_pair_dists"], batch["gmap_visited_masks"],
                batch["gmap_vpids"], batch["vp_img_embeds"], batch["vp_pos_fts"], batch["vp_masks"], batch["vp_nav_masks"],
                batch["vp_obj_masks"], batch["vp_cand_vpids"], batch["front_vp_feats"], batch["front_gmap_feats"],
                flops_count=batch["flops_count"], fuse_idx=batch["fuse_idx"])
        if mode == "instr_zdict_update":
            return self.forward_text(batch["z_txt"], batch["z_txt_mask"],
                                     instr_z_direction_features=batch["instr_z_direction_features"],
                                     instr_z_direction_pzs=batch["instr_z_direction_pzs"],
                                     instr_z_landmark_features=batch["instr_z_landmark_features"],
                                     instr_z_landmark_pzs=batch["instr_z_landmark_pzs"],
                                     front_txt_embeds=batch["front_txt_feats"])
        if mode == "extract_cfp_features":
            txt_embeds = self.forward_text(batch["txt_ids"], batch["txt_masks"])
            split_embeds, split_lens, split_fused = self.img_embeddings(
                batch["traj_view_img_fts"], batch["traj_loc_fts"], batch["traj_nav_types"], batch["traj_step_lens"],
                batch["traj_vp_view_lens"], None)
            gmap_embeds = self.global_encoder.forward_cfp(
                split_embeds, split_lens, batch["traj_vpids"], batch["traj_cand_vpids"], batch["gmap_vpids"],
                batch["gmap_step_ids"], batch["gmap_pos_fts"], batch["gmap_lens"], graph_sprels=batch["gmap_pair_dists"],
                split_traj_fused_embeds=split_fused)
            vp_embeds = self.local_encoder.forward_cfp(split_embeds, split_lens, batch["vp_pos_fts"])
            return {"txt_outputs": G.attn_pool_cfp(self.tim_txt_head(txt_embeds), self.tim_txt_attn),
                    "vp_outputs": G.attn_pool_cfp(self.tim_local_head(vp_embeds), self.tim_local_attn),
                    "gmap_outputs": G.attn_pool_cfp(self.tim_global_head(gmap_embeds), self.tim_global_attn)}
        raise ValueError("unknown mode %r" % (mode,))


class VLNBert(nn.Module):
    """M/models/model.py:12-38.  ``args`` is the fine-tune argparse namespace; ``config`` may be passed directly
    (tests / synthetic benchmarks) instead of being assembled by ``nav_config_from_args``."""

    def __init__(self, args, config=None):
        super().__init__()
        self.args = args
        self.vln_bert = GlocalTextPathNavCMT(config if config is not None else nav_config_from_args(args))
        self.drop_env = nn.Dropout(p=args.feat_dropout)

    def forward(self, mode, batch):
        batch = collections.defaultdict(lambda: None, batch)
        if mode == "panorama" and not batch["already_dropout"]:
            from . import functional as Fn
            batch["view_img_fts"] = Fn.dropout(batch["view_img_fts"].float(), self.drop_env.p, self.training)
        return self.vln_bert(mode, batch)


def nav_config_from_args(args):
    """The HF config the reference assembles in code (M/models/vlnbert_init.py:79-154), on the roberta-base defaults
    it starts from (layer_norm_eps 1e-5, pad_token_id 1)."""
    from .config import GoatConfig
    g = lambda k, d: getattr(args, k, d)
    return GoatConfig(
        layer_norm_eps=1e-5, pad_token_id=1, dataset=g("dataset", "r2r"), mode=g("mode", None), max_action_steps=100,
        image_feat_size=g("image_feat_size", 768), angle_feat_size=g("angle_feat_size", 4), obj_feat_size=g("obj_feat_size", 0),
        num_l_layers=g("num_l_layers", 6), num_pano_layers=g("num_pano_layers", 2), num_x_layers=g("num_x_layers", 3),
        num_top_layer=g("num_x_layers", 3), graph_sprels=g("graph_sprels", True), glocal_fuse=g("fusion", "dynamic") == "dynamic",
        fix_lang_embedding=g("fix_lang_embedding", False), fix_pano_embedding=g("fix_pano_embedding", False),
        fix_local_branch=g("fix_local_branch", False), update_lang_bert=not g("fix_lang_embedding", False),
        feat_dropout=g("feat_dropout", 0.4), adaptive_pano_fusion=g("adaptive_pano_fusion", True),
        do_back_img=g("do_back_img", False), do_back_txt=g("do_back_txt", False), do_front_img=g("do_front_img", False),
        do_front_his=g("do_front_his", False), do_front_txt=g("do_front_txt", False), cfp_temperature=g("cfp_temperature", 1.0),
        do_back_txt_type=g("do_back_txt_type", "type_2"), do_back_img_type=g("do_back_img_type", "type_1"),
        do_add_method=g("do_add_method", "door"), hidden_dropout_prob=g("dropout", 0.1), name="R2R", use_lang2visn_attn=False)
