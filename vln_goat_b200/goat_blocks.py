"""GOAT body blocks above the transformer stacks, backed by libgoat_sm100 (R2R / RxR variant).

Same class names, constructor (``config``), ``state_dict`` keys and forward semantics as the reference:

  RobertaEmbeddings             P/model/Bert_backbone.py:56-121   (M/ copy returns the tensor instead of a tuple)
  LanguageEncoderDo             P/model/vilmodel_goat.py:46-159,  M/models/vilmodel_GOAT.py:55-162   (BACL / FACL text)
  CausalImageEmbeddings         P/model/vilmodel_goat.py:234-364, M/models/vilmodel_GOAT.py:164-316  (BACL image, pano
                                encoder, adaptive panorama fusion)
  LocalVPEncoder                P/model/vilmodel_goat.py:366-410, M/models/vilmodel_GOAT.py:318-385
  GlobalMapEncoder              P/model/vilmodel_goat.py:412-527, M/models/vilmodel_GOAT.py:387-510
  ClsPrediction                 P/model/pretrain_goat.py:27-38,   M/models/vilmodel_GOAT.py:512-524
  FrontDoorEncoder              M/models/vilmodel_GOAT.py:526-554 (FACL)
  BertPooler, BertPredictionHeadTransform, BertLMPredictionHead, BertOnlyMLMHead   P/model/Bert_backbone.py:783-838

(P/ = pretrain_src/, M/ = map_nav_src/.)  ``nn.Linear`` / ``nn.LayerNorm`` / ``nn.Embedding`` objects are parameter
containers; the math runs in the CUDA library.  The REVERIE / SOON object branches are out of scope (SURVEY.md 8a:
MRC / OG are REVERIE-only) and raise.  The host-side Python loops of the reference over viewpoint-id strings
(global-map aggregation, logit fusion) become index lists built once on the host plus one gather-reduce kernel.
"""
import torch
from torch import nn

from . import functional as Fn
from . import modules as M
from . import ops, runtime
from .modules import Act, BertAttention, BertLayerNorm, CrossmodalEncoder, RobertaAttention, RobertaLayer, \
    create_transformer_encoder, extend_neg_masks, gen_seq_masks, layer_norm, linear


def _no_objects(config):
    if getattr(config, "name", "R2R") in ("REVERIE", "SOON"):
        raise NotImplementedError("the REVERIE / SOON object branches are outside the hot-path scope (SURVEY.md 8a)")


# --------------------------------------------------------------------------------------
# embeddings / heads
# --------------------------------------------------------------------------------------
class RobertaEmbeddings(nn.Module):
    """LN(word[ids] + type[0] + pos[arange(L)]) -> dropout.  ``tuple_output`` selects the pretrain flavour, which
    returns (embeddings, z_direction, z_landmark) (P/model/Bert_backbone.py:112-116)."""

    def __init__(self, config, tuple_output=False):
        super().__init__()
        self.config = config
        self.tuple_output = tuple_output
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        # registered here so the state_dict order is the reference's (word, position, type); re-created below
        self.position_embeddings = nn.Embedding(1, 1)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)))
        self.padding_idx = config.pad_token_id
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size,
                                                padding_idx=self.padding_idx)

    def run(self, input_ids):
        if not input_ids.is_cuda:
            raise RuntimeError("vln_goat_b200 blocks need CUDA tensors: there is no CPU fallback on this path")
        B, L = input_ids.shape
        s = Fn.EmbedFn.apply(input_ids, self.word_embeddings.weight, self.position_embeddings.weight,
                             self.token_type_embeddings.weight, self.padding_idx)
        y32, _ = Fn.LayerNormFn.apply(s, self.LayerNorm.weight, self.LayerNorm.bias, self.LayerNorm.eps, torch.float32)
        y32 = Fn.dropout(y32, self.dropout.p, self.training)
        return y32.view(B, L, -1)

    def forward(self, input_ids=None, token_type_ids=None, position_ids=None, inputs_embeds=None,
                instr_z_direction_features=None, instr_z_landmark_features=None):
        if inputs_embeds is not None or position_ids is not None:
            raise NotImplementedError("inputs_embeds / explicit position_ids are not used on the GOAT path")
        if token_type_ids is not None and token_type_ids.numel() and int(token_type_ids.max()) != 0:
            raise NotImplementedError("GOAT passes all-zero token types (P/model/vilmodel_goat.py:557)")
        emb = self.run(input_ids)
        if not self.tuple_output:
            return emb
        zd = zl = None
        if instr_z_direction_features is not None:
            zd = instr_z_direction_features.to(torch.float32)
            zl = instr_z_landmark_features.to(torch.float32)
        return emb, zd, zl


class ClsPrediction(nn.Module):
    """Linear -> ReLU -> LN(1e-12) -> Linear(-> output_size)"""

    def __init__(self, hidden_size, input_size=None, output_size=1):
        super().__init__()
        if input_size is None:
            input_size = hidden_size
        self.net = nn.Sequential(nn.Linear(input_size, hidden_size), nn.ReLU(), BertLayerNorm(hidden_size, eps=1e-12),
                                 nn.Linear(hidden_size, output_size))

    def forward(self, x):
        h = linear(self.net[0], x, ops.ACT_RELU)
        h = layer_norm(self.net[2], h)
        return linear(self.net[3], h)


class BertPooler(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states, location=0):
        return linear(self.dense, hidden_states[:, location], ops.ACT_TANH)


class BertPredictionHeadTransform(nn.Module):
    """LN(gelu(W x + b))  -- the CFP extra heads and the MLM transform"""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        if config.hidden_act != "gelu":
            raise ValueError("libgoat_sm100 implements the reference's erf GELU only")
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)

    def forward(self, hidden_states):
        return layer_norm(self.LayerNorm, linear(self.dense, hidden_states, ops.ACT_GELU))


class BertLMPredictionHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        self.decoder = nn.Linear(config.hidden_size, config.vocab_size, bias=False)   # tied to the word embeddings
        self.bias = nn.Parameter(torch.zeros(config.vocab_size))

    def forward(self, hidden_states):
        h = self.transform(hidden_states)
        shp = h.shape
        cdt = runtime.compute_dtype()
        y = Fn.LinearFn.apply(h.reshape(-1, shp[-1]).contiguous(), None, self.decoder.weight, self.bias,
                              runtime.wc(self.decoder.weight, cdt), ops.ACT_NONE, cdt)
        return y.view(shp[:-1] + (self.decoder.weight.shape[0],))


    def loss(self, hidden_states, labels, ignore_index=-1):
        """cross_entropy(self(hidden_states), labels, reduction='none') without the [n, vocab] logits in memory"""
        h = self.transform(hidden_states)
        cdt = runtime.compute_dtype()
        return Fn.VocabXentFn.apply(h.reshape(-1, h.shape[-1]), self.decoder.weight, self.bias, labels.reshape(-1),
                                    ignore_index, runtime.wc(self.decoder.weight, cdt), cdt)


class BertOnlyMLMHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.predictions = BertLMPredictionHead(config)

    def forward(self, sequence_output):
        return self.predictions(sequence_output)

    def loss(self, sequence_output, labels, ignore_index=-1):
        return self.predictions.loss(sequence_output, labels, ignore_index)


def cross_entropy(logits, labels, ignore_index=-100):
    """F.cross_entropy(..., reduction='none') through goat_xent (accepts -inf logits and strided views)."""
    return Fn.XentFn.apply(logits, labels, ignore_index)


def attn_pool_cfp(x, attn_vec, n_valid=None):
    """tanh(sum_n softmax_n(tanh(x_n) . a) x_n)   P/model/pretrain_goat.py:502-515
    n_valid: CUDA int32 [1], pool over the first n_valid tokens only (statically padded buffers)"""
    return Fn.AttnPoolFn.apply(x, attn_vec, None, 1, n_valid)


def infonce(a, b, temperature):
    """(CE(a b^T / T, diag) + CE((a b^T / T)^T, diag)) / 2   P/model/pretrain_goat.py:519-532"""
    n = a.shape[0]
    target = torch.arange(n, device=a.device)
    sim = Fn.LinearFn.apply(a.contiguous(), None, b, None, b.detach().contiguous(), ops.ACT_NONE, torch.float32)
    if temperature != 1.0:
        sim = sim / temperature
    return (cross_entropy(sim, target) + cross_entropy(sim.t(), target)) / 2.0


def door_gate(aug, ori, aug_linear, ori_linear):
    return Fn.DoorGateFn.apply(aug, ori, aug_linear.weight, aug_linear.bias, ori_linear.weight, ori_linear.bias)


# --------------------------------------------------------------------------------------
# text: language encoder with the back-door / front-door interventions
# --------------------------------------------------------------------------------------
class LanguageEncoderDo(nn.Module):
    """6 x RobertaLayer, then BACL-text (type_1: p(z)-weighted dictionary sums; type_2: cross-attention onto the
    direction / landmark dictionaries) and FACL-text (cross-attention onto the front-door prototypes), merged by
    add / door.  ``pretrain_layout`` selects the parameter set of the pretrain class (P/model/vilmodel_goat.py:61-85)."""

    def __init__(self, config, pretrain_layout=False):
        super().__init__()
        self.config = config
        self.pretrain_layout = pretrain_layout
        self.num_l_layers = config.num_l_layers
        self.update_lang_bert = config.update_lang_bert
        self.layer = nn.ModuleList([RobertaLayer(config) for _ in range(self.num_l_layers)])
        if not self.update_lang_bert:
            for _, param in self.layer.named_parameters():
                param.requires_grad = False
        H = config.hidden_size
        eps = config.layer_norm_eps
        do_any = config.do_back_txt if pretrain_layout else (config.do_back_txt or config.do_front_txt)
        if do_any:
            if pretrain_layout and getattr(config, "z_cross_attn", False):
                self.z_direc_cross_attn = RobertaAttention(config)
                self.z_landm_cross_attn = RobertaAttention(config)
            self.z_txt_linear = nn.Linear(H, H)
            self.z_direct_linear = nn.Linear(H, H)
            self.z_landm_linear = nn.Linear(H, H)
            self.z_concat_layernorm = BertLayerNorm(H, eps=eps)
            self.z_direct_ln = BertLayerNorm(H, eps=eps)
            self.z_landm_ln = BertLayerNorm(H, eps=eps)
            if config.do_back_txt_type == "type_2":
                self.z_direc_cross_attn = RobertaAttention(config)
                self.z_landm_cross_attn = RobertaAttention(config)
                if pretrain_layout:
                    self.txt_self_attn = RobertaAttention(config)
                self.instr_aug_linear = nn.Linear(H, 1)
                self.instr_ori_linear = nn.Linear(H, 1)
                self.instr_sigmoid = nn.Sigmoid()
                self.concat_linear = nn.Linear(H * 3, H)
        if config.do_front_txt:
            self.z_front_cross_attn = RobertaAttention(config)
            self.z_front_linear = nn.Linear(H, H)
            self.z_front_ln = BertLayerNorm(H, eps=eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def _xattn(self, attn, lin, ln, txt, z):
        """LN(W . CrossAttn(txt -> z)) with no key mask (M/models/vilmodel_GOAT.py:121-129)"""
        a = attn.run(txt, None, z.to(torch.float32), None)
        return layer_norm(ln, linear(lin, a.tensor()))

    def forward(self, txt_embeds, txt_masks, z_direc_embeds=None, z_direc_pzs=None, z_landm_embeds=None,
                z_landm_pzs=None, front_txt_embeds=None):
        cfg = self.config
        x = Act.of(txt_embeds)
        m = extend_neg_masks(txt_masks)[:, 0, 0].contiguous()
        for layer in self.layer:
            x = layer.run(x, m)
        txt = x.tensor()
        if not self.update_lang_bert:
            txt = txt.detach()
        if self.pretrain_layout:
            active = z_direc_embeds is not None
            use_back = active
            use_front = active and front_txt_embeds is not None
        else:
            active = cfg.do_back_txt or cfg.do_front_txt
            use_back = cfg.do_back_txt
            use_front = cfg.do_front_txt and front_txt_embeds is not None
        if not active:
            return txt
        if cfg.do_back_txt_type == "type_1":
            if use_back:
                if self.pretrain_layout and getattr(cfg, "z_cross_attn", False):
                    raise NotImplementedError("z_cross_attn is off in every shipped config")
                sd = Fn.WSumFn.apply(z_direc_embeds.to(torch.float32), z_direc_pzs).unsqueeze(1)
                sl = Fn.WSumFn.apply(z_landm_embeds.to(torch.float32), z_landm_pzs).unsqueeze(1)
                txt = linear(self.z_txt_linear, txt) + linear(self.z_direct_linear, sd) + linear(self.z_landm_linear, sl)
            if use_front:
                zf = self._xattn(self.z_front_cross_attn, self.z_front_linear, self.z_front_ln, txt, front_txt_embeds)
                txt = txt + zf
            return layer_norm(self.z_concat_layernorm, txt)
        if cfg.do_back_txt_type != "type_2":
            raise ValueError("unknown do_back_txt_type %r" % (cfg.do_back_txt_type,))
        zd = zl = zf = None
        if use_back:
            zd = self._xattn(self.z_direc_cross_attn, self.z_direct_linear, self.z_direct_ln, txt, z_direc_embeds)
            if z_landm_embeds is not None:
                zl = self._xattn(self.z_landm_cross_attn, self.z_landm_linear, self.z_landm_ln, txt, z_landm_embeds)
        if use_front:
            zf = self._xattn(self.z_front_cross_attn, self.z_front_linear, self.z_front_ln, txt, front_txt_embeds)
        if cfg.do_add_method == "door":
            if use_back:
                aug = zd
                if zl is not None:
                    aug = aug + zl
                if zf is not None:
                    aug = aug + zf
            elif zf is not None:
                aug = zf
            else:
                raise ValueError("door: no intervention features were given")
            txt = door_gate(aug, txt, self.instr_aug_linear, self.instr_ori_linear)
        elif cfg.do_add_method == "add":
            if use_back:
                txt = txt + zd + zl
            if zf is not None:
                txt = txt + zf
        elif cfg.do_add_method == "concat":
            txt = linear(self.concat_linear, torch.cat((txt, zd, zl), -1))
        else:
            raise ValueError("unknown do_add_method %r" % (cfg.do_add_method,))
        return layer_norm(self.z_concat_layernorm, txt)


# --------------------------------------------------------------------------------------
# panorama: view features -> embeddings (+ BACL image) -> pano encoder -> adaptive fusion
# --------------------------------------------------------------------------------------
class CausalImageEmbeddings(nn.Module):
    def __init__(self, config, pretrain_layout=False):
        super().__init__()
        _no_objects(config)
        self.config = config
        self.pretrain_layout = pretrain_layout
        H = config.hidden_size
        self.img_linear = nn.Linear(config.image_feat_size, H)
        self.img_layer_norm = BertLayerNorm(H, eps=1e-12)
        self.loc_linear = nn.Linear(config.angle_feat_size + 3, H)
        self.loc_layer_norm = BertLayerNorm(H, eps=1e-12)
        if pretrain_layout:
            self.img_self_attn = BertAttention(config)     # constructed, never called (P/model/vilmodel_goat.py:248)
        self.img_self_encoder = create_transformer_encoder(config, config.num_pano_layers, norm=True)
        self.do_back_img = config.do_back_img
        if self.do_back_img:
            self.do_img_before_linear = nn.Linear(config.image_feat_size, H)
            self.do_img_layer_norm = BertLayerNorm(H, eps=1e-12)
            self.do_img_attn = BertAttention(config)
            self.do_img_after_linear = nn.Linear(H, H)
            self.img_after_linear = nn.Linear(H, H)
            self.do_img_concat_layernorm = BertLayerNorm(H, eps=1e-12)
            if getattr(config, "do_back_img_type", "type_1") == "type_2" or pretrain_layout:
                if config.do_add_method == "door":
                    self.sigmoid = nn.Sigmoid()
                elif config.do_add_method == "concat":
                    self.do_concat_img_linear = nn.Linear(H * 2, H)
        self.nav_type_embedding = nn.Embedding(2, H)
        if config.adaptive_pano_fusion:
            self.adaptive_pano_attn = nn.Linear(H, 1)
            self.adaptive_softmax = nn.Softmax(dim=1)
        self.layer_norm = BertLayerNorm(H, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def back_door(self, view_img_embeds, z_img_features, z_img_pzs):
        """BACL image: type_1 LN(W_a x + W_b sum_z p(z) LN(W_z z)); type_2 cross-attention onto the dictionary."""
        cfg = self.config
        z = layer_norm(self.do_img_layer_norm, linear(self.do_img_before_linear, z_img_features.to(torch.float32)))
        kind = getattr(cfg, "do_back_img_type", "type_1")
        if kind == "type_1" or self.pretrain_layout:
            s = Fn.WSumFn.apply(z, z_img_pzs).unsqueeze(1)
            x = linear(self.img_after_linear, view_img_embeds) + linear(self.do_img_after_linear, s)
        elif kind == "type_2":
            za = self.do_img_attn.run(view_img_embeds, None, z, None).tensor()
            if cfg.do_add_method == "door":
                g = torch.sigmoid(linear(self.img_after_linear, view_img_embeds) + linear(self.do_img_after_linear, za))
                x = g * view_img_embeds + (1 - g) * za
            elif cfg.do_add_method == "add":
                x = view_img_embeds + za
            elif cfg.do_add_method == "concat":
                x = linear(self.do_concat_img_linear, torch.cat((view_img_embeds, za), -1))
            else:
                raise ValueError("unknown do_add_method %r" % (cfg.do_add_method,))
        else:
            raise ValueError("unknown do_back_img_type %r" % (kind,))
        return layer_norm(self.do_img_concat_layernorm, x)

    def pano_fuse(self, view_img_embeds):
        """sum_v softmax_v(tanh(w . x_v + b)) x_v over ALL view slots (no mask)."""
        return Fn.AttnPoolFn.apply(view_img_embeds, self.adaptive_pano_attn.weight, self.adaptive_pano_attn.bias, 0)

    def encode(self, view_img_fts, loc_fts, view_lens, z_img_features=None, z_img_pzs=None, loc_after_do=False):
        """-> (view_img_embeds [S,V,H], img_masks bool [S,V], fused [S,H] or None)
        loc_after_do: the per-step fine-tune path adds the location embedding after the intervention
        (M/models/vilmodel_GOAT.py:686-688), the trajectory path before it (:232-233)."""
        x = layer_norm(self.img_layer_norm, linear(self.img_linear, view_img_fts))
        loc = layer_norm(self.loc_layer_norm, linear(self.loc_linear, loc_fts))
        if not loc_after_do:
            x = x + loc
        if z_img_features is not None:
            x = self.back_door(x, z_img_features, z_img_pzs)
        if loc_after_do:
            x = x + loc
        img_masks = gen_seq_masks(view_lens, view_img_fts.shape[1])
        x = Fn.dropout(x, self.dropout.p, self.training)
        x = self.img_self_encoder.run(x, img_masks.logical_not()).tensor()
        fused = self.pano_fuse(x) if self.config.adaptive_pano_fusion else None
        return x, img_masks, fused

    def forward(self, traj_view_img_fts, traj_loc_fts, traj_nav_types, traj_step_lens, traj_vp_view_lens,
                type_embed_layer, traj_reverie_obj_fts=None, traj_reverie_obj_lens=None, *extra, z_img_features=None,
                z_img_pzs=None):
        if traj_reverie_obj_fts is not None:
            raise NotImplementedError("object features (REVERIE / SOON) are outside the hot-path scope")
        if self.pretrain_layout and len(extra) >= 3:
            # pretrain passes (traj_reverie_loc_fts, z_img_features, z_img_pzs, obj_names) positionally
            z_img_features, z_img_pzs = extra[1], extra[2]
        elif not self.pretrain_layout and len(extra) >= 2:
            z_img_features, z_img_pzs = extra[0], extra[1]
        x, _, fused = self.encode(traj_view_img_fts, traj_loc_fts, traj_vp_view_lens, z_img_features, z_img_pzs)
        split_traj_embeds = torch.split(x, traj_step_lens, 0)
        split_traj_vp_lens = torch.split(traj_vp_view_lens, traj_step_lens, 0)
        split_fused = torch.split(fused, traj_step_lens, 0) if fused is not None else None
        return split_traj_embeds, split_traj_vp_lens, split_fused


# --------------------------------------------------------------------------------------
# local / global branches
# --------------------------------------------------------------------------------------
def _pos_embed(seq, x):
    """nn.Sequential(Linear, LayerNorm) position embedding"""
    return layer_norm(seq[1], linear(seq[0], x))


def _last_step_views(split_traj_embeds, split_traj_vp_lens):
    """[B, V, H] embeddings and lengths of each sample's current (last) panorama."""
    cur = torch.stack([x[-1] for x in split_traj_embeds], 0)
    lens = torch.stack([x[-1] for x in split_traj_vp_lens], 0)
    return cur, lens


class LocalVPEncoder(nn.Module):
    def __init__(self, config, with_cfp=None):
        super().__init__()
        self.vp_pos_embeddings = nn.Sequential(nn.Linear(config.angle_feat_size * 2 + 6, config.hidden_size),
                                               BertLayerNorm(config.hidden_size, eps=1e-12))
        self.encoder = CrossmodalEncoder(config)
        if with_cfp is None:
            with_cfp = "cfp" in getattr(config, "pretrain_tasks", ()) or getattr(config, "mode", None) == "extract_cfp_features"
        if with_cfp:
            self.tim_self_encoder = BertAttention(config)

    def vp_input_embedding(self, split_traj_embeds, split_traj_vp_lens, vp_pos_fts):
        cur, lens = _last_step_views(split_traj_embeds, split_traj_vp_lens)
        vp_lens = lens + 1
        vp_masks = gen_seq_masks(vp_lens)
        max_vp_len = vp_masks.shape[1]
        B, _, H = cur.shape
        vp_img = torch.cat([cur.new_zeros(B, 1, H), cur], 1)[:, :max_vp_len]    # [stop] token first
        return vp_img + _pos_embed(self.vp_pos_embeddings, vp_pos_fts), vp_masks

    def vp_input_embedding_flat(self, views, view_lens, last_rows, vp_pos_fts):
        """views [S,V,H], view_lens [S], last_rows int32 [B,1] (row of each sample's current panorama) -> the same
        (vp_embeds [B,Nq,H], vp_masks [B,Nq]) with Nq = vp_pos_fts.shape[1], through one gather kernel."""
        S, V, H = views.shape
        B, Nq = vp_pos_fts.shape[0], vp_pos_fts.shape[1]
        cur = Fn.SegmentReduceFn.apply(views.reshape(S, V * H), last_rows, False).view(B, V, H)
        vp_lens = view_lens[last_rows.view(-1).long()] + 1
        vp_img = torch.cat([cur.new_zeros(B, 1, H), cur], 1)[:, :Nq]            # [stop] token first
        if vp_img.shape[1] < Nq:
            vp_img = torch.cat([vp_img, cur.new_zeros(B, Nq - vp_img.shape[1], H)], 1)
        return vp_img + _pos_embed(self.vp_pos_embeddings, vp_pos_fts), gen_seq_masks(vp_lens, Nq)

    def forward(self, txt_embeds, txt_masks, split_traj_embeds, split_traj_vp_lens, vp_pos_fts):
        vp_embeds, vp_masks = self.vp_input_embedding(split_traj_embeds, split_traj_vp_lens, vp_pos_fts)
        return self.encoder(vp_embeds, vp_masks, txt_embeds, txt_masks)

    def forward_cfp(self, split_traj_embeds, split_traj_vp_lens, vp_pos_fts):
        vp_embeds, vp_masks = self.vp_input_embedding(split_traj_embeds, split_traj_vp_lens, vp_pos_fts)
        return self.tim_self_encoder(vp_embeds, extend_neg_masks(vp_masks))[0]


def build_gmap_index(traj_step_lens, view_lens, traj_vpids, traj_cand_vpids, gmap_vpids, use_fused, num_views,
                     start_id=1):
    """Host-side replacement of the string-keyed dict walk in ``_aggregate_gmap_features``
    (P/model/vilmodel_goat.py:430-468).  Sources are rows of cat([fused [S,H], views [S*V,H]]):
    row s = fused panorama of global step s, row S + s*V + j = view j of step s.
    -> int32 [B, Gmax-start_id, K]: per global-map node the rows whose MEAN is its image feature (-1 = empty)."""
    S = sum(traj_step_lens)
    rows = []
    off = 0
    for i, n_steps in enumerate(traj_step_lens):
        visited, unvisited = {}, {}
        for t in range(n_steps):
            s = off + t
            if use_fused:
                visited[traj_vpids[i][t]] = [s]
            else:
                visited[traj_vpids[i][t]] = [S + s * num_views + j for j in range(int(view_lens[s]))]
            for j, vp in enumerate(traj_cand_vpids[i][t]):
                if vp not in visited:
                    unvisited.setdefault(vp, []).append(S + s * num_views + j)
        off += n_steps
        rows.append([visited[vp] if vp in visited else unvisited[vp] for vp in gmap_vpids[i][start_id:]])
    return _pack_index(rows)


def _pack_index(rows, G=None, K=None):
    """list (batch) of lists (nodes) of index lists -> int32 [B, G, K] padded with -1, filled through ONE numpy array
    (no per-node tensor construction)."""
    import numpy as np
    Gn = max(len(r) for r in rows)
    Kn = max(1, max((len(e) for r in rows for e in r), default=1))
    G = Gn if G is None else G
    K = Kn if K is None else K
    if Gn > G or Kn > K:
        raise ValueError("index lists need [%d, %d] slots but only [%d, %d] were given" % (Gn, Kn, G, K))
    arr = np.full((len(rows), G, K), -1, dtype=np.int32)
    for i, r in enumerate(rows):
        for g, e in enumerate(r):
            if e:
                arr[i, g, :len(e)] = e
    return torch.from_numpy(arr)


def split_gmap_index(idx, S):
    """Index lists over cat([fused [S], views [S*V]]) -> (idx_fused over the S fused rows, idx_views over the S*V view
    rows), each int32 with -1 for empty.  A node has entries in only one of the two."""
    idx_f = torch.where(idx < S, idx, torch.full_like(idx, -1))[..., :1].contiguous()
    idx_v = torch.where(idx >= S, idx - S, torch.full_like(idx, -1))
    return idx_f, idx_v


class GlobalMapEncoder(nn.Module):
    def __init__(self, config, with_cfp=None):
        super().__init__()
        self.config = config
        self.gmap_pos_embeddings = nn.Sequential(nn.Linear(config.angle_feat_size + 3, config.hidden_size),
                                                 BertLayerNorm(config.hidden_size, eps=1e-12))
        self.gmap_step_embeddings = nn.Embedding(config.max_action_steps, config.hidden_size)
        self.encoder = CrossmodalEncoder(config)
        if with_cfp is None:
            with_cfp = "cfp" in getattr(config, "pretrain_tasks", ()) or getattr(config, "mode", None) == "extract_cfp_features"
        if with_cfp:
            self.tim_self_encoder = BertAttention(config)
        self.sprel_linear = nn.Linear(1, 1) if config.graph_sprels else None

    def aggregate_flat(self, views, fused, idx_f, idx_v):
        """views [S,V,H], fused [S,H] or None, idx_f int32 [B,G1,1] (rows of fused), idx_v int32 [B,G1,K] (rows of
        views.view(S*V,H)) -> [B, 1+G1, H]: per global-map node the fused panorama of its visit, or the MEAN of the
        candidate views that observed it; [stop] (zeros) first.  The index lists only ever name valid views, so the
        reference's multiplication by the view mask (P/model/vilmodel_goat.py:441) is the identity on what is read."""
        S, V, H = views.shape
        B, G1, K = idx_v.shape
        out = Fn.SegmentReduceFn.apply(views.reshape(S * V, H), idx_v.view(B * G1, K), True)
        if fused is not None:
            out = out + Fn.SegmentReduceFn.apply(fused, idx_f.view(B * G1, 1), False)
        return torch.cat([out.new_zeros(B, 1, H), out.view(B, G1, H)], 1)

    def _aggregate_gmap_features(self, split_traj_embeds, split_traj_vp_lens, traj_vpids, traj_cand_vpids, gmap_vpids,
                                 split_traj_fused_embeds=None):
        step_lens = [int(x.shape[0]) for x in split_traj_embeds]
        views = torch.cat(list(split_traj_embeds), 0)                       # [S,V,H]
        S, V, H = views.shape
        lens_h = torch.cat(list(split_traj_vp_lens), 0).tolist()
        use_fused = split_traj_fused_embeds is not None
        idx = build_gmap_index(step_lens, lens_h, traj_vpids, traj_cand_vpids, gmap_vpids, use_fused, V)
        if use_fused:
            idx_f, idx_v = split_gmap_index(idx, S)
            fused = torch.cat(list(split_traj_fused_embeds), 0)
        else:
            idx_f, idx_v, fused = None, idx - S, None       # every entry is a view row (>= S); -1 - S stays negative
        return self.aggregate_flat(views, fused, None if idx_f is None else idx_f.to(views.device), idx_v.to(views.device))

    def embed_nodes(self, gmap_img_fts, gmap_step_ids, gmap_pos_fts, gmap_lens):
        gmap_embeds = gmap_img_fts + self.step_embed(gmap_step_ids) + _pos_embed(self.gmap_pos_embeddings, gmap_pos_fts)
        return gmap_embeds, gen_seq_masks(gmap_lens, gmap_step_ids.shape[1])

    def gmap_input_embedding(self, split_traj_embeds, split_traj_vp_lens, traj_vpids, traj_cand_vpids, gmap_vpids,
                             gmap_step_ids, gmap_pos_fts, gmap_lens, split_traj_fused_embeds=None):
        gmap_img_fts = self._aggregate_gmap_features(split_traj_embeds, split_traj_vp_lens, traj_vpids, traj_cand_vpids,
                                                     gmap_vpids, split_traj_fused_embeds)
        return self.embed_nodes(gmap_img_fts, gmap_step_ids, gmap_pos_fts, gmap_lens)

    def step_embed(self, gmap_step_ids):
        return Fn.GatherRowsFn.apply(gmap_step_ids, self.gmap_step_embeddings.weight)

    def sprels(self, pair_dists):
        """sprel_linear (1 -> 1) on the pairwise distances -> additive self-attention bias [B,1,G,G]"""
        if self.sprel_linear is None:
            return None
        return Fn.SprelFn.apply(pair_dists, self.sprel_linear.weight, self.sprel_linear.bias).unsqueeze(1)

    def forward(self, txt_embeds, txt_masks, split_traj_embeds, split_traj_vp_lens, traj_vpids, traj_cand_vpids,
                gmap_vpids, gmap_step_ids, gmap_pos_fts, gmap_lens, graph_sprels=None, split_traj_fused_embeds=None):
        gmap_embeds, gmap_masks = self.gmap_input_embedding(
            split_traj_embeds, split_traj_vp_lens, traj_vpids, traj_cand_vpids, gmap_vpids, gmap_step_ids, gmap_pos_fts,
            gmap_lens, split_traj_fused_embeds)
        return self.encoder(gmap_embeds, gmap_masks, txt_embeds, txt_masks, graph_sprels=self.sprels(graph_sprels))

    def forward_cfp(self, split_traj_embeds, split_traj_vp_lens, traj_vpids, traj_cand_vpids, gmap_vpids, gmap_step_ids,
                    gmap_pos_fts, gmap_lens, graph_sprels=None, split_traj_fused_embeds=None):
        gmap_embeds, gmap_masks = self.gmap_input_embedding(
            split_traj_embeds, split_traj_vp_lens, traj_vpids, traj_cand_vpids, gmap_vpids, gmap_step_ids, gmap_pos_fts,
            gmap_lens, split_traj_fused_embeds)
        return self.tim_self_encoder(gmap_embeds, extend_neg_masks(gmap_masks))[0]


def build_fusion_index(gmap_vpids, gmap_visited_masks, cand_vpids, n_local, first_cand, first_node):
    """Host-side replacement of the logit-fusion double loop (P/model/pretrain_goat.py:328-345,
    M/models/vilmodel_GOAT.py:797-813).  -> int32 [B, G, K]: for every global-map node the flat indices
    (i * n_local + j) of the local logits that are ADDED to its global logit.
      node 0 ([stop])          <- local [stop] logit
      unvisited node with a candidate view of the same viewpoint <- that candidate's logit (last one wins)
      other unvisited nodes    <- the sum of the logits of candidates that lead back to visited nodes
    first_cand: local position of candidate 0 (1 in pretrain: [stop]; 2 in fine-tune: [stop],[MEM]);
    first_node: first global-map position that can receive local logits (1 / 2)."""
    B = len(gmap_vpids)
    G = max(len(v) for v in gmap_vpids)
    vis = gmap_visited_masks.tolist() if torch.is_tensor(gmap_visited_masks) else gmap_visited_masks
    rows = []
    for i in range(B):
        visited = set(vp for vp, m in zip(gmap_vpids[i], vis[i]) if m)
        tmp, bw = {}, []
        for j, vp in enumerate(cand_vpids[i]):
            pos = j + first_cand if first_cand == 1 else j
            if first_cand != 1 and j <= 1:
                continue                      # fine-tune lists carry [stop] and [MEM] placeholders at 0 and 1
            if vp in visited:
                bw.append(i * n_local + pos)
            else:
                tmp[vp] = i * n_local + pos
        r = [[i * n_local]]
        for j, vp in enumerate(gmap_vpids[i]):
            if j == 0:
                continue
            if j >= first_node and vp not in visited:
                r.append([tmp[vp]] if vp in tmp else list(bw))
            else:
                r.append([])
        rows.append(r)
    return _pack_index(rows, G=G)


def fuse_logits(global_logits, local_logits, idx):
    """fused = global + gather-sum(local, idx)"""
    B, G = global_logits.shape
    add = Fn.SegmentReduceFn.apply(local_logits.reshape(-1, 1), idx.view(B * G, -1).contiguous(), False)
    return global_logits + add.view(B, G)


# --------------------------------------------------------------------------------------
# FACL
# --------------------------------------------------------------------------------------
class FrontDoorEncoder(nn.Module):
    """LN(SelfAttn(x, mask) + CrossAttn(x -> prototypes)), then the door gate against x."""

    def __init__(self, config):
        super().__init__()
        self.ll_self_attn = BertAttention(config)
        self.lg_cross_attn = BertAttention(config)
        self.ln = BertLayerNorm(config.hidden_size, eps=1e-12)
        self.config = config
        self.aug_linear = nn.Linear(config.hidden_size, 1)
        self.ori_linear = nn.Linear(config.hidden_size, 1)
        self.sigmoid = nn.Sigmoid()

    def forward(self, local_feats, global_feats, local_feats_masks=None):
        if local_feats_masks is not None and local_feats_masks.dim() != 4:
            local_feats_masks = extend_neg_masks(local_feats_masks)
        x = Act.of(local_feats)
        ll = self.ll_self_attn.run(x, local_feats_masks).tensor()
        lg = self.lg_cross_attn.run(x, None, global_feats.to(torch.float32), None).tensor()
        out = layer_norm(self.ln, ll + lg)
        return door_gate(out, x.tensor(), self.aug_linear, self.ori_linear)
