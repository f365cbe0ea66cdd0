"""CPU oracle: a plain restatement of the GOAT cross-modal hot path.

TEST INFRASTRUCTURE ONLY -- not product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module, and only
as the checker / CPU baseline.  Nothing under ``vln_goat_b200/`` imports it; the product path
raises if the CUDA library is missing.

Every function is written from the reference's *behaviour* as explicit tensor algebra
(matmul / exp / sum ...), takes parameters as a flat ``dict`` that uses the reference's own
``state_dict`` key names (so the same dict loads into the reference module, this oracle and the
CUDA modules), and cites the reference lines it follows.  It runs on CPU in fp32 or fp64
(``dtype=`` of the tensors you pass); backward comes from torch autograd over these explicit
formulas (the task allows a torch reference for floating-point kernels).

Pinning: ``tests/test_oracle_vs_reference.py`` checks each function against the unmodified
reference modules imported through ``oracle/ref_shim.py`` (when /root/reference is mounted)
and ``tests/test_oracle_golden.py`` checks it against the committed fixtures in
``tests/golden/`` that ``tests/golden/make_golden.py`` produced from those same reference
modules.  The reference itself ships no tests or golden vectors.

P/ = /root/reference/pretrain_src/, M/ = /root/reference/map_nav_src/.
"""
import math

import torch

NEG_MASK = -10000.0  # P/model/ops.py:25-34 (additive fp32 mask, not -inf)


# --------------------------------------------------------------------------------------
# elementary pieces
# --------------------------------------------------------------------------------------
def gelu_erf(x):
    """x * 0.5 * (1 + erf(x / sqrt 2))  -- P/model/Bert_backbone.py:41-47."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layernorm(x, w, b, eps):
    """nn.LayerNorm over the last dim, biased variance -- e.g. P/model/Bert_backbone.py:303,309."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def linear(x, w, b=None):
    y = x @ w.transpose(-1, -2)
    return y if b is None else y + b


def extend_neg_masks(masks, dtype=torch.float32):
    """bool [B,L] -> additive [B,1,1,L] with 0 / -10000  -- P/model/ops.py:25-34."""
    return (1.0 - masks[:, None, None, :].to(dtype)) * NEG_MASK


def gen_seq_masks(lens, max_len=None):
    """P/model/ops.py:36-44."""
    if max_len is None:
        max_len = int(max(lens))
    return torch.arange(max_len, device=lens.device)[None, :] < lens[:, None]


def attn_core(q, k, v, mask=None, num_heads=12):
    """softmax(Q K^T / sqrt(d) + mask) V per head, heads merged back.

    q [B,Nq,H], k/v [B,Nk,H]; mask broadcastable to [B,1,Nq,Nk] (additive).
    P/model/Bert_backbone.py:194-197 (head split), :247 (QK^T), :265 (/sqrt d), :266-268 (+mask),
    :271 (softmax), :286 (PV), :288-290 (merge).  Dropout (:280) is the identity at p=0 / eval.
    """
    B, Nq, H = q.shape
    Nk = k.shape[1]
    d = H // num_heads
    qh = q.view(B, Nq, num_heads, d).permute(0, 2, 1, 3)
    kh = k.view(B, Nk, num_heads, d).permute(0, 2, 1, 3)
    vh = v.view(B, Nk, num_heads, d).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2)
    s = s / math.sqrt(d)
    if mask is not None:
        s = s + mask
    m = s.max(-1, keepdim=True).values
    e = torch.exp(s - m)
    p = e / e.sum(-1, keepdim=True)
    ctx = p @ vh
    return ctx.permute(0, 2, 1, 3).reshape(B, Nq, H)


def bert_attention(p, pre, x, mask=None, enc=None, enc_mask=None, eps=1e-12, num_heads=12):
    """BertAttention / RobertaAttention: MHA + out-proj + residual + LayerNorm.

    P/model/Bert_backbone.py:199-296 (BertSelfAttention), :299-310 (BertSelfOutput), :313-342.
    When ``enc`` is given K,V come from ``enc`` and the mask is ``enc_mask`` (:221-224) -- the
    first mask argument is ignored for cross-attention.
    """
    g = lambda n: p[pre + n]
    q = linear(x, g("self.query.weight"), g("self.query.bias"))
    src = x if enc is None else enc
    m = mask if enc is None else enc_mask
    k = linear(src, g("self.key.weight"), g("self.key.bias"))
    v = linear(src, g("self.value.weight"), g("self.value.bias"))
    ctx = attn_core(q, k, v, m, num_heads)
    h = linear(ctx, g("output.dense.weight"), g("output.dense.bias"))
    return layernorm(h + x, g("output.LayerNorm.weight"), g("output.LayerNorm.bias"), eps)


def ffn(p, pre_inter, pre_out, x, eps=1e-12):
    """BertIntermediate + BertOutput: LN(W2 gelu(W1 x + b1) + b2 + x).
    P/model/Bert_backbone.py:345-370 (Roberta copies :546-572)."""
    h = gelu_erf(linear(x, p[pre_inter + "dense.weight"], p[pre_inter + "dense.bias"]))
    y = linear(h, p[pre_out + "dense.weight"], p[pre_out + "dense.bias"])
    return layernorm(y + x, p[pre_out + "LayerNorm.weight"], p[pre_out + "LayerNorm.bias"], eps)


def roberta_layer(p, pre, x, mask, eps=1e-12, num_heads=12):
    """RobertaLayer (encoder only): self-attn block + FFN block.  P/model/Bert_backbone.py:574-659."""
    a = bert_attention(p, pre + "attention.", x, mask, eps=eps, num_heads=num_heads)
    return ffn(p, pre + "intermediate.", pre + "output.", a, eps)


def lang_encoder(p, pre, x, txt_masks, num_layers=6, eps=1e-12, num_heads=12):
    """LanguageEncoder: extend mask, 6x RobertaLayer.  P/model/vilmodel_goat.py:24-44."""
    m = extend_neg_masks(txt_masks, x.dtype)
    for i in range(num_layers):
        x = roberta_layer(p, "%slayer.%d." % (pre, i), x, m, eps, num_heads)
    return x


def cross_layer(p, pre, x, enc, mask, enc_mask, graph_sprels=None, eps=1e-12, num_heads=12):
    """BertCrossLayer: self-attn (mask + sprel bias) -> cross-attn -> FFN.
    P/model/Bert_backbone.py:678-727; sprels only touch self-attention (:690-698 vs :221-224)."""
    if graph_sprels is not None:
        mask = mask + graph_sprels
    a = bert_attention(p, pre + "attention.", x, mask, eps=eps, num_heads=num_heads)
    c = bert_attention(p, pre + "crossattention.", a, mask, enc, enc_mask, eps=eps, num_heads=num_heads)
    return ffn(p, pre + "intermediate.", pre + "output.", c, eps)


def crossmodal_encoder(p, pre, q, q_masks, kv, kv_masks, graph_sprels=None, num_layers=3,
                       eps=1e-12, num_heads=12):
    """CrossmodalEncoder: 2-D bool masks get extended, then num_top_layer cross layers.
    P/model/Bert_backbone.py:756-781."""
    if q_masks is not None and q_masks.dim() != 4:
        q_masks = extend_neg_masks(q_masks, q.dtype)
    if kv_masks.dim() != 4:
        kv_masks = extend_neg_masks(kv_masks, q.dtype)
    for i in range(num_layers):
        q = cross_layer(p, "%scrossattention.%d." % (pre, i), q, kv, q_masks, kv_masks,
                        graph_sprels, eps, num_heads)
    return q


def lang2visn_layer(p, pre, lang, lang_mask, visn, visn_mask, eps=1e-12, num_heads=12):
    """BertCrossLayer.forward_lang2visn (pretrain-only extra params).  P/model/Bert_backbone.py:729-754."""
    a = bert_attention(p, pre + "crossattention.", lang, lang_mask, visn, visn_mask, eps, num_heads)
    s = bert_attention(p, pre + "lang_self_attn.", a, lang_mask, eps=eps, num_heads=num_heads)
    return ffn(p, pre + "lang_inter.", pre + "lang_output.", s, eps)


# --------------------------------------------------------------------------------------
# panorama encoder (DETR-style pre-LN layers around nn.MultiheadAttention)
# --------------------------------------------------------------------------------------
def pano_layer(p, pre, x, key_padding_mask=None, num_heads=12, eps=1e-5):
    """TransformerEncoderLayer.forward_pre (batch-first restatement).
    P/model/transformer.py:170-182; nn.MultiheadAttention packs q,k,v in in_proj_weight [3H,H];
    bool key_padding_mask (True = pad) becomes -inf on those keys; activation is F.gelu (erf);
    norm1/norm2 are nn.LayerNorm(d_model) with the torch default eps 1e-5 (:144-145)."""
    H = x.shape[-1]
    x2 = layernorm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"], eps)
    w, b = p[pre + "self_attn.in_proj_weight"], p[pre + "self_attn.in_proj_bias"]
    q = linear(x2, w[:H], b[:H])
    k = linear(x2, w[H:2 * H], b[H:2 * H])
    v = linear(x2, w[2 * H:], b[2 * H:])
    mask = None
    if key_padding_mask is not None:
        mask = torch.zeros(key_padding_mask.shape, dtype=x.dtype, device=x.device)
        mask = mask.masked_fill(key_padding_mask, float("-inf"))[:, None, None, :]
    ctx = attn_core(q, k, v, mask, num_heads)
    x = x + linear(ctx, p[pre + "self_attn.out_proj.weight"], p[pre + "self_attn.out_proj.bias"])
    x2 = layernorm(x, p[pre + "norm2.weight"], p[pre + "norm2.bias"], eps)
    h = gelu_erf(linear(x2, p[pre + "linear1.weight"], p[pre + "linear1.bias"]))
    return x + linear(h, p[pre + "linear2.weight"], p[pre + "linear2.bias"])


def pano_encoder(p, pre, x, key_padding_mask=None, num_layers=2, num_heads=12):
    """TransformerEncoder with final BertLayerNorm(eps=1e-12).
    P/model/transformer.py:62-89, P/model/ops.py:11-23."""
    for i in range(num_layers):
        x = pano_layer(p, "%slayers.%d." % (pre, i), x, key_padding_mask, num_heads)
    return layernorm(x, p[pre + "norm.weight"], p[pre + "norm.bias"], 1e-12)


# --------------------------------------------------------------------------------------
# pooling / gating heads
# --------------------------------------------------------------------------------------
def pano_fuse(x, w, b):
    """Adaptive pano fusion: softmax over views of tanh(x w + b), weighted sum (no mask).
    P/model/vilmodel_goat.py:354-362 (M/models/vilmodel_GOAT.py:728-735)."""
    a = torch.tanh(linear(x, w, b))                  # [S,V,1]
    a = torch.exp(a - a.max(1, keepdim=True).values)
    a = a / a.sum(1, keepdim=True)
    return (x * a).sum(1)


def cfp_pool(x, w):
    """CFP attention pooling: tanh(sum_t softmax_t(tanh(x) w) x), no mask over padding.
    P/model/pretrain_goat.py:502-515 (M/models/vilmodel_GOAT.py:907-920).  w is [H,1]."""
    s = torch.tanh(x) @ w                             # [B,T,1]
    s = torch.exp(s - s.max(1, keepdim=True).values)
    a = s / s.sum(1, keepdim=True)
    return torch.tanh((x * a).sum(1))


def cross_entropy_rows(logits, target):
    """per-row CE(logits, target) with -100 ignored (-> 0 loss), as F.cross_entropy(reduction='none')."""
    m = logits.max(-1, keepdim=True).values
    lse = (m + torch.log(torch.exp(logits - m).sum(-1, keepdim=True))).squeeze(-1)
    valid = target != -100
    t = torch.where(valid, target, torch.zeros_like(target))
    picked = logits.gather(-1, t[:, None]).squeeze(-1)
    return torch.where(valid, lse - picked, torch.zeros_like(lse))


def infonce_sym(a, b, temperature=1.0):
    """Symmetric InfoNCE over in-batch negatives: (CE(sim) + CE(sim^T)) / 2 per row.
    P/model/pretrain_goat.py:519-534."""
    sim = (a @ b.transpose(0, 1)) / temperature
    tgt = torch.arange(a.shape[0], device=a.device)
    return (cross_entropy_rows(sim, tgt) + cross_entropy_rows(sim.transpose(0, 1), tgt)) / 2.0


def head_transform(p, pre, x, eps=1e-12):
    """BertPredictionHeadTransform: LN(gelu(W x + b)).  P/model/Bert_backbone.py:797-811."""
    h = gelu_erf(linear(x, p[pre + "dense.weight"], p[pre + "dense.bias"]))
    return layernorm(h, p[pre + "LayerNorm.weight"], p[pre + "LayerNorm.bias"], eps)


def cls_prediction(p, pre, x):
    """ClsPrediction: Linear -> ReLU -> LN(1e-12) -> Linear(->1).  P/model/pretrain_goat.py:27-38."""
    h = torch.relu(linear(x, p[pre + "net.0.weight"], p[pre + "net.0.bias"]))
    h = layernorm(h, p[pre + "net.2.weight"], p[pre + "net.2.bias"], 1e-12)
    return linear(h, p[pre + "net.3.weight"], p[pre + "net.3.bias"])


def door_gate(aug, ori, w_aug, b_aug, w_ori, b_ori):
    """g = sigmoid(aug w_a + b_a + ori w_o + b_o); g*aug + (1-g)*ori.
    M/models/vilmodel_GOAT.py:147-150 and :549-552."""
    g = torch.sigmoid(linear(aug, w_aug, b_aug) + linear(ori, w_ori, b_ori))
    return g * aug + (1.0 - g) * ori


def front_door_encoder(p, pre, local, glob, local_masks=None, eps=1e-5, num_heads=12):
    """FrontDoorEncoder (FACL): LN(SelfAttn(x,mask) + CrossAttn(x -> prototypes)), door gate vs x.
    M/models/vilmodel_GOAT.py:526-554.  The two BertAttention blocks use config.layer_norm_eps
    (1e-5 in fine-tune), the outer ``ln`` is eps 1e-12; the cross-attention has no key mask."""
    m = None
    if local_masks is not None:
        m = local_masks if local_masks.dim() == 4 else extend_neg_masks(local_masks, local.dtype)
    ll = bert_attention(p, pre + "ll_self_attn.", local, m, eps=eps, num_heads=num_heads)
    lg = bert_attention(p, pre + "lg_cross_attn.", local, None, glob, None, eps=eps, num_heads=num_heads)
    out = layernorm(ll + lg, p[pre + "ln.weight"], p[pre + "ln.bias"], 1e-12)
    return door_gate(out, local, p[pre + "aug_linear.weight"], p[pre + "aug_linear.bias"],
                     p[pre + "ori_linear.weight"], p[pre + "ori_linear.bias"])


def bacl_text_type2_door(p, pre, txt, z_direc=None, z_landm=None, front_txt=None, eps=1e-5,
                         num_heads=12):
    """LanguageEncoderDo causal tail, do_back_txt_type='type_2', do_add_method='door'.
    M/models/vilmodel_GOAT.py:121-160: cross-attn onto each dictionary (no key mask),
    Linear + LN each, summed, door-gated against txt, final z_concat_layernorm."""
    aug = None
    if z_direc is not None:
        d = bert_attention(p, pre + "z_direc_cross_attn.", txt, None, z_direc, None, eps, num_heads)
        aug = layernorm(linear(d, p[pre + "z_direct_linear.weight"], p[pre + "z_direct_linear.bias"]),
                        p[pre + "z_direct_ln.weight"], p[pre + "z_direct_ln.bias"], eps)
        if z_landm is not None:
            l = bert_attention(p, pre + "z_landm_cross_attn.", txt, None, z_landm, None, eps, num_heads)
            aug = aug + layernorm(
                linear(l, p[pre + "z_landm_linear.weight"], p[pre + "z_landm_linear.bias"]),
                p[pre + "z_landm_ln.weight"], p[pre + "z_landm_ln.bias"], eps)
    if front_txt is not None:
        f = bert_attention(p, pre + "z_front_cross_attn.", txt, None, front_txt, None, eps, num_heads)
        f = layernorm(linear(f, p[pre + "z_front_linear.weight"], p[pre + "z_front_linear.bias"]),
                      p[pre + "z_front_ln.weight"], p[pre + "z_front_ln.bias"], eps)
        aug = f if aug is None else aug + f
    out = door_gate(aug, txt, p[pre + "instr_aug_linear.weight"], p[pre + "instr_aug_linear.bias"],
                    p[pre + "instr_ori_linear.weight"], p[pre + "instr_ori_linear.bias"])
    return layernorm(out, p[pre + "z_concat_layernorm.weight"], p[pre + "z_concat_layernorm.bias"], eps)


def bacl_image_type1(p, pre, view_embeds, z_feats, z_pzs):
    """BACL image, type_1: LN(W_a x + W_b sum_z p(z) LN(W_z z)).
    M/models/vilmodel_GOAT.py:661-667 (all LN eps 1e-12)."""
    z = layernorm(linear(z_feats, p[pre + "do_img_before_linear.weight"], p[pre + "do_img_before_linear.bias"]),
                  p[pre + "do_img_layer_norm.weight"], p[pre + "do_img_layer_norm.bias"], 1e-12)
    s = (z * z_pzs.to(z.dtype)).sum(1, keepdim=True)
    y = linear(view_embeds, p[pre + "img_after_linear.weight"], p[pre + "img_after_linear.bias"]) + \
        linear(s, p[pre + "do_img_after_linear.weight"], p[pre + "do_img_after_linear.bias"])
    return layernorm(y, p[pre + "do_img_concat_layernorm.weight"], p[pre + "do_img_concat_layernorm.bias"], 1e-12)


def roberta_embeddings(p, pre, ids, eps=1e-12):
    """word + token-type(0) + position(arange from 0) -> LN.  P/model/Bert_backbone.py:85-121."""
    L = ids.shape[1]
    e = p[pre + "word_embeddings.weight"][ids] + p[pre + "token_type_embeddings.weight"][0] + \
        p[pre + "position_embeddings.weight"][:L][None]
    return layernorm(e, p[pre + "LayerNorm.weight"], p[pre + "LayerNorm.bias"], eps)


def sap_fuse_logits(global_logits, local_logits, gmap_vpids, gmap_visited_masks, cand_vpids, skip=1):
    """Scatter local action logits onto graph nodes.
    P/model/pretrain_goat.py:328-345 (skip=1: [stop]) / M/models/vilmodel_GOAT.py:794-813
    (skip=2: [stop],[MEM]).  cand_vpids[i][j] is the viewpoint id of local token j+skip
    (pretrain passes traj_cand_vpids[i][-1]; fine-tune passes vp_cand_vpids[i][skip:])."""
    fused = global_logits.clone()
    fused[:, 0] = fused[:, 0] + local_logits[:, 0]
    for i in range(global_logits.shape[0]):
        visited = set(vp for vp, m in zip(gmap_vpids[i], gmap_visited_masks[i]) if m)
        tmp, bw = {}, 0
        for j, c in enumerate(cand_vpids[i]):
            if c in visited:
                bw = bw + local_logits[i, j + skip]
            else:
                tmp[c] = local_logits[i, j + skip]
        for j, vp in enumerate(gmap_vpids[i]):
            if j >= skip and vp not in visited:
                fused[i, j] = fused[i, j] + (tmp[vp] if vp in tmp else bw)
    return fused


# --------------------------------------------------------------------------------------
# the BASELINE.json config-2 workload: 6 text layers + 3 cross layers (one branch)
# --------------------------------------------------------------------------------------
def c2_forward(p, txt_embeds, txt_masks, vp_embeds, vp_masks, eps=1e-12):
    """LanguageEncoder(6) then CrossmodalEncoder(3) with q = vp tokens, kv = text.
    P/model/vilmodel_goat.py:563-564 and :399 (the local branch of GlocalTextPathCMT.forward)."""
    t = lang_encoder(p, "lang_encoder.", txt_embeds, txt_masks, 6, eps)
    v = crossmodal_encoder(p, "local_encoder.encoder.", vp_embeds, vp_masks, t, txt_masks, None, 3, eps)
    return t, v


# --------------------------------------------------------------------------------------
# seeded parameter recipe shared by fixtures, tests, smoke() and bench.py
# --------------------------------------------------------------------------------------
def seeded_params(named_shapes, seed=0, std=0.02, dtype=torch.float32):
    """Deterministic parameters for a {state_dict key: shape} mapping (or an iterable of pairs).

    Each tensor gets its own generator seeded from (seed, crc32(key)), so the values do not
    depend on iteration order or on which other keys exist.  1-D ``*weight`` tensors are
    LayerNorm gains (1 + 0.1 N(0,1)); everything else is std * N(0,1) -- unlike the reference
    init (zero bias, unit gain) so parity tests exercise biases and LN affine terms.
    Deterministic for a given torch build (the GPU box runs the same image)."""
    import zlib
    items = named_shapes.items() if hasattr(named_shapes, "items") else named_shapes
    out = {}
    for name, shape in items:
        shape = tuple(shape)
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
        if len(shape) == 1 and name.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            t = std * torch.randn(shape, generator=g)
        out[name] = t.to(dtype)
    return out


def attn_shapes(pre, H=768):
    d = {}
    for n in ("query", "key", "value"):
        d[pre + "self.%s.weight" % n] = (H, H)
        d[pre + "self.%s.bias" % n] = (H,)
    d[pre + "output.dense.weight"] = (H, H)
    d[pre + "output.dense.bias"] = (H,)
    d[pre + "output.LayerNorm.weight"] = (H,)
    d[pre + "output.LayerNorm.bias"] = (H,)
    return d


def ffn_shapes(pi, po, H=768, F=3072):
    return {pi + "dense.weight": (F, H), pi + "dense.bias": (F,), po + "dense.weight": (H, F),
            po + "dense.bias": (H,), po + "LayerNorm.weight": (H,), po + "LayerNorm.bias": (H,)}


def cross_layer_shapes(pre="", H=768, F=3072):
    d = attn_shapes(pre + "attention.", H)
    d.update(attn_shapes(pre + "crossattention.", H))
    d.update(ffn_shapes(pre + "intermediate.", pre + "output.", H, F))
    return d


def roberta_layer_shapes(pre="", H=768, F=3072):
    d = attn_shapes(pre + "attention.", H)
    d.update(ffn_shapes(pre + "intermediate.", pre + "output.", H, F))
    return d


def pano_encoder_shapes(pre="", num_layers=2, H=768, F=3072):
    d = {}
    for i in range(num_layers):
        l = "%slayers.%d." % (pre, i)
        d.update({l + "self_attn.in_proj_weight": (3 * H, H), l + "self_attn.in_proj_bias": (3 * H,),
                  l + "self_attn.out_proj.weight": (H, H), l + "self_attn.out_proj.bias": (H,),
                  l + "linear1.weight": (F, H), l + "linear1.bias": (F,),
                  l + "linear2.weight": (H, F), l + "linear2.bias": (H,),
                  l + "norm1.weight": (H,), l + "norm1.bias": (H,),
                  l + "norm2.weight": (H,), l + "norm2.bias": (H,)})
    d[pre + "norm.weight"] = (H,)
    d[pre + "norm.bias"] = (H,)
    return d


def c2_shapes():
    d = {}
    for i in range(6):
        d.update(roberta_layer_shapes("lang_encoder.layer.%d." % i))
    for i in range(3):
        d.update(cross_layer_shapes("local_encoder.encoder.crossattention.%d." % i))
    return d


# --------------------------------------------------------------------------------------
# optimizer step (for the CPU baseline of a full training step and the fused-AdamW parity test)
# --------------------------------------------------------------------------------------
def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ semantics (P/train_r2r_goat.py:352): total L2 norm over all grads,
    scale by max_norm / (norm + 1e-6) when that is < 1.  Returns (norm, coefficient)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = max_norm / (total + 1e-6)
    coef = torch.clamp(coef, max=1.0)
    return total, coef


def adamw_step(p, g, m, v, step, lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
    """One AdamW update of a single tensor, in place.  P/optim/adamw.py:85-110:
    m,v EMAs; denom = sqrt(v) + eps; step_size = lr * sqrt(1-b2^t) / (1-b1^t); p -= step_size * m/denom;
    then decoupled decay p -= lr * wd * p."""
    b1, b2 = betas
    m.mul_(b1).add_(g, alpha=1.0 - b1)
    v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
    denom = v.sqrt().add_(eps)
    step_size = lr
    if correct_bias:
        step_size = step_size * math.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)
    return p


def c2_loss(txt_out, vp_out):
    """The synthetic training objective bench.py / smoke() put on the C2 workload: mean square of both
    output streams (any differentiable scalar exercises the same forward + backward kernels)."""
    return 0.5 * (txt_out.float() ** 2).mean() + 0.5 * (vp_out.float() ** 2).mean()
