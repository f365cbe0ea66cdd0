"""Drop-in transformer blocks of the GOAT cross-modal path, backed by libgoat_sm100.

Same class names, constructor (``config``), ``forward`` signatures, return conventions and
``state_dict`` keys as the reference blocks, so a reference checkpoint loads unchanged and these
classes can be swapped into the reference model files:

  BertAttention / RobertaAttention   P/model/Bert_backbone.py:313-342, :515-543
  BertIntermediate / BertOutput      P/model/Bert_backbone.py:345-370
  RobertaLayer                       P/model/Bert_backbone.py:574-659
  BertCrossLayer                     P/model/Bert_backbone.py:661-754
  CrossmodalEncoder                  P/model/Bert_backbone.py:756-781
  LanguageEncoder                    P/model/vilmodel_goat.py:24-44
  TransformerEncoder(Layer)          P/model/transformer.py:62-89, :127-189 (pre-LN pano encoder)
  create_transformer_encoder, extend_neg_masks, gen_seq_masks, pad_tensors_wgrad   P/model/ops.py

(P/ = pretrain_src/, M/ = map_nav_src/ of CrystalSixone/VLN-GOAT; the M/ copies are identical in math.)

``nn.Linear`` / ``nn.LayerNorm`` / ``nn.Embedding`` objects are used as *parameter containers only*
(that is what fixes the state_dict names and the reference initialisation); their own forward is
never called -- all math runs in the CUDA library through ``functional``.  There is no CPU path:
calling a block with CPU tensors raises.

Inside a stack the activations travel as an ``Act`` (fp32 residual stream + 16-bit operand copy
written by the LayerNorm kernel) so that no cast kernels sit between blocks.
"""
import torch
from torch import nn

from . import functional as Fn
from . import ops, runtime

BertLayerNorm = nn.LayerNorm


# --------------------------------------------------------------------------------------
# helpers from P/model/ops.py (host-side mask plumbing, same semantics)
# --------------------------------------------------------------------------------------
def extend_neg_masks(masks, dtype=None):
    """bool [N,L] -> additive fp32 [N,1,1,L] with 0 / -10000.  P/model/ops.py:25-34"""
    if dtype is None:
        dtype = torch.float
    return (1.0 - masks[:, None, None, :].to(dtype=dtype)) * -10000.0


def gen_seq_masks(seq_lens, max_len=None):
    """P/model/ops.py:36-44"""
    if max_len is None:
        max_len = int(max(seq_lens))
    return torch.arange(max_len, device=seq_lens.device)[None, :] < seq_lens[:, None]


def pad_tensors_wgrad(tensors, lens=None):
    """B x [T, ...] -> [B, max T, ...] zero padded, differentiable.  P/model/ops.py:46-68"""
    sizes = [t.size(0) for t in tensors]
    if lens is None or [int(x) for x in lens] == sizes:
        # one fused pad-and-stack (differentiable) instead of a cat per sample
        return torch.nn.utils.rnn.pad_sequence(list(tensors), batch_first=True)
    max_len = int(max(lens))
    parts = []
    for i, t in enumerate(tensors):
        n = int(lens[i])
        parts.append(torch.nn.functional.pad(t[:n], (0, 0) * (t.dim() - 1) + (0, max_len - n)))
    return torch.stack(parts, 0)


def init_weights(module):
    """P/model/Bert_backbone.py:840-849"""
    if isinstance(module, (nn.Linear, nn.Embedding)):
        module.weight.data.normal_(mean=0.0, std=0.02)
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)
    if isinstance(module, nn.Linear) and module.bias is not None:
        module.bias.data.zero_()


# --------------------------------------------------------------------------------------
# activation carrier
# --------------------------------------------------------------------------------------
class Act(object):
    """[B, N, H] activation as 2-D token-major tensors: fp32 ``x32`` [B*N, H] (+ optional 16-bit copy)."""
    __slots__ = ("x32", "x16", "B", "N")

    def __init__(self, x32, x16, B, N):
        self.x32, self.x16, self.B, self.N = x32, x16, B, N

    @staticmethod
    def of(t):
        if isinstance(t, Act):
            return t
        if not t.is_cuda:
            raise RuntimeError("vln_goat_b200 blocks need CUDA tensors: there is no CPU fallback on this path")
        if t.dim() != 3:
            raise ValueError("expected [B, N, H] hidden states, got %s" % (tuple(t.shape),))
        B, N, H = t.shape
        x = t if t.dtype == torch.float32 else t.float()
        return Act(x.contiguous().view(B * N, H), None, B, N)

    def tensor(self):
        return self.x32.view(self.B, self.N, -1)


def _mask_parts(mask, B, Nq, Nk):
    """Reference additive mask ([B,1,1,Nk], [B,1,Nq,Nk], [B,Nk] or None) -> (kmask [B,Nk], bias [B,Nq,Nk])."""
    if mask is None:
        return None, None
    if mask.dim() == 2:
        return mask.to(torch.float32).contiguous(), None
    if mask.dim() != 4 or mask.shape[1] != 1:
        raise ValueError("attention mask must be [B,1,1,Nk] or [B,1,Nq,Nk] (per-head masks are not on this path), got %s"
                         % (tuple(mask.shape),))
    m = mask[:, 0].to(torch.float32)
    if m.shape[1] == 1:
        return m[:, 0].expand(B, Nk).contiguous(), None
    return None, m.expand(B, Nq, Nk).contiguous()


class KVCache(object):
    """Rollout-level cache of cross-attention K|V projections.  While a cache is active
    (``kv_cache_scope``) a cross-attention whose key/value tensor is the one it saw last time -- same storage, shape and
    version counter: the instruction embeddings a navigation rollout passes to every step, or the FACL prototypes --
    reuses its projection instead of recomputing it; gradients of all uses are summed by autograd and the projection's
    backward runs once.  An entry keeps its source tensor alive, so a pointer match cannot be a recycled allocation.
    ``clear()`` at the start of a rollout (GlocalTextPathNavCMT does it in ``language`` mode)."""

    def __init__(self):
        self.entries = {}
        self.hits = 0
        self.misses = 0

    def clear(self):
        self.entries.clear()

    def get(self, attn, enc, cdt):
        s = attn.self
        tag = (enc.x32.data_ptr(), tuple(enc.x32.shape), enc.x32._version, cdt, torch.is_grad_enabled(),
               s.key.weight._version, s.value.weight._version, s.key.weight.data_ptr(), runtime.generation())
        e = self.entries.get(id(attn))
        if e is not None and e[0] == tag:
            self.hits += 1
            return e[2], e[3]
        self.misses += 1
        H = s.all_head_size
        wkv = runtime.wc_cat((s.query.weight, s.key.weight, s.value.weight), cdt)[H:]
        bkv = runtime.wc_cat((s.query.bias, s.key.bias, s.value.bias), torch.float32)[H:]
        kvp32, kvp16 = Fn.KVProjFn.apply(enc.x32, enc.x16, s.key.weight, s.key.bias, s.value.weight, s.value.bias,
                                         wkv, bkv, cdt)
        self.entries[id(attn)] = (tag, enc.x32, kvp32, kvp16)
        return kvp32, kvp16


_ACTIVE_KV_CACHE = [None]


class kv_cache_scope(object):
    """``with kv_cache_scope(cache):`` -- cross-attentions run inside reuse / fill ``cache`` (None switches it off)."""

    def __init__(self, cache):
        self.cache = cache

    def __enter__(self):
        self.prev = _ACTIVE_KV_CACHE[0]
        _ACTIVE_KV_CACHE[0] = self.cache
        return self.cache

    def __exit__(self, *exc):
        _ACTIVE_KV_CACHE[0] = self.prev
        return False


def _p_drop(module_training, p):
    return float(p) if (module_training and p > 0.0) else 0.0


# --------------------------------------------------------------------------------------
# attention / FFN blocks
# --------------------------------------------------------------------------------------
class BertSelfAttention(nn.Module):
    """Parameter container (query/key/value) -- P/model/Bert_backbone.py:157-180"""

    def __init__(self, config):
        super().__init__()
        if config.hidden_size % config.num_attention_heads != 0:
            raise ValueError("hidden size %d is not a multiple of the number of heads %d"
                             % (config.hidden_size, config.num_attention_heads))
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = config.hidden_size // config.num_attention_heads
        self.all_head_size = config.hidden_size
        if self.attention_head_size != 64:
            raise ValueError("libgoat_sm100 attention needs head size 64 (got %d)" % self.attention_head_size)
        if getattr(config, "position_embedding_type", "absolute") != "absolute":
            raise ValueError("only absolute position embeddings are on the GOAT path (SURVEY.md 8a note 1)")
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        self.key = nn.Linear(config.hidden_size, self.all_head_size)
        self.value = nn.Linear(config.hidden_size, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)


class BertSelfOutput(nn.Module):
    """Parameter container (dense, LayerNorm) -- P/model/Bert_backbone.py:299-310"""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class BertAttention(nn.Module):
    """LN(dropout(W_o MHA(x [, enc])) + x).  One fused-QKV GEMM, the attention core, one out-proj GEMM with
    bias+dropout+residual epilogue and one LayerNorm kernel."""

    def __init__(self, config, position_embedding_type=None):
        super().__init__()
        self.self = BertSelfAttention(config)
        self.output = BertSelfOutput(config)
        self.pruned_heads = set()

    def run(self, x, mask=None, enc=None, enc_mask=None, bias=None):
        """Act in, Act out (stack-internal entry point).  ``bias``: optional additive [B,Nq,Nk] / [B,1,Nq,Nk] score
        bias of a self-attention (graph_sprels), kept apart from the key mask so neither is materialised per query."""
        x = Act.of(x)
        s, o = self.self, self.output
        cdt = runtime.compute_dtype()
        cross = enc is not None
        if cross:
            enc = Act.of(enc)
            if enc.B != x.B:
                raise ValueError("cross-attention batch mismatch: %d vs %d" % (x.B, enc.B))
            Nk = enc.N
            kmask, mbias = _mask_parts(enc_mask, x.B, x.N, Nk)
        else:
            Nk = x.N
            kmask, mbias = _mask_parts(mask, x.B, x.N, Nk)
        if bias is not None:
            if bias.dim() == 4:
                bias = bias[:, 0]
            bias = bias.to(torch.float32).expand(x.B, x.N, Nk).contiguous()
            bias = bias if mbias is None else bias + mbias
        else:
            bias = mbias
        seed = Fn.next_seed() if self.training else 0
        cfg = Fn.AttnCfg(x.B, x.N, Nk, s.num_attention_heads, o.LayerNorm.eps,
                         _p_drop(self.training, s.dropout.p), _p_drop(self.training, o.dropout.p), seed,
                         Fn.seed_ptr(), cross, cdt)
        qkv = (s.query.weight, s.key.weight, s.value.weight)
        kvp32 = kvp16 = None
        cache = _ACTIVE_KV_CACHE[0] if cross else None
        if cache is not None:
            kvp32, kvp16 = cache.get(self, enc, cdt)
        y32, y16 = Fn.AttnBlockFn.apply(
            x.x32, x.x16, enc.x32 if cross else None, enc.x16 if cross else None, kmask, bias,
            s.query.weight, s.query.bias, s.key.weight, s.key.bias, s.value.weight, s.value.bias,
            o.dense.weight, o.dense.bias, o.LayerNorm.weight, o.LayerNorm.bias,
            runtime.wc_cat(qkv, cdt), runtime.wc_cat((s.query.bias, s.key.bias, s.value.bias), torch.float32),
            runtime.wc(o.dense.weight, cdt), cfg, kvp32, kvp16)
        return Act(y32, y16, x.B, x.N)

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_value=None, output_attentions=False):
        if head_mask is not None or past_key_value is not None:
            raise NotImplementedError("head_mask / past_key_value are not used on the GOAT path")
        return (self.run(hidden_states, attention_mask, encoder_hidden_states, encoder_attention_mask).tensor(),)


RobertaAttention = BertAttention
RobertaSelfAttention = BertSelfAttention
RobertaSelfOutput = BertSelfOutput


class BertIntermediate(nn.Module):
    """Parameter container -- P/model/Bert_backbone.py:345-357 (erf GELU only)"""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)
        if config.hidden_act not in ("gelu",):
            raise ValueError("libgoat_sm100 FFN implements the reference's erf GELU only (got %r)" % (config.hidden_act,))


class BertOutput(nn.Module):
    """Parameter container -- P/model/Bert_backbone.py:359-370"""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


RobertaIntermediate = BertIntermediate
RobertaOutput = BertOutput


def ffn_block(inter, out, x, training):
    """LN(dropout(W2 gelu(W1 x + b1) + b2) + x)  (BertIntermediate + BertOutput)"""
    x = Act.of(x)
    cdt = runtime.compute_dtype()
    cfg = Fn.FFNCfg(out.LayerNorm.eps, _p_drop(training, out.dropout.p), Fn.next_seed() if training else 0,
                    Fn.seed_ptr(), cdt)
    y32, y16 = Fn.FFNBlockFn.apply(x.x32, x.x16, inter.dense.weight, inter.dense.bias, out.dense.weight, out.dense.bias,
                                   out.LayerNorm.weight, out.LayerNorm.bias, runtime.wc(inter.dense.weight, cdt),
                                   runtime.wc(out.dense.weight, cdt), cfg)
    return Act(y32, y16, x.B, x.N)


class RobertaLayer(nn.Module):
    """Self-attention block + FFN block (encoder only).  P/model/Bert_backbone.py:574-659"""

    def __init__(self, config):
        super().__init__()
        if getattr(config, "is_decoder", False) or getattr(config, "add_cross_attention", False):
            raise ValueError("decoder / cross-attention RobertaLayer is not on the GOAT path")
        self.attention = RobertaAttention(config)
        self.intermediate = RobertaIntermediate(config)
        self.output = RobertaOutput(config)

    def run(self, x, mask):
        a = self.attention.run(x, mask)
        return ffn_block(self.intermediate, self.output, a, self.training)

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_value=None, output_attentions=False):
        return (self.run(hidden_states, attention_mask).tensor(),)


class LanguageEncoder(nn.Module):
    """num_l_layers x RobertaLayer over the instruction tokens.  P/model/vilmodel_goat.py:24-44"""

    def __init__(self, config):
        super().__init__()
        self.num_l_layers = config.num_l_layers
        self.update_lang_bert = config.update_lang_bert
        self.layer = nn.ModuleList([RobertaLayer(config) for _ in range(self.num_l_layers)])
        if not self.update_lang_bert:
            for _, param in self.layer.named_parameters():
                param.requires_grad = False

    def run(self, txt_embeds, txt_masks):
        x = Act.of(txt_embeds)
        m = extend_neg_masks(txt_masks)[:, 0, 0].contiguous()
        for layer in self.layer:
            x = layer.run(x, m)
        if not self.update_lang_bert:
            x = Act(x.x32.detach(), x.x16, x.B, x.N)
        return x

    def forward(self, txt_embeds, txt_masks):
        return self.run(txt_embeds, txt_masks).tensor()


class BertCrossLayer(nn.Module):
    """self-attn (+ graph_sprels bias) -> cross-attn onto the other modality -> FFN.
    P/model/Bert_backbone.py:661-754"""

    def __init__(self, config):
        super().__init__()
        self.attention = BertAttention(config)
        self.crossattention = BertAttention(config)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)
        if getattr(config, "use_lang2visn_attn", False):
            self.lang_self_attn = BertAttention(config)
            self.lang_inter = RobertaIntermediate(config)
            self.lang_output = RobertaOutput(config)

    def run(self, x, enc, mask=None, enc_mask=None, graph_sprels=None):
        # the sprel bias is added to the query-side mask and so only reaches self-attention (:690-698)
        a = self.attention.run(x, mask, bias=graph_sprels)
        c = self.crossattention.run(a, None, enc, enc_mask)
        return ffn_block(self.intermediate, self.output, c, self.training)

    def forward(self, hidden_states, encoder_hidden_states, attention_mask=None, encoder_attention_mask=None,
                output_attentions=False, graph_sprels=None):
        return (self.run(hidden_states, encoder_hidden_states, attention_mask, encoder_attention_mask,
                         graph_sprels).tensor(),)

    def run_lang2visn(self, lang, lang_mask, visn, visn_mask):
        a = self.crossattention.run(lang, None, visn, visn_mask)
        s = self.lang_self_attn.run(a, lang_mask)
        return ffn_block(self.lang_inter, self.lang_output, s, self.training)

    def forward_lang2visn(self, lang_feats, lang_attention_mask, visn_feats, visn_attention_mask):
        return self.run_lang2visn(lang_feats, lang_attention_mask, visn_feats, visn_attention_mask).tensor()


class CrossmodalEncoder(nn.Module):
    """num_top_layer x BertCrossLayer.  P/model/Bert_backbone.py:756-781"""

    def __init__(self, config):
        super().__init__()
        self.num_top_layer = config.num_top_layer
        self.crossattention = nn.ModuleList([BertCrossLayer(config) for _ in range(self.num_top_layer)])
        self.crossattention.apply(init_weights)

    def run(self, q_embeds, q_masks, kv_embeds, kv_masks, graph_sprels=None):
        if q_masks is not None and q_masks.dim() != 4:
            q_masks = extend_neg_masks(q_masks)
        if kv_embeds is not None and kv_masks.dim() != 4:
            kv_masks = extend_neg_masks(kv_masks)
        q = Act.of(q_embeds)
        kv = Act.of(kv_embeds)
        for layer in self.crossattention:
            q = layer.run(q, kv, q_masks, kv_masks, graph_sprels)
        return q

    def forward(self, q_embeds, q_masks, kv_embeds, kv_masks, graph_sprels=None):
        return self.run(q_embeds, q_masks, kv_embeds, kv_masks, graph_sprels).tensor()


# --------------------------------------------------------------------------------------
# panorama encoder: pre-LN layers around a packed-in_proj multi-head attention
# --------------------------------------------------------------------------------------
class _MHAParams(nn.Module):
    """Parameter container with nn.MultiheadAttention's names (in_proj_weight/bias, out_proj.*) and init."""

    def __init__(self, d_model, nhead):
        super().__init__()
        self.embed_dim, self.num_heads = d_model, nhead
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = nn.Linear(d_model, d_model)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.0)


class TransformerEncoderLayer(nn.Module):
    """forward_pre of P/model/transformer.py:170-182 (normalize_before=True is the only mode GOAT builds,
    P/model/ops.py:11-23)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__()
        if not normalize_before:
            raise ValueError("GOAT's pano encoder is pre-LN (normalize_before=True); post-LN is not on the path")
        if activation != "gelu":
            raise ValueError("GOAT's pano encoder uses gelu (config.hidden_act)")
        if d_model // nhead != 64:
            raise ValueError("libgoat_sm100 attention needs head size 64")
        self.self_attn = _MHAParams(d_model, nhead)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.normalize_before = normalize_before

    def run(self, x, kmask):
        cdt = runtime.compute_dtype()
        a = self.self_attn
        cfg = Fn.PanoCfg(x.B, x.N, a.num_heads, self.norm1.eps, _p_drop(self.training, self.dropout.p),
                         Fn.next_seed() if self.training else 0, Fn.seed_ptr(), cdt)
        y = Fn.PanoLayerFn.apply(x.x32, kmask, a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias,
                                 self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias,
                                 self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias,
                                 runtime.wc(a.in_proj_weight, cdt), runtime.wc(a.out_proj.weight, cdt),
                                 runtime.wc(self.linear1.weight, cdt), runtime.wc(self.linear2.weight, cdt), cfg)
        return Act(y, None, x.B, x.N)


class TransformerEncoder(nn.Module):
    """Stack of pre-LN layers (+ optional final norm); batch_first API.  P/model/transformer.py:62-89"""

    def __init__(self, encoder_layer, num_layers, norm=None, batch_first=False):
        super().__init__()
        import copy
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = norm
        self.batch_first = batch_first

    def run(self, src, src_key_padding_mask=None):
        x = Act.of(src)
        kmask = None
        if src_key_padding_mask is not None:
            # bool, True = padding -> -inf on those keys (nn.MultiheadAttention key_padding_mask)
            kmask = torch.zeros(src_key_padding_mask.shape, dtype=torch.float32, device=src_key_padding_mask.device)
            kmask = kmask.masked_fill(src_key_padding_mask, float("-inf")).contiguous()
        for layer in self.layers:
            x = layer.run(x, kmask)
        if self.norm is not None:
            y32, y16 = Fn.LayerNormFn.apply(x.x32, self.norm.weight, self.norm.bias, self.norm.eps,
                                            runtime.compute_dtype())
            x = Act(y32, y16, x.B, x.N)
        return x

    def forward(self, src, mask=None, src_key_padding_mask=None, pos=None):
        if mask is not None or pos is not None:
            raise NotImplementedError("attn mask / pos are not used by GOAT's pano encoder")
        if not self.batch_first:
            src = src.transpose(0, 1)
        out = self.run(src, src_key_padding_mask).tensor()
        return out if self.batch_first else out.transpose(0, 1)


def create_transformer_encoder(config, num_layers, norm=False):
    """P/model/ops.py:11-23"""
    enc_layer = TransformerEncoderLayer(config.hidden_size, config.num_attention_heads,
                                        dim_feedforward=config.intermediate_size, dropout=config.hidden_dropout_prob,
                                        activation=config.hidden_act, normalize_before=True)
    norm_layer = BertLayerNorm(config.hidden_size, eps=1e-12) if norm else None
    return TransformerEncoder(enc_layer, num_layers, norm=norm_layer, batch_first=True)


# --------------------------------------------------------------------------------------
# small building blocks used by the embeddings / heads
# --------------------------------------------------------------------------------------
def linear(lin, x, act=ops.ACT_NONE):
    """act(x W^T + b) on [.., in] fp32 -> [.., out] fp32 through goat_gemm."""
    shp = x.shape
    x2 = x.reshape(-1, shp[-1])
    cdt = runtime.compute_dtype()
    if x2.dtype == cdt and cdt != torch.float32 and not x2.requires_grad:
        # already a 16-bit GEMM operand (rows gathered from the GPU-resident feature bank): no fp32 copy, no cast kernel
        y = Fn.LinearFn.apply(None, x2.contiguous(), lin.weight, lin.bias, runtime.wc(lin.weight, cdt), act, cdt)
        return y.view(shp[:-1] + (lin.weight.shape[0],))
    if x2.dtype != torch.float32:
        x2 = x2.float()
    y = Fn.LinearFn.apply(x2.contiguous(), None, lin.weight, lin.bias, runtime.wc(lin.weight, cdt), act, cdt)
    return y.view(shp[:-1] + (lin.weight.shape[0],))


def layer_norm(ln, x):
    """nn.LayerNorm(x) through goat_layernorm (fp32 in / fp32 out)."""
    shp = x.shape
    x2 = x.reshape(-1, shp[-1])
    if x2.dtype != torch.float32:
        x2 = x2.float()
    y32, _ = Fn.LayerNormFn.apply(x2.contiguous(), ln.weight, ln.bias, ln.eps, torch.float32)
    return y32.view(shp)

