"""Import shim that makes the *unmodified* reference importable in this container.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vln_goat_b200/`` may import this file.
It is used by ``tests/golden/make_golden.py`` (fixture generation, run in the build
container where ``/root/reference`` is mounted) and by the CPU tests that pin
``oracle/goat_oracle.py`` against the real reference modules.  On the GPU box
``/root/reference`` does not exist, so everything here is guarded by ``available()``.

Why a shim is needed: the reference pins transformers 4.34.1 /
torch 1.9 (``requirements.txt:16,18``) while this image has transformers 5.5 / torch 2.11:
  * ``pretrain_src/model/Bert_backbone.py:10-13`` imports ``apply_chunking_to_forward``
    from ``transformers.modeling_utils`` (moved to ``transformers.pytorch_utils``);
  * ``BertPreTrainedModel.init_weights/tie_weights`` changed, so the three top-level
    classes cannot run ``self.init_weights()``;
  * ``PretrainedConfig`` lost the v4 defaults (``is_decoder`` ...).
The shim patches those three things in ``sys.modules`` and never touches reference files.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GOAT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pretrain_src", "model"))


_installed = None


def install(which: str = "pretrain"):
    """Patch transformers and put one of the two reference trees on sys.path.

    which = "pretrain" -> package ``model`` (pretrain_src), "nav" -> package ``models``
    (map_nav_src).  Their top-level names collide (utils, parser), so one per process.
    """
    global _installed
    if _installed is not None:
        if _installed != which:
            raise RuntimeError("reference shim already installed for %r" % _installed)
        return
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    import torch
    from torch import nn
    import transformers
    import transformers.modeling_utils as mu
    from transformers.pytorch_utils import apply_chunking_to_forward

    mu.apply_chunking_to_forward = apply_chunking_to_forward

    class _StubBertPreTrainedModel(nn.Module):
        """v4-era BertPreTrainedModel behaviour: N(0, initializer_range) Linear/Embedding,
        zero bias, unit LayerNorm; weight tying by sharing the Parameter."""
        base_model_prefix = "bert"

        def __init__(self, config, *a, **k):
            super().__init__()
            self.config = config

        def _init_weights(self, module):
            if isinstance(module, nn.Linear):
                module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
                if module.bias is not None:
                    module.bias.data.zero_()
            elif isinstance(module, nn.Embedding):
                module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
                if module.padding_idx is not None:
                    module.weight.data[module.padding_idx].zero_()
            elif isinstance(module, nn.LayerNorm):
                module.bias.data.zero_()
                module.weight.data.fill_(1.0)

        def init_weights(self):
            self.apply(self._init_weights)

        def _tie_or_clone_weights(self, out, inp):
            out.weight = inp.weight

        @classmethod
        def from_pretrained(cls, pretrained_model_name_or_path=None, config=None, state_dict=None, **kw):
            m = cls(config)
            if state_dict:
                m.load_state_dict(state_dict, strict=False)
            return m

    import transformers.models.roberta.modeling_roberta  # noqa: F401  (must be imported before rebinding)
    import transformers.models.bert.modeling_bert as _mb

    _mb.BertPreTrainedModel = _StubBertPreTrainedModel
    transformers.BertPreTrainedModel = _StubBertPreTrainedModel

    sub = "pretrain_src" if which == "pretrain" else "map_nav_src"
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, sub))
    if which == "pretrain":
        # pretrain_goat.py imports data.common.check_gpu_mem_usedRate (needs pynvml + h5py chain);
        # only the symbol is needed and it is never called with empty_cache=False.
        if "data" not in sys.modules:
            data_pkg = types.ModuleType("data")
            data_pkg.__path__ = []
            common = types.ModuleType("data.common")
            common.check_gpu_mem_usedRate = lambda *a, **k: (0, 0.0, 1)
            data_pkg.common = common
            sys.modules["data"] = data_pkg
            sys.modules["data.common"] = common
    _installed = which


def v4_defaults(cfg, pad_token_id=None):
    for k, v in dict(pad_token_id=pad_token_id, is_decoder=False, add_cross_attention=False,
                     chunk_size_feed_forward=0, initializer_range=0.02).items():
        if not hasattr(cfg, k) or (k == "pad_token_id" and getattr(cfg, k, None) is None and v is not None):
            setattr(cfg, k, v)
    return cfg


def pretrain_config(**overrides):
    """Config of P/train_r2r_goat.py:102-107,189 built from the shipped JSON."""
    from transformers import PretrainedConfig
    cfg = PretrainedConfig.from_json_file(
        os.path.join(REFERENCE_ROOT, "pretrain_src/config/r2r_GOAT_model_config.json"))
    cfg.pretrain_tasks = {"mlm", "sap", "cfp"}
    cfg.name = "R2R"
    cfg.cuda_first_device = 0
    cfg.empty_cache = False
    v4_defaults(cfg)
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def nav_config(**overrides):
    """Fine-tune config assembled as M/models/vlnbert_init.py:89-154 does, on top of the
    roberta-base defaults it starts from (hidden_act gelu, layer_norm_eps 1e-5, pad 1)."""
    from transformers import PretrainedConfig
    cfg = PretrainedConfig()
    base = dict(hidden_act="gelu", layer_norm_eps=1e-5, pad_token_id=1, initializer_range=0.02,
                type_vocab_size=1, max_position_embeddings=514, vocab_size=50265,
                dataset="r2r", mode="train", max_action_steps=100, image_feat_size=768,
                angle_feat_size=4, obj_feat_size=0, obj_loc_size=3, obj_name_vocab_size=45,
                num_l_layers=6, num_pano_layers=2, num_x_layers=3, graph_sprels=True,
                glocal_fuse=True, fix_lang_embedding=False, fix_pano_embedding=False,
                fix_local_branch=False, update_lang_bert=True, output_attentions=True,
                pred_head_dropout_prob=0.1, max_instr_len=200, feat_dropout=0.4,
                adaptive_pano_fusion=True, do_back_img=True, do_back_txt=True,
                do_front_img=True, do_front_his=True, do_front_txt=True, cfp_temperature=1.0,
                do_back_txt_type="type_2", do_back_img_type="type_1", do_add_method="door",
                mlm_prob=0.15, draw_false_text=0, num_top_layer=3, input_image_embed_size=768,
                input_text_embed_size=768, hidden_size=768, num_attention_heads=12,
                num_hidden_layers=6, mlp_ratio=4, hidden_dropout_prob=0.1,
                attention_probs_dropout_prob=0.1, intermediate_size=3072, name="R2R",
                use_lang2visn_attn=False)
    base.update(overrides)
    for k, v in base.items():
        setattr(cfg, k, v)
    v4_defaults(cfg, pad_token_id=1)
    return cfg
