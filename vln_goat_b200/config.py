"""Model constants of the GOAT path as a plain config object.

The reference passes a HuggingFace ``PretrainedConfig`` built from
P/config/r2r_GOAT_model_config.json (pretrain, P/train_r2r_goat.py:102-107) or assembled in code
(fine-tune, M/models/vlnbert_init.py:79-154).  The blocks in ``modules`` only read attributes, so
any object with these names works -- including the reference's own config instance.
"""


class GoatConfig(object):
    """Defaults = P/config/r2r_GOAT_model_config.json:5-58 (+ the v4 PretrainedConfig defaults the
    reference relied on: pad_token_id, is_decoder, add_cross_attention, chunk_size_feed_forward)."""

    def __init__(self, **overrides):
        self.hidden_size = 768
        self.num_attention_heads = 12
        self.intermediate_size = 3072
        self.hidden_act = "gelu"
        self.hidden_dropout_prob = 0.1
        self.attention_probs_dropout_prob = 0.1
        self.pred_head_dropout_prob = 0.1
        self.layer_norm_eps = 1e-12
        self.initializer_range = 0.02
        self.vocab_size = 50265
        self.max_position_embeddings = 514
        self.type_vocab_size = 1
        self.pad_token_id = None
        self.num_l_layers = 6
        self.num_x_layers = 3
        self.num_top_layer = 3
        self.num_pano_layers = 2
        self.max_action_steps = 100
        self.image_feat_size = 768
        self.angle_feat_size = 4
        self.obj_feat_size = 0
        self.update_lang_bert = True
        self.use_lang2visn_attn = True
        self.graph_sprels = True
        self.glocal_fuse = True
        self.adaptive_pano_fusion = True
        self.cfp_extra_head = True
        self.cfp_temperature = 1.0
        self.do_back_txt = False
        self.do_back_img = False
        self.do_back_txt_type = "type_1"
        self.do_back_imgobj_type = "type_1"
        self.do_add_method = "add"
        self.do_front_img = False
        self.do_front_his = False
        self.do_front_txt = False
        self.front_n_clusters = 24
        self.z_cross_attn = False
        self.is_decoder = False
        self.add_cross_attention = False
        self.chunk_size_feed_forward = 0
        self.empty_cache = False
        self.name = "R2R"
        self.pretrain_tasks = ("mlm", "sap", "cfp")
        for k, v in overrides.items():
            setattr(self, k, v)
