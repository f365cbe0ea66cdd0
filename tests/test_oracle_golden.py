"""CPU: the oracle restatement against the committed golden fixtures (made from the unmodified
reference by tests/golden/make_golden.py).  fp32 oracle vs fp32 reference: tolerance 1e-5
(BASELINE.json north_star, fp32); the fp64 oracle must agree to the same bound."""
import pytest
import torch

from oracle import goat_oracle as O
from tests.helpers import assert_digests, golden, maxerr

TOL = 1e-5


def _leaf(t, dtype):
    return t.to(dtype).clone().requires_grad_(True)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_c1_cross_layer(dtype):
    g = golden("c1_cross_layer")
    P = {k: _leaf(v, dtype) for k, v in O.seeded_params(O.cross_layer_shapes(), seed=1).items()}
    q, kv = _leaf(g["q"], dtype), _leaf(g["kv"], dtype)
    kvm = O.extend_neg_masks(O.gen_seq_masks(g["txt_lens"], 80), dtype)
    qm = torch.zeros(2, 1, 1, 36, dtype=dtype)
    out = O.cross_layer(P, "", q, kv, qm, kvm)
    assert maxerr(out, g["out"]) < TOL
    (out * g["w_out"].to(dtype)).sum().backward()
    assert maxerr(q.grad, g["dq"]) < TOL
    assert maxerr(kv.grad, g["dkv"]) < TOL
    assert_digests(g, {k: v.grad for k, v in P.items()}, rtol=2e-5)
    att = O.bert_attention(P, "crossattention.", g["q"].to(dtype), None, g["kv"].to(dtype), kvm)
    assert maxerr(att, g["cross_attn_out"]) < TOL


def test_xenc_sprels():
    g = golden("xenc_sprels")
    shapes = {}
    for i in range(3):
        shapes.update(O.cross_layer_shapes("crossattention.%d." % i))
    # the reference module also owns lang_* extras (use_lang2visn_attn); they do not enter forward()
    P = {k: _leaf(v, torch.float32) for k, v in O.seeded_params(shapes, seed=2).items()}
    gm, tx, sp = _leaf(g["gmap"], torch.float32), _leaf(g["txt"], torch.float32), _leaf(g["sprels"], torch.float32)
    out = O.crossmodal_encoder(P, "", gm, O.gen_seq_masks(g["gmap_lens"], 12), tx,
                               O.gen_seq_masks(g["txt_lens"], 40), sp)
    assert maxerr(out, g["out"]) < TOL
    (out * g["w_out"]).sum().backward()
    assert maxerr(gm.grad, g["dgmap"]) < TOL
    assert maxerr(tx.grad, g["dtxt"]) < TOL
    assert maxerr(sp.grad, g["dsprels"]) < TOL
    assert_digests({k: v for k, v in g.items() if "lang_" not in k}, {k: v.grad for k, v in P.items()}, rtol=2e-5)


def test_lang_encoder():
    g = golden("lang_encoder")
    shapes = {}
    for i in range(6):
        shapes.update(O.roberta_layer_shapes("layer.%d." % i))
    P = {k: _leaf(v, torch.float32) for k, v in O.seeded_params(shapes, seed=3).items()}
    x = _leaf(g["x"], torch.float32)
    out = O.lang_encoder(P, "", x, O.gen_seq_masks(g["txt_lens"], 80))
    assert maxerr(out, g["out"]) < 2 * TOL
    (out * g["w_out"]).sum().backward()
    assert maxerr(x.grad, g["dx"]) < 2 * TOL
    assert_digests(g, {k: v.grad for k, v in P.items()}, rtol=2e-5)


def test_pano_encoder():
    g = golden("pano_encoder")
    P = {k: _leaf(v, torch.float32) for k, v in O.seeded_params(O.pano_encoder_shapes(), seed=4).items()}
    x = _leaf(g["x"], torch.float32)
    kpm = O.gen_seq_masks(g["view_lens"], 36).logical_not()
    out = O.pano_encoder(P, "", x, kpm)
    assert maxerr(out, g["out"]) < TOL
    (out * g["w_out"]).sum().backward()
    assert maxerr(x.grad, g["dx"]) < 2 * TOL
    assert_digests(g, {k: v.grad for k, v in P.items()}, rtol=2e-5)


def test_heads():
    g = golden("heads")
    x = g["x"]
    H = 768
    ht = O.seeded_params({"dense.weight": (H, H), "dense.bias": (H,), "LayerNorm.weight": (H,),
                          "LayerNorm.bias": (H,)}, seed=5)
    assert maxerr(O.head_transform(ht, "", x), g["ht_out"]) < TOL
    cp = O.seeded_params({"net.0.weight": (H, H), "net.0.bias": (H,), "net.2.weight": (H,), "net.2.bias": (H,),
                          "net.3.weight": (1, H), "net.3.bias": (1,)}, seed=6)
    assert maxerr(O.cls_prediction(cp, "", x), g["cp_out"]) < TOL
    assert maxerr(O.pano_fuse(x, g["fuse_w"], g["fuse_b"]), g["fuse_out"]) < TOL
    pool = O.cfp_pool(x, g["pool_w"])
    assert maxerr(pool, g["pool_out"]) < TOL
    assert maxerr(O.infonce_sym(pool, g["nce_y"], 1.0), g["nce"]) < TOL


def test_front_door():
    g = golden("front_door")
    shapes = {}
    shapes.update(O.attn_shapes("ll_self_attn."))
    shapes.update(O.attn_shapes("lg_cross_attn."))
    shapes.update({"ln.weight": (768,), "ln.bias": (768,), "aug_linear.weight": (1, 768), "aug_linear.bias": (1,),
                   "ori_linear.weight": (1, 768), "ori_linear.bias": (1,)})
    P = {k: _leaf(v, torch.float32) for k, v in O.seeded_params(shapes, seed=7).items()}
    x = _leaf(g["x"], torch.float32)
    out = O.front_door_encoder(P, "", x, g["proto"], O.gen_seq_masks(g["lens"], 38), eps=1e-5)
    assert maxerr(out, g["out"]) < TOL
    (out * g["w_out"]).sum().backward()
    assert maxerr(x.grad, g["dx"]) < TOL
    assert_digests(g, {k: v.grad for k, v in P.items()}, rtol=2e-5)


def test_lang_encoder_do_tail_and_bacl_image():
    g = golden("lang_encoder_do")
    shapes = {}
    for i in range(6):
        shapes.update(O.roberta_layer_shapes("layer.%d." % i))
    for n in ("z_direc_cross_attn.", "z_landm_cross_attn.", "z_front_cross_attn."):
        shapes.update(O.attn_shapes(n))
    for n in ("z_txt_linear", "z_direct_linear", "z_landm_linear", "z_front_linear"):
        shapes.update({n + ".weight": (768, 768), n + ".bias": (768,)})
    for n in ("z_concat_layernorm", "z_direct_ln", "z_landm_ln", "z_front_ln"):
        shapes.update({n + ".weight": (768,), n + ".bias": (768,)})
    shapes.update({"instr_aug_linear.weight": (1, 768), "instr_aug_linear.bias": (1,),
                   "instr_ori_linear.weight": (1, 768), "instr_ori_linear.bias": (1,)})
    P = O.seeded_params(shapes, seed=8)
    txt = _leaf(g["txt"], torch.float32)
    t = O.lang_encoder(P, "", txt, O.gen_seq_masks(g["txt_lens"], 44), 6, eps=1e-5)
    out = O.bacl_text_type2_door(P, "", t, g["z_direc"], g["z_landm"], g["front_txt"], eps=1e-5)
    assert maxerr(out, g["out"]) < 2 * TOL
    (out * g["w_out"]).sum().backward()
    assert maxerr(txt.grad, g["dtxt"]) < 2 * TOL

    g = golden("bacl_image")
    H = 768
    shapes = {}
    for n in ("do_img_before_linear", "do_img_after_linear", "img_after_linear"):
        shapes.update({n + ".weight": (H, H), n + ".bias": (H,)})
    for n in ("do_img_layer_norm", "do_img_concat_layernorm"):
        shapes.update({n + ".weight": (H,), n + ".bias": (H,)})
    P = O.seeded_params(shapes, seed=9)
    out = O.bacl_image_type1(P, "", g["view"], g["zf"], g["pz"])
    assert maxerr(out, g["out"]) < TOL


def test_pretrain_oracle_matches_reference_fixture():
    """oracle/goat_pretrain_oracle.py (full MLM / SAP / CFP forward) vs the fixture the UNMODIFIED reference model produced
    on the same seeded batch and seeded parameters (tests/golden/make_golden.py --tree pretrain_full)."""
    from oracle import goat_pretrain_oracle as PO
    from tests import synth
    g = golden("pretrain_full")
    keys = str(g["state_dict_keys"]).split("\n")
    shapes = dict(pretrain_state_shapes())
    assert set(keys) == set(shapes), (set(keys) ^ set(shapes))
    P = O.seeded_params(shapes, seed=20)
    # tied weights: load_state_dict copies both keys into the one shared tensor, the later key (the decoder's) stays
    P["bert.embeddings.word_embeddings.weight"] = P["mlm_head.predictions.decoder.weight"]
    batch = synth.pretrain_batch(B=3, L=24, seed=5)
    with torch.no_grad():
        scores, mlm_loss = PO.forward_mlm(P, batch)
        gl, ll, fl, sap_loss = PO.forward_sap(P, batch)
        go, vo, fo, to, cfp_loss = PO.forward_cfp(P, batch)

    def rel(a, b):
        fin = torch.isfinite(b)
        assert torch.equal(torch.isfinite(a), fin)
        return (a[fin] - b[fin]).abs().max().item() / max(1.0, b[fin].abs().max().item())
    assert rel(scores[:, :64], g["mlm_scores_head"]) < TOL and rel(mlm_loss, g["mlm_loss"]) < TOL
    for a, k in ((gl, "sap_global_logits"), (ll, "sap_local_logits"), (fl, "sap_fused_logits"), (sap_loss, "sap_loss")):
        assert rel(a, g[k]) < TOL, k
    for a, k in ((go, "cfp_gmap"), (vo, "cfp_vp"), (fo, "cfp_fused"), (to, "cfp_txt"), (cfp_loss, "cfp_loss")):
        assert rel(a, g[k]) < TOL, k


def pretrain_state_shapes():
    """{state_dict key: shape} of the full pretraining model, from the product module's own parameter containers (the
    key list is checked against the reference's in tests/test_host_logic.py)."""
    from vln_goat_b200 import pretrain_model
    from vln_goat_b200.config import GoatConfig
    with torch.device("meta"):
        m = pretrain_model.GlocalTextPathCMTPreTraining(GoatConfig(pretrain_tasks=("mlm", "sap", "cfp")))
    return [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
