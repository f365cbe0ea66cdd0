"""Generate the committed golden fixtures from the UNMODIFIED reference modules.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

Every fixture is: seeded inputs + seeded parameters (oracle.goat_oracle.seeded_params over the
reference module's own state_dict keys/shapes, loaded with load_state_dict) -> reference forward
and backward on CPU fp32 with dropout off (.eval()) -> outputs, input grads and parameter-grad
digests saved to tests/golden/*.npz.  Parameters are NOT stored (one cross layer is 38 MB);
they are regenerated from (seed, key) by the same recipe wherever the fixture is consumed.
The reference cannot travel to the GPU box, these files can.  The two reference trees need
separate processes (their top-level package names collide), hence --tree.
"""
import argparse
import os
import subprocess
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import goat_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402


def digest(t):
    """Small, order-sensitive summary of a big tensor: [sum, abs-sum, cos-weighted sum, first, last]."""
    t = t.detach().double().flatten()
    idx = torch.arange(t.numel(), dtype=torch.float64)
    return np.array([t.sum().item(), t.abs().sum().item(), (t * torch.cos(idx * 0.37)).sum().item(),
                     t[0].item(), t[-1].item()], dtype=np.float64)


def grads_digest(module):
    return {("gdig." + n): digest(p.grad) for n, p in module.named_parameters() if p.grad is not None}


def load_seeded(module, seed):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    module.load_state_dict(O.seeded_params(shapes, seed=seed), strict=True)
    return module


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def gen_pretrain():
    ref_shim.install("pretrain")
    from model.Bert_backbone import BertAttention, BertCrossLayer, CrossmodalEncoder, \
        BertPredictionHeadTransform
    from model.vilmodel_goat import LanguageEncoder
    from model.ops import create_transformer_encoder, extend_neg_masks
    from model.pretrain_goat import ClsPrediction
    cfg = ref_shim.pretrain_config()

    # ---- C1 (BASELINE.json configs[0]): one BertCrossLayer, pano [2,36,768] x text [2,80,768],
    #      text lengths {80,57}; plus the bare cross BertAttention of the same layer
    g = torch.Generator().manual_seed(0)
    q = torch.randn(2, 36, 768, generator=g)
    kv = torch.randn(2, 80, 768, generator=g)
    layer = load_seeded(BertCrossLayer(cfg).eval(), seed=1)
    q_ = q.clone().requires_grad_(True)
    kv_ = kv.clone().requires_grad_(True)
    txt_lens = torch.tensor([80, 57])
    kvm = extend_neg_masks(O.gen_seq_masks(txt_lens, 80))
    qm = extend_neg_masks(torch.ones(2, 36, dtype=torch.bool))
    out = layer(q_, kv_, attention_mask=qm, encoder_attention_mask=kvm)[0]
    w_out = torch.randn(out.shape, generator=g)
    (out * w_out).sum().backward()
    att = BertAttention(cfg).eval()
    att.load_state_dict({k[len("crossattention."):]: v for k, v in layer.state_dict().items()
                         if k.startswith("crossattention.")})
    att_out = att(q, None, None, kv, kvm)[0]
    save("c1_cross_layer", q=q, kv=kv, txt_lens=txt_lens.numpy(), out=out, w_out=w_out, dq=q_.grad,
         dkv=kv_.grad, cross_attn_out=att_out, **grads_digest(layer))

    # ---- CrossmodalEncoder(3) with graph_sprels, ragged gmap / text lengths
    g = torch.Generator().manual_seed(3)
    B, G, L = 3, 12, 40
    gm = torch.randn(B, G, 768, generator=g)
    tx = torch.randn(B, L, 768, generator=g)
    gl = torch.tensor([12, 7, 3])
    tl = torch.tensor([40, 33, 21])
    sp = torch.rand(B, G, G, generator=g) * 10.0
    sp = (sp + sp.transpose(1, 2)) / 2
    sprels = (0.3 * sp - 0.1)[:, None]       # what sprel_linear (1->1) would emit
    enc = load_seeded(CrossmodalEncoder(cfg).eval(), seed=2)
    gm_ = gm.clone().requires_grad_(True)
    tx_ = tx.clone().requires_grad_(True)
    sp_ = sprels.clone().requires_grad_(True)
    out = enc(gm_, O.gen_seq_masks(gl, G), tx_, O.gen_seq_masks(tl, L), graph_sprels=sp_)
    w_out = torch.randn(out.shape, generator=g)
    (out * w_out).sum().backward()
    save("xenc_sprels", gmap=gm, txt=tx, gmap_lens=gl.numpy(), txt_lens=tl.numpy(), sprels=sprels, out=out,
         w_out=w_out, dgmap=gm_.grad, dtxt=tx_.grad, dsprels=sp_.grad, **grads_digest(enc))

    # ---- LanguageEncoder (6 RobertaLayers), ragged lengths
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 80, 768, generator=g)
    tl = torch.tensor([80, 41])
    le = load_seeded(LanguageEncoder(cfg).eval(), seed=3)
    x_ = x.clone().requires_grad_(True)
    out = le(x_, O.gen_seq_masks(tl, 80))
    w_out = torch.randn(out.shape, generator=g)
    (out * w_out).sum().backward()
    save("lang_encoder", x=x, txt_lens=tl.numpy(), out=out, w_out=w_out, dx=x_.grad, **grads_digest(le))

    # ---- pano encoder (2 pre-LN layers + final LN), key padding {36, 29}
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 36, 768, generator=g)
    vl = torch.tensor([36, 29])
    pe = load_seeded(create_transformer_encoder(cfg, cfg.num_pano_layers, norm=True).eval(), seed=4)
    x_ = x.clone().requires_grad_(True)
    out = pe(x_, src_key_padding_mask=O.gen_seq_masks(vl, 36).logical_not())
    w_out = torch.randn(out.shape, generator=g)
    (out * w_out).sum().backward()
    save("pano_encoder", x=x, view_lens=vl.numpy(), out=out, w_out=w_out, dx=x_.grad, **grads_digest(pe))

    # ---- heads: BertPredictionHeadTransform, ClsPrediction, pano fusion, CFP pooling, InfoNCE
    g = torch.Generator().manual_seed(8)
    x = torch.randn(4, 11, 768, generator=g)
    ht = load_seeded(BertPredictionHeadTransform(cfg).eval(), seed=5)
    ht_out = ht(x)
    cp = load_seeded(ClsPrediction(768).eval(), seed=6)
    cp_out = cp(x)
    # adaptive pano fusion exactly as P/model/vilmodel_goat.py:354-362
    fw = 0.05 * torch.randn(1, 768, generator=g)
    fb = 0.05 * torch.randn(1, generator=g)
    a = torch.softmax(torch.tanh(torch.nn.functional.linear(x, fw, fb)), dim=1)
    fuse_out = torch.sum(torch.mul(x, a), dim=1)
    # CFP pooling exactly as P/model/pretrain_goat.py:502-505
    aw = (torch.rand(768, 1, generator=g) - 0.5) * 0.2
    M1 = torch.tanh(x)
    a1 = torch.softmax(torch.matmul(M1, aw), 1)
    pool_out = torch.tanh(torch.sum(x * a1, 1))
    # InfoNCE as :519-524 (without the hard-coded .cuda())
    y = torch.tanh(torch.randn(4, 768, generator=g))
    tgt = torch.arange(4)
    sim = (pool_out @ y.T) / 1.0
    nce = (torch.nn.functional.cross_entropy(sim, tgt, reduction="none") +
           torch.nn.functional.cross_entropy(sim.T, tgt, reduction="none")) / 2.0
    save("heads", x=x, ht_out=ht_out, cp_out=cp_out, fuse_w=fw, fuse_b=fb, fuse_out=fuse_out, pool_w=aw,
         pool_out=pool_out, nce_y=y, nce=nce)


def gen_nav():
    ref_shim.install("nav")
    import models.vilmodel_GOAT as V
    cfg = ref_shim.nav_config()

    # ---- FrontDoorEncoder (FACL): vp tokens [3,38,768], prototypes [3,24,768], ragged mask
    g = torch.Generator().manual_seed(11)
    fd = load_seeded(V.FrontDoorEncoder(cfg).eval(), seed=7)
    x = torch.randn(3, 38, 768, generator=g)
    proto = torch.tanh(torch.randn(3, 24, 768, generator=g))
    lens = torch.tensor([38, 30, 17])
    x_ = x.clone().requires_grad_(True)
    out = fd(x_, proto, O.gen_seq_masks(lens, 38))
    w_out = torch.randn(out.shape, generator=g)
    (out * w_out).sum().backward()
    save("front_door", x=x, proto=proto, lens=lens.numpy(), out=out, w_out=w_out, dx=x_.grad, **grads_digest(fd))

    # ---- LanguageEncoderDo causal tail (BACL text type_2 + FACL text, door), on given txt embeds
    g = torch.Generator().manual_seed(12)
    le = load_seeded(V.LanguageEncoderDo(cfg).eval(), seed=8)
    txt = torch.randn(2, 44, 768, generator=g)
    zd = torch.randn(2, 35, 768, generator=g)
    zl = torch.randn(2, 39, 768, generator=g)
    ft = torch.tanh(torch.randn(2, 24, 768, generator=g))
    tl = torch.tensor([44, 30])
    txt_ = txt.clone().requires_grad_(True)
    out = le(txt_, O.gen_seq_masks(tl, 44), zd, None, zl, None, ft)
    w_out = torch.randn(out.shape, generator=g)
    (out * w_out).sum().backward()
    save("lang_encoder_do", txt=txt, z_direc=zd, z_landm=zl, front_txt=ft, txt_lens=tl.numpy(), out=out,
         w_out=w_out, dtxt=txt_.grad, **grads_digest(le))

    # ---- BACL image type_1 exactly as M/models/vilmodel_GOAT.py:661-667
    g = torch.Generator().manual_seed(13)
    ie = load_seeded(V.CausalImageEmbeddings(cfg).eval(), seed=9)
    view = torch.randn(2, 36, 768, generator=g)
    zf = torch.randn(2, 50, 768, generator=g)
    pz = torch.rand(2, 50, 1, generator=g, dtype=torch.float64)
    pz = pz / pz.sum(1, keepdim=True)
    z = ie.do_img_layer_norm(ie.do_img_before_linear(zf))
    s = torch.sum(z * pz.to(torch.float32), 1).unsqueeze(1)
    y = ie.do_img_concat_layernorm(ie.img_after_linear(view) + ie.do_img_after_linear(s))
    save("bacl_image", view=view, zf=zf, pz=pz, out=y)


def _keys_blob(module):
    return np.array("\n".join(module.state_dict().keys()))


def gen_pretrain_full():
    """Full pretrain model (208 M parameters, seeded): MLM / SAP / CFP on one synthetic batch (tests/synth.py)."""
    ref_shim.install("pretrain")
    from model.pretrain_goat import GlocalTextPathCMTPreTraining
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    cfg = ref_shim.pretrain_config()
    model = load_seeded(GlocalTextPathCMTPreTraining(cfg).eval(), seed=20)
    model.tie_weights()
    batch = synth.pretrain_batch(B=3, L=24, seed=5)
    out = {"state_dict_keys": _keys_blob(model)}
    # MLM
    model.zero_grad()
    scores = model(batch, "mlm", compute_loss=False)
    loss = model(batch, "mlm", compute_loss=True)
    loss.sum().backward()
    out.update(mlm_scores_digest=digest(scores), mlm_scores_head=scores[:, :64], mlm_loss=loss)
    out.update({"mlm." + k: v for k, v in grads_digest(model).items()})
    # SAP
    model.zero_grad()
    gl, ll, fl, _, _ = model(batch, "sap", compute_loss=False)
    loss = model(batch, "sap", compute_loss=True)
    loss.sum().backward()
    out.update(sap_global_logits=gl, sap_local_logits=ll, sap_fused_logits=fl, sap_loss=loss)
    out.update({"sap." + k: v for k, v in grads_digest(model).items()})
    # CFP (the reference hard-codes .cuda() on the target, P/model/pretrain_goat.py:520)
    model.zero_grad()
    go, vo, fo, to = model(batch, "cfp", compute_loss=False)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        loss = model(batch, "cfp", compute_loss=True)
    finally:
        torch.Tensor.cuda = orig_cuda
    loss.sum().backward()
    out.update(cfp_gmap=go, cfp_vp=vo, cfp_fused=fo, cfp_txt=to, cfp_loss=loss)
    out.update({"cfp." + k: v for k, v in grads_digest(model).items()})
    save("pretrain_full", **out)


def gen_nav_full():
    """Full fine-tune model with BACL + FACL on: language -> panorama -> navigation on synthetic step inputs."""
    ref_shim.install("nav")
    import models.vilmodel_GOAT as V
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    from collections import defaultdict
    cfg = ref_shim.nav_config()
    model = load_seeded(V.GlocalTextPathNavCMT(cfg).eval(), seed=21)
    lang, pano, nav = synth.nav_inputs(B=3, L=20, seed=6)
    dd = lambda d: defaultdict(lambda: None, d)
    txt = model("language", dd(lang))
    pe, pm, pf = model("panorama", dd(pano))
    navb = dict(nav)
    mem = navb.pop("mem_embeds")
    navb["txt_embeds"] = txt
    navb["vp_img_embeds"] = torch.cat([torch.zeros_like(pe[:, :1]), mem.unsqueeze(1), pe], 1)
    outs = model("navigation", dd(navb))
    g = torch.Generator().manual_seed(99)
    w_cls = torch.randn(outs["cls_embeds"].shape, generator=g)
    w_pf = torch.randn(pf.shape, generator=g)
    target = torch.tensor([4, 0, 5])      # an unvisited node / [stop] / an unvisited node
    loss = torch.nn.functional.cross_entropy(outs["fused_logits"], target, reduction="sum") + \
        (outs["cls_embeds"] * w_cls).sum() + (pf * w_pf).sum()
    loss.backward()
    save("nav_full", state_dict_keys=_keys_blob(model), txt_embeds=txt, pano_embeds=pe, pano_masks=pm, pano_fused=pf,
         global_logits=outs["global_logits"], local_logits=outs["local_logits"], fused_logits=outs["fused_logits"],
         cls_embeds=outs["cls_embeds"], gmap_embeds=outs["gmap_embeds"], vp_embeds=outs["vp_embeds"], w_cls=w_cls,
         w_pf=w_pf, target=target, loss=loss.detach(), **grads_digest(model))


def gen_nav_rollout():
    """3-step teacher-forced rollout through the UNMODIFIED reference model (BACL + FACL on), one backward."""
    ref_shim.install("nav")
    import models.vilmodel_GOAT as V
    cfg = ref_shim.nav_config()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    model = load_seeded(V.GlocalTextPathNavCMT(cfg).eval(), seed=23)
    lang, per_step, targets = synth.rollout_inputs()
    loss, logits, clss, txt = synth.run_rollout(model, lang, per_step, targets)
    loss.backward()
    arrs = {"loss": loss.detach(), "txt_embeds": txt}
    for t, (lg, c) in enumerate(zip(logits, clss)):
        arrs["fused_logits_%d" % t] = lg
        arrs["cls_embeds_%d" % t] = c
    save("nav_rollout", **arrs, **grads_digest(model))


def gen_nav_branches():
    """The configuration branches the shipped scripts leave off but the code carries: BACL text type_1 and type_2 with
    the 'add' / 'concat' merges (M/models/vilmodel_GOAT.py:107-160), BACL image type_2 with door / add / concat (:661-683,
    run through the reference's own ``forward_panorama_do_per_step`` on a stand-in object that owns the reference
    ``CausalImageEmbeddings``).  Outputs are stored on every 4th token (plus a digest of the whole tensor) to keep the
    fixture small; input-gradient digests are stored too."""
    ref_shim.install("nav")
    import types
    import models.vilmodel_GOAT as V
    out = {}
    g = torch.Generator().manual_seed(31)
    txt = torch.randn(2, 28, 768, generator=g)
    zd = torch.randn(2, 35, 768, generator=g)
    zl = torch.randn(2, 39, 768, generator=g)
    ft = torch.tanh(torch.randn(2, 24, 768, generator=g))
    tl = torch.tensor([28, 19])

    def pz(n):
        p = torch.rand(2, n, 1, generator=g, dtype=torch.float64)
        return p / p.sum(1, keepdim=True)
    pzd, pzl = pz(35), pz(39)
    w_txt = torch.randn(2, 28, 768, generator=g)
    out.update(txt=txt, z_direc=zd, z_landm=zl, front_txt=ft, txt_lens=tl.numpy(), pz_direc=pzd, pz_landm=pzl, w_txt=w_txt)
    for tag, kw in (("t1", dict(do_back_txt_type="type_1")),
                    ("t2add", dict(do_back_txt_type="type_2", do_add_method="add")),
                    ("t2cat", dict(do_back_txt_type="type_2", do_add_method="concat"))):
        cfg = ref_shim.nav_config(**kw)
        le = load_seeded(V.LanguageEncoderDo(cfg).eval(), seed=41)
        t_ = txt.clone().requires_grad_(True)
        y = le(t_, O.gen_seq_masks(tl, 28), zd, pzd, zl, pzl, ft)
        (y * w_txt).sum().backward()
        out["txt_%s_out" % tag] = y[:, ::4]
        out["txt_%s_dig" % tag] = digest(y)
        out["txt_%s_dx_dig" % tag] = digest(t_.grad)
    view = torch.randn(2, 36, 768, generator=g)
    loc = torch.randn(2, 36, 7, generator=g)
    zf = torch.randn(2, 50, 768, generator=g)
    pzi = pz(50)
    vl = torch.tensor([36, 31])
    w_v = torch.randn(2, 36, 768, generator=g)
    out.update(view=view, loc=loc, z_img=zf, pz_img=pzi, view_lens=vl.numpy(), w_view=w_v)
    for tag, kw in (("door", dict(do_back_img_type="type_2", do_add_method="door")),
                    ("add", dict(do_back_img_type="type_2", do_add_method="add")),
                    ("cat", dict(do_back_img_type="type_2", do_add_method="concat"))):
        cfg = ref_shim.nav_config(**kw)
        ie = load_seeded(V.CausalImageEmbeddings(cfg).eval(), seed=42)
        holder = types.SimpleNamespace(img_embeddings=ie, config=cfg)
        v_ = view.clone().requires_grad_(True)
        pe, pm, pf = V.GlocalTextPathNavCMT.forward_panorama_do_per_step(holder, v_, loc, None, vl, zf, pzi)
        ((pe * w_v).sum() + pf.sum()).backward()
        out["img_%s_out" % tag] = pe[:, ::4]
        out["img_%s_dig" % tag] = digest(pe)
        out["img_%s_fused" % tag] = pf
        out["img_%s_dx_dig" % tag] = digest(v_.grad)
    save("nav_branches", **out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tree", choices=["pretrain", "nav", "pretrain_full", "nav_full", "nav_rollout", "nav_branches", "all"], default="all")
    a = ap.parse_args()
    if a.tree == "all":
        for t in ("pretrain", "nav", "pretrain_full", "nav_full", "nav_rollout", "nav_branches"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--tree", t])
    elif a.tree == "pretrain":
        gen_pretrain()
    elif a.tree == "pretrain_full":
        gen_pretrain_full()
    elif a.tree == "nav_full":
        gen_nav_full()
    elif a.tree == "nav_rollout":
        gen_nav_rollout()
    elif a.tree == "nav_branches":
        gen_nav_branches()
    else:
        gen_nav()
