"""GPU parity of the head kernels (heads.cu) and of the two full models against the golden fixtures produced by
the unmodified reference (tests/golden/make_golden.py --tree pretrain_full / nav_full) on the same seeded
synthetic batches (tests/synth.py) and seeded parameters.

Tolerances (BASELINE.json north_star): 1e-5 (fp32 mode) / 1e-3 class (16-bit operand modes) on action logits, CFP
embeddings, MLM scores and losses, relative to max(1, max|ref|).  The full models stack ~15 blocks, so fp32 mode
is held to 5e-5 and bf16 to 3e-2 (bf16 has 8 mantissa bits; the per-block bound is in test_gpu_blocks.py).
"""
from collections import defaultdict

import pytest
import torch
import torch.nn.functional as F

from oracle import goat_oracle as O
from tests import synth
from tests.helpers import assert_digests, golden, maxerr

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    ref = ref.float()
    fin = torch.isfinite(ref)
    assert torch.equal(torch.isfinite(got.cpu()), fin), "-inf pattern differs"
    if fin.sum() == 0:
        return 0.0
    return (got.cpu().float()[fin] - ref[fin]).abs().max().item() / max(1.0, ref[fin].abs().max().item())


@pytest.fixture(autouse=True)
def _lib_loaded():
    from vln_goat_b200 import _lib
    assert _lib.lib().goat_device_supported() == 1, "needs an sm_100 device"


# ----------------------------------------------------------------------------------------------
# head kernels vs torch autograd on the CPU restatement
# ----------------------------------------------------------------------------------------------
def _leaf(t):
    return t.clone().float().cuda().requires_grad_(True)


@pytest.mark.parametrize("B,N", [(4, 11), (3, 37), (2, 80), (1, 1)])
def test_attn_pool_both_modes(B, N):
    from vln_goat_b200 import functional as Fn
    g = torch.Generator().manual_seed(B * 100 + N)
    x = torch.randn(B, N, 768, generator=g)
    w = 0.05 * torch.randn(1, 768, generator=g)
    b = 0.05 * torch.randn(1, generator=g)
    aw = (torch.rand(768, 1, generator=g) - 0.5) * 0.2
    wo = torch.randn(B, 768, generator=g)
    for mode in (0, 1):
        xr = x.clone().requires_grad_(True)
        wr = (w if mode == 0 else aw).clone().requires_grad_(True)
        br = b.clone().requires_grad_(True)
        ref = O.pano_fuse(xr, wr, br) if mode == 0 else O.cfp_pool(xr, wr)
        (ref * wo).sum().backward()
        xc, wc, bc = _leaf(x), _leaf(w if mode == 0 else aw), _leaf(b)
        out = Fn.AttnPoolFn.apply(xc, wc, bc if mode == 0 else None, mode)
        (out * wo.cuda()).sum().backward()
        assert maxerr(out, ref) < 2e-5
        assert maxerr(xc.grad, xr.grad) < 2e-5
        assert maxerr(wc.grad, wr.grad) < 2e-4
        if mode == 0:
            assert maxerr(bc.grad, br.grad) < 2e-4


def test_attn_pool_matches_golden_heads():
    from vln_goat_b200 import functional as Fn
    g = golden("heads")
    x = g["x"].cuda()
    fuse = Fn.AttnPoolFn.apply(x, g["fuse_w"].cuda(), g["fuse_b"].cuda(), 0)
    pool = Fn.AttnPoolFn.apply(x, g["pool_w"].cuda(), None, 1)
    assert maxerr(fuse, g["fuse_out"]) < 1e-5
    assert maxerr(pool, g["pool_out"]) < 1e-5
    from vln_goat_b200 import goat_blocks as G
    nce = G.infonce(pool, g["nce_y"].cuda(), 1.0)
    assert maxerr(nce, g["nce"]) < 1e-5


def test_door_gate_and_wsum():
    from vln_goat_b200 import functional as Fn
    g = torch.Generator().manual_seed(3)
    aug, ori = torch.randn(5, 13, 768, generator=g), torch.randn(5, 13, 768, generator=g)
    wa, wo_ = 0.05 * torch.randn(1, 768, generator=g), 0.05 * torch.randn(1, 768, generator=g)
    ba, bo = torch.randn(1, generator=g), torch.randn(1, generator=g)
    w = torch.randn(5, 13, 768, generator=g)
    refs = [t.clone().requires_grad_(True) for t in (aug, ori, wa, ba, wo_, bo)]
    ref = O.door_gate(*refs)
    (ref * w).sum().backward()
    cs = [_leaf(t) for t in (aug, ori, wa, ba, wo_, bo)]
    out = Fn.DoorGateFn.apply(*cs)
    (out * w.cuda()).sum().backward()
    assert maxerr(out, ref) < 1e-5
    for c, r, tol in zip(cs, refs, (1e-5, 1e-5, 2e-4, 2e-4, 2e-4, 2e-4)):
        assert maxerr(c.grad, r.grad) < tol
    # p(z)-weighted dictionary sum, float64 probabilities as the agent passes them (M/r2r/agent.py:53-56)
    z = torch.randn(4, 50, 768, generator=g)
    pz = torch.rand(4, 50, 1, generator=g, dtype=torch.float64)
    zr = z.clone().requires_grad_(True)
    ref = torch.sum(zr * pz.to(torch.float32), 1)
    (ref * w[:4, 0]).sum().backward()
    zc = _leaf(z)
    out = Fn.WSumFn.apply(zc, pz.cuda())
    (out * w[:4, 0].cuda()).sum().backward()
    assert maxerr(out, ref) < 1e-5 and maxerr(zc.grad, zr.grad) < 1e-6


@pytest.mark.parametrize("M,N", [(7, 13), (64, 64), (5, 50265)])
def test_xent_rows_inf_ignore_and_transposed(M, N):
    from vln_goat_b200 import functional as Fn
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, N, generator=g) * 3
    lab = torch.randint(0, N, (M,), generator=g)
    if N < 100:
        x[:, 1] = -float("inf")          # a masked action everywhere
        x[0, 2:5] = -float("inf")
        lab[lab == 1] = 0
        lab[0] = 0
        lab[M - 1] = -100                # ignored sample
    w = torch.rand(M, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = F.cross_entropy(xr, lab, reduction="none", ignore_index=-100)
    (ref * w).sum().backward()
    xc = _leaf(x)
    out = Fn.XentFn.apply(xc, lab.cuda(), -100)
    (out * w.cuda()).sum().backward()
    assert maxerr(out, ref) < 2e-5
    assert maxerr(xc.grad, xr.grad) < 2e-6
    if M == N:                            # transposed view, as the symmetric InfoNCE uses it
        x = torch.randn(M, N, generator=g) * 3      # no masked column: its transpose would be an all -inf row
        xr2 = x.clone().requires_grad_(True)
        tgt = torch.arange(M)
        ref2 = F.cross_entropy(xr2.t(), tgt, reduction="none")
        (ref2 * w).sum().backward()
        xc2 = _leaf(x)
        out2 = Fn.XentFn.apply(xc2.t(), tgt.cuda(), -100)
        (out2 * w.cuda()).sum().backward()
        assert maxerr(out2, ref2) < 2e-5 and maxerr(xc2.grad, xr2.grad) < 2e-6


def test_segment_reduce_and_embed():
    from vln_goat_b200 import functional as Fn
    g = torch.Generator().manual_seed(9)
    src = torch.randn(40, 768, generator=g)
    idx = torch.randint(-1, 40, (17, 5), generator=g).to(torch.int32)
    idx[3] = -1
    w = torch.randn(17, 768, generator=g)
    for mean in (True, False):
        sr = src.clone().requires_grad_(True)
        rows = []
        for r in range(17):
            sel = [int(i) for i in idx[r] if i >= 0]
            if not sel:
                rows.append(torch.zeros(768))
            else:
                t = sr[sel].sum(0)
                rows.append(t / len(sel) if mean else t)
        ref = torch.stack(rows, 0)
        (ref * w).sum().backward()
        sc = _leaf(src)
        out = Fn.SegmentReduceFn.apply(sc, idx.cuda(), mean)
        (out * w.cuda()).sum().backward()
        assert maxerr(out, ref) < 1e-5 and maxerr(sc.grad, sr.grad) < 1e-5
    # embeddings
    ids = torch.randint(0, 300, (3, 11), generator=g)
    word, pos, typ = (torch.randn(300, 768, generator=g), torch.randn(20, 768, generator=g), torch.randn(1, 768, generator=g))
    wr, pr, tr = (t.clone().requires_grad_(True) for t in (word, pos, typ))
    ref = wr[ids] + pr[torch.arange(11)][None] + tr[0]
    wo = torch.randn(3, 11, 768, generator=g)
    (ref * wo).sum().backward()
    wc, pc, tc = _leaf(word), _leaf(pos), _leaf(typ)
    out = Fn.EmbedFn.apply(ids.cuda(), wc, pc, tc)
    (out.view(3, 11, 768) * wo.cuda()).sum().backward()
    assert maxerr(out.view(3, 11, 768), ref) < 1e-6
    assert maxerr(wc.grad, wr.grad) < 1e-5 and maxerr(pc.grad, pr.grad) < 1e-5 and maxerr(tc.grad, tr.grad) < 1e-4


# ----------------------------------------------------------------------------------------------
# full models vs the reference fixtures
# ----------------------------------------------------------------------------------------------
FULL_TOL = {torch.float32: 5e-5, torch.bfloat16: 3e-2}
FULL_DIG = {torch.float32: 1e-4, torch.bfloat16: 5e-2}


def _seed_model(model, seed):
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(O.seeded_params(shapes, seed=seed), strict=True)
    return model


def _grads(module):
    return {k: v.grad for k, v in module.named_parameters() if v.grad is not None}


def _sub(gold, prefix):
    return {k[len(prefix):]: v for k, v in gold.items() if k.startswith(prefix)}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pretrain_full_mlm_sap_cfp(dtype):
    from vln_goat_b200 import pretrain_model, runtime
    from vln_goat_b200.config import GoatConfig
    gold = golden("pretrain_full")
    cfg = GoatConfig(pretrain_tasks=("mlm", "sap", "cfp"))
    model = pretrain_model.GlocalTextPathCMTPreTraining(cfg)
    assert list(model.state_dict().keys()) == str(gold["state_dict_keys"]).split("\n")
    _seed_model(model, 20)
    model.tie_weights()
    model = model.cuda().eval()
    batch = synth.batch_to(synth.pretrain_batch(B=3, L=24, seed=5), "cuda")
    tol, dig = FULL_TOL[dtype], FULL_DIG[dtype]
    with runtime.compute(dtype):
        # MLM
        model.zero_grad()
        scores = model(batch, "mlm", compute_loss=False)
        loss = model(batch, "mlm", compute_loss=True)
        loss.sum().backward()
        assert _rel(scores[:, :64], gold["mlm_scores_head"]) < tol
        assert _rel(loss, gold["mlm_loss"]) < tol
        assert_digests(_sub(gold, "mlm."), _grads(model), rtol=dig, atol=1e-3 if dtype == torch.float32 else 1e-1,
                       key_bias_atol=None if dtype == torch.float32 else 64.0)
        # SAP
        model.zero_grad()
        gl, ll, fl, _, _ = model(batch, "sap", compute_loss=False)
        loss = model(batch, "sap", compute_loss=True)
        loss.sum().backward()
        assert _rel(gl, gold["sap_global_logits"]) < tol
        assert _rel(ll, gold["sap_local_logits"]) < tol
        assert _rel(fl, gold["sap_fused_logits"]) < tol
        assert _rel(loss, gold["sap_loss"]) < tol
        assert_digests(_sub(gold, "sap."), _grads(model), rtol=dig, atol=1e-3 if dtype == torch.float32 else 1e-1,
                       key_bias_atol=None if dtype == torch.float32 else 64.0)
        # CFP
        model.zero_grad()
        go, vo, fo, to = model(batch, "cfp", compute_loss=False)
        loss = model(batch, "cfp", compute_loss=True)
        loss.sum().backward()
        for got, name in ((go, "cfp_gmap"), (vo, "cfp_vp"), (fo, "cfp_fused"), (to, "cfp_txt")):
            assert _rel(got, gold[name]) < tol, name
        assert _rel(loss, gold["cfp_loss"]) < tol
        assert_digests(_sub(gold, "cfp."), _grads(model), rtol=dig, atol=1e-3 if dtype == torch.float32 else 1e-1,
                       key_bias_atol=None if dtype == torch.float32 else 64.0)


def _nav_cfg():
    from vln_goat_b200.config import GoatConfig
    return GoatConfig(layer_norm_eps=1e-5, pad_token_id=1, dataset="r2r", mode="train", obj_feat_size=0, feat_dropout=0.4,
                      do_back_img=True, do_back_txt=True, do_front_img=True, do_front_his=True, do_front_txt=True,
                      do_back_txt_type="type_2", do_back_img_type="type_1", do_add_method="door",
                      use_lang2visn_attn=False, fix_lang_embedding=False, fix_pano_embedding=False, fix_local_branch=False)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_nav_full_language_panorama_navigation(dtype):
    """BACL (text type_2 + image type_1) and FACL (text, vp, gmap) on: the three per-step modes chained."""
    from vln_goat_b200 import nav_model, runtime
    gold = golden("nav_full")
    model = nav_model.GlocalTextPathNavCMT(_nav_cfg())
    assert list(model.state_dict().keys()) == str(gold["state_dict_keys"]).split("\n")
    model = _seed_model(model, 21).cuda().eval()
    lang, pano, nav = synth.nav_inputs(B=3, L=20, seed=6)
    dd = lambda d: defaultdict(lambda: None, synth.batch_to(d, "cuda"))
    tol, dig = FULL_TOL[dtype], FULL_DIG[dtype]
    with runtime.compute(dtype):
        txt = model("language", dd(lang))
        pe, pm, pf = model("panorama", dd(pano))
        navb = dict(nav)
        mem = navb.pop("mem_embeds").cuda()
        navb["txt_embeds"] = txt
        navb["vp_img_embeds"] = torch.cat([torch.zeros_like(pe[:, :1]), mem.unsqueeze(1), pe], 1)
        outs = model("navigation", dd(navb))
        loss = F.cross_entropy(outs["fused_logits"], gold["target"].cuda(), reduction="sum") + \
            (outs["cls_embeds"] * gold["w_cls"].cuda()).sum() + (pf * gold["w_pf"].cuda()).sum()
        loss.backward()
    assert _rel(txt, gold["txt_embeds"]) < tol
    assert _rel(pe, gold["pano_embeds"]) < tol
    assert torch.equal(pm.cpu(), gold["pano_masks"].bool())
    assert _rel(pf, gold["pano_fused"]) < tol
    for k in ("global_logits", "local_logits", "fused_logits", "cls_embeds", "gmap_embeds", "vp_embeds"):
        assert _rel(outs[k], gold[k]) < tol, k
    assert abs(loss.item() - gold["loss"].item()) < tol * max(1.0, abs(gold["loss"].item())) * 10
    assert_digests(gold, _grads(model), rtol=dig, atol=1e-3 if dtype == torch.float32 else 1e-1,
                   key_bias_atol=None if dtype == torch.float32 else 64.0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("kv_cache", [True, False])
def test_nav_rollout_three_steps_vs_reference(dtype, kv_cache):
    """a teacher-forced rollout (language once, then panorama + navigation for three steps,
    the [MEM] token chaining cls_embeds from step to step, cross-entropy summed, ONE backward) against the fixture the
    UNMODIFIED reference produced for the same inputs (tests/golden/make_golden.py --tree nav_rollout).  With the
    rollout-level K|V cache on (the default) the six instruction projections and the two FACL prototype projections are
    computed at step 0 and reused at steps 1 and 2."""
    from vln_goat_b200 import nav_model, runtime
    gold = golden("nav_rollout")
    cfg = _nav_cfg()
    cfg.kv_cache = kv_cache
    model = _seed_model(nav_model.GlocalTextPathNavCMT(cfg), 23).cuda().eval()
    lang, per_step, targets = synth.rollout_inputs()
    # device copies made once; the FACL prototypes are the SAME tensors at every step, as in the agent
    fv, fg = per_step[0][1]["front_vp_feats"].cuda(), per_step[0][1]["front_gmap_feats"].cuda()
    dev_steps = []
    for pano, nav in per_step:
        n = synth.batch_to(nav, "cuda")
        n["front_vp_feats"], n["front_gmap_feats"] = fv, fg
        dev_steps.append((synth.batch_to(pano, "cuda"), n))
    tol, dig = FULL_TOL[dtype], FULL_DIG[dtype]
    with runtime.compute(dtype):
        loss, logits, clss, txt = synth.run_rollout(model, synth.batch_to(lang, "cuda"), dev_steps, targets)
        loss.backward()
    assert _rel(txt, gold["txt_embeds"]) < tol
    for t in range(3):
        assert _rel(logits[t], gold["fused_logits_%d" % t]) < tol * (1 + t), t      # errors chain through [MEM]
        assert _rel(clss[t], gold["cls_embeds_%d" % t]) < tol * (1 + t), t
    assert abs(loss.item() - gold["loss"].item()) < tol * max(1.0, abs(gold["loss"].item())) * 10
    assert_digests(gold, _grads(model), rtol=dig, atol=1e-3 if dtype == torch.float32 else 1e-1,
                   key_bias_atol=None if dtype == torch.float32 else 64.0)
    if kv_cache:
        assert (model._kv_cache.misses, model._kv_cache.hits) == (8, 16)
    else:
        assert model._kv_cache is None


def test_vlnbert_wrapper_feature_dropout_and_modes():
    """VLNBert(mode, batch): panorama applies the environment feature dropout unless already_dropout (model.py:28-32)."""
    from types import SimpleNamespace
    from vln_goat_b200 import nav_model, runtime
    args = SimpleNamespace(feat_dropout=0.4)
    torch.manual_seed(0)
    wrap = nav_model.VLNBert(args, config=_nav_cfg()).cuda()
    lang, pano, _ = synth.nav_inputs(B=2, L=12, seed=1)
    with runtime.compute(torch.bfloat16):
        wrap.eval()
        a = wrap("panorama", synth.batch_to(dict(pano, already_dropout=False), "cuda"))[0]
        b = wrap("panorama", synth.batch_to(pano, "cuda"))[0]
        assert torch.equal(a, b)                       # eval: dropout is the identity
        wrap.train()
        c = wrap("panorama", synth.batch_to(dict(pano, already_dropout=False), "cuda"))[0]
        assert torch.isfinite(c).all() and not torch.equal(c, b)
        t = wrap("language", synth.batch_to(lang, "cuda"))
        assert t.shape == (2, 12, 768) and torch.isfinite(t).all()
