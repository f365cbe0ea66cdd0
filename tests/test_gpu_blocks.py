"""GPU parity: the drop-in blocks (vln_goat_b200.modules -> C ABI -> sm_100a kernels) against the golden
fixtures produced by the unmodified reference (tests/golden/make_golden.py) and against the CPU oracle.

Tolerances (BASELINE.json north_star): 1e-5 in fp32 mode, 1e-3 in fp16 mode, both relative to the
output scale max(1, max|ref|); bf16 operands carry 8x the fp16 rounding step, so bf16 is held to 8e-3.
Parameter gradients are compared through the 5-number digests stored in the fixtures.
"""
import pytest
import torch

from oracle import goat_oracle as O
from tests.helpers import assert_digests, golden, maxerr

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.float16: 1e-3, torch.bfloat16: 8e-3}
DIG = {torch.float32: 2e-5, torch.float16: 2e-3, torch.bfloat16: 1.6e-2}
KB = {torch.float32: None, torch.float16: 2.0, torch.bfloat16: 16.0}   # abs-sum bound over 768 noise-only elements
DTYPES = [torch.float32, torch.float16, torch.bfloat16]


def _rel(got, ref):
    return maxerr(got, ref) / max(1.0, ref.abs().max().item())


def _cuda_leaf(t):
    return t.float().cuda().requires_grad_(True)


def _load(module, params):
    missing, unexpected = module.load_state_dict(params, strict=False)
    assert not unexpected, unexpected
    return module.cuda().eval()


def _grads(module):
    return {k: v.grad for k, v in module.named_parameters() if v.grad is not None}


@pytest.fixture(autouse=True)
def _lib_loaded():
    from vln_goat_b200 import _lib
    assert _lib.lib().goat_device_supported() == 1, "needs an sm_100 device"


@pytest.mark.parametrize("dtype", DTYPES)
def test_c1_cross_layer(dtype):
    """BASELINE.json configs[0]: one BertCrossLayer, pano [2,36,768] x text [2,80,768], lengths {80,57}."""
    from vln_goat_b200 import modules as M, runtime
    from vln_goat_b200.config import GoatConfig
    g = golden("c1_cross_layer")
    layer = _load(M.BertCrossLayer(GoatConfig()), O.seeded_params(O.cross_layer_shapes(), seed=1))
    q, kv = _cuda_leaf(g["q"]), _cuda_leaf(g["kv"])
    kvm = M.extend_neg_masks(M.gen_seq_masks(g["txt_lens"].cuda(), 80))
    qm = torch.zeros(2, 1, 1, 36, device="cuda")
    with runtime.compute(dtype):
        out = layer(q, kv, attention_mask=qm, encoder_attention_mask=kvm)[0]
        (out * g["w_out"].cuda()).sum().backward()
        att = layer.crossattention(g["q"].cuda(), None, None, g["kv"].cuda(), kvm)[0]
    assert _rel(out, g["out"]) < TOL[dtype]
    assert _rel(att, g["cross_attn_out"]) < TOL[dtype]
    assert _rel(q.grad, g["dq"]) < TOL[dtype]
    assert _rel(kv.grad, g["dkv"]) < TOL[dtype]
    grads = _grads(layer)
    assert_digests({k: v for k, v in g.items() if "lang_" not in k}, grads, rtol=DIG[dtype], key_bias_atol=KB[dtype])


@pytest.mark.parametrize("dtype", DTYPES)
def test_xenc_sprels(dtype):
    """CrossmodalEncoder(3) with the graph_sprels bias and ragged lengths; d(sprels) flows back."""
    from vln_goat_b200 import modules as M, runtime
    from vln_goat_b200.config import GoatConfig
    g = golden("xenc_sprels")
    shapes = {}
    for i in range(3):
        shapes.update(O.cross_layer_shapes("crossattention.%d." % i))
    enc = _load(M.CrossmodalEncoder(GoatConfig()), O.seeded_params(shapes, seed=2))
    gm, tx, sp = _cuda_leaf(g["gmap"]), _cuda_leaf(g["txt"]), _cuda_leaf(g["sprels"])
    with runtime.compute(dtype):
        out = enc(gm, M.gen_seq_masks(g["gmap_lens"].cuda(), 12), tx, M.gen_seq_masks(g["txt_lens"].cuda(), 40),
                  graph_sprels=sp)
        (out * g["w_out"].cuda()).sum().backward()
    assert _rel(out, g["out"]) < TOL[dtype]
    assert _rel(gm.grad, g["dgmap"]) < TOL[dtype]
    assert _rel(tx.grad, g["dtxt"]) < TOL[dtype]
    assert _rel(sp.grad, g["dsprels"]) < TOL[dtype]
    assert_digests({k: v for k, v in g.items() if "lang_" not in k}, _grads(enc), rtol=DIG[dtype], key_bias_atol=KB[dtype])


@pytest.mark.parametrize("dtype", DTYPES)
def test_lang_encoder(dtype):
    from vln_goat_b200 import modules as M, runtime
    from vln_goat_b200.config import GoatConfig
    g = golden("lang_encoder")
    shapes = {}
    for i in range(6):
        shapes.update(O.roberta_layer_shapes("layer.%d." % i))
    le = _load(M.LanguageEncoder(GoatConfig()), O.seeded_params(shapes, seed=3))
    x = _cuda_leaf(g["x"])
    with runtime.compute(dtype):
        out = le(x, M.gen_seq_masks(g["txt_lens"].cuda(), 80))
        (out * g["w_out"].cuda()).sum().backward()
    assert _rel(out, g["out"]) < 2 * TOL[dtype]
    assert _rel(x.grad, g["dx"]) < 2 * TOL[dtype]
    assert_digests(g, _grads(le), rtol=DIG[dtype], key_bias_atol=KB[dtype])


@pytest.mark.parametrize("dtype", DTYPES)
def test_pano_encoder(dtype):
    """2 pre-LN layers + final LN, -inf key padding {36, 29}."""
    from vln_goat_b200 import modules as M, runtime
    from vln_goat_b200.config import GoatConfig
    g = golden("pano_encoder")
    pe = _load(M.create_transformer_encoder(GoatConfig(), 2, norm=True), O.seeded_params(O.pano_encoder_shapes(), seed=4))
    x = _cuda_leaf(g["x"])
    with runtime.compute(dtype):
        out = pe(x, src_key_padding_mask=M.gen_seq_masks(g["view_lens"].cuda(), 36).logical_not())
        (out * g["w_out"].cuda()).sum().backward()
    assert _rel(out, g["out"]) < TOL[dtype]
    assert _rel(x.grad, g["dx"]) < 2 * TOL[dtype]
    assert_digests(g, _grads(pe), rtol=DIG[dtype], key_bias_atol=KB[dtype])


def test_state_dict_keys_match_reference_fixture():
    """Every parameter name the reference BertCrossLayer / pano encoder owns exists here with the same shape."""
    from vln_goat_b200 import modules as M
    from vln_goat_b200.config import GoatConfig
    layer = M.BertCrossLayer(GoatConfig())
    sd = layer.state_dict()
    for k, shp in O.cross_layer_shapes().items():
        assert tuple(sd[k].shape) == tuple(shp), k
    pe = M.create_transformer_encoder(GoatConfig(), 2, norm=True)
    sd = pe.state_dict()
    for k, shp in O.pano_encoder_shapes().items():
        assert tuple(sd[k].shape) == tuple(shp), k


@pytest.mark.parametrize("B,Nq,Nk", [(1, 1, 1), (2, 3, 24), (2, 37, 80), (3, 38, 35), (2, 60, 200), (1, 128, 512),
                                     (16, 37, 50), (2, 38, 39)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_attention_block_shape_sweep(B, Nq, Nk, dtype):
    """BertAttention (cross) vs the CPU oracle over the shape sweep of SURVEY.md Appendix B.4, ragged key lengths."""
    from vln_goat_b200 import modules as M, runtime
    from vln_goat_b200.config import GoatConfig
    P = O.seeded_params(O.attn_shapes(""), seed=11)
    att = _load(M.BertAttention(GoatConfig()), P)
    g = torch.Generator().manual_seed(B * 1000 + Nq * 10 + Nk)
    x = torch.randn(B, Nq, 768, generator=g)
    enc = torch.randn(B, Nk, 768, generator=g)
    lens = torch.randint(1, Nk + 1, (B,), generator=g)
    lens[0] = Nk
    km = O.extend_neg_masks(O.gen_seq_masks(lens, Nk))
    w = torch.randn(B, Nq, 768, generator=g)
    xr, er = x.clone().requires_grad_(True), enc.clone().requires_grad_(True)
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    ref = O.bert_attention(Pr, "", xr, None, er, km)
    (ref * w).sum().backward()
    xc, ec = _cuda_leaf(x), _cuda_leaf(enc)
    with runtime.compute(dtype):
        out = att(xc, None, None, ec, km.cuda())[0]
        (out * w.cuda()).sum().backward()
    assert _rel(out, ref.detach()) < TOL[dtype]
    assert _rel(xc.grad, xr.grad) < TOL[dtype]
    assert _rel(ec.grad, er.grad) < TOL[dtype]
    for k, v in att.named_parameters():
        assert _rel(v.grad, Pr[k].grad) < 4 * TOL[dtype], k


@pytest.mark.parametrize("dtype", DTYPES)
def test_kv_cache_rollout_matches_uncached(dtype):
    """Rollout-level K|V projection cache: three navigation-like steps attend to the SAME instruction
    embeddings; with the cache the text K|V projection of each cross layer is computed once and its backward runs once
    on the fp32-summed gradient.  Outputs are bit-identical, gradients agree to accumulation-order noise; the golden
    xenc_sprels fixture (reference-generated) still matches when its single step runs through the cached path."""
    from vln_goat_b200 import modules as M, runtime
    from vln_goat_b200.config import GoatConfig
    g = golden("xenc_sprels")
    shapes = {}
    for i in range(3):
        shapes.update(O.cross_layer_shapes("crossattention.%d." % i))
    params = O.seeded_params(shapes, seed=2)
    gm_len, tx_len = g["gmap_lens"].cuda(), g["txt_lens"].cuda()
    torch.manual_seed(11)
    steps = [g["gmap"].float()] + [torch.randn_like(g["gmap"].float()) for _ in range(2)]

    def run(cache):
        enc = _load(M.CrossmodalEncoder(GoatConfig()), params)
        tx = _cuda_leaf(g["txt"])
        qs = [_cuda_leaf(q) for q in steps]
        outs = []
        with runtime.compute(dtype), M.kv_cache_scope(cache):
            loss = 0.0
            for q in qs:
                out = enc(q, M.gen_seq_masks(gm_len, 12), tx, M.gen_seq_masks(tx_len, 40))
                outs.append(out)
                loss = loss + (out * g["w_out"].cuda()).sum()
            loss.backward()
        return outs, tx.grad, [q.grad for q in qs], _grads(enc)

    o0, dtx0, dq0, gr0 = run(None)
    cache = M.KVCache()
    o1, dtx1, dq1, gr1 = run(cache)
    assert (cache.misses, cache.hits) == (3, 6)
    for a, b in zip(o0, o1):
        assert torch.equal(a, b)
    tol = DIG[dtype]
    assert _rel(dtx1, dtx0) < tol
    for a, b in zip(dq0, dq1):
        assert _rel(b, a) < tol
    assert set(gr0) == set(gr1)
    for k in gr0:
        if ".key.bias" in k:
            continue                                   # mathematically zero: pure rounding noise in both runs
        assert _rel(gr1[k], gr0[k]) < tol, k
    # single step through the cached path against the reference-generated fixture
    enc = _load(M.CrossmodalEncoder(GoatConfig()), params)
    gm, tx, sp = _cuda_leaf(g["gmap"]), _cuda_leaf(g["txt"]), _cuda_leaf(g["sprels"])
    with runtime.compute(dtype), M.kv_cache_scope(M.KVCache()):
        out = enc(gm, M.gen_seq_masks(gm_len, 12), tx, M.gen_seq_masks(tx_len, 40), graph_sprels=sp)
        (out * g["w_out"].cuda()).sum().backward()
    assert _rel(out, g["out"]) < TOL[dtype]
    assert _rel(tx.grad, g["dtxt"]) < TOL[dtype]
    assert_digests({k: v for k, v in g.items() if "lang_" not in k}, _grads(enc), rtol=DIG[dtype], key_bias_atol=KB[dtype])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_c2_full_size_batch_matches_oracle_on_a_slice(dtype):
    """BASELINE.json configs[1] at FULL size (batch 64, 80 tokens, [stop]+36 views, 6 + 3 layers) on the GPU.  Samples
    are independent on this path, so samples {0, 1, 63} of the batch-64 outputs and the gradient of a loss that only
    touches those samples must equal the CPU oracle run on just those three samples (seconds on CPU)."""
    from vln_goat_b200 import runtime, workloads
    from vln_goat_b200.config import GoatConfig
    B, L, Nq, H = 64, 80, 37, 768
    gen = torch.Generator().manual_seed(5)
    txt = torch.randn(B, L, H, generator=gen)
    vp = torch.randn(B, Nq, H, generator=gen)
    vp[:, 0] = 0.0
    lens = torch.randint(L // 2, L + 1, (B,), generator=gen)
    lens[0] = L
    tm = O.gen_seq_masks(lens, L)
    vm = torch.ones(B, Nq, dtype=torch.bool)
    pick = torch.tensor([0, 1, 63])
    params = O.seeded_params(O.c2_shapes(), seed=0)
    P = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    torch.set_num_threads(max(1, (torch.get_num_threads())))
    t_ref, v_ref = O.c2_forward(P, txt[pick], tm[pick], vp[pick], vm[pick])
    O.c2_loss(t_ref, v_ref).backward()
    model = workloads.C2CrossEncoder(GoatConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0))
    model.load_state_dict(params, strict=False)
    model = model.cuda().train()
    with runtime.compute(dtype):
        t, v = model(txt.cuda(), tm.cuda(), vp.cuda(), vm.cuda())
        loss = 0.5 * (t[pick.cuda()] ** 2).mean() + 0.5 * (v[pick.cuda()] ** 2).mean()
        loss.backward()
    tol = TOL[dtype] * (2 if dtype == torch.float32 else 1)      # 9 layers deep
    assert _rel(t[pick.cuda()], t_ref.detach()) < tol
    assert _rel(v[pick.cuda()], v_ref.detach()) < tol
    assert abs(loss.item() - O.c2_loss(t_ref, v_ref).item()) < 1e-4 * (1 if dtype == torch.float32 else 50)
    got = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    for k in ("lang_encoder.layer.0.attention.self.query.weight", "lang_encoder.layer.5.output.dense.weight",
              "local_encoder.encoder.crossattention.2.output.dense.weight",
              "local_encoder.encoder.crossattention.0.crossattention.self.value.weight",
              "local_encoder.encoder.crossattention.1.intermediate.dense.bias"):
        ref = P[k].grad
        err = (got[k].cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)
        assert err < (1e-3 if dtype == torch.float32 else 5e-2), (k, err)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_c5_long_instruction_cross_layer(dtype):
    """BASELINE.json configs[4] shape per GPU: RxR-length text (512 tokens, ragged) x [stop]+36 views, batch 4 -- one
    BertCrossLayer forward + backward against the CPU oracle (keys beyond 128 run the chunked tcgen05 attention)."""
    from vln_goat_b200 import modules as M, runtime
    from vln_goat_b200.config import GoatConfig
    B, Nq, Nk, H = 4, 37, 512, 768
    gen = torch.Generator().manual_seed(8)
    q0 = torch.randn(B, Nq, H, generator=gen)
    kv0 = torch.randn(B, Nk, H, generator=gen)
    w = torch.randn(B, Nq, H, generator=gen)
    lens = torch.tensor([512, 300, 129, 477])
    params = O.seeded_params(O.cross_layer_shapes(), seed=4)
    P = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    q64, kv64 = q0.clone().requires_grad_(True), kv0.clone().requires_grad_(True)
    qm_ref = O.extend_neg_masks(torch.ones(B, Nq, dtype=torch.bool))
    km_ref = O.extend_neg_masks(O.gen_seq_masks(lens, Nk))
    ref = O.cross_layer(P, "", q64, kv64, qm_ref, km_ref)
    (ref * w).sum().backward()
    layer = _load(M.BertCrossLayer(GoatConfig()), params)
    q, kv = _cuda_leaf(q0), _cuda_leaf(kv0)
    with runtime.compute(dtype):
        out = layer(q, kv, attention_mask=qm_ref.cuda(), encoder_attention_mask=km_ref.cuda())[0]
        (out * w.cuda()).sum().backward()
    assert _rel(out, ref.detach()) < TOL[dtype]
    assert _rel(q.grad, q64.grad) < TOL[dtype] * 2
    assert _rel(kv.grad, kv64.grad) < TOL[dtype] * 2
    got = _grads(layer)
    for k in ("crossattention.self.value.weight", "output.dense.weight", "attention.self.query.weight"):
        refg = P[k].grad
        err = (got[k].cpu() - refg).abs().max().item() / max(refg.abs().max().item(), 1e-12)
        assert err < DIG[dtype] * 2, (k, err)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_c5_long_instruction_text_layer(dtype):
    """BASELINE.json configs[4]: the TEXT side of the 512-token stress -- one RobertaLayer (self-attention over 512 / 300
    tokens, i.e. more than 128 QUERY rows, + FFN), forward + backward against the CPU oracle.  16-bit modes run the
    query-tiled tcgen05 attention kernels (attention_tc.cu), no SIMT fallback."""
    from vln_goat_b200 import modules as M, runtime
    from vln_goat_b200.config import GoatConfig
    B, L, H = 2, 512, 768
    gen = torch.Generator().manual_seed(18)
    x0 = torch.randn(B, L, H, generator=gen)
    w = torch.randn(B, L, H, generator=gen)
    lens = torch.tensor([512, 300])
    params = O.seeded_params(O.roberta_layer_shapes(), seed=14)
    P = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    x64 = x0.clone().requires_grad_(True)
    m_ref = O.extend_neg_masks(O.gen_seq_masks(lens, L))
    ref = O.roberta_layer(P, "", x64, m_ref)
    (ref * w).sum().backward()
    layer = _load(M.RobertaLayer(GoatConfig()), params)
    x = _cuda_leaf(x0)
    with runtime.compute(dtype):
        out = layer(x, attention_mask=m_ref.cuda())[0]
        (out * w.cuda()).sum().backward()
    assert _rel(out, ref.detach()) < TOL[dtype]
    assert _rel(x.grad, x64.grad) < TOL[dtype] * 2
    got = _grads(layer)
    for k in ("attention.self.query.weight", "attention.self.key.weight", "attention.self.value.weight", "output.dense.weight"):
        refg = P[k].grad
        err = (got[k].cpu() - refg).abs().max().item() / max(refg.abs().max().item(), 1e-12)
        assert err < DIG[dtype] * 2, (k, err)


# ----------------------------------------------------------------------------------------------
# causal-intervention blocks in isolation, against fixtures of the unmodified reference classes
# ----------------------------------------------------------------------------------------------
def _nav_cfg(**kw):
    from vln_goat_b200.config import GoatConfig
    base = dict(layer_norm_eps=1e-5, pad_token_id=1, dataset="r2r", mode="train", obj_feat_size=0, feat_dropout=0.4,
                do_back_img=True, do_back_txt=True, do_front_img=True, do_front_his=True, do_front_txt=True,
                do_back_txt_type="type_2", do_back_img_type="type_1", do_add_method="door", use_lang2visn_attn=False)
    base.update(kw)
    return GoatConfig(**base)


def _seeded(module, seed):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    module.load_state_dict(O.seeded_params(shapes, seed=seed), strict=True)
    return module.cuda().eval()


def _digest_close(t, ref, rtol):
    from tests.helpers import digest
    d = digest(t)
    scale = ref[1].abs().item() + 1e-30
    return bool(((d[:3] - ref.double()[:3]).abs() <= rtol * scale + 1e-4).all())


@pytest.mark.parametrize("dtype", DTYPES)
def test_front_door_encoder_block(dtype):
    """FACL FrontDoorEncoder (M/models/vilmodel_GOAT.py:526-554): self-attention + cross-attention onto the 24 prototypes,
    LN, door gate; output, input gradient and the 26 parameter-gradient digests of the reference."""
    from vln_goat_b200 import goat_blocks as G, runtime
    g = golden("front_door")
    fd = _seeded(G.FrontDoorEncoder(_nav_cfg()), 7)
    x = _cuda_leaf(g["x"])
    with runtime.compute(dtype):
        out = fd(x, g["proto"].cuda(), O.gen_seq_masks(g["lens"], 38).cuda())
        (out * g["w_out"].cuda()).sum().backward()
    assert _rel(out, g["out"]) < TOL[dtype]
    assert _rel(x.grad, g["dx"]) < TOL[dtype] * 2
    assert_digests(g, _grads(fd), rtol=DIG[dtype] * 2, key_bias_atol=KB[dtype])


@pytest.mark.parametrize("dtype", DTYPES)
def test_language_encoder_do_block(dtype):
    """LanguageEncoderDo (6 RobertaLayers + BACL text type_2 + FACL text, door merge; M/models/vilmodel_GOAT.py:55-162)."""
    from vln_goat_b200 import goat_blocks as G, runtime
    g = golden("lang_encoder_do")
    le = _seeded(G.LanguageEncoderDo(_nav_cfg()), 8)
    txt = _cuda_leaf(g["txt"])
    with runtime.compute(dtype):
        out = le(txt, O.gen_seq_masks(g["txt_lens"], 44).cuda(), g["z_direc"].cuda(), None, g["z_landm"].cuda(), None,
                 g["front_txt"].cuda())
        (out * g["w_out"].cuda()).sum().backward()
    assert _rel(out, g["out"]) < TOL[dtype] * 2           # 6 layers + two intervention stages deep
    assert _rel(txt.grad, g["dtxt"]) < TOL[dtype] * 4
    assert_digests(g, _grads(le), rtol=DIG[dtype] * 3, atol=1e-3 if dtype == torch.float32 else 3e-2, key_bias_atol=KB[dtype])


@pytest.mark.parametrize("dtype", DTYPES)
def test_bacl_image_type1_block(dtype):
    """CausalImageEmbeddings.back_door, type_1 (M/models/vilmodel_GOAT.py:661-667)."""
    from vln_goat_b200 import goat_blocks as G, runtime
    g = golden("bacl_image")
    ie = _seeded(G.CausalImageEmbeddings(_nav_cfg()), 9)
    with runtime.compute(dtype), torch.no_grad():
        out = ie.back_door(g["view"].cuda(), g["zf"].cuda(), g["pz"].cuda())
    assert _rel(out, g["out"]) < TOL[dtype]


BRANCH_TOL = {torch.float32: 2e-5, torch.float16: 2e-3}      # 6 text layers / 2 pano layers below the branch under test


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("tag,kw", [("t1", dict(do_back_txt_type="type_1")),
                                    ("t2add", dict(do_back_txt_type="type_2", do_add_method="add")),
                                    ("t2cat", dict(do_back_txt_type="type_2", do_add_method="concat"))])
def test_language_encoder_do_other_branches(tag, kw, dtype):
    """BACL text type_1 (p(z)-weighted dictionary sums) and the type_2 'add' / 'concat' merges, which the shipped scripts
    leave off: fixtures from the reference class under those configs (tests/golden/make_golden.py --tree nav_branches)."""
    from vln_goat_b200 import goat_blocks as G, runtime
    g = golden("nav_branches")
    le = _seeded(G.LanguageEncoderDo(_nav_cfg(**kw)), 41)
    txt = _cuda_leaf(g["txt"])
    with runtime.compute(dtype):
        out = le(txt, O.gen_seq_masks(g["txt_lens"], 28).cuda(), g["z_direc"].cuda(), g["pz_direc"].cuda(),
                 g["z_landm"].cuda(), g["pz_landm"].cuda(), g["front_txt"].cuda())
        (out * g["w_txt"].cuda()).sum().backward()
    assert _rel(out[:, ::4], g["txt_%s_out" % tag]) < BRANCH_TOL[dtype]
    assert _digest_close(out, g["txt_%s_dig" % tag], DIG[dtype] * 2)
    assert _digest_close(txt.grad, g["txt_%s_dx_dig" % tag], DIG[dtype] * 4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("tag,kw", [("door", dict(do_back_img_type="type_2", do_add_method="door")),
                                    ("add", dict(do_back_img_type="type_2", do_add_method="add")),
                                    ("cat", dict(do_back_img_type="type_2", do_add_method="concat"))])
def test_bacl_image_type2_branches_through_panorama(tag, kw, dtype):
    """BACL image type_2 (cross-attention onto the image dictionary, door / add / concat merge) through the whole per-step
    panorama path (embedding + location + 2-layer pano encoder + adaptive fusion), vs the reference's
    forward_panorama_do_per_step under those configs."""
    from vln_goat_b200 import goat_blocks as G, runtime
    g = golden("nav_branches")
    ie = _seeded(G.CausalImageEmbeddings(_nav_cfg(**kw)), 42)
    view = _cuda_leaf(g["view"])
    with runtime.compute(dtype):
        pe, pm, pf = ie.encode(view, g["loc"].cuda(), g["view_lens"].cuda(), g["z_img"].cuda(), g["pz_img"].cuda(),
                               loc_after_do=True)
        ((pe * g["w_view"].cuda()).sum() + pf.sum()).backward()
    assert _rel(pe[:, ::4], g["img_%s_out" % tag]) < BRANCH_TOL[dtype]
    assert _rel(pf, g["img_%s_fused" % tag]) < BRANCH_TOL[dtype]
    assert _digest_close(pe, g["img_%s_dig" % tag], DIG[dtype] * 2)
    assert _digest_close(view.grad, g["img_%s_dx_dig" % tag], DIG[dtype] * 4)


def test_no_cpu_fallback():
    from vln_goat_b200 import modules as M
    from vln_goat_b200.config import GoatConfig
    att = M.BertAttention(GoatConfig())
    with pytest.raises(RuntimeError):
        att(torch.randn(1, 4, 768))
