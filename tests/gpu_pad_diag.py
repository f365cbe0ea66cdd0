"""Diagnostic (not a test): where do the padded and unpadded prepared forwards first differ in a 16-bit mode?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import synth
from tests.test_gpu_models import _pretrain_model
from vln_goat_b200 import batching, runtime
from vln_goat_b200.modules import extend_neg_masks

model = _pretrain_model()
batch = synth.pretrain_batch(B=3, L=24, seed=5)
pad = batching.PadSpec(S=16, G=16, NM=32, K=8, KF=8)
dt = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[sys.argv[1] if len(sys.argv) > 1 else "fp16"]


def stages(P):
    bert = model.bert
    out = {}
    txt, txt_masks, views, fused = bert.encode_prepared(P)
    out["txt"], out["views"], out["fused"] = txt, views, fused
    gmap_in, gmap_masks = bert.gmap_inputs_prepared(P, views, fused)
    out["gmap_in"] = gmap_in
    ge = bert.global_encoder
    sp = ge.sprels(P["gmap_pair_dists"])
    out["sprels"] = sp
    x = gmap_in
    for i, layer in enumerate(ge.encoder.crossattention):
        a = layer.attention.run(x, extend_neg_masks(gmap_masks), bias=sp)
        out["g%d_self" % i] = a.tensor()
        c = layer.crossattention.run(a, None, txt, extend_neg_masks(txt_masks))
        out["g%d_cross" % i] = c.tensor()
        from vln_goat_b200.modules import ffn_block
        x = ffn_block(layer.intermediate, layer.output, c, False).tensor()
        out["g%d_ffn" % i] = x
    vp_in, vp_masks = bert.vp_inputs_prepared(P, views)
    out["vp_in"] = vp_in
    out["vp"] = bert.local_encoder.encoder(vp_in, vp_masks, txt, txt_masks)
    out["ghead"] = model.global_sap_head(x)
    return out


with runtime.compute(dt), torch.no_grad():
    P0 = synth.batch_to(batching.prepare_pretrain(batch, "sap", pad=None), "cuda")
    P1 = synth.batch_to(batching.prepare_pretrain(batch, "sap", pad=pad), "cuda")
    a, b = stages(P0), stages(P1)
    S, G = P0["view_fts"].shape[0], P0["gmap_step_ids"].shape[1]
    for k in a:
        x, y = a[k], b[k]
        if k in ("views", "fused"):
            y = y[:S]
        elif k == "sprels":
            y = y[:, :, :G, :G]
        elif y.dim() == 3 and y.shape[1] != x.shape[1]:
            y = y[:, :x.shape[1]]
        d = (x.float() - y.float()).abs().max().item()
        print("%-10s max|diff| %.3e  (max|x| %.3e)" % (k, d, x.float().abs().max().item()))
