"""Scratch benchmark (not a test): BASELINE.json configs[3] shape -- a teacher-forced fine-tune rollout of 16 parallel
episodes x 15 steps with BACL + FACL on (language once, then panorama + navigation per step, cross-entropy summed over
the steps, ONE backward), eager mode, bf16.  Compares the rollout-level K|V projection cache on / off
and prints the per-rollout device time (CUDA events) and wall time."""
import os, sys, time
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tests import synth
from vln_goat_b200 import nav_model, runtime
from vln_goat_b200.config import GoatConfig

B, L, T, G = int(os.environ.get("B", 16)), int(os.environ.get("L", 80)), int(os.environ.get("T", 15)), 12
dev = "cuda"


def cfg(kv_cache):
    return GoatConfig(layer_norm_eps=1e-5, pad_token_id=1, dataset="r2r", mode="train", obj_feat_size=0, feat_dropout=0.4,
                      do_back_img=True, do_back_txt=True, do_front_img=True, do_front_his=True, do_front_txt=True,
                      do_back_txt_type="type_2", do_back_img_type="type_1", do_add_method="door", use_lang2visn_attn=False,
                      fix_lang_embedding=False, fix_pano_embedding=False, fix_local_branch=False, kv_cache=kv_cache)


def rollout(model, lang, steps, targets):
    dd = lambda d: defaultdict(lambda: None, d)
    txt = model("language", dd(lang))
    loss = 0.0
    mem = None
    for (pano, nav), tgt in zip(steps, targets):
        pe, pm, pf = model("panorama", dd(pano))
        navb = dict(nav)
        m0 = navb.pop("mem_embeds")
        mem_t = m0 if mem is None else mem
        navb["txt_embeds"] = txt
        navb["vp_img_embeds"] = torch.cat([torch.zeros_like(pe[:, :1]), mem_t.unsqueeze(1), pe], 1)
        outs = model("navigation", dd(navb))
        mem = outs["cls_embeds"]
        loss = loss + F.cross_entropy(outs["fused_logits"], tgt, reduction="sum")
    (loss / B).backward()
    return loss.detach()


def main():
    runtime.set_compute_dtype(torch.bfloat16)
    lang, _, _ = synth.nav_inputs(B=B, L=L, seed=1, G=G)
    lang = synth.batch_to(lang, dev)
    steps, targets = [], []
    for t in range(T):
        _, pano, nav = synth.nav_inputs(B=B, L=L, seed=100 + t, G=G)
        nav["txt_masks"] = lang["txt_masks"]
        if steps:       # the agent passes the SAME front-door prototypes at every step of a rollout (M/r2r/agent.py:567-583)
            nav["front_vp_feats"], nav["front_gmap_feats"] = steps[0][1]["front_vp_feats"], steps[0][1]["front_gmap_feats"]
        steps.append((synth.batch_to(pano, dev), synth.batch_to(nav, dev)))
        targets.append(torch.zeros(B, dtype=torch.int64, device=dev))      # [stop] is always a valid action
    res = {}
    for kv in (False, True):
        torch.manual_seed(0)
        model = nav_model.GlocalTextPathNavCMT(cfg(kv)).to(dev).train()
        for _ in range(2):
            model.zero_grad(set_to_none=True)
            l = rollout(model, lang, steps, targets)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            model.zero_grad(set_to_none=True)
            l = rollout(model, lang, steps, targets)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / reps * 1e3
        devms = e0.elapsed_time(e1) / reps
        c = model._kv_cache
        res[kv] = (devms, wall, float(l))
        print("kv_cache=%-5s rollout (B=%d, L=%d, %d steps, fwd+bwd): %.1f ms device-span, %.1f ms wall, loss %.4f%s"
              % (kv, B, L, T, devms, wall, float(l), "" if c is None else "  cache hits/misses %d/%d" % (c.hits, c.misses)),
              flush=True)
    print("speed-up from the K|V cache: %.2fx" % (res[False][0] / res[True][0]))


if __name__ == "__main__":
    if os.environ.get("PROFILE"):
        import cProfile, pstats
        cProfile.run("main()", "/tmp/rollout.prof")
        pstats.Stats("/tmp/rollout.prof").sort_stats("tottime").print_stats(35)
    else:
        main()
