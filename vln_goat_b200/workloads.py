"""The BASELINE.json workloads as small nn.Modules built from the drop-in blocks.

C2 ("full 9-layer GOAT cross-encoder fwd+bwd, batch=64"): the local branch of
GlocalTextPathCMT.forward -- LanguageEncoder (6 RobertaLayers) over the instruction tokens, then
LocalVPEncoder.encoder = CrossmodalEncoder (3 BertCrossLayers) with the [stop]+36 view tokens as
queries and the text as keys/values (P/model/vilmodel_goat.py:563-564 and :399).  Parameter names
follow the reference model (``lang_encoder.layer.N.*``, ``local_encoder.encoder.crossattention.N.*``).
"""
import torch
from torch import nn

from . import modules as M


class _HalfMeanSquare(torch.autograd.Function):
    """0.5 * mean(x^2) as one dot product forward and one scaled copy backward (the plain torch expression costs eight
    elementwise / reduction launches over the output streams per step)."""

    @staticmethod
    def forward(ctx, x):
        xf = x.reshape(-1)
        ctx.save_for_backward(x)
        return torch.dot(xf, xf) * (0.5 / xf.numel())

    @staticmethod
    def backward(ctx, go):
        x, = ctx.saved_tensors
        return x * (go * (1.0 / x.numel()))


def c2_loss(txt_out, vp_out):
    """The synthetic objective of the C2 workload (oracle.goat_oracle.c2_loss): mean square of both output streams."""
    return _HalfMeanSquare.apply(txt_out) + _HalfMeanSquare.apply(vp_out)


class _LocalBranch(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.encoder = M.CrossmodalEncoder(config)


class C2CrossEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.lang_encoder = M.LanguageEncoder(config)
        self.local_encoder = _LocalBranch(config)

    def forward(self, txt_embeds, txt_masks, vp_embeds, vp_masks):
        """txt_embeds [B,L,H], txt_masks bool [B,L], vp_embeds [B,Nq,H], vp_masks bool [B,Nq]
        -> (txt_out [B,L,H], vp_out [B,Nq,H])"""
        t = self.lang_encoder.run(txt_embeds, txt_masks)
        v = self.local_encoder.encoder.run(vp_embeds, vp_masks, t, txt_masks)
        return t.tensor(), v.tensor()
