"""Shared test helpers: golden fixtures, seeded parameters, digests."""
import os

import numpy as np
import torch

from oracle import goat_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: (torch.from_numpy(z[k]) if z[k].dtype.kind in "fiub" else z[k]) for k in z.files}


def digest(t):
    t = t.detach().double().flatten().cpu()
    idx = torch.arange(t.numel(), dtype=torch.float64)
    return torch.tensor([t.sum().item(), t.abs().sum().item(), (t * torch.cos(idx * 0.37)).sum().item(),
                         t[0].item(), t[-1].item()], dtype=torch.float64)


def assert_digests(gold, grads, rtol, prefix="gdig.", atol=1e-4, key_bias_atol=None):
    """grads: {param name: grad tensor}.  Compares against the 5-number digests stored in a fixture.
    The tolerance is relative to the abs-sum (digest[1]), the natural scale of the cancelling sums,
    plus an absolute floor: key-bias grads are analytically zero (softmax is shift-invariant), so the
    reference holds only rounding noise there."""
    n = 0
    for k, v in gold.items():
        if not k.startswith(prefix):
            continue
        name = k[len(prefix):]
        assert name in grads, "missing grad for %s" % name
        d = digest(grads[name])
        ref = v.double()
        scale = ref[1].abs().item() + 1e-30
        err = (d - ref).abs()
        if key_bias_atol is not None and name.endswith("key.bias"):
            # analytically zero: only rounding noise on both sides (16-bit operands make it ~1e-3 per element)
            assert d[1].item() <= key_bias_atol, (name, "abssum of an analytically-zero grad", d)
            n += 1
            continue
        # sum / cos-sum are cancelling sums over numel terms: allow rtol * abs-sum
        assert err[0].item() <= rtol * scale + atol, (name, "sum", d, ref)
        assert err[1].item() <= rtol * scale + atol, (name, "abssum", d, ref)
        assert err[2].item() <= rtol * scale + atol, (name, "cossum", d, ref)
        n += 1
    assert n > 0
    return n


def params_like(shapes, seed, dtype=torch.float32):
    return O.seeded_params(shapes, seed=seed, dtype=dtype)


def maxerr(a, b):
    return (a.double().cpu() - b.double().cpu()).abs().max().item()
