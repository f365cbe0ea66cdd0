"""Argument validation of the C ABI (include/goat_sm100.h), runnable without a GPU: every entry point rejects bad
arguments with a non-zero code and a message in goat_last_error() BEFORE touching the device -- it never throws, exits or
falls back.  (No compute call is made here; the kernels themselves are covered by the -m gpu parity tests.)"""
import ctypes as C

import pytest

from vln_goat_b200 import _lib

INVALID, UNSUPPORTED = 1, 3


def _err():
    return _lib.lib().goat_last_error().decode()


def test_version_and_device_probe():
    L = _lib.lib()
    assert L.goat_version() >= 100
    assert L.goat_device_supported() in (0, 1)      # 0 on a box without an sm_100 device; never raises


def _gemm_args(**kw):
    a = _lib.GemmArgs()
    a.M, a.N, a.K = 128, 128, 64
    a.dtype, a.out_dtype = _lib.BF16, _lib.BF16
    a.lda = a.ldb = 64
    a.ldc = 128
    a.A = a.B = a.out = 0x1000      # never dereferenced: validation fails first
    a.alpha = 1.0
    for k, v in kw.items():
        setattr(a, k, v)
    return a


@pytest.mark.parametrize("kw,needle", [
    (dict(K=0), "K must be > 0"),
    (dict(M=-1), "negative dimension"),
    (dict(A=None), "null A/B/out"),
    (dict(dtype=7), "bad dtype"),
    (dict(out_dtype=_lib.F16), "out_dtype"),
    (dict(act=9), "bad act"),
    (dict(act=_lib.ACT_DGELU), "need aux_in"),
    (dict(drop_p=1.5), "drop_p"),
    (dict(lda=8), "lda too small"),
    (dict(ldc=64), "ldc too small"),
    (dict(accumulate=1), "accumulate mode"),                  # needs an fp32 output
])
def test_gemm_rejects_bad_arguments(kw, needle):
    L = _lib.lib()
    rc = L.goat_gemm(C.byref(_gemm_args(**kw)), None)
    assert rc == INVALID
    assert needle in _err(), _err()


def test_gemm_null_struct_and_empty_problem():
    L = _lib.lib()
    assert L.goat_gemm(None, None) == INVALID and "null args" in _err()
    assert L.goat_gemm(C.byref(_gemm_args(M=0)), None) == 0          # empty problem: nothing to do, success


def test_layernorm_and_colsum_validation():
    L = _lib.lib()
    rc = L.goat_layernorm_fwd(None, 0, None, None, 1e-5, None, None, 1, None, None, 4, 768, None)
    assert rc == INVALID and "null" in _err()
    rc = L.goat_layernorm_fwd(0x1000, 0, 0x1000, 0x1000, 1e-5, 0x1000, None, 1, None, None, 4, 4096, None)
    assert rc == INVALID and "unsupported" in _err()
    rc = L.goat_layernorm_bwd_acc(0x1000, 0x1000, 0, 0x1000, 0x1000, 0x1000, None, 0x1000, None, 1, 0.0, 0, None,
                                  0x1000, 0x1000, None, 4, 512, None)
    assert rc == UNSUPPORTED and "768" in _err()
    rc = L.goat_colsum_acc(0x1000, _lib.BF16, 16, 12, 12, 0x1000, None)     # 12 is not a multiple of 8 columns
    assert rc == INVALID and "multiples" in _err()
    assert L.goat_colsum_acc(0x1000, _lib.BF16, 0, 16, 16, 0x1000, None) == 0   # no rows: success, no launch


def _ptrs(vals):
    return (C.c_void_p * len(vals))(*vals)


def test_peer_exchange_validation():
    """csrc/exchange.cu: world / rank / alignment / null checks happen on the host before any launch."""
    L = _lib.lib()
    assert L.goat_peer_signal_bytes() >= 4 * (_lib.MAX_PEERS + 2 * _lib.MAX_PEERS)
    n = C.c_int(0)
    two = _ptrs([0x1000, 0x2000])
    # reduce: world outside 1..GOAT_MAX_PEERS, unaligned shard start, null peer pointer
    assert L.goat_peer_reduce_sumsq(two, _lib.MAX_PEERS + 1, 0, 64, 0x1000, 0x1000, C.byref(n), None) == INVALID
    assert "world" in _err()
    assert L.goat_peer_reduce_sumsq(two, 2, 2, 64, 0x1000, 0x1000, C.byref(n), None) == INVALID
    assert "aligned" in _err()
    assert L.goat_peer_reduce_sumsq(_ptrs([0x1000, None]), 2, 0, 64, 0x1000, 0x1000, C.byref(n), None) == INVALID
    assert "bad peer pointer 1" in _err()
    assert L.goat_peer_reduce_sumsq(None, 2, 0, 64, 0x1000, 0x1000, C.byref(n), None) == INVALID
    # AdamW over peers: rank outside the world, shard start not a multiple of 8, shadow_lo without shadow, bad dtype
    args = dict(p=two, s=two, slo=two, dt=_lib.F16, world=2, rank=0, lo=0, n=64, n_decay=64, n_fp32=0, nparts=1)

    def adamw(**kw):
        a = dict(args, **kw)
        return L.goat_adamw_step_peers(a["p"], a["s"], a["slo"], a["dt"], a["world"], a["rank"], 0x1000, 0x1000, 0x1000,
                                       a["lo"], a["n"], a["n_decay"], a["n_fp32"], 0x1000, 0x1000, a["nparts"], None, None,
                                       None)
    assert adamw(rank=2) == INVALID and "world / rank" in _err()
    assert adamw(lo=4) == INVALID and "multiples of 8" in _err()
    assert adamw(s=None) == INVALID and "shadow_lo needs shadow" in _err()
    assert adamw(dt=_lib.F32) == INVALID and "shadow dtype" in _err()
    assert adamw(nparts=0) == INVALID and "nparts" in _err()
    assert adamw(p=_ptrs([0x1000, 0x1004])) == INVALID and "parameter pointer of rank 1" in _err()
    assert adamw(n=0) == 0                                   # empty shard: nothing to do
    # flag barrier / scalar exchange
    assert L.goat_peer_barrier(None, 2, 0, 1, None) == INVALID
    assert L.goat_peer_barrier(two, 2, 5, 1, None) == INVALID and "world / rank" in _err()
    assert L.goat_peer_sum_scalar(two, 2, 0, 1, None, 1, 0x1000, None) == INVALID
    assert L.goat_peer_sum_scalar(two, 2, 0, 1, 0x1000, 0, 0x1000, None) == INVALID and "nparts" in _err()
    # split cast and handle export
    assert L.goat_split_cast(0x1000, 0x1000, None, _lib.F32, 8, None) == INVALID and "dtype" in _err()
    assert L.goat_split_cast(0x1004, 0x1000, None, _lib.F16, 8, None) == INVALID and "aligned" in _err()
    assert L.goat_split_cast(None, 0x1000, None, _lib.F16, 8, None) == INVALID
    assert L.goat_split_cast(0x1000, 0x1000, None, _lib.F16, 0, None) == 0
    off = C.c_ulonglong(0)
    assert L.goat_peer_export(None, (C.c_ubyte * _lib.PEER_HANDLE_BYTES)(), C.byref(off)) == INVALID
    assert L.goat_peer_open(None, C.byref(C.c_void_p(0))) == INVALID
    assert L.goat_peer_close(None) == 0
