"""ctypes binding of libgoat_sm100.so (the C ABI declared in include/goat_sm100.h).

The product path has no fallback: if the shared library is missing or a call fails this module
raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``vln_goat_b200/csrc/build.sh``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgoat_sm100.so")

F32, F16, BF16 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_RELU, ACT_DGELU, ACT_DRELU, ACT_TANH = 0, 1, 2, 3, 4, 5


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("dtype", C.c_int),
        ("a_mn_major", C.c_int), ("b_mn_major", C.c_int),
        ("lda", C.c_int), ("ldb", C.c_int),
        ("A", C.c_void_p), ("B", C.c_void_p),
        ("bias", C.c_void_p), ("res", C.c_void_p), ("ldres", C.c_int),
        ("aux_in", C.c_void_p), ("aux_out", C.c_void_p), ("ldaux", C.c_int),
        ("out", C.c_void_p), ("ldc", C.c_int), ("out_dtype", C.c_int),
        ("out2", C.c_void_p), ("ldc2", C.c_int),
        ("act", C.c_int), ("alpha", C.c_float), ("drop_p", C.c_float),
        ("drop_seed", C.c_uint64), ("drop_seed_ptr", C.c_void_p),
        ("force_simt", C.c_int),
        ("accumulate", C.c_int),
        ("B_lo", C.c_void_p),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("heads", C.c_int), ("Nq", C.c_int), ("Nk", C.c_int), ("D", C.c_int),
        ("dtype", C.c_int),
        ("Q", C.c_void_p), ("K", C.c_void_p), ("V", C.c_void_p),
        ("ldq", C.c_int), ("ldk", C.c_int), ("ldv", C.c_int),
        ("sbq", C.c_longlong), ("sbk", C.c_longlong), ("sbv", C.c_longlong),
        ("kmask", C.c_void_p), ("bias", C.c_void_p),
        ("scale", C.c_float),
        ("O", C.c_void_p), ("ldo", C.c_int), ("sbo", C.c_longlong),
        ("lse", C.c_void_p),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint64), ("drop_seed_ptr", C.c_void_p),
        ("dO", C.c_void_p), ("dQ", C.c_void_p), ("dK", C.c_void_p), ("dV", C.c_void_p),
        ("dbias", C.c_void_p),
        ("force_simt", C.c_int),
    ]


# name -> (restype, argtypes); every symbol include/goat_sm100.h declares
MAX_PEERS = 8              # GOAT_MAX_PEERS (include/goat_sm100.h)
PEER_HANDLE_BYTES = 64     # GOAT_PEER_HANDLE_BYTES

SYMBOLS = {
    "goat_version": (C.c_int, []),
    "goat_last_error": (C.c_char_p, []),
    "goat_device_supported": (C.c_int, []),
    "goat_gemm": (C.c_int, [C.POINTER(GemmArgs), C.c_void_p]),
    "goat_attn_core_fwd": (C.c_int, [C.POINTER(AttnArgs), C.c_void_p]),
    "goat_attn_core_bwd": (C.c_int, [C.POINTER(AttnArgs), C.c_void_p]),
    "goat_layernorm_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "goat_layernorm_bwd_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "goat_layernorm_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "goat_layernorm_bwd_acc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "goat_colsum_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "goat_colsum": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "goat_colsum_acc": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "goat_cast": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]),
    "goat_dropout_cast": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_uint64,
                                    C.c_void_p, C.c_void_p]),
    "goat_sumsq_workspace_bytes": (C.c_size_t, []),
    "goat_sumsq": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.POINTER(C.c_int), C.c_void_p]),
    "goat_adamw_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong,
                                  C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "goat_scaler_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p]),
    "goat_peer_export": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_ulonglong)]),
    "goat_peer_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "goat_peer_close": (C.c_int, [C.c_void_p]),
    "goat_split_cast": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]),
    "goat_peer_signal_bytes": (C.c_size_t, []),
    "goat_peer_barrier": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint, C.c_void_p]),
    "goat_peer_sum_scalar": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_void_p]),
    "goat_peer_reduce_sumsq": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_int), C.c_void_p]),
    "goat_adamw_step_peers": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong,
                                        C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "goat_act_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]),
    "goat_gather_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p]),
    "goat_act_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]),
    "goat_sprel_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "goat_sprel_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "goat_attn_pool_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "goat_attn_pool_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "goat_wsum_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "goat_wsum_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "goat_door_gate_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "goat_door_gate_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "goat_xent_fwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_longlong,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "goat_xent_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_int, C.c_longlong, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p]),
    "goat_xent_chunk_fwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "goat_xent_chunk_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_longlong, C.c_longlong, C.c_longlong, C.c_void_p, C.c_int, C.c_longlong,
                                      C.c_void_p]),
    "goat_segment_reduce_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p]),
    "goat_segment_reduce_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p]),
    "goat_embed_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                 C.c_void_p]),
    "goat_embed_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """Load (once) and return the CDLL with argtypes set.  Raises if the library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libgoat_sm100.so not found at %s -- build it first (__graft_entry__.build() or "
                "vln_goat_b200/csrc/build.sh); there is no CPU / PyTorch fallback for this path" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().goat_last_error()
        raise RuntimeError("%s failed (status %d): %s" % (what, rc, msg.decode() if msg else ""))
