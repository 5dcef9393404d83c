"""ctypes binding of the C ABI in include/ldpc_b200.h (libldpc_b200.so, built in-tree by csrc/build.py).

There is no CPU fallback anywhere in this package: if the shared library is missing the import fails
with instructions, and if no CUDA device is usable every compute call raises ``LdpcError``.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# LDPC_B200_LIB selects a tuning build of the same library (csrc/build.py variant ...); default: the in-tree build
LIB_PATH = os.environ.get("LDPC_B200_LIB") or os.path.join(os.path.dirname(PKG_DIR), "ldpc_b200", "libldpc_b200.so")

LDPC_OK, LDPC_ERR_INVALID, LDPC_ERR_CUDA, LDPC_ERR_NOMEM, LDPC_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
LDPC_F32, LDPC_F64 = 0, 1
RULE_NORMALIZED, RULE_OFFSET = 0, 1
SCHEDULE_FLOODING, SCHEDULE_LAYERED = 0, 1
(GRAPH_N, GRAPH_M, GRAPH_E, GRAPH_CHECK_CLASSES, GRAPH_VAR_CLASSES, GRAPH_MAX_DC, GRAPH_MAX_DV,
 GRAPH_DEVICE, GRAPH_LAYER_LEVELS, GRAPH_LAYER_PIPED) = range(10)

# every symbol include/ldpc_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "ldpc_version", "ldpc_last_error", "ldpc_device_count", "ldpc_host_alloc", "ldpc_host_free",
    "ldpc_graph_create", "ldpc_graph_destroy", "ldpc_graph_query", "ldpc_graph_slot_of_edge",
    "ldpc_decoder_create", "ldpc_decoder_set_weights", "ldpc_decoder_destroy", "ldpc_decoder_reserve",
    "ldpc_decode_device", "ldpc_decode_host", "ldpc_decode_device_packed", "ldpc_decode_host_packed", "ldpc_host_chunk_plan", "ldpc_train_forward", "ldpc_train_backward", "ldpc_awgn_llr", "ldpc_mc_round", "ldpc_count_errors",
    "ldpc_decoder_profile_mode", "ldpc_decoder_profile_read",
]


class LdpcError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"ldpc_b200 error {code}: {message}")
        self.code = code


class DecoderConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("dtype", C.c_int32), ("max_iterations", C.c_int32),
        ("early_stop", C.c_int32), ("n_beta", C.c_int32), ("n_alpha", C.c_int32), ("bc", C.c_int32),
        ("n_quantizers", C.c_int32), ("check_rule", C.c_int32), ("schedule", C.c_int32),
        ("beta_index", C.c_void_p), ("beta", C.c_void_p),
        ("alpha_index", C.c_void_p), ("alpha", C.c_void_p), ("thresholds", C.c_void_p),
        ("quantizer_of_iter", C.c_void_p),
    ]


class Profile(C.Structure):
    _fields_ = [
        ("launches", C.c_int64), ("cn_launches", C.c_int64), ("vn_launches", C.c_int64),
        ("cn_ms", C.c_double), ("vn_ms", C.c_double), ("other_ms", C.c_double),
        ("frames_padded", C.c_int64), ("compactions", C.c_int64), ("early_exits", C.c_int64),
        ("graph_replays", C.c_int64), ("small_decodes", C.c_int64), ("resident_decodes", C.c_int64),
    ]


_lib = None


def load():
    """Load libldpc_b200.so (once).  Raises ImportError with build instructions if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  This package has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    lib.ldpc_version.restype = C.c_int
    lib.ldpc_last_error.restype = C.c_char_p
    lib.ldpc_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.ldpc_host_alloc.argtypes = [C.POINTER(vp), i64]
    lib.ldpc_host_free.argtypes = [vp]
    lib.ldpc_graph_create.argtypes = [C.c_int, i32, i32, vp, vp, C.POINTER(vp)]
    lib.ldpc_graph_destroy.argtypes = [vp]
    lib.ldpc_graph_query.argtypes = [vp, C.c_int, C.POINTER(i64)]
    lib.ldpc_graph_slot_of_edge.argtypes = [vp, vp]
    lib.ldpc_decoder_create.argtypes = [vp, C.POINTER(DecoderConfig), C.POINTER(vp)]
    lib.ldpc_decoder_set_weights.argtypes = [vp, vp, vp]
    lib.ldpc_decoder_destroy.argtypes = [vp]
    lib.ldpc_decoder_reserve.argtypes = [vp, i64]
    lib.ldpc_decode_device.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp]
    lib.ldpc_decode_host.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.ldpc_decode_device_packed.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp]
    lib.ldpc_decode_host_packed.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.ldpc_train_forward.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp]
    lib.ldpc_train_backward.argtypes = [vp, vp, vp, vp, vp]
    lib.ldpc_host_chunk_plan.argtypes = [i64, i64, i32, vp, i32, C.POINTER(i32)]
    lib.ldpc_awgn_llr.argtypes = [C.c_int, i32, i64, u64, u64, C.c_float, i32, vp, vp, vp]
    lib.ldpc_mc_round.argtypes = [vp, C.c_float, i32, u64, u64, i64, vp, vp, vp, vp, vp]
    lib.ldpc_count_errors.argtypes = [C.c_int, i32, i64, vp, vp, vp, vp, vp, vp]
    lib.ldpc_decoder_profile_mode.argtypes = [vp, i32]
    lib.ldpc_decoder_profile_read.argtypes = [vp, C.POINTER(Profile), i32]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("ldpc_version", "ldpc_last_error"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != LDPC_OK:
        msg = load().ldpc_last_error()
        raise LdpcError(rc, msg.decode() if msg else "unknown")


def device_count() -> int:
    c = C.c_int(0)
    rc = load().ldpc_device_count(C.byref(c))
    return int(c.value) if rc == LDPC_OK else 0
