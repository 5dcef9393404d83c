"""ctypes front-end of oracle/minsum_oracle.c (TEST INFRASTRUCTURE -- see restatement.py)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import build as _build
from .restatement import MODE_NMS, MODE_OFFSET, MODE_RCQ, MODE_WRCQ, OracleResult, SparseGraph  # noqa: F401

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.OUT
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(_build.SRC):
            path = _build.build()
        _lib = C.CDLL(path)
        _lib.oracle_torch_sum_f32.restype = C.c_float
        _lib.oracle_np_sum_f64.restype = C.c_double
        _lib.oracle_dequantize_f32.restype = C.c_float
        _lib.oracle_quantize_f32.argtypes = [C.c_float, C.c_void_p, C.c_int, C.c_int]
        _lib.oracle_dequantize_f32.argtypes = [C.c_int, C.c_void_p, C.c_int]
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def decode(graph: SparseGraph, llr: np.ndarray, *, T: int, mode: int = MODE_NMS, dtype=np.float32,
           beta: Optional[np.ndarray] = None, alpha: Optional[np.ndarray] = None, bc: int = 0,
           thresholds: Optional[np.ndarray] = None, quantizer_of_iter: Optional[np.ndarray] = None,
           early_stop: bool = True, want_posterior: bool = True, nthreads: int = 1) -> OracleResult:
    """Same contract as restatement.decode (beta [T,E] check-major, alpha [T,n])."""
    dt = np.dtype(dtype)
    llr = np.ascontiguousarray(np.asarray(llr, dtype=dt))
    if llr.ndim == 1:
        llr = llr[None, :]
    B, n = llr.shape
    assert n == graph.n
    cp = np.ascontiguousarray(graph.check_ptr, dtype=np.int64)
    cv = np.ascontiguousarray(graph.check_var, dtype=np.int32)
    vp = np.ascontiguousarray(graph.var_ptr, dtype=np.int64)
    ve = np.ascontiguousarray(graph.var_edge, dtype=np.int64)
    b = None if beta is None else np.ascontiguousarray(beta, dtype=dt)
    a = None if alpha is None else np.ascontiguousarray(alpha, dtype=dt)
    if b is not None:
        assert b.shape == (T, graph.E)
    if a is not None:
        assert a.shape == (T, n)
    th = None if thresholds is None else np.ascontiguousarray(thresholds, dtype=np.float32)
    qi = None if quantizer_of_iter is None else np.ascontiguousarray(quantizer_of_iter, dtype=np.int32)
    bits = np.zeros((B, n), dtype=np.uint8)
    post = np.zeros((B, n), dtype=dt) if want_posterior else None
    iters = np.zeros(B, dtype=np.int32)
    succ = np.zeros(B, dtype=np.uint8)
    fn = lib().oracle_decode_f32 if dt == np.float32 else lib().oracle_decode_f64
    rc = fn(C.c_int(n), C.c_int(graph.m), _p(cp), _p(cv), _p(vp), _p(ve), C.c_int(mode), C.c_int(T),
            C.c_int(1 if early_stop else 0), _p(b), _p(a), C.c_int(bc), _p(th), _p(qi),
            C.c_int64(B), _p(llr), _p(bits), _p(post), _p(iters), _p(succ), C.c_int(nthreads))
    if rc != 0:
        raise MemoryError("oracle_decode failed")
    return OracleResult(bits=bits, posterior=post, iterations=iters, success=succ.astype(bool))


def decode_layered_rcq(graph: SparseGraph, llr: np.ndarray, *, T: int, bc: int, thresholds: np.ndarray,
                       quantizer_of_iter: np.ndarray, nthreads: int = 1) -> OracleResult:
    """Same contract as restatement.decode_layered_rcq."""
    llr = np.ascontiguousarray(np.asarray(llr, dtype=np.float32))
    if llr.ndim == 1:
        llr = llr[None, :]
    B, n = llr.shape
    cp = np.ascontiguousarray(graph.check_ptr, dtype=np.int64)
    cv = np.ascontiguousarray(graph.check_var, dtype=np.int32)
    th = np.ascontiguousarray(thresholds, dtype=np.float32)
    qi = np.ascontiguousarray(quantizer_of_iter, dtype=np.int32)
    bits = np.zeros((B, n), dtype=np.uint8)
    post = np.zeros((B, n), dtype=np.float32)
    iters = np.zeros(B, dtype=np.int32)
    succ = np.zeros(B, dtype=np.uint8)
    rc = lib().oracle_decode_layered_rcq(C.c_int(n), C.c_int(graph.m), _p(cp), _p(cv), C.c_int(T), C.c_int(bc), _p(th),
                                         _p(qi), C.c_int64(B), _p(llr), _p(bits), _p(post), _p(iters), _p(succ),
                                         C.c_int(nthreads))
    if rc != 0:
        raise MemoryError("oracle_decode_layered_rcq failed")
    return OracleResult(bits=bits, posterior=post, iterations=iters, success=succ.astype(bool))
