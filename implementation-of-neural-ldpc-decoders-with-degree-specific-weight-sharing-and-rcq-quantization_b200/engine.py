"""Host-side plumbing between the reference-shaped Python classes and the C ABI.

``TannerGraph``  : CSR view of a parity-check matrix in the reference's neighbour order.
``Engine``       : one ``ldpc_decoder`` handle (graph + weight tables + quantisers + workspace).

PyTorch appears here only as the owner of device memory / streams for the device-pointer entry
points; the host-buffer entry points take numpy arrays and never touch torch.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import LdpcError  # noqa: F401


def _ptr(a) -> Optional[int]:
    if a is None:
        return None
    return a.ctypes.data


def default_device() -> int:
    try:
        import torch

        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except Exception:
        pass
    return 0


class TannerGraph:
    """Sparse Tanner graph.  Edges are the entries ``H == 1`` (ldpc_decoder.py:92: ``np.where(H[i, :] == 1)``),
    numbered check-major with ascending variable index inside a check."""

    def __init__(self, n: int, m: int, check_ptr: np.ndarray, check_var: np.ndarray):
        self.n = int(n)
        self.m = int(m)
        self.check_ptr = np.ascontiguousarray(check_ptr, dtype=np.int64)
        self.check_var = np.ascontiguousarray(check_var, dtype=np.int32)
        self.E = int(self.check_ptr[-1]) if self.m > 0 else 0
        self.check_degree = np.diff(self.check_ptr).astype(np.int64) if self.m > 0 else np.zeros(0, np.int64)
        self.var_degree = np.bincount(self.check_var, minlength=self.n).astype(np.int64)
        self.edge_check = np.repeat(np.arange(self.m, dtype=np.int64), self.check_degree)
        self._handles = {}
        self._lock = threading.Lock()

    @staticmethod
    def from_H(H) -> "TannerGraph":
        try:
            import scipy.sparse as sp

            if sp.issparse(H):
                Hc = H.tocsr()
                Hc.sort_indices()
                keep = Hc.data == 1
                rows = np.repeat(np.arange(Hc.shape[0]), np.diff(Hc.indptr))[keep]
                cols = Hc.indices[keep]
                return TannerGraph.from_coo(Hc.shape[1], Hc.shape[0], rows, cols)
        except ImportError:
            pass
        H = np.asarray(H)
        m, n = H.shape
        rows, cols = np.nonzero(H == 1)
        return TannerGraph.from_coo(n, m, rows, cols)

    @staticmethod
    def from_coo(n: int, m: int, rows, cols) -> "TannerGraph":
        rows = np.asarray(rows, dtype=np.int64)
        cols = np.asarray(cols, dtype=np.int64)
        order = np.lexsort((cols, rows))
        rows, cols = rows[order], cols[order]
        if rows.size > 1:
            dup = (rows[1:] == rows[:-1]) & (cols[1:] == cols[:-1])
            if dup.any():
                raise ValueError("duplicate (check, variable) entries")
        ptr = np.zeros(m + 1, dtype=np.int64)
        np.add.at(ptr, rows + 1, 1)
        return TannerGraph(n, m, np.cumsum(ptr), cols.astype(np.int32))

    # the graph is immutable: copies of a decoder share it; a pickled one re-creates its device handles on demand
    def __deepcopy__(self, memo):
        return self

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handles"] = {}
        state.pop("_lock", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._lock = threading.Lock()

    def handle(self, device: int) -> int:
        """Device-side graph (degree-sorted slot layout), created once per device."""
        with self._lock:
            h = self._handles.get(device)
            if h is None:
                lib = _lib.load()
                out = C.c_void_p()
                _lib.check(lib.ldpc_graph_create(device, self.n, self.m, _ptr(self.check_ptr), _ptr(self.check_var),
                                                 C.byref(out)))
                h = out.value
                self._handles[device] = h
            return h

    def query(self, device: int, what: int) -> int:
        v = C.c_int64(0)
        _lib.check(_lib.load().ldpc_graph_query(self.handle(device), what, C.byref(v)))
        return int(v.value)

    def slot_of_edge(self, device: int) -> np.ndarray:
        out = np.zeros(self.E, dtype=np.int32)
        _lib.check(_lib.load().ldpc_graph_slot_of_edge(self.handle(device), _ptr(out)))
        return out

    def syndrome(self, bits: np.ndarray) -> np.ndarray:
        """H . bits mod 2 for a batch [B, n] (host helper for tests / properties)."""
        bits = np.asarray(bits).astype(np.uint8)
        if bits.ndim == 1:
            bits = bits[None]
        par = np.zeros((bits.shape[0], self.m), dtype=np.uint8)
        contrib = bits[:, self.check_var]
        np.bitwise_xor.at(par, (slice(None), self.edge_check), contrib)
        return par

    def __del__(self):
        try:
            if self._handles:          # device handles exist only if the library was loaded
                lib = _lib.load()
                for h in self._handles.values():
                    lib.ldpc_graph_destroy(h)
        except Exception:
            pass


class EngineCache:
    """Mixin of the decoder classes: ``self._engines`` (device engines by configuration) is per object -- a copy or
    an unpickled decoder starts without engines and builds its own on first use."""

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engines"] = {}
        if "_pushed" in state:
            state["_pushed"] = {}
        return state


class Engine:
    """A configured decoder on one device (wraps ``ldpc_decoder_create`` .. ``ldpc_decoder_destroy``)."""

    def __init__(self, graph: TannerGraph, *, dtype=np.float32, max_iterations: int, early_stop: bool = True,
                 beta: Optional[np.ndarray] = None, beta_index: Optional[np.ndarray] = None,
                 alpha: Optional[np.ndarray] = None, alpha_index: Optional[np.ndarray] = None,
                 bc: int = 0, thresholds: Optional[np.ndarray] = None,
                 quantizer_of_iter: Optional[np.ndarray] = None, check_rule: int = 0, schedule: int = 0,
                 device: Optional[int] = None):
        self.graph = graph
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("dtype must be float32 or float64")
        self.T = int(max_iterations)
        self.device = default_device() if device is None else int(device)
        self.bc = int(bc)
        cfg = _lib.DecoderConfig()
        cfg.struct_size = C.sizeof(_lib.DecoderConfig)
        cfg.dtype = _lib.LDPC_F32 if self.dtype == np.float32 else _lib.LDPC_F64
        cfg.max_iterations = self.T
        cfg.early_stop = 1 if early_stop else 0
        cfg.check_rule = int(check_rule)
        cfg.schedule = int(schedule)
        keep = []

        def table(a, width_name):
            a = np.ascontiguousarray(a, dtype=self.dtype)
            if a.ndim != 2 or a.shape[0] != self.T:
                raise ValueError(f"{width_name} must have shape [T, width]")
            keep.append(a)
            return a

        if beta is not None:
            b = table(beta, "beta")
            cfg.n_beta = b.shape[1]
            cfg.beta = _ptr(b)
            if beta_index is not None:
                bi = np.ascontiguousarray(beta_index, dtype=np.int32)
                if bi.shape != (graph.E,):
                    raise ValueError("beta_index must have one entry per edge")
                keep.append(bi)
                cfg.beta_index = _ptr(bi)
        if alpha is not None:
            a = table(alpha, "alpha")
            cfg.n_alpha = a.shape[1]
            cfg.alpha = _ptr(a)
            if alpha_index is not None:
                ai = np.ascontiguousarray(alpha_index, dtype=np.int32)
                if ai.shape != (graph.n,):
                    raise ValueError("alpha_index must have one entry per variable")
                keep.append(ai)
                cfg.alpha_index = _ptr(ai)
        if self.bc:
            th = np.ascontiguousarray(thresholds, dtype=np.float32)
            if th.ndim != 2 or th.shape[1] != 2 ** (self.bc - 1):
                raise ValueError("thresholds must have shape [Q, 2^(bc-1)]")
            qi = np.ascontiguousarray(quantizer_of_iter, dtype=np.int32)
            if qi.shape != (self.T,):
                raise ValueError("quantizer_of_iter must have T entries")
            keep += [th, qi]
            cfg.bc = self.bc
            cfg.n_quantizers = th.shape[0]
            cfg.thresholds = _ptr(th)
            cfg.quantizer_of_iter = _ptr(qi)
        self._n_beta = int(cfg.n_beta)
        self._n_alpha = int(cfg.n_alpha)
        lib = _lib.load()
        out = C.c_void_p()
        _lib.check(lib.ldpc_decoder_create(graph.handle(self.device), C.byref(cfg), C.byref(out)))
        self._h = out.value
        self._lock = threading.Lock()

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.load().ldpc_decoder_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_weights(self, beta: Optional[np.ndarray] = None, alpha: Optional[np.ndarray] = None):
        b = None if beta is None else np.ascontiguousarray(beta, dtype=self.dtype)
        a = None if alpha is None else np.ascontiguousarray(alpha, dtype=self.dtype)
        if b is not None and b.shape != (self.T, self._n_beta):
            raise ValueError("beta shape changed")
        if a is not None and a.shape != (self.T, self._n_alpha):
            raise ValueError("alpha shape changed")
        _lib.check(_lib.load().ldpc_decoder_set_weights(self._h, _ptr(b), _ptr(a)))

    def reserve(self, frames: int):
        _lib.check(_lib.load().ldpc_decoder_reserve(self._h, int(frames)))

    # ------------------------------------------------------------------ decode
    @property
    def row_words(self) -> int:
        """uint32 words per frame of the packed decision rows."""
        return (self.graph.n + 31) // 32

    def unpack_rows(self, packed: np.ndarray) -> np.ndarray:
        """Packed decision rows [B, row_words] uint32 -> one byte per bit [B, n] uint8 (host)."""
        rows = np.ascontiguousarray(packed, dtype=np.uint32)
        return np.unpackbits(rows.view(np.uint8), axis=1, bitorder="little")[:, :self.graph.n]

    def decode_host(self, llr: np.ndarray, want_posterior: bool = False, out: Optional[dict] = None,
                    packed_bits: bool = False) -> Tuple[np.ndarray, Optional[np.ndarray], np.ndarray, np.ndarray]:
        """llr: host [B, n] (pinned memory gives full PCIe rate).  Returns bits u8, posterior|None,
        iterations i32, success u8 as host arrays (``out`` may supply preallocated ones).  ``packed_bits``: the
        decisions come back as packed rows [B, row_words] uint32 (bit j & 31 of word j >> 5 = variable j), an eighth
        of the bytes over the host link; ``unpack_rows`` expands them."""
        llr = np.ascontiguousarray(llr, dtype=self.dtype)
        if llr.ndim != 2 or llr.shape[1] != self.graph.n:
            raise IndexError(f"llr must have shape [B, {self.graph.n}], got {llr.shape}")
        B = llr.shape[0]
        out = out or {}

        def buffer(name, shape, dtype, wanted=True):
            a = out.get(name)
            if a is None:
                return np.empty(shape, dtype=dtype) if wanted else None
            # the C side writes through the raw pointer: the array must be exactly what it expects
            if not isinstance(a, np.ndarray) or a.shape != shape or a.dtype != np.dtype(dtype) or \
                    not a.flags.c_contiguous or not a.flags.writeable:
                raise ValueError(f"out[{name!r}] must be a writable C-contiguous {np.dtype(dtype).name} array of shape {shape}")
            return a

        if packed_bits:
            bits = buffer("bits", (B, self.row_words), np.uint32)
        else:
            bits = buffer("bits", (B, self.graph.n), np.uint8)
        post = buffer("posterior", (B, self.graph.n), self.dtype, wanted=want_posterior)
        iters = buffer("iterations", (B,), np.int32)
        succ = buffer("success", (B,), np.uint8)
        fn = _lib.load().ldpc_decode_host_packed if packed_bits else _lib.load().ldpc_decode_host
        with self._lock:
            _lib.check(fn(self._h, _ptr(llr), B, _ptr(bits), _ptr(post), _ptr(iters), _ptr(succ)))
        return bits, post, iters, succ

    def decode_device(self, llr, want_posterior: bool = False, want_bits: bool = True, packed_bits: bool = False):
        """llr: CUDA torch tensor [B, n] on this engine's device.  Enqueued on the current torch stream.
        ``packed_bits``: decisions as packed rows [B, row_words] int32 (the uint32 words of ``decode_host``)."""
        import torch

        tdt = torch.float32 if self.dtype == np.float32 else torch.float64
        if llr.device.type != "cuda" or llr.device.index != self.device:
            raise ValueError(f"llr must live on cuda:{self.device}")
        if llr.dim() != 2 or llr.shape[1] != self.graph.n:
            raise IndexError(f"llr must have shape [B, {self.graph.n}], got {tuple(llr.shape)}")
        llr = llr.to(tdt).contiguous()
        B = llr.shape[0]
        dev = llr.device
        if packed_bits:
            bits = torch.empty((B, self.row_words), dtype=torch.int32, device=dev) if want_bits else None
        else:
            bits = torch.empty((B, self.graph.n), dtype=torch.uint8, device=dev) if want_bits else None
        post = torch.empty((B, self.graph.n), dtype=tdt, device=dev) if want_posterior else None
        iters = torch.empty(B, dtype=torch.int32, device=dev)
        succ = torch.empty(B, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        fn = _lib.load().ldpc_decode_device_packed if packed_bits else _lib.load().ldpc_decode_device
        with self._lock:
            _lib.check(fn(
                self._h, llr.data_ptr(), B, bits.data_ptr() if bits is not None else None,
                post.data_ptr() if post is not None else None, iters.data_ptr(), succ.data_ptr(), stream))
        return bits, post, iters, succ

    # ------------------------------------------------------------------ posterior training
    def train_forward(self, llr):
        """forward() of a training step: llr CUDA float32 [B, n]; keeps every iteration's messages in the handle for
        ``train_backward``.  Returns bits u8, posterior f32, iterations i32, success u8 (CUDA)."""
        import torch

        if llr.device.type != "cuda" or llr.device.index != self.device:
            raise ValueError(f"llr must live on cuda:{self.device}")
        if llr.dim() != 2 or llr.shape[1] != self.graph.n:
            raise IndexError(f"llr must have shape [B, {self.graph.n}], got {tuple(llr.shape)}")
        llr = llr.to(torch.float32).contiguous()
        B, dev = llr.shape[0], llr.device
        bits = torch.empty((B, self.graph.n), dtype=torch.uint8, device=dev)
        post = torch.empty((B, self.graph.n), dtype=torch.float32, device=dev)
        iters = torch.empty(B, dtype=torch.int32, device=dev)
        succ = torch.empty(B, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with self._lock:
            _lib.check(_lib.load().ldpc_train_forward(self._h, llr.data_ptr(), B, bits.data_ptr(), post.data_ptr(),
                                                      iters.data_ptr(), succ.data_ptr(), stream))
            self._train_serial = getattr(self, "_train_serial", 0) + 1
            self._train_frames = B
        return bits, post, iters, succ

    def train_backward(self, grad_posterior):
        """d loss / d posterior [B, n] (CUDA float32, the frames of the last ``train_forward``) ->
        (d loss / d beta [T, n_beta] | None, d loss / d alpha [T, n_alpha] | None) as CUDA float32 tensors."""
        import torch

        g = grad_posterior.to(torch.float32).contiguous()
        if g.shape != (getattr(self, "_train_frames", -1), self.graph.n):
            raise ValueError("grad_posterior must match the batch of the last train_forward")
        dev = g.device
        gb = torch.empty((self.T, self._n_beta), dtype=torch.float32, device=dev) if self._n_beta else None
        ga = torch.empty((self.T, self._n_alpha), dtype=torch.float32, device=dev) if self._n_alpha else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        with self._lock:
            _lib.check(_lib.load().ldpc_train_backward(self._h, g.data_ptr(), gb.data_ptr() if gb is not None else None,
                                                       ga.data_ptr() if ga is not None else None, stream))
        return gb, ga

    # ------------------------------------------------------------------ Monte-Carlo
    def mc_round(self, snr_db: float, frames: int, *, seed: int, frame0: int, llr_sign: int, counters,
                 codeword=None, frame_bit_errors=None, frame_iterations=None):
        """One device-side round (AWGN -> decode -> count); accumulates into ``counters`` (CUDA int64[4])."""
        import torch

        stream = torch.cuda.current_stream(counters.device).cuda_stream
        with self._lock:
            _lib.check(_lib.load().ldpc_mc_round(
                self._h, float(snr_db), int(llr_sign), int(seed) & (2 ** 64 - 1), int(frame0), int(frames),
                codeword.data_ptr() if codeword is not None else None, counters.data_ptr(),
                frame_bit_errors.data_ptr() if frame_bit_errors is not None else None,
                frame_iterations.data_ptr() if frame_iterations is not None else None, stream))

    # ------------------------------------------------------------------ instrumentation
    def profile_mode(self, mode: int):
        _lib.check(_lib.load().ldpc_decoder_profile_mode(self._h, int(mode)))

    def profile_read(self, reset: bool = True) -> dict:
        p = _lib.Profile()
        _lib.check(_lib.load().ldpc_decoder_profile_read(self._h, C.byref(p), 1 if reset else 0))
        return {k: getattr(p, k) for k, _ in _lib.Profile._fields_}


def awgn_llr(n: int, frames: int, snr_db: float, *, seed: int = 0, frame0: int = 0, llr_sign: int = 1,
             codeword=None, device: Optional[int] = None):
    """Device AWGN LLRs [frames, n] float32 (same noise as Engine.mc_round for the same seed / frame index)."""
    import torch

    dev = default_device() if device is None else int(device)
    out = torch.empty((frames, n), dtype=torch.float32, device=f"cuda:{dev}")
    stream = torch.cuda.current_stream(out.device).cuda_stream
    _lib.check(_lib.load().ldpc_awgn_llr(dev, n, frames, int(frame0), int(seed) & (2 ** 64 - 1), float(snr_db),
                                         int(llr_sign), codeword.data_ptr() if codeword is not None else None,
                                         out.data_ptr(), stream))
    return out


def count_errors(bits, codeword=None, iterations=None, counters=None, frame_bit_errors=None):
    """Accumulate {frame_errors, bit_errors, total_iterations, total_frames} for CUDA uint8 bits [B, n]."""
    import torch

    dev = bits.device
    if counters is None:
        counters = torch.zeros(4, dtype=torch.int64, device=dev)
    bits = bits.contiguous()
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.load().ldpc_count_errors(
        dev.index, bits.shape[1], bits.shape[0], bits.data_ptr(),
        codeword.data_ptr() if codeword is not None else None,
        iterations.data_ptr() if iterations is not None else None, counters.data_ptr(),
        frame_bit_errors.data_ptr() if frame_bit_errors is not None else None, stream))
    return counters


class PinnedBuffer:
    """Page-locked host array (ldpc_host_alloc) so that the host entry points run at full PCIe rate."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        _lib.check(_lib.load().ldpc_host_alloc(C.byref(p), nbytes))
        self._p = p.value
        buf = (C.c_char * max(nbytes, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        p, self._p = self._p, None
        if p:
            self.array = None
            _lib.load().ldpc_host_free(p)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
