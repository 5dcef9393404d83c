"""Code container, basic normalised min-sum decoder and the AWGN channel -- the B200-native
counterparts of the reference's ``ldpc_decoder.py`` (same public names and call signatures).

Reference surface mirrored here (reference file:line):
  LDPCCode                 ldpc_decoder.py:26-54
  BasicMinSumDecoder       ldpc_decoder.py:56-153   decode(llr) -> (decoded int64, success bool, iterations int)
  create_test_ldpc_code    ldpc_decoder.py:274-284  the (7,4) matrix, max_iterations=10
  simulate_awgn_channel    ldpc_decoder.py:286-302

Everything that computes runs in the CUDA library (include/ldpc_b200.h); there is no CPU fallback.
Additive extensions: ``decode`` also accepts a batch ``[B, n]`` (returns ``[B, n]`` int64, ``[B]`` bool,
``[B]`` int32) and CUDA tensors (returns CUDA tensors); ``H`` may be a scipy sparse matrix.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

from .engine import Engine, EngineCache, TannerGraph, awgn_llr, default_device


@dataclass
class LDPCCode:
    """LDPC code parameters (ldpc_decoder.py:26-54).  ``H`` is the m x n parity-check matrix; only
    entries equal to 1 are edges.  The degree maps hold the same values as the reference's
    properties but are computed once instead of on every access."""
    n: int
    k: int
    H: object
    max_iterations: int = 50
    _graph: Optional[TannerGraph] = field(default=None, repr=False, compare=False)
    _deg: Optional[tuple] = field(default=None, repr=False, compare=False)

    @property
    def rate(self) -> float:
        return self.k / self.n

    def invalidate(self):
        """Forget the cached degree maps and graph (call after editing ``H`` IN PLACE; assigning a new ``H`` object is
        noticed by itself).  The reference recomputes both from ``H`` on every access."""
        self._graph = None
        self._deg = None
        self._h_id = id(self.H)

    def _check_h(self):
        if getattr(self, "_h_id", None) != id(self.H):
            self.invalidate()

    def _degree_maps(self):
        self._check_h()
        if self._deg is None:
            try:
                import scipy.sparse as sp
                sparse = sp.issparse(self.H)
            except ImportError:
                sparse = False
            if sparse:
                row = np.asarray(self.H.sum(axis=1)).ravel()
                col = np.asarray(self.H.sum(axis=0)).ravel()
            else:
                Hd = np.asarray(self.H)
                row, col = Hd.sum(axis=1), Hd.sum(axis=0)
            self._deg = ({i: int(d) for i, d in enumerate(row)}, {j: int(d) for j, d in enumerate(col)})
        return self._deg

    @property
    def check_node_degrees(self) -> Dict[int, int]:
        """{check index: row sum of H} -- ldpc_decoder.py:38-45"""
        return self._degree_maps()[0]

    @property
    def variable_node_degrees(self) -> Dict[int, int]:
        """{variable index: column sum of H} -- ldpc_decoder.py:47-54"""
        return self._degree_maps()[1]

    @property
    def graph(self) -> TannerGraph:
        self._check_h()
        if self._graph is None:
            self._graph = TannerGraph.from_H(self.H)
            if self._graph.n != self.n:
                raise ValueError(f"H has {self._graph.n} columns but n = {self.n}")
        return self._graph


def _as_batch(llr):
    """Returns (array-or-tensor 2-D, single: bool, is_torch: bool)."""
    try:
        import torch
        if isinstance(llr, torch.Tensor):
            return (llr[None] if llr.dim() == 1 else llr), llr.dim() == 1, True
    except ImportError:
        pass
    a = np.asarray(llr)
    return (a[None] if a.ndim == 1 else a), a.ndim == 1, False


class BasicMinSumDecoder(EngineCache):
    """Normalised min-sum, flooding schedule, float64 (ldpc_decoder.py:56-153).

    ``c2v = (factor * raw) * prod(other signs)``, iterations = ``code.max_iterations`` (read at call
    time, like the reference), early stop on a zero syndrome."""

    def __init__(self, code: LDPCCode, factor: float = 0.7):
        self.code = code
        self.factor = factor
        self._engines = {}

    def _engine(self, device: int) -> Engine:
        T = int(self.code.max_iterations)
        key = (device, T, float(self.factor))
        eng = self._engines.get(key)
        if eng is None:
            beta = np.full((T, 1), float(self.factor), dtype=np.float64)
            eng = Engine(self.code.graph, dtype=np.float64, max_iterations=T, beta=beta, device=device)
            self._engines = {key: eng}
        return eng

    def decode(self, llr) -> Tuple[np.ndarray, bool, int]:
        batch, single, is_torch = _as_batch(llr)
        if self.code.max_iterations < 1:
            raise ValueError("max_iterations must be >= 1")
        if is_torch and batch.device.type == "cuda":
            import torch
            eng = self._engine(batch.device.index)
            bits, _, iters, succ = eng.decode_device(batch.to(torch.float64))
            decoded = bits.to(torch.int64)
            if single:
                return decoded[0], bool(succ[0].item()), int(iters[0].item())
            return decoded, succ.bool(), iters
        arr = batch.numpy() if is_torch else batch
        eng = self._engine(default_device())
        bits, _, iters, succ = eng.decode_host(np.asarray(arr, dtype=np.float64))
        decoded = bits.astype(np.int64)
        if single:
            return decoded[0], bool(succ[0]), int(iters[0])
        return decoded, succ.astype(bool), iters


def create_test_ldpc_code() -> LDPCCode:
    """The (7,4) test code of ldpc_decoder.py:274-284 (13 edges, check degrees 3,3,3,4)."""
    rows = ["1101000", "0110100", "1010010", "1110001"]
    H = np.array([[int(c) for c in r] for r in rows])
    return LDPCCode(n=7, k=4, H=H, max_iterations=10)


def simulate_awgn_channel(codeword: np.ndarray, snr_db: float) -> np.ndarray:
    """BPSK over AWGN, LLR out (ldpc_decoder.py:286-302): symbol = 2*c - 1, noise power = 10^(-snr/10),
    llr = 2*y / noise power -- the reference's own sign convention (bit 0 -> negative LLR).

    The noise comes from the device Philox generator; the 64-bit seed is drawn from numpy's global
    RNG so ``np.random.seed`` still makes a run reproducible (the values differ from MT19937's)."""
    import torch

    cw = np.ascontiguousarray(codeword).astype(np.uint8)
    seed = int(np.random.randint(0, 2 ** 63 - 1, dtype=np.int64))
    dev = default_device()
    cw_dev = torch.from_numpy(cw).to(f"cuda:{dev}")
    llr = awgn_llr(cw.shape[0], 1, snr_db, seed=seed, frame0=0, llr_sign=-1, codeword=cw_dev, device=dev)
    return llr[0].double().cpu().numpy()


def __getattr__(name):
    # ``from ldpc_decoder import NeuralMinSumDecoder`` gives the reference's legacy class of that name
    # (ldpc_decoder.py:155-272); resolved lazily because it is built on the nn.Module base of the neural decoders
    if name == "NeuralMinSumDecoder":
        from .neural_minsum_decoder import LegacyNeuralMinSumDecoder
        return LegacyNeuralMinSumDecoder
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
