"""RCQ (reconstruction-computation-quantisation) decoders -- B200-native counterparts of the
reference's ``rcq_decoder.py`` (same constructors, attributes and return tuples).

  NonUniformQuantizer   rcq_decoder.py:22-121   thresholds tau_j = C * (j / (2^(bc-1) - 1))^gamma
  RCQMinSumDecoder      rcq_decoder.py:123-279  decode(llr) -> (decoded int32, success bool, iterations int)
  WeightedRCQDecoder    rcq_decoder.py:352-597  forward(llr) -> (decoded int32, posterior f32, iterations int)

What the reference's code actually does and what is reproduced: only the
check-to-variable messages are quantised (to a bc-bit sign-magnitude code, reconstructed as the lower
bin edge); variable-to-check messages and posteriors stay float32; ``bv`` is stored and unused.  On
the device the C2V messages are held as the integer codes themselves (1 byte per edge per frame).
``RCQMinSumDecoder(layered=True)`` runs the layered schedule exactly as the reference executes it
(rcq_decoder.py:281-350): posteriors updated in place check by check, the "subtract previous C2V" step
being a no-op there.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from ._neural_base import DecoderModule, build_2d_tables
from .engine import Engine, EngineCache, default_device
from .ldpc_decoder import LDPCCode


class NonUniformQuantizer:
    """Power-function threshold quantiser (rcq_decoder.py:22-121).  ``thresholds`` is a list of Python
    floats like the reference's; comparisons happen in float32."""

    def __init__(self, bc: int, C: float, gamma: float):
        self.bc = bc
        self.C = C
        self.gamma = gamma
        levels = 2 ** (bc - 1)
        top = levels - 1
        self.thresholds: List[float] = [C * (j / top) ** gamma for j in range(levels)]

    def _thr32(self, device) -> torch.Tensor:
        return torch.tensor(self.thresholds, dtype=torch.float64, device=device).to(torch.float32)

    def quantize(self, x: torch.Tensor) -> torch.Tensor:
        """Sign-magnitude code in [0, 2^bc): magnitude index = last threshold the magnitude reaches,
        plus 2^(bc-1) when x < 0 (rcq_decoder.py:59-91)."""
        thr = self._thr32(x.device)
        mag = x.abs().to(torch.float32)
        passed = mag.unsqueeze(-1) >= thr                       # [..., levels]
        order = torch.arange(thr.numel(), device=x.device)
        idx = torch.where(passed, order, torch.zeros_like(order)).amax(dim=-1)
        neg = (x < 0).to(torch.long)
        return neg * (2 ** (self.bc - 1)) + idx.to(torch.long)

    def dequantize(self, quantized: torch.Tensor) -> torch.Tensor:
        """+-float32(threshold[idx]) -- the lower edge of the bin (rcq_decoder.py:93-121)."""
        levels = 2 ** (self.bc - 1)
        thr = self._thr32(quantized.device)
        neg = (quantized >= levels).to(torch.float32)
        return (1 - 2 * neg) * thr[quantized % levels]


def _schedule(T: int, Q: int) -> np.ndarray:
    """Quantiser per iteration (rcq_decoder.py:156-167): thirds of the iteration budget."""
    t = np.arange(T)
    if Q == 1:
        return np.zeros(T, dtype=np.int32)
    return np.where(t < T // 3, 0, np.where(t < 2 * T // 3, 1, Q - 1)).astype(np.int32)


def _threshold_table(quantizers) -> np.ndarray:
    return np.array([q.thresholds for q in quantizers], dtype=np.float64).astype(np.float32)


class RCQMinSumDecoder(EngineCache):
    """Min-sum whose C2V messages pass through quantise -> reconstruct (rcq_decoder.py:123-279)."""

    def __init__(self, code: LDPCCode, bc: int, bv: int, quantizer_params: List[Tuple[float, float]],
                 max_iterations: int = 50, layered: bool = False):
        self.code = code
        self.bc = bc
        self.bv = bv
        self.max_iterations = max_iterations
        self.layered = layered
        self.quantizers = [NonUniformQuantizer(bc, C, gamma) for C, gamma in quantizer_params]
        self._engines = {}

    def _get_quantizer(self, iteration: int) -> NonUniformQuantizer:
        return self.quantizers[int(_schedule(self.max_iterations, len(self.quantizers))[iteration])]

    def _engine(self, device: int) -> Engine:
        # the reference reads max_iterations, bc, layered and the quantisers at call time (rcq_decoder.py:169-208)
        thr = _threshold_table(self.quantizers)
        key = (device, bool(self.layered), int(self.max_iterations), int(self.bc), thr.tobytes())
        eng = self._engines.get(key)
        if eng is None:
            if self.max_iterations < 1:
                raise ValueError("max_iterations must be >= 1")
            eng = Engine(self.code.graph, dtype=np.float32, max_iterations=self.max_iterations, bc=self.bc,
                         thresholds=thr, quantizer_of_iter=_schedule(self.max_iterations, len(self.quantizers)),
                         schedule=1 if self.layered else 0, device=device)
            for old in [k for k in self._engines if k[:2] == key[:2]]:   # one engine per (device, schedule)
                self._engines.pop(old).close()
            self._engines[key] = eng
        return eng

    def decode(self, llr: torch.Tensor) -> Tuple[torch.Tensor, bool, int]:
        if not isinstance(llr, torch.Tensor):
            llr = torch.as_tensor(np.asarray(llr))
        single = llr.dim() == 1
        batch = llr[None] if single else llr
        if batch.device.type == "cuda":
            bits, _, iters, succ = self._engine(batch.device.index).decode_device(batch.to(torch.float32))
        else:
            arr = batch.detach().to(torch.float32).contiguous().numpy()
            b, _, i, s = self._engine(default_device()).decode_host(arr)
            bits, iters, succ = torch.from_numpy(b), torch.from_numpy(i), torch.from_numpy(s)
        decoded = bits.to(torch.int32)
        if single:
            return decoded[0], bool(succ[0].item()), int(iters[0].item())
        return decoded, succ.bool(), iters


class WeightedRCQDecoder(DecoderModule):
    """Degree-shared neural weights + RCQ quantisation (rcq_decoder.py:352-597); flooding only (the
    reference stores ``layered`` and never reads it, rcq_decoder.py:377)."""

    def __init__(self, code: LDPCCode, bc: int, bv: int, quantizer_params: List[Tuple[float, float]],
                 weight_sharing_type: int = 2, max_iterations: int = 50, layered: bool = False):
        super().__init__()
        self._init_base(code, max_iterations)
        self.bc = bc
        self.bv = bv
        self.weight_sharing_type = weight_sharing_type
        self.layered = layered
        self.quantizers = [NonUniformQuantizer(bc, C, gamma) for C, gamma in quantizer_params]
        # like the reference, an invalid sharing type builds no weights and only fails when called
        build_2d_tables(self, code, weight_sharing_type, max_iterations, validate=False)

    def _get_quantizer(self, iteration: int) -> NonUniformQuantizer:
        return self.quantizers[int(_schedule(self.max_iterations, len(self.quantizers))[iteration])]

    def _quant_config(self):
        return self.bc, _threshold_table(self.quantizers), _schedule(self.max_iterations, len(self.quantizers))

    def forward(self, llr: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, int]:
        if self.weight_sharing_type not in (1, 2, 3, 4):
            raise StopIteration  # rcq_decoder.py:435 ``next(self.parameters())`` on a weightless module
        return self._forward_impl(llr)
