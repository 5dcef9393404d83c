"""Neural 2-D min-sum decoders with node-degree-based weight sharing -- B200-native counterparts of the
reference's ``neural_2d_decoder.py`` (``Neural2DMinSumDecoder`` :16-225, ``Neural2DOffsetMinSumDecoder``
:227-434; same constructors, attributes and return triples).

forward(llr) -> (decoded int32, posterior float32, iterations int); a batch ``[B, n]`` returns
``([B, n], [B, n], [B] int32)``.  Weights live in dense ``[T, W]`` parameter tables; ``beta_weights`` /
``alpha_weights`` expose them under the reference's ParameterDict keys (``iter_{t}_dc{dc}_dv{dv}``,
``iter_{t}_dc{dc}``, ``iter_{t}_dv{dv}``) and ``load_reference_state_dict`` imports a reference
checkpoint.  The decode itself runs in the CUDA library; there is no backward pass (training is out
of scope)."""
from __future__ import annotations

from typing import Tuple

import torch

from ._neural_base import DecoderModule, build_2d_tables
from .ldpc_decoder import LDPCCode


class Neural2DMinSumDecoder(DecoderModule):
    """Weight sharing types (neural_2d_decoder.py:19-25):
    1: one beta per (check degree, variable degree); 2: beta per check degree and alpha per variable
    degree; 3: beta per check degree only; 4: alpha per variable degree only (beta = 0.7)."""

    def __init__(self, code: LDPCCode, weight_sharing_type: int = 2, max_iterations: int = 50):
        super().__init__()
        self._init_base(code, max_iterations)
        self.weight_sharing_type = weight_sharing_type
        build_2d_tables(self, code, weight_sharing_type, max_iterations, validate=True)

    def forward(self, llr: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, int]:
        return self._forward_impl(llr)


class Neural2DOffsetMinSumDecoder(DecoderModule):
    """Offset min-sum with degree-shared offsets (neural_2d_decoder.py:227-434):
    ``c2v = prod(other signs) * (relu(raw - beta) - alpha)`` with beta by check degree (types 1-3; type 1
    also by variable degree) and alpha by the edge's VARIABLE degree (types 2, 4), both applied at the
    check node; missing ones are 0; variable-node sums are unweighted (:403-410)."""

    _check_rule = 1  # LDPC_RULE_OFFSET

    def __init__(self, code: LDPCCode, weight_sharing_type: int = 2, max_iterations: int = 50):
        super().__init__()
        self._init_base(code, max_iterations)
        self.weight_sharing_type = weight_sharing_type
        build_2d_tables(self, code, weight_sharing_type, max_iterations, validate=True)
        self._beta_const = None   # type 4: beta = 0.0 (neural_2d_decoder.py:313), i.e. no subtraction

    def forward(self, llr: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, int]:
        return self._forward_impl(llr)
