"""Neural min-sum decoder with one weight per (iteration, edge) -- B200-native counterpart of the
reference's ``neural_minsum_decoder.py:19-150``.

``beta_weights`` exposes the dense ``[T, E]`` table under the reference keys ``iter_{t}_c{i}_v{j}``
(initialised ``0.7 + 0.1 * randn`` in the reference's creation order, neural_minsum_decoder.py:47-53)."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn as nn

from ._neural_base import DecoderModule, WeightView, seeded_normal
from .ldpc_decoder import LDPCCode


class NeuralMinSumDecoder(DecoderModule):
    def __init__(self, code: LDPCCode, max_iterations: int = 50):
        super().__init__()
        self._init_base(code, max_iterations)
        g = code.graph
        T, E = max_iterations, g.E
        self._beta_table = nn.Parameter(seeded_normal(T * E, 0.1, 0.7).reshape(T, E).clone())
        self._alpha_table = None
        self._beta_index = np.arange(E, dtype=np.int32)   # check-major edge order == creation order
        self._alpha_index = None
        self._beta_const = None
        self._edge_check = g.edge_check
        self._edge_var = g.check_var
        self.beta_weights = _EdgeKeyView(self._beta_table, T, g)

    def forward(self, llr: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, int]:
        return self._forward_impl(llr)


class NeuralOffsetMinSumDecoder(DecoderModule):
    """Offset min-sum with one offset per (iteration, edge) (neural_minsum_decoder.py:152-286):
    ``c2v = prod(other signs) * relu(raw - beta)``, offsets initialised ``0.1 * randn`` (:185)."""

    _check_rule = 1  # LDPC_RULE_OFFSET

    def __init__(self, code: LDPCCode, max_iterations: int = 50):
        super().__init__()
        self._init_base(code, max_iterations)
        g = code.graph
        T, E = max_iterations, g.E
        self._beta_table = nn.Parameter(seeded_normal(T * E, 0.1, 0.0).reshape(T, E).clone())
        self._alpha_table = None
        self._beta_index = np.arange(E, dtype=np.int32)
        self._alpha_index = None
        self._beta_const = None
        self.beta_weights = _EdgeKeyView(self._beta_table, T, g)

    def forward(self, llr: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, int]:
        return self._forward_impl(llr)


class _EdgeKeyView(WeightView):
    """Key view for T*E per-edge weights without materialising T*E strings up front."""

    def __init__(self, table, T, graph):
        self._table = table
        self._T = T
        self._g = graph
        self._edge = None

    def _edges(self):
        if self._edge is None:
            self._edge = {(int(i), int(j)): e for e, (i, j) in enumerate(zip(self._g.edge_check, self._g.check_var))}
        return self._edge

    def _parse(self, key: str):
        try:
            it, c, v = key.split("_")[1:]
            return int(it), self._edges()[(int(c[1:]), int(v[1:]))]
        except (ValueError, KeyError, IndexError):
            raise KeyError(key)

    def __getitem__(self, key):
        t, e = self._parse(key)
        if not 0 <= t < self._T:
            raise KeyError(key)
        return self._table[t, e:e + 1]

    def __iter__(self):
        for t in range(self._T):
            for i, j in zip(self._g.edge_check, self._g.check_var):
                yield f"iter_{t}_c{int(i)}_v{int(j)}"

    def position(self, key):
        try:
            t, e = self._parse(key)
        except KeyError:
            return None
        return (t, e) if 0 <= t < self._T else None

    def items_index(self):
        for t in range(self._T):
            for e, (i, j) in enumerate(zip(self._g.edge_check, self._g.check_var)):
                yield f"iter_{t}_c{int(i)}_v{int(j)}", (t, e)

    def __len__(self):
        return self._T * self._g.E

    def __contains__(self, key):
        try:
            t, _ = self._parse(key)
            return 0 <= t < self._T
        except KeyError:
            return False


class LegacyNeuralMinSumDecoder(NeuralMinSumDecoder):
    """The reference's second class of the same name, ``ldpc_decoder.NeuralMinSumDecoder``
    (ldpc_decoder.py:155-272): the same decoder with weights initialised ``0.1 * randn`` (:173) and an
    ``alpha_weights`` dict that stays empty (:165).  Exported as ``ldpc_decoder.NeuralMinSumDecoder``."""

    def __init__(self, code: LDPCCode, max_iterations: int = 50):
        DecoderModule.__init__(self)
        self._init_base(code, max_iterations)
        g = code.graph
        T, E = max_iterations, g.E
        self._beta_table = nn.Parameter(seeded_normal(T * E, 0.1, 0.0).reshape(T, E).clone())
        self._alpha_table = None
        self._beta_index = np.arange(E, dtype=np.int32)
        self._alpha_index = None
        self._beta_const = None
        self.beta_weights = _EdgeKeyView(self._beta_table, T, g)
        self.alpha_weights = WeightView(None, [], {})


def analyze_weight_patterns(decoder: NeuralMinSumDecoder, code: LDPCCode) -> Dict:
    """Weight statistics of an N-NMS decoder (neural_minsum_decoder.py:288-349), same result layout:
    ``iteration_patterns[t]`` = mean / std / min / max of iteration t's weights, and
    ``node_degree_correlations['check_degree_<dc>']`` = mean / std / count of the per-edge averages over the
    iterations, for the edges of checks of degree dc.  Pure host code over the dense ``[T, E]`` table."""
    analysis = {'weight_statistics': {}, 'iteration_patterns': {}, 'node_degree_correlations': {}}
    table = decoder._beta_table.detach().cpu().numpy().astype(np.float64)[:decoder.max_iterations]
    g = code.graph
    if table.shape[1] == 0:
        return analysis
    for t in range(table.shape[0]):
        w = table[t]
        analysis['iteration_patterns'][t] = {'mean': np.mean(w), 'std': np.std(w), 'min': np.min(w), 'max': np.max(w)}
    check_degrees = code.check_node_degrees
    edge_dc = np.array([check_degrees[int(i)] for i in g.edge_check])
    per_edge = table.mean(axis=0)
    for dc in set(check_degrees.values()):
        w = per_edge[edge_dc == dc]
        if w.size:
            analysis['node_degree_correlations'][f'check_degree_{dc}'] = {'mean': np.mean(w), 'std': np.std(w), 'count': int(w.size)}
    return analysis
