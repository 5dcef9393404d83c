"""Shared host logic of the neural / RCQ decoder classes: dense weight tables with a reference-keyed
view, degree-class index maps, engine caching and the forward() plumbing."""
from __future__ import annotations

from collections import OrderedDict
from collections.abc import Mapping
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from .engine import Engine, default_device


class WeightView(Mapping):
    """Dict-like view that exposes one dense ``[T, W]`` parameter table under the reference's
    ``ParameterDict`` keys (e.g. ``iter_3_dc6``).  ``view[key]`` is a shape-[1] tensor sharing storage
    with the table, so in-place edits through the view reach the decoder."""

    def __init__(self, table: Optional[nn.Parameter], keys: List[str], index: Dict[str, Tuple[int, int]]):
        self._table = table
        self._keys = keys
        self._index = index

    def __getitem__(self, key: str) -> torch.Tensor:
        t, c = self._index[key]
        return self._table[t, c:c + 1]

    def __iter__(self):
        return iter(self._keys)

    def __len__(self) -> int:
        return len(self._keys)

    def __contains__(self, key) -> bool:
        return key in self._index


def seeded_normal(count: int, scale: float, shift: float) -> torch.Tensor:
    """``count`` draws of ``torch.randn(1) * scale + shift`` in call order.  Up to 2^17 draws are
    made one by one so that, under the same ``torch.manual_seed``, the values equal the reference's
    per-key initialisation (neural_2d_decoder.py:54, neural_minsum_decoder.py:53); beyond that a
    single vectorised draw is used."""
    if count <= (1 << 17):
        out = torch.empty(count)
        for i in range(count):
            out[i] = torch.randn(1)[0]
    else:
        out = torch.randn(count)
    return out * scale + shift


class DecoderModule(nn.Module):
    """Base of the nn.Module decoders.  Subclasses fill:
         self._beta_table / self._alpha_table : nn.Parameter [T, W] or None
         self._beta_index [E] / self._alpha_index [n] : int32 column maps (numpy) or None
         self._beta_const : float32 value used when there is no beta table (type 4)
    """

    def _init_base(self, code, max_iterations: int):
        self.code = code
        self.max_iterations = max_iterations
        self._engines = {}
        self._pushed = {}

    # ---- quantiser hooks (overridden by the RCQ classes) ----
    def _quant_config(self):
        return 0, None, None

    def _tables(self):
        T = self.max_iterations
        beta = self._beta_table
        if beta is not None:
            b = beta.detach().cpu().numpy().astype(np.float32)
        elif getattr(self, "_beta_const", None) is not None:
            b = np.full((T, 1), np.float32(self._beta_const), dtype=np.float32)
        else:
            b = None
        alpha = self._alpha_table
        a = alpha.detach().cpu().numpy().astype(np.float32) if alpha is not None else None
        return b, a

    def _versions(self):
        return tuple((p._version, p.data_ptr()) if p is not None else None
                     for p in (self._beta_table, self._alpha_table))

    def _engine(self, device: int) -> Engine:
        if self.max_iterations < 1:
            raise ValueError("max_iterations must be >= 1")
        eng = self._engines.get(device)
        if eng is None:
            b, a = self._tables()
            bc, thr, qoi = self._quant_config()
            eng = Engine(self.code.graph, dtype=np.float32, max_iterations=self.max_iterations,
                         beta=b, beta_index=self._beta_index if self._beta_table is not None else None,
                         alpha=a, alpha_index=self._alpha_index if a is not None else None,
                         bc=bc, thresholds=thr, quantizer_of_iter=qoi,
                         check_rule=getattr(self, "_check_rule", 0), device=device)
            self._engines[device] = eng
            self._pushed[device] = self._versions()
        elif self._pushed[device] != self._versions():
            b, a = self._tables()
            eng.set_weights(b if self._beta_table is not None else None, a)
            self._pushed[device] = self._versions()
        return eng

    def _run(self, llr, want_posterior: bool):
        """llr: torch tensor [n] or [B, n] (any device) -> (bits u8, posterior|None, iters i32, success u8),
        all torch tensors on llr.device, plus `single`."""
        if not isinstance(llr, torch.Tensor):
            llr = torch.as_tensor(np.asarray(llr))
        single = llr.dim() == 1
        batch = llr[None] if single else llr
        if batch.device.type == "cuda":
            eng = self._engine(batch.device.index)
            bits, post, iters, succ = eng.decode_device(batch.to(torch.float32), want_posterior=want_posterior)
        else:
            eng = self._engine(default_device())
            arr = batch.detach().to(torch.float32).contiguous().numpy()
            b, p, i, s = eng.decode_host(arr, want_posterior=want_posterior)
            bits, iters, succ = torch.from_numpy(b), torch.from_numpy(i), torch.from_numpy(s)
            post = torch.from_numpy(p) if p is not None else None
        return bits, post, iters, succ, single

    def _forward_impl(self, llr):
        bits, post, iters, _, single = self._run(llr, want_posterior=True)
        decoded = bits.to(torch.int32)  # (posterior < 0).int() in the reference
        if single:
            return decoded[0], post[0], int(iters[0].item())
        return decoded, post, iters

    # ---- checkpoint interchange with the reference's ParameterDict layout ----
    def reference_state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        """State dict keyed exactly like the reference module's (``beta_weights.iter_0_dc3`` ...)."""
        out = OrderedDict()
        for name in ("beta_weights", "alpha_weights"):
            view = getattr(self, name, None)
            if view is None:
                continue
            for key in view:
                out[f"{name}.{key}"] = view[key].detach().clone()
        return out

    def load_reference_state_dict(self, state: Mapping, strict: bool = True):
        """Load weights saved by the reference module's ``state_dict()``."""
        seen = set()
        with torch.no_grad():
            for full, value in state.items():
                name, _, key = full.partition(".")
                view = getattr(self, name, None)
                if view is None or key not in view:
                    if strict:
                        raise KeyError(f"unexpected key {full}")
                    continue
                view[key].copy_(torch.as_tensor(value).reshape(1))
                seen.add(full)
        if strict:
            missing = [k for k in self.reference_state_dict() if k not in seen]
            if missing:
                raise KeyError(f"missing keys: {missing[:5]}{'...' if len(missing) > 5 else ''}")


def degree_lists(code):
    """``list(set(...))`` exactly as neural_2d_decoder.py:34-35 builds them (same element order)."""
    cnd = list(set(code.check_node_degrees.values()))
    vnd = list(set(code.variable_node_degrees.values()))
    return cnd, vnd


def build_2d_tables(module: DecoderModule, code, weight_sharing_type: int, T: int, validate: bool):
    """Create the dense tables + key views + index maps for the four sharing types
    (neural_2d_decoder.py:46-131 / rcq_decoder.py:398-480):
        1: beta[t, dc, dv]            alpha = 1
        2: beta[t, dc], alpha[t, dv]
        3: beta[t, dc]                alpha = 1
        4: beta = float32(0.7)        alpha[t, dv]
    Initialisation draws follow the reference's creation order so equal seeds give equal weights."""
    g = code.graph
    cnd, vnd = degree_lists(code)
    module.check_node_degrees = cnd
    module.variable_node_degrees = vnd
    ci = {d: i for i, d in enumerate(cnd)}
    vi = {d: i for i, d in enumerate(vnd)}
    # weight lookups use H's row/column SUMS (code.check_node_degrees), which equal the edge counts
    # for 0/1 matrices; keep the reference's definition.
    cdeg = np.array([code.check_node_degrees[i] for i in range(g.m)], dtype=np.int64) if g.m else np.zeros(0, np.int64)
    vdeg = np.array([code.variable_node_degrees[j] for j in range(g.n)], dtype=np.int64)
    e_ci = np.array([ci[int(d)] for d in cdeg[g.edge_check]], dtype=np.int32) if g.E else np.zeros(0, np.int32)
    e_vi = np.array([vi[int(d)] for d in vdeg[g.check_var]], dtype=np.int32) if g.E else np.zeros(0, np.int32)
    v_vi = np.array([vi[int(d)] for d in vdeg], dtype=np.int32)
    Ndc, Ndv = len(cnd), len(vnd)
    module._beta_table = None
    module._alpha_table = None
    module._beta_index = None
    module._alpha_index = None
    module._beta_const = None
    bkeys, bidx, akeys, aidx = [], {}, [], {}
    wt = weight_sharing_type
    if wt == 1:
        draws = seeded_normal(T * Ndc * Ndv, 0.1, 0.0)
        module._beta_table = nn.Parameter(draws.reshape(T, Ndc * Ndv).clone())
        for t in range(T):
            for a, dc in enumerate(cnd):
                for b, dv in enumerate(vnd):
                    key = f"iter_{t}_dc{dc}_dv{dv}"
                    bkeys.append(key)
                    bidx[key] = (t, a * Ndv + b)
        module._beta_index = (e_ci.astype(np.int64) * Ndv + e_vi).astype(np.int32)
    elif wt == 2:
        draws = seeded_normal(T * (Ndc + Ndv), 0.1, 0.0).reshape(T, Ndc + Ndv)
        module._beta_table = nn.Parameter(draws[:, :Ndc].clone())
        module._alpha_table = nn.Parameter(draws[:, Ndc:].clone())
        module._beta_index = e_ci
        module._alpha_index = v_vi
    elif wt == 3:
        module._beta_table = nn.Parameter(seeded_normal(T * Ndc, 0.1, 0.0).reshape(T, Ndc).clone())
        module._beta_index = e_ci
    elif wt == 4:
        module._alpha_table = nn.Parameter(seeded_normal(T * Ndv, 0.1, 0.0).reshape(T, Ndv).clone())
        module._alpha_index = v_vi
        module._beta_const = np.float32(0.7)   # torch.tensor(0.7): neural_2d_decoder.py:104
    elif validate:
        raise ValueError(f"Invalid weight sharing type: {weight_sharing_type}")
    if wt in (2, 3):
        for t in range(T):
            for a, dc in enumerate(cnd):
                key = f"iter_{t}_dc{dc}"
                bkeys.append(key)
                bidx[key] = (t, a)
    if wt in (2, 4):
        for t in range(T):
            for b, dv in enumerate(vnd):
                key = f"iter_{t}_dv{dv}"
                akeys.append(key)
                aidx[key] = (t, b)
    module.beta_weights = WeightView(module._beta_table, bkeys, bidx)
    module.alpha_weights = WeightView(module._alpha_table, akeys, aidx)
