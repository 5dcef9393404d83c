"""Shared host logic of the neural / RCQ decoder classes: dense weight tables with a reference-keyed
view, degree-class index maps, engine caching and the forward() plumbing."""
from __future__ import annotations

from collections import OrderedDict
from collections.abc import Mapping
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from .engine import Engine, EngineCache, default_device


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

    def position(self, key: str) -> Optional[Tuple[int, int]]:
        """(row, column) of ``key`` in the table, or None."""
        return self._index.get(key)

    def items_index(self):
        """(key, (row, column)) in the reference's creation order."""
        return ((k, self._index[k]) for k in self._keys)


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


class DecoderModule(EngineCache, nn.Module):
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
        # forward() under autograd is differentiable with respect to the weights only when asked for (the trainer
        # classes switch it on): the training forward keeps 2*T*E floats per frame, which inference batches of
        # tens of thousands of frames must never pay just because torch's grad mode is on by default
        self.differentiable = False
        # state_dict() / load_state_dict() speak the reference's ParameterDict layout (see the hooks below)
        self._register_state_dict_hook(_export_reference_keys)
        self._register_load_state_dict_pre_hook(_import_reference_keys, with_module=True)

    # ---- quantiser hooks (overridden by the RCQ classes) ----
    def _quant_config(self):
        return 0, None, None

    def _tables(self):
        """float32 host copies of the weight tables, cut to the current ``max_iterations`` rows.  The reference
        reads ``max_iterations`` at call time and looks its weights up by key: fewer iterations than the tables
        hold simply use the first rows, more fail with the missing key."""
        T = self.max_iterations

        def rows(table):
            if table.shape[0] < T:
                raise KeyError(f"iter_{table.shape[0]}_...: the decoder was built with {table.shape[0]} iterations of "
                               f"weights, max_iterations is now {T}")
            return np.array(table.detach().cpu().numpy()[:T], dtype=np.float32, order="C", copy=True)   # a snapshot, not an alias

        beta = self._beta_table
        if beta is not None:
            b = rows(beta)
        elif getattr(self, "_beta_const", None) is not None:
            b = np.full((T, 1), np.float32(self._beta_const), dtype=np.float32)
        else:
            b = None
        alpha = self._alpha_table
        a = rows(alpha) if alpha is not None else None
        return b, a

    def _engine_key(self, device: int):
        bc, thr, _ = self._quant_config()
        return (device, int(self.max_iterations), int(bc), thr.tobytes() if thr is not None else None)

    def _engine(self, device: int) -> Engine:
        """The device engine for the module's CURRENT configuration.  Everything the reference reads at call time is
        re-read here: ``max_iterations``, ``bc`` and the quantiser thresholds select the engine, and the weight
        tables are compared BY CONTENT with what the device holds (in-place edits through ``param.data`` or a numpy
        alias do not bump a tensor's version counter), and pushed again when they differ."""
        if self.max_iterations < 1:
            raise ValueError("max_iterations must be >= 1")
        key = self._engine_key(device)
        b, a = self._tables()
        eng = self._engines.get(key)
        if eng is None:
            bc, thr, qoi = self._quant_config()
            eng = Engine(self.code.graph, dtype=np.float32, max_iterations=self.max_iterations,
                         beta=b, beta_index=self._beta_index if self._beta_table is not None else None,
                         alpha=a, alpha_index=self._alpha_index if a is not None else None,
                         bc=bc, thresholds=thr, quantizer_of_iter=qoi,
                         check_rule=getattr(self, "_check_rule", 0), device=device)
            # one engine per device: a changed configuration replaces it (its workspace goes with it)
            for old in [k for k in self._engines if k[0] == device]:
                self._engines.pop(old).close()
                self._pushed.pop(old, None)
            self._engines[key] = eng
            self._pushed[key] = (b, a)
        else:
            pb, pa = self._pushed[key]
            new_b = b is not None and self._beta_table is not None and not np.array_equal(b, pb)
            new_a = a is not None and not np.array_equal(a, pa)
            if new_b or new_a:
                eng.set_weights(b if new_b else None, a if new_a else None)
                self._pushed[key] = (b, a)
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

    def _trainable(self) -> bool:
        """Posterior training exists for the normalised float rule (N-NMS, N-2D-NMS): the quantiser of W-RCQ passes no
        gradient in the reference either, and the offset rule's backward pass is not built."""
        if not self.differentiable or getattr(self, "_check_rule", 0) != 0 or self._quant_config()[0] != 0:
            return False
        return any(p is not None and p.requires_grad for p in (self._beta_table, self._alpha_table))

    def _forward_impl(self, llr):
        if torch.is_grad_enabled() and self._trainable():
            return self._forward_train(llr)
        bits, post, iters, _, single = self._run(llr, want_posterior=True)
        decoded = bits.to(torch.int32)  # (posterior < 0).int() in the reference
        if single:
            return decoded[0], post[0], int(iters[0].item())
        return decoded, post, iters

    def _forward_train(self, llr):
        """forward() under autograd: the same decode, with ``posterior`` differentiable with respect to the weight
        tables (the reference's forward is differentiable through torch's autograd; ours through the backward
        kernels of csrc/ldpc_train.cu)."""
        if not isinstance(llr, torch.Tensor):
            llr = torch.as_tensor(np.asarray(llr))
        single = llr.dim() == 1
        batch = (llr[None] if single else llr).detach()
        home = batch.device
        dev = home if home.type == "cuda" else torch.device("cuda", default_device())
        self._engine(dev.index)     # no device -> LdpcError here, like the inference path (there is no CPU fallback)
        bits, post, iters = _PosteriorTraining.apply(self, batch.to(dev, torch.float32), self._beta_table, self._alpha_table)
        bits, post, iters = bits.to(home), post.to(home), iters.to(home)
        decoded = bits.to(torch.int32)
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


class _PosteriorTraining(torch.autograd.Function):
    """Decode with the message history kept on the device; backward = ldpc_train_backward."""

    @staticmethod
    def forward(ctx, module, llr, beta_table, alpha_table):
        eng = module._engine(llr.device.index)          # pushes the current weights (compared by content)
        bits, post, iters, _ = eng.train_forward(llr)
        ctx.eng, ctx.serial = eng, eng._train_serial
        ctx.shapes = tuple((None if t is None else (tuple(t.shape), t.device, t.dtype)) for t in (beta_table, alpha_table))
        ctx.mark_non_differentiable(bits, iters)
        return bits, post, iters

    @staticmethod
    def backward(ctx, _gbits, gpost, _giters):
        eng = ctx.eng
        if eng._train_serial != ctx.serial:
            raise RuntimeError("the decoder ran another training forward pass before this backward pass: its message "
                               "history was overwritten")
        gb, ga = eng.train_backward(gpost)
        out = []
        for g, spec in zip((gb, ga), ctx.shapes):
            if spec is None or g is None or not ctx.needs_input_grad[2 + len(out)]:
                out.append(None)
                continue
            shape, device, dtype = spec
            full = torch.zeros(shape, dtype=dtype, device=device)   # rows beyond max_iterations were not used
            full[:g.shape[0]] = g.to(device=device, dtype=dtype)
            out.append(full)
        return None, None, out[0], out[1]


def _weight_views(module):
    for name, table_name in (("beta_weights", "_beta_table"), ("alpha_weights", "_alpha_table")):
        view = getattr(module, name, None)
        if isinstance(view, WeightView) and getattr(module, table_name, None) is not None:
            yield name, table_name, view


def _export_reference_keys(module, state_dict, prefix, local_metadata):
    """state_dict() hook: replace the dense tables by the reference's ``ParameterDict`` entries
    (``beta_weights.iter_0_dc3`` ... each a shape-[1] tensor sharing storage with the table, in the reference's
    creation order), so that ``reference_module.load_state_dict(ours.state_dict())`` works
    (neural_2d_decoder.py:46-82, neural_minsum_decoder.py:47-53, rcq_decoder.py:398-431)."""
    for name, table_name, view in _weight_views(module):
        table = state_dict.pop(prefix + table_name, None)
        if table is None:
            continue
        for key, (t, c) in view.items_index():
            state_dict[f"{prefix}{name}.{key}"] = table[t, c:c + 1]
    return state_dict


def _import_reference_keys(module, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
    """load_state_dict() pre-hook: fold ``beta_weights.<key>`` / ``alpha_weights.<key>`` entries (a reference
    module's state_dict, or our own) into the dense tables.  Unknown keys stay in ``state_dict`` and are reported
    as unexpected; absent ones are reported as missing (both only matter under ``strict``).  The native
    ``_beta_table`` / ``_alpha_table`` entries are accepted as well."""
    for name, table_name, view in _weight_views(module):
        head = f"{prefix}{name}."
        keys = [k for k in state_dict if k.startswith(head)]
        if not keys:
            continue
        table = getattr(module, table_name).detach().clone()
        seen = 0
        for full in keys:
            pos = view.position(full[len(head):])
            if pos is None:
                continue                      # unexpected: left for load_state_dict to report
            value = torch.as_tensor(state_dict.pop(full))
            if value.numel() != 1:
                error_msgs.append(f"size mismatch for {full}: expected one element, got {tuple(value.shape)}")
                continue
            table[pos[0], pos[1]] = value.reshape(()).to(table.dtype)
            seen += 1
        if seen != len(view):
            have = {full[len(head):] for full in keys}
            missing_keys.extend(f"{head}{k}" for k in view if k not in have)
        state_dict[prefix + table_name] = table


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
