"""CPU restatement of the reference's flooding min-sum / RCQ hot path (TEST INFRASTRUCTURE).

This is the oracle, not the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product path
(the CUDA library behind ``include/ldpc_b200.h``) never routes through this file.

Parity status: the reference has no golden vectors of its own, so this
restatement is PINNED against outputs of the live, unmodified reference run in the build container
(``oracle/ref_shim.py``) -- committed as ``tests/golden/*.npz`` together with the generating script
``tests/golden/make_golden.py`` -- and against the known-answer vectors of SURVEY.md appendix B.

It is a *sparse* restatement: same arithmetic, same operation order, but the Tanner graph is walked
through adjacency lists instead of ``np.where`` scans over a dense ``H``, and frames are vectorised
with numpy (element-wise IEEE float32/float64 ops are identical whether done one frame at a time or
many).  Every function cites the reference lines it follows.

Reference lines followed:
  BasicMinSumDecoder.decode          ldpc_decoder.py:63-153
  NeuralMinSumDecoder.forward        neural_minsum_decoder.py:58-150
  Neural2DMinSumDecoder.forward      neural_2d_decoder.py:133-225   (weights :84-131)
  NonUniformQuantizer                rcq_decoder.py:48-121
  RCQMinSumDecoder._decode_flooding  rcq_decoder.py:190-279         (schedule :156-167)
  WeightedRCQDecoder.forward         rcq_decoder.py:495-597
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np


# --------------------------------------------------------------------------------------------
# Graph: adjacency in the reference's order
# --------------------------------------------------------------------------------------------
@dataclass
class SparseGraph:
    n: int
    m: int
    check_ptr: np.ndarray  # [m+1] int64, CSR by check
    check_var: np.ndarray  # [E] int32, ascending variable index inside a check
    var_ptr: np.ndarray    # [n+1] int64, CSR by variable
    var_edge: np.ndarray   # [E] int64: edge ids (check-major numbering) in ascending check index
    var_chk: np.ndarray    # [E] int32: the check of each entry of var_edge

    @property
    def E(self) -> int:
        return int(self.check_var.shape[0])

    @staticmethod
    def from_dense(H: np.ndarray) -> "SparseGraph":
        """Only entries == 1 are edges, exactly like ``np.where(H[i, :] == 1)``
        (ldpc_decoder.py:92,124)."""
        H = np.asarray(H)
        m, n = H.shape
        rows, cols = np.nonzero(H == 1)  # row-major order -> ascending var inside each check
        return SparseGraph.from_coo(n, m, rows, cols)

    @staticmethod
    def from_coo(n: int, m: int, rows: np.ndarray, cols: np.ndarray) -> "SparseGraph":
        rows = np.asarray(rows, dtype=np.int64)
        cols = np.asarray(cols, dtype=np.int64)
        order = np.lexsort((cols, rows))
        rows, cols = rows[order], cols[order]
        E = rows.shape[0]
        check_ptr = np.zeros(m + 1, dtype=np.int64)
        np.add.at(check_ptr, rows + 1, 1)
        check_ptr = np.cumsum(check_ptr)
        vorder = np.lexsort((rows, cols))  # by variable, then ascending check
        var_ptr = np.zeros(n + 1, dtype=np.int64)
        np.add.at(var_ptr, cols + 1, 1)
        var_ptr = np.cumsum(var_ptr)
        return SparseGraph(n=n, m=m, check_ptr=check_ptr, check_var=cols.astype(np.int32),
                           var_ptr=var_ptr, var_edge=vorder.astype(np.int64),
                           var_chk=rows[vorder].astype(np.int32))

    def check_degrees(self) -> np.ndarray:
        return np.diff(self.check_ptr).astype(np.int64)

    def var_degrees(self) -> np.ndarray:
        return np.diff(self.var_ptr).astype(np.int64)

    def edge_check(self) -> np.ndarray:
        return np.repeat(np.arange(self.m, dtype=np.int64), np.diff(self.check_ptr))


# --------------------------------------------------------------------------------------------
# Library reduction orders -- these decide hard decisions near zero
# --------------------------------------------------------------------------------------------
def torch_sum_f32(xs: Sequence[np.ndarray]) -> np.ndarray:
    """Order of ``torch.sum`` on a contiguous float32 k-vector (torch 2.11 CPU), applied
    element-wise to k arrays.  neural_2d_decoder.py:203,209 / rcq_decoder.py:257,263 call it on
    ``c2v_messages[neighbors, j]``.

    k <= 7 : four accumulators over full groups of 4, leftovers into acc[0], ((a0+a1)+a2)+a3.
    k >= 8 : floor(k/8) 8-lane vectors reduced lane-wise with the same 4-accumulator scheme, then
             r = 0 + tail scalars in order, then r += lane_0 .. lane_7.
    """
    k = len(xs)
    f32 = np.float32
    zero = np.zeros_like(np.asarray(xs[0], dtype=f32)) if k else f32(0.0)
    if k == 0:
        return f32(0.0)
    xs = [np.asarray(x, dtype=f32) for x in xs]

    def four_acc(items, z):
        acc = [z, z, z, z]
        g = len(items) // 4
        for i in range(g):
            for q in range(4):
                acc[q] = acc[q] + items[4 * i + q]
        for r in range(4 * g, len(items)):
            acc[0] = acc[0] + items[r]
        return ((acc[0] + acc[1]) + acc[2]) + acc[3]

    if k < 8:
        return four_acc(xs, zero)
    nv = k // 8
    lanes = []
    for lane in range(8):
        lanes.append(four_acc([xs[8 * v + lane] for v in range(nv)], zero))
    r = zero
    for t in range(8 * nv, k):
        r = r + xs[t]
    for lane in range(8):
        r = r + lanes[lane]
    return r


def np_sum_f64(xs: Sequence[np.ndarray]) -> np.ndarray:
    """Order of ``np.sum`` on a contiguous float64 k-vector (numpy pairwise, k <= 128), used by
    ldpc_decoder.py:131,137."""
    k = len(xs)
    if k == 0:
        return np.float64(0.0)
    xs = [np.asarray(x, dtype=np.float64) for x in xs]
    if k < 8:
        r = np.zeros_like(xs[0])
        for x in xs:
            r = r + x
        return r
    assert k <= 128, "pairwise blocks beyond 128 terms are not modelled"
    r = [xs[q] for q in range(8)]
    full = k - (k % 8)
    for i in range(8, full, 8):
        for q in range(8):
            r[q] = r[q] + xs[i + q]
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    for i in range(full, k):
        res = res + xs[i]
    return res


# --------------------------------------------------------------------------------------------
# Quantiser (rcq_decoder.py:48-121)
# --------------------------------------------------------------------------------------------
def quantizer_thresholds(bc: int, C: float, gamma: float) -> List[float]:
    """rcq_decoder.py:48-57 -- Python doubles, exactly the same expression."""
    max_idx = 2 ** (bc - 1) - 1
    return [C * (j / (2 ** (bc - 1) - 1)) ** gamma for j in range(max_idx + 1)]


def quantizer_schedule(T: int, Q: int) -> np.ndarray:
    """rcq_decoder.py:156-167 -- which quantiser each iteration uses."""
    out = np.zeros(T, dtype=np.int32)
    for t in range(T):
        if Q == 1:
            out[t] = 0
        elif t < T // 3:
            out[t] = 0
        elif t < 2 * T // 3:
            out[t] = 1 if Q > 1 else 0
        else:
            out[t] = Q - 1
    return out


def quantize(x: np.ndarray, thresholds_f32: np.ndarray, bc: int) -> np.ndarray:
    """rcq_decoder.py:59-91.  Comparison is float32 against float32(threshold); index = last j whose
    test passes (the reference's loop overwrites in order); sign bit = (x < 0) strictly."""
    x = np.asarray(x, dtype=np.float32)
    mag = np.abs(x)
    idx = np.zeros(x.shape, dtype=np.int64)
    for j, th in enumerate(thresholds_f32):
        idx = np.where(mag >= np.float32(th), j, idx)
    idx = np.where(mag >= np.float32(thresholds_f32[-1]), 2 ** (bc - 1) - 1, idx)
    sign_bit = (np.sign(x) < 0).astype(np.int64)
    return sign_bit * (2 ** (bc - 1)) + idx


def dequantize(q: np.ndarray, thresholds_f32: np.ndarray, bc: int) -> np.ndarray:
    """rcq_decoder.py:93-121.  Reconstruction value = +-float32(threshold[idx]) (lower bin edge)."""
    q = np.asarray(q, dtype=np.int64)
    sign_bit = (q >= 2 ** (bc - 1)).astype(np.float32)
    idx = q % (2 ** (bc - 1))
    mag = np.asarray(thresholds_f32, dtype=np.float32)[idx]
    return ((np.float32(1.0) - np.float32(2.0) * sign_bit) * mag).astype(np.float32)


# --------------------------------------------------------------------------------------------
# The decoder restatement
# --------------------------------------------------------------------------------------------
MODE_NMS = 0    # c2v = (beta * raw) * sp            Basic / N-NMS / N-2D   (neural_2d_decoder.py:188-191)
MODE_RCQ = 1    # c2v = Qinv(Q(sp * raw))            RCQ                    (rcq_decoder.py:242-246)
MODE_WRCQ = 2   # c2v = Qinv(Q((beta * sp) * raw))   W-RCQ                  (rcq_decoder.py:559-563)
MODE_OFFSET = 3 # c2v = sp * (relu(raw - beta) - alpha_j)   N-OMS / N-2D-OMS (neural_minsum_decoder.py:252-253,
                #                                            neural_2d_decoder.py:400-401); VN sums are unweighted


@dataclass
class OracleResult:
    bits: np.ndarray        # [B, n] uint8
    posterior: np.ndarray   # [B, n] float32 / float64
    iterations: np.ndarray  # [B] int32
    success: np.ndarray     # [B] bool
    c2v_codes: Optional[np.ndarray] = None  # [T_executed..] not kept; see decode(..., trace=)


def decode(graph: SparseGraph, llr: np.ndarray, *, T: int, mode: int = MODE_NMS,
           dtype=np.float32, beta: Optional[np.ndarray] = None, alpha: Optional[np.ndarray] = None,
           bc: int = 0, thresholds: Optional[np.ndarray] = None,
           quantizer_of_iter: Optional[np.ndarray] = None, early_stop: bool = True,
           trace: Optional[list] = None) -> OracleResult:
    """Flooding decode of a batch ``llr[B, n]``.

    beta  : [T, E] per-edge check-side weight in check-major edge order (None -> 1)
    alpha : [T, n] per-variable weight (None -> no alpha multiply, as for Basic / N-NMS / RCQ)
    thresholds : [Q, 2^(bc-1)] float32-rounded thresholds, quantizer_of_iter : [T]
    trace : optional list; per executed iteration appends the [E, B] integer RCQ codes (or c2v floats)
    """
    dt = np.dtype(dtype)
    llr = np.ascontiguousarray(np.asarray(llr, dtype=dt))
    if llr.ndim == 1:
        llr = llr[None, :]
    B, n = llr.shape
    assert n == graph.n
    E, m = graph.E, graph.m
    cp, cv = graph.check_ptr, graph.check_var
    vp, ve = graph.var_ptr, graph.var_edge
    ssum = torch_sum_f32 if dt == np.float32 else np_sum_f64
    llrT = np.ascontiguousarray(llr.T)  # [n, B]

    # message state before iteration 0 (ldpc_decoder.py:80-87): v2c = llr on every edge, c2v = 0
    v2c = llrT[cv].copy()               # [E, B]
    c2v = np.zeros((E, B), dtype=dt)

    done = np.zeros(B, dtype=bool)
    iters = np.full(B, T, dtype=np.int32)
    success = np.zeros(B, dtype=bool)
    out_bits = np.zeros((B, n), dtype=np.uint8)
    out_post = np.zeros((B, n), dtype=dt)
    ar = np.arange(B)

    for t in range(T):
        act = ~done
        if not act.any():
            break
        if bc:
            th = np.asarray(thresholds[int(quantizer_of_iter[t])], dtype=np.float32)
        new_c2v = c2v.copy()
        codes = np.zeros((E, B), dtype=np.int64) if (trace is not None and bc) else None
        # ---- check node update (ldpc_decoder.py:91-120, neural_2d_decoder.py:161-191) ----
        for i in range(m):
            e0, e1 = int(cp[i]), int(cp[i + 1])
            dc = e1 - e0
            if dc == 0:
                continue
            inc = v2c[e0:e1]                       # [dc, B]
            signs = np.sign(inc)                    # three-valued
            mags = np.abs(inc)
            k0 = np.argmin(mags, axis=0)            # first index of the minimum
            m1 = mags[k0, ar]
            if dc > 1:
                tmp = mags.copy()
                tmp[k0, ar] = np.inf
                m2 = tmp.min(axis=0)
            else:
                m2 = m1
            for k in range(dc):
                others = [signs[q] for q in range(dc) if q != k]
                sp = np.ones(B, dtype=dt)
                for s in others:                    # prod of exact {-1,0,1}: order-free
                    sp = sp * s
                raw = np.where(k0 == k, m2, m1)
                e = e0 + k
                if mode == MODE_NMS:
                    b = dt.type(1.0) if beta is None else dt.type(beta[t, e])
                    val = (b * raw).astype(dt) * sp
                elif mode == MODE_OFFSET:
                    b = dt.type(0.0) if beta is None else dt.type(beta[t, e])
                    off = np.maximum((raw - b).astype(dt), dt.type(0.0))          # F.relu(raw_msg - beta)
                    if alpha is not None:
                        off = (off - dt.type(alpha[t, int(cv[e])])).astype(dt)    # ... - alpha (2-D variant)
                    val = sp * off
                else:
                    if mode == MODE_RCQ:
                        x = (sp * raw).astype(np.float32)
                    else:
                        b = np.float32(beta[t, e])
                        x = ((b * sp).astype(np.float32) * raw).astype(np.float32)
                    q = quantize(x, th, bc)
                    if codes is not None:
                        codes[e] = q
                    val = dequantize(q, th, bc)
                new_c2v[e] = val.astype(dt)
        c2v = np.where(act[None, :], new_c2v, c2v)
        if trace is not None:
            trace.append(codes if codes is not None else c2v.copy())
        # ---- variable node update + posterior (ldpc_decoder.py:123-137) ----
        new_v2c = v2c.copy()
        post = llrT.copy()
        for j in range(n):
            es = ve[int(vp[j]):int(vp[j + 1])]
            dv = es.shape[0]
            if dv == 0:
                continue
            msgs = [c2v[int(e)] for e in es]
            for d in range(dv):
                s = ssum([msgs[q] for q in range(dv) if q != d])
                if alpha is not None and mode != MODE_OFFSET:
                    s = (dt.type(alpha[t, j]) * s)
                new_v2c[int(es[d])] = llrT[j] + s
            post[j] = llrT[j] + ssum(msgs)
        v2c = np.where(act[None, :], new_v2c, v2c)
        # ---- decision, syndrome, early stop (ldpc_decoder.py:140-144) ----
        bits = (post < 0).astype(np.uint8)          # [n, B]
        syn = np.zeros(B, dtype=np.int64)
        chk_of_edge = graph.edge_check()
        par = np.zeros((m, B), dtype=np.int64)
        np.add.at(par, chk_of_edge, bits[cv].astype(np.int64))
        syn = (par % 2).sum(axis=0)
        ok = (syn == 0)
        upd = act  # frames still running take this iteration's outputs
        out_bits[upd] = bits.T[upd]
        out_post[upd] = post.T[upd]
        if early_stop:
            newly = act & ok
            iters[newly] = t + 1
            success[newly] = True
            done |= newly
        elif t == T - 1:
            success[:] = ok
    if T == 0:
        out_post[:] = llr
        out_bits[:] = (llr < 0)
    return OracleResult(bits=out_bits, posterior=out_post, iterations=iters, success=success)


# --------------------------------------------------------------------------------------------
# Weight-table expansion shared by tests (mirrors neural_2d_decoder.py:84-131 semantics)
# --------------------------------------------------------------------------------------------
def expand_2d_weights(graph: SparseGraph, weight_sharing_type: int, T: int, beta_fn, alpha_fn):
    """beta_fn(t, dc, dv) / alpha_fn(t, dv) -> float; returns per-edge beta[T,E], per-variable
    alpha[T,n] (or None) float32 following the four sharing types:
      1: beta[t,dc,dv], alpha absent(=1)   2: beta[t,dc], alpha[t,dv]
      3: beta[t,dc], alpha=1               4: beta=float32(0.7), alpha[t,dv]."""
    dcs = graph.check_degrees()
    dvs = graph.var_degrees()
    e_chk = graph.edge_check()
    e_dc = dcs[e_chk]
    e_dv = dvs[graph.check_var]
    beta = np.zeros((T, graph.E), dtype=np.float32)
    alpha = np.ones((T, graph.n), dtype=np.float32)
    for t in range(T):
        for e in range(graph.E):
            if weight_sharing_type == 4:
                beta[t, e] = np.float32(0.7)
            else:
                beta[t, e] = np.float32(beta_fn(t, int(e_dc[e]), int(e_dv[e])))
        if weight_sharing_type in (2, 4):
            for j in range(graph.n):
                alpha[t, j] = np.float32(alpha_fn(t, int(dvs[j])))
    return beta, alpha


def decode_layered_rcq(graph: SparseGraph, llr: np.ndarray, *, T: int, bc: int, thresholds: np.ndarray,
                       quantizer_of_iter: np.ndarray) -> OracleResult:
    """RCQMinSumDecoder._decode_layered as the reference actually behaves (rcq_decoder.py:281-350,
    SURVEY appendix C6): posteriors start at the LLRs; checks are visited in index order; the
    "subtract the previous C2V" step always subtracts 0 because ``c2v_messages`` is re-created as zeros
    inside the check loop (:323) -- except on a graph with a single non-empty check, where the previous
    visit's row survives (that degenerate case is restated too); the quantised new C2V are ADDED to the
    posteriors in place; the syndrome is tested once per iteration."""
    llr = np.ascontiguousarray(np.asarray(llr, dtype=np.float32))
    if llr.ndim == 1:
        llr = llr[None, :]
    B, n = llr.shape
    cp, cv = graph.check_ptr, graph.check_var
    m = graph.m
    post = np.ascontiguousarray(llr.T).copy()           # [n, B]
    done = np.zeros(B, dtype=bool)
    iters = np.full(B, T, dtype=np.int32)
    success = np.zeros(B, dtype=bool)
    out_bits = (llr < 0).astype(np.uint8)
    ar = np.arange(B)
    nonempty = [i for i in range(m) if cp[i + 1] > cp[i]]
    prev_check, prev_vals = None, None
    chk_of_edge = graph.edge_check()
    for t in range(T):
        act = ~done
        if not act.any():
            break
        th = np.asarray(thresholds[int(quantizer_of_iter[t])], dtype=np.float32)
        newp = post.copy()
        for i in nonempty:
            e0, e1 = int(cp[i]), int(cp[i + 1])
            vs = cv[e0:e1]
            dc = e1 - e0
            if prev_check == i:                          # only possible with one non-empty check
                for k in range(dc):
                    newp[vs[k]] = newp[vs[k]] - prev_vals[k]
            inc = newp[vs]
            signs = np.sign(inc)
            mags = np.abs(inc)
            k0 = np.argmin(mags, axis=0)
            m1 = mags[k0, ar]
            if dc > 1:
                tmp = mags.copy()
                tmp[k0, ar] = np.inf
                m2 = tmp.min(axis=0)
            else:
                m2 = m1
            vals = []
            for k in range(dc):
                sp = np.ones(B, dtype=np.float32)
                for q in range(dc):
                    if q != k:
                        sp = sp * signs[q]
                raw = np.where(k0 == k, m2, m1)
                x = (sp * raw).astype(np.float32)
                vals.append(dequantize(quantize(x, th, bc), th, bc))
            for k in range(dc):
                newp[vs[k]] = (newp[vs[k]] + vals[k]).astype(np.float32)
            prev_check, prev_vals = i, vals
        post = np.where(act[None, :], newp, post)
        bits = (post < 0).astype(np.uint8)
        par = np.zeros((m, B), dtype=np.int64)
        np.add.at(par, chk_of_edge, bits[cv].astype(np.int64))
        ok = ((par % 2).sum(axis=0) == 0)
        out_bits[act] = bits.T[act]
        newly = act & ok
        iters[newly] = t + 1
        success[newly] = True
        done |= newly
    return OracleResult(bits=out_bits, posterior=post.T.copy(), iterations=iters, success=success)
