"""Generate the committed golden vectors by running the LIVE, UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  Usage:
    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

What is recorded, per (code, decoder) case: the parity-check matrix, the LLR frames fed in, the
weights as (key -> value) pairs exactly as the reference's ParameterDict holds them, the same
weights expanded per edge / per variable *by calling the reference's own _get_beta_weight /
_get_alpha_weight*, the quantiser thresholds the reference computed, and the outputs of the
reference's decode()/forward(): bits, posterior (where returned), iterations, success.

The only accommodations are those of oracle/ref_shim.py (matplotlib stub, cached degree dicts),
both value-neutral.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402


def code_hamming(ref):
    return ref.ldpc_decoder.create_test_ldpc_code()  # ldpc_decoder.py:274-284, max_iterations=10


def code_irregular(ref, seed=7, m=24, n=48):
    """Irregular code with variable degrees {1,2,3,8,9,12} and ragged check degrees; also one
    all-zero column (isolated variable) and entries equal to 2 that must NOT count as edges."""
    rng = np.random.default_rng(seed)
    H = np.zeros((m, n), dtype=np.int64)
    dvs = rng.choice([1, 2, 3, 8, 9, 12], size=n, p=[0.1, 0.35, 0.3, 0.1, 0.1, 0.05])
    for j in range(n):
        rows = rng.choice(m, size=int(dvs[j]), replace=False)
        H[rows, j] = 1
    H[:, 5] = 0            # isolated variable
    H[3, 5] = 2            # not an edge: the reference tests == 1
    for i in range(m):     # no empty checks except one we force
        if H[i].sum() == 0:
            H[i, rng.integers(n)] = 1
    H[m - 1, :] = 0        # an empty check row (skipped by the reference)
    H[m - 2, :] = 0
    H[m - 2, 11] = 1       # a degree-1 check
    return ref.ldpc_decoder.LDPCCode(n=n, k=n - m, H=H, max_iterations=8)


def code_regular36(ref, seed=3, n=60):
    """(3,6)-regular-ish code from permutations (duplicate edges collapse, giving a few lighter nodes)."""
    rng = np.random.default_rng(seed)
    m = n // 2
    H = np.zeros((m, n), dtype=np.int64)
    for _ in range(3):
        perm = rng.permutation(n)
        for idx, j in enumerate(perm):
            H[idx // 2, j] = 1
    return ref.ldpc_decoder.LDPCCode(n=n, k=n - m, H=H, max_iterations=12)


def make_llrs(rng, n, frames, dtype):
    """AWGN LLRs in both sign conventions plus hand-made edge cases (zeros, ties, huge values)."""
    out = []
    for f in range(frames):
        snr_db = [0.0, 2.0, 4.0, 6.0, 8.0][f % 5]
        sigma2 = 10 ** (-snr_db / 10)
        s = 1.0 if (f % 3) != 2 else -1.0      # +1: converges to all-zero; -1: reference convention
        y = s + np.sqrt(sigma2) * rng.standard_normal(n)
        out.append(2 * y / sigma2)
    out = np.array(out)
    if frames >= 4:
        out[1, : n // 3] = np.round(out[1, : n // 3] * 2) / 2     # ties between magnitudes
        out[3, ::5] = 0.0                                           # exact zeros -> three-valued sign
        out[3, 1::7] = -0.0
    return out.astype(dtype)


def set_weights(module, rng, style):
    kv = {}
    for name in ("beta_weights", "alpha_weights"):
        pd = getattr(module, name, None)
        if pd is None:
            continue
        for key in pd.keys():
            if style == "init":          # the reference's own scale, 0.1*randn: small, may be negative
                v = 0.1 * rng.standard_normal()
            elif name == "beta_weights":
                v = 0.55 + 0.4 * rng.random()
            else:
                v = 0.8 + 0.35 * rng.random()
            with torch.no_grad():
                pd[key].fill_(float(np.float32(v)))
            kv[f"{name}.{key}"] = float(pd[key].item())
    return kv


def expand_weights(module, code, T):
    H = code.H
    rows, cols = np.nonzero(H == 1)
    E = rows.shape[0]
    beta = np.ones((T, E), dtype=np.float32)
    alpha = np.ones((T, code.n), dtype=np.float32)
    has_beta = hasattr(module, "_get_beta_weight") or hasattr(module, "beta_weights")
    for t in range(T):
        for e in range(E):
            i, j = int(rows[e]), int(cols[e])
            if hasattr(module, "_get_beta_weight"):
                beta[t, e] = float(module._get_beta_weight(t, i, j))
            elif has_beta:
                beta[t, e] = float(module.beta_weights[f"iter_{t}_c{i}_v{j}"])
        if hasattr(module, "_get_alpha_weight"):
            for j in range(code.n):
                alpha[t, j] = float(module._get_alpha_weight(t, 0, j))
    return beta, alpha


def run_case(ref, code, kind, llrs, rng, **kw):
    T = kw.get("T", code.max_iterations)
    rec = {"H": code.H.astype(np.int8), "kind": kind, "T": T, "n": code.n, "k": code.k}
    torch.manual_seed(0)
    if kind == "basic":
        dec = ref.ldpc_decoder.BasicMinSumDecoder(code, factor=kw.get("factor", 0.7))
        rec["factor"] = kw.get("factor", 0.7)
        rec["T"] = code.max_iterations
        outs = [dec.decode(l.astype(np.float64)) for l in llrs]
        rec["llr"] = llrs.astype(np.float64)
        rec["bits"] = np.array([o[0] for o in outs], dtype=np.uint8)
        rec["success"] = np.array([o[1] for o in outs], dtype=bool)
        rec["iterations"] = np.array([o[2] for o in outs], dtype=np.int32)
        return rec
    llr32 = llrs.astype(np.float32)
    rec["llr"] = llr32
    if kind == "nnms":
        dec = ref.neural_minsum_decoder.NeuralMinSumDecoder(code, max_iterations=T)
    elif kind == "n2d":
        dec = ref.neural_2d_decoder.Neural2DMinSumDecoder(code, weight_sharing_type=kw["wtype"], max_iterations=T)
        rec["wtype"] = kw["wtype"]
    elif kind == "rcq":
        dec = ref.rcq_decoder.RCQMinSumDecoder(code, bc=kw["bc"], bv=8, quantizer_params=kw["qp"], max_iterations=T)
    elif kind == "wrcq":
        dec = ref.rcq_decoder.WeightedRCQDecoder(code, bc=kw["bc"], bv=8, quantizer_params=kw["qp"],
                                                 weight_sharing_type=kw["wtype"], max_iterations=T)
        rec["wtype"] = kw["wtype"]
    else:
        raise ValueError(kind)
    if kind in ("rcq", "wrcq"):
        rec["bc"] = kw["bc"]
        rec["qp"] = np.array(kw["qp"], dtype=np.float64)
        rec["thresholds"] = np.array([q.thresholds for q in dec.quantizers], dtype=np.float64)
        rec["quantizer_of_iter"] = np.array(
            [next(i for i, q in enumerate(dec.quantizers) if q is dec._get_quantizer(t)) for t in range(T)],
            dtype=np.int32)
    if kind != "rcq":
        kv = set_weights(dec, rng, kw.get("style", "trained"))
        rec["weight_keys"] = np.array(list(kv.keys()))
        rec["weight_vals"] = np.array(list(kv.values()), dtype=np.float32)
        beta, alpha = expand_weights(dec, code, T)
        rec["beta_edge"] = beta
        rec["alpha_var"] = alpha
    bits, post, its, succ = [], [], [], []
    with torch.no_grad():
        for l in llr32:
            t = torch.tensor(l, dtype=torch.float32)
            if kind == "rcq":
                b, s, it = dec.decode(t)
                succ.append(bool(s))
            else:
                b, p, it = dec(t)
                post.append(p.numpy().astype(np.float32))
            bits.append(b.numpy().astype(np.uint8))
            its.append(int(it))
    rec["bits"] = np.array(bits, dtype=np.uint8)
    rec["iterations"] = np.array(its, dtype=np.int32)
    if post:
        rec["posterior"] = np.array(post, dtype=np.float32)
    if succ:
        rec["success"] = np.array(succ, dtype=bool)
    return rec


def main():
    ref = ref_shim.load(cache_degrees=True)
    rng = np.random.default_rng(20260101)
    QP = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]   # simulation_framework.py:408
    codes = {
        "hamming74": (code_hamming(ref), 16),
        "irregular48": (code_irregular(ref), 8),
        "regular60": (code_regular36(ref), 6),
    }
    for cname, (code, frames) in codes.items():
        llrs = make_llrs(rng, code.n, frames, np.float64)
        T = code.max_iterations
        cases = [
            ("basic", dict()),
            ("nnms", dict(T=T)),
            ("n2d_t1", dict(kind="n2d", wtype=1, T=T)),
            ("n2d_t2", dict(kind="n2d", wtype=2, T=T)),
            ("n2d_t3", dict(kind="n2d", wtype=3, T=T)),
            ("n2d_t4", dict(kind="n2d", wtype=4, T=T)),
            ("n2d_t2_init", dict(kind="n2d", wtype=2, T=T, style="init")),
            ("rcq_b3", dict(kind="rcq", bc=3, qp=QP, T=T)),
            ("rcq_b4_q1", dict(kind="rcq", bc=4, qp=[(6.0, 1.0)], T=T)),
            ("rcq_b5_q2", dict(kind="rcq", bc=5, qp=[(4.0, 1.5), (9.0, 0.8)], T=T)),
            ("wrcq_t1", dict(kind="wrcq", wtype=1, bc=3, qp=QP, T=T)),
            ("wrcq_t2", dict(kind="wrcq", wtype=2, bc=3, qp=QP, T=T)),
            ("wrcq_t3", dict(kind="wrcq", wtype=3, bc=4, qp=QP, T=T)),
            ("wrcq_t4", dict(kind="wrcq", wtype=4, bc=3, qp=QP, T=T)),
        ]
        out = {}
        for name, kw in cases:
            kind = kw.pop("kind", name)
            rec = run_case(ref, code, kind, llrs, rng, **kw)
            for k, v in rec.items():
                out[f"{name}/{k}"] = np.asarray(v)
            print(cname, name, "iters", rec["iterations"].tolist())
        np.savez_compressed(os.path.join(HERE, f"{cname}.npz"), **out)

    # quantiser known answers (comprehensive_test.py:252-268 prints this round trip)
    q = ref.rcq_decoder.NonUniformQuantizer(3, 5.0, 1.5)
    x = np.array([-3.2, -1.1, 0.5, 2.8, 4.1, 0.0, -0.0, 5.0, 100.0, np.nan, np.inf, -1e-30, 0.9622504, 0.96225053],
                 dtype=np.float32)
    codes_q = q.quantize(torch.tensor(x)).numpy()
    vals = q.dequantize(torch.tensor(codes_q)).numpy()
    np.savez_compressed(os.path.join(HERE, "quantizer_kat.npz"), x=x, codes=codes_q, values=vals,
                        thresholds=np.array(q.thresholds, dtype=np.float64))
    print("quantizer", codes_q.tolist(), vals.tolist())


if __name__ == "__main__":
    main()
