"""Seeded synthetic parity-check matrices with the SHAPES of the codes the reference's paper quotes
(SURVEY.md appendix D).  They are not the standardised matrices (the reference has none either: its
``create_dvbs2_code`` is a dense random matrix, training_framework.py:379-400); only the degree
profile, size and edge count are matched, which is what sets the decoder's memory traffic.

  dvbs2_shaped()  n=16200, m=9000, E=48599: variables 1800 x dv8, 5400 x dv3, 8999 x dv2, 1 x dv1;
                  checks 1441 x dc4, 3239 x dc5, 3600 x dc6, 720 x dc7 (IRA: random information part +
                  dual-diagonal parity part)
  qc_shaped()     n=9472, m=1280, Z=128, base 10 x 74: every variable dv4, 512 checks dc29 + 768 dc30,
                  E=37888 (random circulant shifts)
Both return an ``LDPCCode`` whose ``H`` is a scipy CSR matrix (a dense 9000 x 16200 int64 H is 1.2 GB).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import scipy.sparse as sp

from .ldpc_decoder import LDPCCode


def _pair_sockets(rng, var_deg: np.ndarray, chk_deg: np.ndarray):
    """Configuration-model pairing of variable and check sockets without repeated (check, variable)."""
    assert int(var_deg.sum()) == int(chk_deg.sum()), (int(var_deg.sum()), int(chk_deg.sum()))
    vs = np.repeat(np.arange(var_deg.size), var_deg)
    cs = np.repeat(np.arange(chk_deg.size), chk_deg)
    rng.shuffle(cs)
    for _ in range(200):
        key = vs.astype(np.int64) * chk_deg.size + cs
        order = np.argsort(key, kind="stable")
        dup = np.zeros(key.size, dtype=bool)
        dup[order[1:]] = key[order][1:] == key[order][:-1]
        bad = np.nonzero(dup)[0]
        if bad.size == 0:
            return vs, cs
        partners = rng.integers(0, key.size, size=bad.size)
        for a, b in zip(bad, partners):   # swap the check ends of a duplicated socket with a random one
            cs[a], cs[b] = cs[b], cs[a]
    raise RuntimeError("could not remove duplicate edges")


def ira_code(info_deg: Dict[int, int], chk_deg: Dict[int, int], max_iterations: int = 50, seed: int = 0) -> LDPCCode:
    """IRA-style code: k information variables with the given degree histogram {dv: count}, m checks
    with the given TOTAL degree histogram {dc: count}, and an m-variable dual-diagonal parity part
    (parity variable i joins checks i and i+1; the last one only check m-1)."""
    rng = np.random.default_rng(seed)
    m = int(sum(chk_deg.values()))
    k = int(sum(info_deg.values()))
    n = k + m
    total = np.concatenate([np.full(c, d, dtype=np.int64) for d, c in sorted(chk_deg.items())])
    rng.shuffle(total)
    parity_edges = np.full(m, 2, dtype=np.int64)
    parity_edges[0] = 1
    if (total < parity_edges).any():
        raise ValueError("check degree below the parity part's contribution")
    vdeg = np.concatenate([np.full(c, d, dtype=np.int64) for d, c in sorted(info_deg.items(), reverse=True)])
    vs, cs = _pair_sockets(rng, vdeg, total - parity_edges)
    prow = np.concatenate([np.arange(m), np.arange(1, m)])
    pcol = np.concatenate([k + np.arange(m), k + np.arange(m - 1)])
    rows = np.concatenate([cs, prow])
    cols = np.concatenate([vs, pcol])
    H = sp.csr_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(m, n))
    H.sort_indices()
    assert H.data.max() == 1
    return LDPCCode(n=n, k=k, H=H, max_iterations=max_iterations)


def dvbs2_shaped(max_iterations: int = 50, seed: int = 0, scale: int = 1) -> LDPCCode:
    """(16200, 7200)-shaped code; ``scale`` > 1 divides every count (for fast tests), with the check
    histogram's last class absorbing the rounding so that the socket counts still match."""
    info = {8: 1800 // scale, 3: 5400 // scale}
    m = 9000 // scale
    chk = {4: 1441 // scale, 5: 3239 // scale, 6: 3600 // scale}
    chk[7] = m - sum(chk.values())
    need = sum(d * c for d, c in info.items()) + 2 * m - 1
    have = sum(d * c for d, c in chk.items())
    # move checks between the dc6 and dc7 classes until the edge counts agree
    diff = need - have
    if diff > 0:
        chk[6] -= diff
        chk[7] += diff
    elif diff < 0:
        chk[7] -= -diff
        chk[6] += -diff
    return ira_code(info, chk, max_iterations=max_iterations, seed=seed)


def qc_shaped(max_iterations: int = 50, seed: int = 0, Z: int = 128, base_rows: int = 10, base_cols: int = 74,
              dv: int = 4) -> LDPCCode:
    """(9472, 8192)-shaped quasi-cyclic code: base matrix with column weight ``dv`` and row weights as
    equal as possible (4 x 29 + 6 x 30 for the default shape), each base edge a Z x Z circulant."""
    rng = np.random.default_rng(seed)
    edges = base_cols * dv
    lo, rem = divmod(edges, base_rows)
    row_w = np.full(base_rows, lo, dtype=np.int64)
    row_w[base_rows - rem:] += 1
    vs, cs = _pair_sockets(rng, np.full(base_cols, dv, dtype=np.int64), row_w)
    shifts = rng.integers(0, Z, size=vs.size)
    z = np.arange(Z)
    rows = (cs[:, None] * Z + z[None, :]).ravel()
    cols = (vs[:, None] * Z + (z[None, :] + shifts[:, None]) % Z).ravel()
    m, n = base_rows * Z, base_cols * Z
    H = sp.csr_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(m, n))
    H.sort_indices()
    return LDPCCode(n=n, k=n - m, H=H, max_iterations=max_iterations)


def regular_code(n: int, dv: int, dc: int, max_iterations: int = 50, seed: int = 0) -> LDPCCode:
    """(dv, dc)-regular code from the configuration model (handy for tests and sweeps)."""
    assert (n * dv) % dc == 0
    m = n * dv // dc
    rng = np.random.default_rng(seed)
    vs, cs = _pair_sockets(rng, np.full(n, dv, dtype=np.int64), np.full(m, dc, dtype=np.int64))
    H = sp.csr_matrix((np.ones(vs.size, dtype=np.int8), (cs, vs)), shape=(m, n))
    H.sort_indices()
    return LDPCCode(n=n, k=n - m, H=H, max_iterations=max_iterations)
