"""Soak: many decodes through the asynchronous-copy kernels (row ring, shared-memory stage) against the plain
kernels (LDPC_WIDE_RING=0) on fresh noise each time -- looks for rare ordering bugs (mbarrier phases, cp.async
waits) that a handful of test runs would not hit.  Usage: python tools/soak.py [rounds]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import ldpc_b200 as L

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 100
T = 8
qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
cases = {
    "qc n2d2": (L.codes.qc_shaped(max_iterations=T), lambda c: L.Neural2DMinSumDecoder(c, 2, T), 6.0),
    "qc wrcq1": (L.codes.qc_shaped(max_iterations=T), lambda c: L.WeightedRCQDecoder(c, 3, 8, qp, 1, T), 6.5),
    "dv12 n2d2": (L.codes.ira_code({12: 1620, 3: 4860}, {5: 4861, 6: 4859}, max_iterations=T), lambda c: L.Neural2DMinSumDecoder(c, 2, T), 2.5),
    "dv12 rcq": (L.codes.ira_code({12: 1620, 3: 4860}, {5: 4861, 6: 4859}, max_iterations=T), lambda c: L.RCQMinSumDecoder(c, 3, 8, qp, max_iterations=T), 3.5),
}
for name, (code, make, snr) in cases.items():
    decs = []
    for ring in ("0", "1"):
        os.environ["LDPC_WIDE_RING"] = ring
        torch.manual_seed(0)
        d = make(code)
        with torch.no_grad():
            for tname in ("_beta_table", "_alpha_table"):
                t = getattr(d, tname, None)
                if t is not None:
                    t.fill_(0.85 if tname == "_beta_table" else 1.0)
        eng = d._engine(0) if hasattr(d, "_engine") else None
        decs.append((d, eng))
    bad = 0
    for r in range(rounds):
        B = int(np.random.default_rng(r).integers(1, 3000))
        llr = L.awgn_llr(code.n, B, snr, seed=1000 + r, llr_sign=1)
        outs = []
        for d, eng in decs:
            if eng is not None:
                outs.append(eng.decode_device(llr, want_posterior=(r % 2 == 0)))
            else:
                b, s, i = d.decode(llr)
                outs.append((b, None, i, s))
        a, b = outs
        same = torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]) and (a[1] is None or torch.equal(a[1], b[1]))
        bad += 0 if same else 1
    print(f"{name:10s} rounds {rounds}  mismatches {bad}", flush=True)
    assert bad == 0
print("soak ok")
