"""Soak of the round-2 on-chip decodes: many random batches through the CTA-resident decode and the one-launch small
decode against the per-iteration kernels (fresh noise, random batch sizes, early stop on, posteriors and packed rows
at random) -- looks for rare ordering bugs a handful of test runs would not hit.   python tools/soak_onchip.py [rounds]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import ldpc_b200 as L

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 150
T = 12
qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
rng = np.random.default_rng(11)


def weights(d):
    with torch.no_grad():
        for tname, lo, hi in (("_beta_table", 0.6, 0.95), ("_alpha_table", 0.9, 1.0)):
            t = getattr(d, tname, None)
            if t is not None:
                g = torch.Generator().manual_seed(3)
                t.copy_(lo + (hi - lo) * torch.rand(t.shape, generator=g))
    return d


cases = {
    "dvbs2/20 n2d1": (L.codes.dvbs2_shaped(max_iterations=T, scale=20), lambda c: L.Neural2DMinSumDecoder(c, 1, T), 3.0, 3000),
    "dvbs2/20 wrcq2": (L.codes.dvbs2_shaped(max_iterations=T, scale=20), lambda c: L.WeightedRCQDecoder(c, 3, 8, qp, 2, T), 3.5, 3000),
    "dvbs2 n2d2": (L.codes.dvbs2_shaped(max_iterations=T), lambda c: L.Neural2DMinSumDecoder(c, 2, T), 2.8, 500),
    "qc oms2": (L.codes.qc_shaped(max_iterations=T), lambda c: L.Neural2DOffsetMinSumDecoder(c, 2, T), 6.0, 500),
    "h74 nnms": (L.create_test_ldpc_code(), lambda c: L.NeuralMinSumDecoder(c, T), 4.0, 20000),
}
bad = 0
for name, (code, make, snr, bmax) in cases.items():
    engines = []
    for onchip in ("0", "1"):
        os.environ["LDPC_RESIDENT"] = onchip
        os.environ["LDPC_SMALL"] = onchip
        torch.manual_seed(0)
        d = weights(make(code))
        engines.append((d, d._engine(0)))
    used = 0
    for r in range(rounds):
        B = int(rng.integers(1, bmax))
        llr = L.awgn_llr(code.n, B, snr + float(rng.normal(0, 0.6)), seed=1000 + r, llr_sign=1)
        post, packed = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        outs = [e.decode_device(llr, want_posterior=post, packed_bits=packed) for _, e in engines]
        for a, b in zip(*outs):
            if (a is None) != (b is None) or (a is not None and not torch.equal(a, b)):
                bad += 1
                print("MISMATCH", name, "round", r, "frames", B, flush=True)
                break
    prof = engines[1][1].profile_read()
    print(f"{name}: {rounds} rounds, on-chip decodes {prof['resident_decodes'] + prof['small_decodes']}, mismatches so far {bad}", flush=True)
    for _, e in engines:
        e.close()
print("SOAK", "FAILED" if bad else "OK")
sys.exit(1 if bad else 0)
