"""Round-2 A/B probe (context for DESIGN.md; not a test): decode vs forward at the benchmark size under the
posterior modes and with / without speculative spans.  One JSON line per case on stdout.

    python tools/r02_probe.py [frames]
Environment knobs are read when a decoder handle is created, so every case builds its own decoder."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import ldpc_b200 as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
code = L.codes.dvbs2_shaped(max_iterations=10)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


cases = [("decode", {}, False), ("decode", {"LDPC_SPECULATE": "0"}, False), ("decode", {"LDPC_COMPACT": "0"}, False),
         ("forward", {"LDPC_POST_MODE": "1"}, True), ("forward", {"LDPC_POST_MODE": "2"}, True), ("forward", {}, True),
         ("forward", {"LDPC_SPECULATE": "0"}, True)]
for tag, sign, snr, T in (("no frame stops, T=10", -1, 2.0, 10), ("frames stop, 3 dB, T=10", 1, 3.0, 10),
                          ("frames stop, 2 dB, T=50", 1, 2.0, 50)):
    llr = L.awgn_llr(code.n, B, snr, seed=1, llr_sign=sign)
    for name, env, post in cases:
        for k, v in env.items():
            os.environ[k] = v
        bench.T_ITERS = T
        c = L.codes.dvbs2_shaped(max_iterations=T) if T != 10 else code
        dec = bench.build_decoder(L, c, "n2d2")
        if T != 10:   # bench.det_weights is meant for 10 iterations: keep the T = 50 weights in a sane range
            with torch.no_grad():
                dec._beta_table.clamp_(max=1.0)
                dec._alpha_table.fill_(1.0)
        eng = dec._engine(0)
        ms, out = timed(lambda: eng.decode_device(llr, want_posterior=post))
        prof = eng.profile_read()
        print(json.dumps({"case": tag, "call": name, "env": env, "ms": round(ms, 3), "kfps": round(B / ms, 1),
                          "avg_iterations": round(out[2].float().mean().item(), 3),
                          "compactions": prof["compactions"], "early_exits": prof["early_exits"]}), flush=True)
        for k in env:
            os.environ.pop(k)
        del dec, eng
