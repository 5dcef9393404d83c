"""forward() (posterior wanted: frozen-message kernel variants, final pass, posterior transpose) against the
decisions-only decode at full batch size (context for DESIGN.md; not a test)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import ldpc_b200 as L
import bench

code = L.codes.dvbs2_shaped(max_iterations=10)
dec = bench.build_decoder(L, code, "n2d2")
eng = dec._engine(0)
B = 65536
for tag, sign, snr in (("no frame stops", -1, 2.0), ("frames stop (3 dB)", 1, 3.0)):
    llr = L.awgn_llr(code.n, B, snr, seed=1, llr_sign=sign)
    for name, fn in (("decode (bits, iterations, success)", lambda: eng.decode_device(llr, want_posterior=False)),
                     ("forward (bits, posterior, iterations)", lambda: eng.decode_device(llr, want_posterior=True))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(5):
            out = fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / 5
        print(f"{tag:22s} {name:40s} {dt * 1e3:8.2f} ms  {B / dt / 1e3:8.1f} K frames/s  avg iterations {out[2].float().mean().item():.2f}")
