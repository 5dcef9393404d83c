"""A/B probe: CTA-resident decode (LDPC_RESIDENT=1) against the per-iteration kernels (LDPC_RESIDENT=0).
    python tools/resident_probe.py"""
import json, os, sys, gc
import torch
sys.path.insert(0, ".")
import bench
import ldpc_b200 as L


def timed(fn, reps=3):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


cases = [("dvbs2", "n2d2", 65536, 10, -1, 2.0, False), ("dvbs2", "n2d2", 65536, 10, -1, 2.0, True), ("dvbs2", "n2d2", 8192, 10, -1, 2.0, False),
         ("dvbs2", "rcq", 65536, 10, -1, 2.0, False), ("qc", "wrcq1", 65536, 10, -1, 2.0, True), ("qc", "n2d2", 65536, 10, -1, 2.0, False),
         ("r504", "n2d2", 1 << 20, 10, -1, 2.0, False), ("r504", "rcq", 1 << 20, 10, -1, 2.0, False),
         ("dvbs2", "n2d2", 65536, 50, 1, 2.0, False), ("dvbs2", "n2d2", 65536, 50, 1, 3.0, False)]
for cname, kind, B, T, sign, snr, post in cases:
    bench.T_ITERS = T
    code = bench.make_code(L, cname, T)
    llr = L.awgn_llr(code.n, B, snr, seed=1, llr_sign=sign)
    for mode in ("0", "1"):
        os.environ["LDPC_RESIDENT"] = mode
        dec = bench.build_decoder(L, code, kind, T)
        if T != 10 and kind == "n2d2":
            with torch.no_grad():
                dec._beta_table.clamp_(max=1.0)
                dec._alpha_table.fill_(1.0)
        eng = dec._engine(0)
        gc.collect(); torch.cuda.synchronize()
        ms, out = timed(lambda: eng.decode_device(llr, want_posterior=post))
        prof = eng.profile_read()
        print(json.dumps({"code": cname, "decoder": kind, "frames": B, "T": T, "snr": snr, "sign": sign, "posterior": post, "resident": mode,
                          "ms": round(ms, 3), "kfps": round(B / ms, 1), "avg_it": round(out[2].float().mean().item(), 2),
                          "resident_decodes": prof["resident_decodes"]}), flush=True)
        eng.close()
        del dec, eng
        gc.collect(); torch.cuda.empty_cache()
