"""Per-call latency of small batches: CTA-resident decode against the per-iteration kernels (graph replay), to place
the batch-size threshold of the policy (ldpc_api.cu: resident_max_frames).   python tools/resident_latency_probe.py"""
import json, os, sys, gc
import torch
sys.path.insert(0, ".")
import bench
import ldpc_b200 as L


def timed(fn, reps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for cname, kind in (("dvbs2", "n2d2"), ("dvbs2", "rcq"), ("qc", "wrcq1"), ("r504", "n2d2"), ("r504", "rcq")):
    code = bench.make_code(L, cname, 10)
    for B in (1, 16, 148, 296, 592, 1184, 4096):
        llr = L.awgn_llr(code.n, B, 2.0, seed=1, llr_sign=-1)
        row = {"code": cname, "decoder": kind, "frames": B}
        for mode in ("0", "1"):
            os.environ["LDPC_RESIDENT"] = mode
            dec = bench.build_decoder(L, code, kind, 10)
            eng = dec._engine(0)
            row["resident_us" if mode == "1" else "per_iteration_us"] = round(1e3 * timed(lambda: eng.decode_device(llr)), 1)
            eng.close()
            del dec, eng
            gc.collect()
        print(json.dumps(row), flush=True)
