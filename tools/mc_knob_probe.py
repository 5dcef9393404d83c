"""Early-stop decode at T = 50 (the Monte-Carlo regime) under the compaction / checkpoint knobs: where the time
between the ideal (average iterations x per-iteration rate) and the measured rate goes.   python tools/mc_knob_probe.py"""
import json, os, sys, gc
import torch
sys.path.insert(0, ".")
import bench
import ldpc_b200 as L

B, T = 65536, 50
bench.T_ITERS = T
code = L.codes.dvbs2_shaped(max_iterations=T)


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


envs = [{}, {"LDPC_COMPACT_PERCENT": "70"},
        {"LDPC_SPECULATE": "1"}, {"LDPC_SPECULATE": "0"}, {"LDPC_CHECKPOINT_STEP": "2"}]
for snr in (2.0, 3.0):
    llr = L.awgn_llr(code.n, B, snr, seed=1, llr_sign=1)
    for env in envs:
        os.environ.update(env)
        dec = bench.build_decoder(L, code, "n2d2", T)
        with torch.no_grad():
            dec._beta_table.clamp_(max=1.0)
            dec._alpha_table.fill_(1.0)
        eng = dec._engine(0)
        ms, out = timed(lambda: eng.decode_device(llr))
        prof = eng.profile_read()
        print(json.dumps({"snr": snr, "env": env, "ms": round(ms, 3), "kfps": round(B / ms, 1), "avg_it": round(out[2].float().mean().item(), 3),
                          "compactions": prof["compactions"], "launches": prof["launches"]}), flush=True)
        for k in env:
            os.environ.pop(k)
        eng.close()
        del dec, eng
        gc.collect()
