"""Per-step device times of one bench workload:  python tools/step_times.py DECODER CODE [steps] [post]"""
import sys, torch
sys.path.insert(0, ".")
import bench
import ldpc_b200 as L
kind, cname = sys.argv[1], sys.argv[2]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
post = len(sys.argv) > 4 and sys.argv[4] == "post"
code = bench.make_code(L, cname)
dec = bench.build_decoder(L, code, kind)
eng = dec._engine(0)
llr = L.awgn_llr(code.n, 65536, 2.0, seed=1234, llr_sign=-1)
for _ in range(3):
    eng.decode_device(llr, want_posterior=post)
torch.cuda.synchronize()
for mode in (0, 1):
    eng.profile_mode(mode)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for k in range(steps):
        out = eng.decode_device(llr, want_posterior=post)
        ev[k + 1].record()
    torch.cuda.synchronize()
    print(kind, cname, "posterior" if post else "decode-only", "profile mode", mode, [round(ev[k].elapsed_time(ev[k + 1]), 2) for k in range(steps)])
    print("   ", {k: round(v, 2) if isinstance(v, float) else v for k, v in eng.profile_read().items() if k in ("launches", "cn_ms", "vn_ms", "other_ms", "cn_launches")})
