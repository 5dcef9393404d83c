"""One CTA-resident decode (for ncu): python tools/resident_one.py [dvbs2|qc|r504] [n2d2|rcq|wrcq1] [frames]"""
import os, sys, torch
sys.path.insert(0, ".")
os.environ["LDPC_RESIDENT"] = "1"
import bench
import ldpc_b200 as L
cname = sys.argv[1] if len(sys.argv) > 1 else "dvbs2"
kind = sys.argv[2] if len(sys.argv) > 2 else "n2d2"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 148
code = bench.make_code(L, cname, 10)
dec = bench.build_decoder(L, code, kind, 10)
eng = dec._engine(0)
llr = L.awgn_llr(code.n, B, 2.0, seed=1, llr_sign=-1)
for _ in range(3):
    out = eng.decode_device(llr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = eng.decode_device(llr)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(cname, kind, B, "%.3f ms per call, %.1f K frames/s, avg it %.2f, resident decodes %d" %
      (ms, B / ms, out[2].float().mean().item(), eng.profile_read()["resident_decodes"]))
