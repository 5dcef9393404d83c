"""One on-chip decode of the (7,4) code (for ncu): python tools/small_one.py [basic|n2d2|rcq] [frames]"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import ldpc_b200 as L
kind = sys.argv[1] if len(sys.argv) > 1 else "basic"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
code = L.create_test_ldpc_code()
llr = L.awgn_llr(7, B, 2.0, seed=1, llr_sign=-1)
if kind == "basic":
    eng = L.BasicMinSumDecoder(code, 0.7)._engine(0)
    llr = llr.double()
elif kind == "rcq":
    eng = L.RCQMinSumDecoder(code, 3, 8, [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)], max_iterations=10)._engine(0)
else:
    dec = L.Neural2DMinSumDecoder(code, 2, 10)
    with torch.no_grad():
        dec._beta_table.fill_(0.8); dec._alpha_table.fill_(0.9)
    eng = dec._engine(0)
for _ in range(3):
    out = eng.decode_device(llr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = eng.decode_device(llr)
e1.record(); torch.cuda.synchronize()
print(kind, B, "%.3f ms per call, %.1f M frames/s, avg it %.2f" % (e0.elapsed_time(e1) / 5, B / (e0.elapsed_time(e1) / 5) / 1e3, out[2].float().mean().item()))
