"""Two forward() calls (decisions + posterior) of B frames of the (16200,7200)-shaped code, device-resident row-major
buffers (ncu target for the layout kernels): python tools/forward_one.py [frames]"""
import sys, torch
sys.path.insert(0, ".")
import bench
import ldpc_b200 as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
code = bench.make_code(L, "dvbs2", 10)
dec = bench.build_decoder(L, code, "n2d2", 10)
llr = L.awgn_llr(code.n, B, 2.0, seed=1, llr_sign=-1)
for _ in range(2):
    out = dec(llr)
torch.cuda.synchronize()
print("ok", float(out[2].float().mean()))
