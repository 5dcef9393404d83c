"""One layered RCQ decode of B frames of the (16200,7200)-shaped chain code (ncu target for layered_pipe_kernel):
    ncu --set full -k regex:layered_pipe -s 3 -c 1 python tools/layered_one.py 32768
"""
import sys, torch
sys.path.insert(0, ".")
import ldpc_b200 as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
code = L.codes.dvbs2_shaped(max_iterations=10)
dec = L.RCQMinSumDecoder(code, 3, 8, qp, max_iterations=10, layered=True)
llr = L.awgn_llr(code.n, B, 2.0, seed=1, llr_sign=-1)
out = dec.decode(llr)
torch.cuda.synchronize()
print("ok", out[2].float().mean().item())
