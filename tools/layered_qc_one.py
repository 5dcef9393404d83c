"""A few level-parallel layered decodes of the QC shape (for ncu): python tools/layered_qc_one.py [frames]"""
import sys, torch
sys.path.insert(0, ".")
import ldpc_b200 as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
code = L.codes.qc_shaped(max_iterations=2)
dec = L.RCQMinSumDecoder(code, 3, 8, [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)], max_iterations=2, layered=True)
llr = L.awgn_llr(code.n, B, 2.0, seed=1, llr_sign=-1)
for _ in range(2):
    dec.decode(llr)
torch.cuda.synchronize()
