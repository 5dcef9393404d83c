"""One training step (for ncu): python tools/train_one.py [frames]"""
import sys, torch
sys.path.insert(0, ".")
import bench
import ldpc_b200 as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
code = bench.make_code(L, "dvbs2", 10)
dec = bench.build_decoder(L, code, "n2d2", 10)
eng = dec._engine(0)
llr = L.awgn_llr(code.n, B, 2.0, seed=1, llr_sign=1)
for _ in range(2):
    _, post, _, _ = eng.train_forward(llr)
    eng.train_backward(torch.sigmoid(-post) * (-1.0 / post.numel()))
torch.cuda.synchronize()
