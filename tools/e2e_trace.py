"""Per-chunk timeline of the host-buffer pipeline (ldpc_decode_host) on stderr: LDPC_PIPE_TRACE=1."""
import sys, os, numpy as np, torch
sys.path.insert(0, ".")
import ldpc_b200 as L
import bench
code = L.codes.dvbs2_shaped(max_iterations=10)
dec = bench.build_decoder(L, code, "n2d2")
eng = dec._engine(0)
B = 65536
llr = L.awgn_llr(code.n, B, 2.0, seed=1, llr_sign=-1)
pin = L.PinnedBuffer((B, code.n), np.float32)
for s in range(0, B, 8192):
    pin.array[s:s + 8192] = llr[s:s + 8192].cpu().numpy()
keep = [L.PinnedBuffer((B, code.n), np.uint8), L.PinnedBuffer((B,), np.int32), L.PinnedBuffer((B,), np.uint8)]   # keep the owners alive
outs = dict(bits=keep[0].array, iterations=keep[1].array, success=keep[2].array)
eng.decode_host(pin.array, out=outs)
os.environ["LDPC_PIPE_TRACE"] = "1"
eng.decode_host(pin.array, out=outs)
