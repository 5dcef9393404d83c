"""Per-chunk timeline of the host-buffer pipeline (ldpc_decode_host[_packed]) on stderr: LDPC_PIPE_TRACE=1.
    python tools/e2e_trace.py [post] [packed]"""
import sys, os, time, numpy as np, torch
sys.path.insert(0, ".")
import ldpc_b200 as L
import bench
post, packed = "post" in sys.argv, "packed" in sys.argv
code = L.codes.dvbs2_shaped(max_iterations=10)
dec = bench.build_decoder(L, code, "n2d2")
eng = dec._engine(0)
B = 65536
llr = L.awgn_llr(code.n, B, 2.0, seed=1, llr_sign=-1)
pin = L.PinnedBuffer((B, code.n), np.float32)
for s in range(0, B, 8192):
    pin.array[s:s + 8192] = llr[s:s + 8192].cpu().numpy()
keep = {"bits": L.PinnedBuffer((B, eng.row_words), np.uint32) if packed else L.PinnedBuffer((B, code.n), np.uint8),
        "iterations": L.PinnedBuffer((B,), np.int32), "success": L.PinnedBuffer((B,), np.uint8)}
if post:
    keep["posterior"] = L.PinnedBuffer((B, code.n), np.float32)
outs = {k: v.array for k, v in keep.items()}
for _ in range(2):
    eng.decode_host(pin.array, want_posterior=post, out=outs, packed_bits=packed)
t = time.perf_counter()
eng.decode_host(pin.array, want_posterior=post, out=outs, packed_bits=packed)
print("untraced call: %.2f ms" % ((time.perf_counter() - t) * 1e3), file=sys.stderr)
os.environ["LDPC_PIPE_TRACE"] = "1"
eng.decode_host(pin.array, want_posterior=post, out=outs, packed_bits=packed)
