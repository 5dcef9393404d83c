"""Throughput on small codes at large batches (per-iteration kernels, messages through HBM)."""
import sys, time, torch, numpy as np
sys.path.insert(0, ".")
import ldpc_b200 as L
T = 10
def run(name, dec, llr, fn):
    import gc; gc.collect(); torch.cuda.synchronize()   # (a decoder freed inside the timed region stalls it)
    for _ in range(2): fn(llr)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3): out = fn(llr)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 3
    it = out[2].float().mean().item() if torch.is_tensor(out[2]) else float(np.mean(out[2]))
    print(name, llr.shape[0], f"{dt*1e3:.2f} ms  {llr.shape[0]/dt/1e6:.1f} M frames/s  avg it {it:.2f}", flush=True)
c74 = L.create_test_ldpc_code()
rng = np.random.default_rng(0)
H = np.zeros((126, 252), dtype=np.int64)
# (3,6)-regular-ish n=252 code
cols = np.repeat(np.arange(252), 3); rng.shuffle(cols)
for i in range(126):
    H[i, np.unique(cols[6 * i:6 * i + 6])] = 1
c252 = L.LDPCCode(252, 126, H, max_iterations=T)
for cname, code in (("(7,4)", c74), ("n252", c252)):
    for B in (65536, 1048576):
        for sign in (-1, 1):
            llr = L.awgn_llr(code.n, B, 2.0, seed=1, llr_sign=sign)
            dec = L.Neural2DMinSumDecoder(code, 2, T)
            run(f"{cname} N-2D type 2 sign {sign:+d}", dec, llr, lambda x: dec._engine(0).decode_device(x))
            rcq = L.RCQMinSumDecoder(code, 3, 8, [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)], max_iterations=T)
            run(f"{cname} RCQ bc=3 sign {sign:+d}", rcq, llr, rcq.decode)
