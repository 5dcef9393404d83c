"""Per-call latency of small batches (context for DESIGN.md; not a test)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import ldpc_b200 as L

def probe(name, dec, llr, n=60):
    for _ in range(5):
        dec(llr)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        out = dec(llr)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / n
    print(f"{name:50s} B={llr.shape[0] if llr.dim() == 2 else 1:6d}  {dt * 1e6:9.1f} us/call  iterations max {int(torch.as_tensor(out[2]).max())}")

for cname, code in (("(7,4)", L.create_test_ldpc_code()), ("dvbs2-shaped", L.codes.dvbs2_shaped(max_iterations=10))):
    dec = L.Neural2DMinSumDecoder(code, 2, 10)
    with torch.no_grad():
        dec._beta_table.fill_(0.8); dec._alpha_table.fill_(1.0)
    for B in (1, 128, 1024, 2048, 4096):
        for snr, tag in ((1.0, "no stop"), (6.0, "early stop")):
            llr = L.awgn_llr(code.n, B, snr, seed=1, llr_sign=1 if tag == "early stop" else -1)
            probe(f"{cname} N-2D T=10 {tag}", dec, llr if B > 1 else llr[0])
