"""Layered RCQ schedule: throughput against batch size (one thread per frame walks the checks of a chain-structured
code; quasi-cyclic codes run level-parallel).  `python tools/layered_probe.py chain` = the chain code only."""
import sys, time, torch, numpy as np
sys.path.insert(0, ".")
import ldpc_b200 as L
qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
cases = [("dvbs2", True, B) for B in (8192, 32768, 65536, 131072, 262144)]
if "chain" not in sys.argv:
    cases = [("dvbs2", False, 32768)] + cases + [("qc", False, 32768), ("qc", True, 32768)]
for cname, layered, B in cases:
    code = L.codes.dvbs2_shaped(max_iterations=10) if cname == "dvbs2" else L.codes.qc_shaped(max_iterations=10)
    dec = L.RCQMinSumDecoder(code, 3, 8, qp, max_iterations=10, layered=layered)
    llr = L.awgn_llr(code.n, B, 2.0, seed=1, llr_sign=-1)
    for _ in range(2): dec.decode(llr)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3): out = dec.decode(llr)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 3
    print(cname, "layered" if layered else "flooding", B, f"{dt*1e3:.1f} ms  {B/dt/1e3:.0f} K frames/s  avg it {out[2].float().mean().item():.2f}", flush=True)
    del llr, dec, out
    torch.cuda.empty_cache()
