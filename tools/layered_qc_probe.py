"""Level-parallel layered RCQ on the (9472,8192)-shaped QC code: staged in shared memory against the unstaged kernel.
    python tools/layered_qc_probe.py [sizes,comma-separated [stages]]"""
import os, sys, torch
sys.path.insert(0, ".")
import ldpc_b200 as L
qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
code = L.codes.qc_shaped(max_iterations=10)
E = code.graph.E
SIZES = tuple(int(x) for x in sys.argv[1].split(",")) if len(sys.argv) > 1 else (8192, 32768, 131072)
STAGES = tuple(sys.argv[2].split(",")) if len(sys.argv) > 2 else ("0", "1", "2")
for B in SIZES:
    llr = L.awgn_llr(code.n, B, 2.0, seed=1, llr_sign=-1)
    for stage in STAGES:
        os.environ["LDPC_LAYERED_STAGE"] = stage
        dec = L.RCQMinSumDecoder(code, 3, 8, qp, max_iterations=10, layered=True)
        for _ in range(2):
            dec.decode(llr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = dec.decode(llr)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        it = out[2].float().mean().item()
        print(f"qc layered frames {B} stage {stage}: {ms:.2f} ms  {B / ms:.0f} K frames/s  avg it {it:.2f}  "
              f"{8 * E * it * B / ms / 1e6:.0f} GB/s of the 8E roofline", flush=True)
        dec._engine(0).close()
        del dec
    del llr
