"""Small end-to-end exercise of every kernel family on ragged sizes (wide checks / variables, tiny batches,
compaction, the Monte-Carlo round), each result compared with the oracle:
    python tools/sanitizer_probe.py
compute-sanitizer is closed on this GPU pool (gpurun refuses it), so out-of-bounds accesses are hunted with
small ragged cases like these plus the oracle comparison, not with the sanitizer."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
os.environ["LDPC_COMPACT_MIN_FRAMES"] = "128"
import ldpc_b200 as L

rng = np.random.default_rng(0)
# wide checks / wide variables / tiny + ragged batch
m, degs = 70, [8, 9, 12, 16, 17, 33, 64, 65] + [2, 3] * 20
n = len(degs)
H = np.zeros((m, n), dtype=np.int64)
for j, d in enumerate(degs):
    H[rng.choice(m, d, replace=False), j] = 1
H[0, rng.choice(n, 40, replace=False)] = 1
code = L.LDPCCode(n, 1, H, max_iterations=6)
llr = torch.from_numpy((rng.standard_normal((300, n)) * 2 + 1).astype(np.float32)).cuda()
qp = [(3.0, 1.3), (5.0, 1.3)]
for dec in (L.Neural2DMinSumDecoder(code, 1, 6), L.NeuralMinSumDecoder(code, 6), L.Neural2DOffsetMinSumDecoder(code, 2, 6),
            L.WeightedRCQDecoder(code, 3, 8, qp, 2, 6)):
    dec(llr); dec(llr[:1]); dec(llr.cpu())
L.RCQMinSumDecoder(code, 4, 8, qp, max_iterations=6).decode(llr)
L.RCQMinSumDecoder(code, 3, 8, qp, max_iterations=6, layered=True).decode(llr)
L.BasicMinSumDecoder(code, 0.5).decode(llr.double().cpu().numpy())
# compaction + Monte-Carlo + graphs on a mid-size code
c2 = L.codes.dvbs2_shaped(max_iterations=30, scale=20)
d2 = L.Neural2DMinSumDecoder(c2, 2, 30)
with torch.no_grad():
    d2._beta_table.fill_(0.8); d2._alpha_table.fill_(1.0)
x = torch.cat([L.awgn_llr(c2.n, 500, s, seed=k, llr_sign=1) for k, s in enumerate((1.0, 2.5, 4.0))])
d2(x); d2._engine(0).decode_device(x)
# layered RCQ on a chain-structured code: software-pipelined walk (shared-memory rings, cp.async), frames that stop early
L.RCQMinSumDecoder(c2, 3, 8, qp + [(7.0, 1.3)], max_iterations=9, layered=True).decode(x)
cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
fbe = torch.zeros(1500, dtype=torch.int32, device="cuda"); fit = torch.zeros_like(fbe)
d2._engine(0).mc_round(2.5, 1500, seed=1, frame0=7, llr_sign=1, counters=cnt, frame_bit_errors=fbe, frame_iterations=fit)
d3 = L.Neural2DMinSumDecoder(L.create_test_ldpc_code(), 2, 10)
for _ in range(3):
    d3(torch.randn(77, 7, device="cuda"))
torch.cuda.synchronize()
print("probe done", d2._engine(0).profile_read(), cnt.tolist())
