"""One posterior-training step (forward with message history + backward) at a BASELINE code size:
    python tools/train_probe.py [frames]"""
import json, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import bench
import ldpc_b200 as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = 10
code = bench.make_code(L, "dvbs2", T)
E, n = code.graph.E, code.n
for kind, wt in (("n2d2", 2), ("n2d1", 1), ("nnms", 0)):
    dec = L.NeuralMinSumDecoder(code, T) if kind == "nnms" else L.Neural2DMinSumDecoder(code, wt, T)
    dec.differentiable = True
    dec = dec.cuda() if hasattr(dec, "cuda") else dec
    llr = L.awgn_llr(n, B, 2.0, seed=1, llr_sign=1)
    eng = dec._engine(0)

    def fwd():
        return eng.train_forward(llr)

    def step():
        bits, post, iters, _ = eng.train_forward(llr)
        g = torch.sigmoid(-post) * (-1.0 / post.numel())          # d BCE-with-logits(-post, 0) / d post
        return eng.train_backward(g)

    def timed(fn, reps=3):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    t_f = timed(fwd)
    t_s = timed(step)
    it = float(eng.train_forward(llr)[2].float().mean())
    # forward: 16E + 4n per frame-iteration (history slices are the message arrays themselves); backward: v2c and c2v
    # history read (8E), gradient rows read and written (16E)
    fwd_bytes = (16 * E + 4 * n) * it * B
    bwd_bytes = 24 * E * it * B
    print(json.dumps({"decoder": kind, "frames": B, "avg_iterations": round(it, 2), "forward_ms": round(t_f, 2), "step_ms": round(t_s, 2),
                      "backward_ms": round(t_s - t_f, 2), "frames_per_s": round(B / t_s * 1e3),
                      "forward_gbs": round(fwd_bytes / t_f / 1e6), "backward_gbs": round(bwd_bytes / (t_s - t_f) / 1e6)}), flush=True)
    eng.close()
    del dec, eng
