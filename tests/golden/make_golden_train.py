"""Golden gradients for posterior training, from the LIVE reference decoders differentiated by torch.autograd.

Runs only in the build container (needs /root/reference).  Usage:
    python tests/golden/make_golden_train.py            # rewrites tests/golden/train_*.npz

The reference's trainer (training_framework.py:106-172) crashes as shipped; with the two repairs of
oracle/reference_training_repairs.patch a training step is: forward() of every frame of the batch (the module takes
one LLR vector at a time), loss = binary_cross_entropy_with_logits(-stack(posteriors), targets), loss.backward().
That is what is run here on the reference's own modules, and what is recorded per case: H, the LLR batch, the weights
(key -> value), loss, per-key gradients, posteriors and iteration counts.  A second record per case runs three Adam
steps (lr 0.01) of the repaired reference trainer's inner loop and stores the weights after them.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from oracle import ref_shim  # noqa: E402
from make_golden import code_hamming, code_irregular  # noqa: E402


def batch_llrs(rng, n, frames):
    """Frames that stop after different numbers of iterations: mixed SNRs in the converging convention, one frame in
    the reference convention (never converges), one with exact zeros and ties."""
    out = []
    for f in range(frames):
        snr_db = [1.0, 3.0, 5.0, 7.0, 2.0, 9.0][f % 6]
        sigma2 = 10 ** (-snr_db / 10)
        s = -1.0 if f == 2 else 1.0
        out.append(2 * (s + np.sqrt(sigma2) * rng.standard_normal(n)) / sigma2)
    out = np.array(out)
    if frames >= 5:
        out[4, ::4] = 0.0
        out[4, 1::6] = np.round(out[4, 1::6])          # ties between magnitudes
    return out.astype(np.float32)


def build(ref, kind, code, T, rng):
    if kind == "nnms":
        m = ref.neural_minsum_decoder.NeuralMinSumDecoder(code, T)
    else:
        m = ref.neural_2d_decoder.Neural2DMinSumDecoder(code, int(kind[-1]), T)
    with torch.no_grad():
        for name, p in m.named_parameters():
            lo, hi = (0.45, 0.95) if name.startswith("beta") else (0.6, 1.0)
            p.copy_(torch.tensor([rng.uniform(lo, hi)], dtype=torch.float32))
    return m


def step(model, llrs, targets):
    outs = [model(llrs[i]) for i in range(llrs.shape[0])]           # the repaired train_epoch body
    decoded = torch.stack([o[0] for o in outs])
    posteriors = torch.stack([o[1] for o in outs])
    loss = F.binary_cross_entropy_with_logits(-posteriors, targets.float())
    return loss, decoded, posteriors, [int(o[2]) for o in outs]


def main():
    ref = ref_shim.load()
    rng = np.random.default_rng(2024)
    cases = [("hamming74", code_hamming(ref), k, 6, 12) for k in ("n2d1", "n2d2", "n2d3", "n2d4", "nnms")]
    cases += [("irregular48", code_irregular(ref), k, 4, 6) for k in ("n2d2", "n2d1")]
    out = {}
    for stem, code, kind, T, B in cases:
        name = f"{stem}_{kind}"
        model = build(ref, kind, code, T, rng)
        llrs = torch.from_numpy(batch_llrs(rng, code.n, B))
        targets = torch.zeros(B, code.n)
        keys = [k for k, _ in model.named_parameters()]
        w0 = np.array([float(p.detach()) for _, p in model.named_parameters()], dtype=np.float32)
        loss, decoded, post, iters = step(model, llrs, targets)
        loss.backward()
        grads = np.array([0.0 if p.grad is None else float(p.grad) for _, p in model.named_parameters()], dtype=np.float32)
        out[f"{name}/H"] = np.asarray(code.H)
        out[f"{name}/T"] = np.int64(T)
        out[f"{name}/kind"] = np.str_(kind)
        out[f"{name}/llr"] = llrs.numpy()
        out[f"{name}/keys"] = np.array(keys)
        out[f"{name}/weights"] = w0
        out[f"{name}/loss"] = np.float32(loss.item())
        out[f"{name}/grads"] = grads
        out[f"{name}/posterior"] = post.detach().numpy()
        out[f"{name}/iterations"] = np.array(iters, dtype=np.int32)
        out[f"{name}/bits"] = decoded.numpy().astype(np.uint8)
        # three Adam steps on the same batch
        opt = torch.optim.Adam(model.parameters(), lr=0.01)
        losses = []
        for _ in range(3):
            opt.zero_grad()
            loss, *_ = step(model, llrs, targets)
            loss.backward()
            opt.step()
            losses.append(loss.item())
        out[f"{name}/adam_losses"] = np.array(losses, dtype=np.float32)
        out[f"{name}/adam_weights"] = np.array([float(p.detach()) for _, p in model.named_parameters()], dtype=np.float32)
        print(name, "loss", out[f"{name}/loss"], "iterations", iters, "|grad|", float(np.linalg.norm(grads)), flush=True)
    np.savez_compressed(os.path.join(HERE, "train_golden.npz"), **out)


if __name__ == "__main__":
    main()
