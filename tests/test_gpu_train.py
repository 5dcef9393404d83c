"""Posterior training (csrc/ldpc_train.cu, training_framework.py) against gradients of the LIVE reference decoders
differentiated by torch.autograd (tests/golden/train_golden.npz, made by tests/golden/make_golden_train.py with the
repaired train_epoch body of oracle/reference_training_repairs.patch).  Forward results are bit-exact; gradients are
float32 sums over frames in a different order (atomics), hence the tolerance rtol 2e-4 / atol 2e-6."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_golden.npz")
CASES = ["hamming74_n2d1", "hamming74_n2d2", "hamming74_n2d3", "hamming74_n2d4", "hamming74_nnms",
         "irregular48_n2d2", "irregular48_n2d1"]
RTOL, ATOL = 2e-4, 2e-6


def _load(L, z, name):
    H, T, kind = z[f"{name}/H"], int(z[f"{name}/T"]), str(z[f"{name}/kind"])
    code = L.LDPCCode(H.shape[1], H.shape[1] - H.shape[0], H, max_iterations=T)
    dec = L.NeuralMinSumDecoder(code, T) if kind == "nnms" else L.Neural2DMinSumDecoder(code, int(kind[-1]), T)
    keys = [str(k) for k in z[f"{name}/keys"]]
    dec.load_state_dict({k: torch.tensor([v]) for k, v in zip(keys, z[f"{name}/weights"])})
    dec.differentiable = True      # (off by default: inference batches must not keep the message history)
    return dec, keys


def _by_key(dec, keys, tables):
    """Values of (beta table, alpha table)-shaped tensors at the reference's keys, in the golden order."""
    out = []
    for full in keys:
        name, _, key = full.partition(".")
        t, c = getattr(dec, name).position(key)
        out.append(float(tables[name][t, c]))
    return np.array(out, dtype=np.float32)


@pytest.mark.parametrize("name", CASES)
def test_gradients_match_the_reference_autograd(built_lib, name):
    L = built_lib
    z = np.load(GOLD)
    dec, keys = _load(L, z, name)
    llr = torch.from_numpy(z[f"{name}/llr"]).cuda()
    decoded, post, iters = dec(llr)
    assert post.requires_grad and decoded.dtype == torch.int32
    assert np.array_equal(decoded.cpu().numpy().astype(np.uint8), z[f"{name}/bits"])
    assert np.array_equal(iters.cpu().numpy(), z[f"{name}/iterations"])
    assert np.array_equal(post.detach().cpu().numpy(), z[f"{name}/posterior"])
    loss = F.binary_cross_entropy_with_logits(-post, torch.zeros_like(post))
    assert float(loss) == pytest.approx(float(z[f"{name}/loss"]), rel=1e-6)
    loss.backward()
    grads = {"beta_weights": dec._beta_table.grad if dec._beta_table is not None else None,
             "alpha_weights": dec._alpha_table.grad if dec._alpha_table is not None else None}
    got = _by_key(dec, keys, grads)
    want = z[f"{name}/grads"]
    assert np.abs(want).max() > 1e-3
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)
    # a CPU batch and a single frame take the same path
    dec.zero_grad()
    d2, p2, i2 = dec(torch.from_numpy(z[f"{name}/llr"]))
    assert p2.device.type == "cpu" and torch.equal(p2.detach(), post.detach().cpu())
    F.binary_cross_entropy_with_logits(-p2, torch.zeros_like(p2)).backward()
    np.testing.assert_allclose(_by_key(dec, keys, {"beta_weights": dec._beta_table.grad if dec._beta_table is not None else None,
                                                   "alpha_weights": dec._alpha_table.grad if dec._alpha_table is not None else None}),
                               want, rtol=RTOL, atol=ATOL)
    d1, p1, i1 = dec(llr[2])
    assert isinstance(i1, int) and i1 == int(z[f"{name}/iterations"][2]) and p1.requires_grad
    with torch.no_grad():
        _, p0, _ = dec(llr)
    assert not p0.requires_grad and torch.equal(p0, post.detach())


@pytest.mark.parametrize("name", ["hamming74_n2d2", "hamming74_nnms", "irregular48_n2d1"])
def test_three_adam_steps_follow_the_reference(built_lib, name):
    L = built_lib
    z = np.load(GOLD)
    dec, keys = _load(L, z, name)
    llr = torch.from_numpy(z[f"{name}/llr"]).cuda()
    opt = torch.optim.Adam(dec.parameters(), lr=0.01)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        _, post, _ = dec(llr)
        loss = F.binary_cross_entropy_with_logits(-post, torch.zeros_like(post))
        loss.backward()
        opt.step()
        losses.append(float(loss))
    np.testing.assert_allclose(losses, z[f"{name}/adam_losses"], rtol=2e-5)
    tables = {"beta_weights": dec._beta_table.detach() if dec._beta_table is not None else None,
              "alpha_weights": dec._alpha_table.detach() if dec._alpha_table is not None else None}
    np.testing.assert_allclose(_by_key(dec, keys, tables), z[f"{name}/adam_weights"], rtol=0, atol=3e-5)


def test_trainer_and_analyzer_surface(built_lib):
    L = built_lib
    from ldpc_b200.training_framework import GradientExplosionAnalyzer, PosteriorJointTrainer, TrainingConfig
    code = L.create_test_ldpc_code()
    torch.manual_seed(0)
    np.random.seed(0)
    dec = L.Neural2DMinSumDecoder(code, 2, 5)
    with torch.no_grad():
        dec._beta_table.fill_(0.5)
        dec._alpha_table.fill_(0.7)
    before = dec._beta_table.detach().clone()
    tr = PosteriorJointTrainer(dec, TrainingConfig(batch_size=64, num_epochs=4, learning_rate=0.02, snr_range=(1.0, 5.0)))
    llrs, targets = tr.generate_training_data(code, 16)
    assert llrs.shape == (16, 7) and llrs.dtype == torch.float32 and not targets.any()
    hist = tr.train(code, num_train_samples=256, num_val_samples=64)
    assert set(hist) == {"train_losses", "train_accuracies", "gradient_norms"} and len(hist["train_losses"]) == 4
    assert all(np.isfinite(hist["train_losses"])) and all(g > 0 for g in hist["gradient_norms"])
    assert not torch.equal(before, dec._beta_table.detach())
    assert hist["train_losses"][-1] < hist["train_losses"][0]            # Adam on a fixed data set: the loss goes down
    res = GradientExplosionAnalyzer(dec, code).analyze_gradient_explosion(num_samples=5)
    assert set(res) == {"gradient_magnitudes", "iteration_counts", "mean_gradient", "std_gradient", "max_gradient"}
    assert len(res["gradient_magnitudes"]) == 5 and res["max_gradient"] >= res["mean_gradient"] >= 0
    # decoders without a backward pass decode under autograd like under no_grad
    w = L.WeightedRCQDecoder(code, 3, 8, [(3.0, 1.3)], weight_sharing_type=2, max_iterations=4)
    _, p, _ = w(llrs[0])
    assert not p.requires_grad
    eng = L.RCQMinSumDecoder(code, 3, 8, [(3.0, 1.3)], max_iterations=4)._engine(0)
    with pytest.raises(L.LdpcError):
        eng.train_forward(llrs.cuda())


def test_backward_needs_its_own_forward(built_lib):
    L = built_lib
    code = L.create_test_ldpc_code()
    dec = L.Neural2DMinSumDecoder(code, 3, 4)
    x = torch.randn(8, 7, device="cuda") * 2
    assert not dec(x)[1].requires_grad         # differentiable only when asked for
    dec.differentiable = True
    _, p1, _ = dec(x)
    _, p2, _ = dec(x + 1)                     # overwrites the message history of the first pass
    with pytest.raises(RuntimeError):
        p1.sum().backward()
    p2.sum().backward()
    assert dec._beta_table.grad is not None and dec._beta_table.grad.shape == dec._beta_table.shape
    dec.zero_grad()
    dec.max_iterations = 2                    # fewer iterations than rows: the unused rows get zero gradient
    _, p3, _ = dec(x)
    p3.sum().backward()
    assert not dec._beta_table.grad[2:].any() and dec._beta_table.grad[:2].any()
