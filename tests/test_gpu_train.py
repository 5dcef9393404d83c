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


# ---------------------------------------------------------------------------------------------------------------
# The backward kernels against torch.autograd on a plain-PyTorch restatement of the reference's forward pass
# (neural_2d_decoder.py:133-225, neural_minsum_decoder.py:58-150: the same torch ops -- sign, abs, argmin, the
# inf-masked torch.min, prod of the other signs, sum -- on sparse neighbour lists), one frame at a time on the CPU.
# Inputs on a coarse grid, so exact zeros (three-valued sign), ties of both minima and a degree-1 check all occur:
# the cases the three-pass check-side kernel hands to its general form.
# ---------------------------------------------------------------------------------------------------------------
def _autograd_forward(H, llr, weights, kind, T):
    m, n = H.shape
    rows = [np.nonzero(H[i])[0] for i in range(m)]
    cols = [np.nonzero(H[:, j])[0] for j in range(n)]
    dcs, dvs = H.sum(1), H.sum(0)
    Hf = torch.tensor(H, dtype=torch.float32)

    def beta(t, i, j):
        if kind == "nnms":
            return weights[f"beta_weights.iter_{t}_c{i}_v{j}"]
        if kind == "n2d1":
            return weights[f"beta_weights.iter_{t}_dc{dcs[i]}_dv{dvs[j]}"]
        if kind in ("n2d2", "n2d3"):
            return weights[f"beta_weights.iter_{t}_dc{dcs[i]}"]
        return torch.tensor(0.7)

    def alpha(t, j):
        if kind in ("n2d2", "n2d4"):
            return weights[f"alpha_weights.iter_{t}_dv{dvs[j]}"]
        return torch.tensor(1.0)

    v2c = {(j, i): llr[j] for j in range(n) for i in cols[j]}
    for t in range(T):
        c2v = {}
        for i in range(m):
            nb = rows[i]
            if len(nb) == 0:
                continue
            inc = torch.stack([v2c[(j, i)] for j in nb])
            signs, mags = torch.sign(inc), torch.abs(inc)
            k0 = torch.argmin(mags)
            mn = mags[k0]
            if len(nb) > 1:
                tmp = mags.clone()
                tmp[k0] = float("inf")
                mn2 = torch.min(tmp)
            else:
                mn2 = mn
            for k, j in enumerate(nb):
                sp = torch.prod(signs[torch.arange(len(nb)) != k])
                c2v[(i, j)] = beta(t, i, j) * (mn2 if k == int(k0) else mn) * sp
        for j in range(n):
            for i in cols[j]:
                others = [c2v[(i2, j)] for i2 in cols[j] if i2 != i]
                s = torch.sum(torch.stack(others)) if others else torch.tensor(0.0)
                v2c[(j, i)] = llr[j] + alpha(t, j) * s
        post = torch.stack([llr[j] + (torch.sum(torch.stack([c2v[(i, j)] for i in cols[j]])) if len(cols[j]) else 0.0)
                            for j in range(n)])
        bits = (post < 0).float()
        if float(torch.sum(torch.matmul(Hf, bits) % 2)) == 0:
            return post, t + 1
    return post, T


def _grid_case(seed):
    rng = np.random.default_rng(seed)
    m, n = 10, 20
    H = np.zeros((m, n), dtype=np.int64)
    H[0, 3] = 1                                            # a degree-1 check
    H[1, rng.choice(n, 11, replace=False)] = 1             # a wide one
    for i in range(2, m):
        H[i, rng.choice(n, int(rng.integers(2, 7)), replace=False)] = 1
    for j in range(n):                                     # no isolated variables
        if not H[:, j].any():
            H[int(rng.integers(1, m)), j] = 1
    llr = (rng.integers(0, 9, size=(24, n)) * 0.5).astype(np.float32)    # zeros and equal magnitudes everywhere
    llr *= np.where(rng.random((24, n)) < 0.12, -1.0, 1.0).astype(np.float32)   # a few wrong signs: frames stop at different iterations
    return H, llr


@pytest.mark.parametrize("kind", ["n2d2", "n2d1", "n2d3", "n2d4", "nnms"])
def test_backward_against_torch_autograd_with_zeros_and_ties(built_lib, kind):
    L = built_lib
    T = 4
    H, llr_np = _grid_case(11)
    code = L.LDPCCode(H.shape[1], H.shape[1] - H.shape[0], H, max_iterations=T)
    dec = L.NeuralMinSumDecoder(code, T) if kind == "nnms" else L.Neural2DMinSumDecoder(code, int(kind[-1]), T)
    g = torch.Generator().manual_seed(5)
    sd = {k: (0.4 + 0.6 * torch.rand(1, generator=g)) * (-1.0 if i % 7 == 3 else 1.0) for i, k in enumerate(dec.state_dict())}
    dec.load_state_dict(sd)
    dec.differentiable = True
    keys = list(sd)
    # plain PyTorch, frame by frame
    weights = {k: v.clone().squeeze(0).requires_grad_(True) for k, v in sd.items()}
    posts, its = [], []
    for f in range(llr_np.shape[0]):
        p, it = _autograd_forward(H, torch.from_numpy(llr_np[f]), weights, kind, T)
        posts.append(p)
        its.append(it)
    want_post = torch.stack(posts)
    F.binary_cross_entropy_with_logits(-want_post, torch.zeros_like(want_post)).backward()
    want = np.array([0.0 if weights[k].grad is None else float(weights[k].grad) for k in keys], dtype=np.float32)
    assert len(set(its)) > 1 and np.abs(want).max() > 1e-3
    # the kernels
    _, post, iters = dec(torch.from_numpy(llr_np).cuda())
    assert iters.cpu().tolist() == its
    np.testing.assert_allclose(post.detach().cpu().numpy(), want_post.detach().numpy(), rtol=1e-6, atol=1e-6)
    F.binary_cross_entropy_with_logits(-post, torch.zeros_like(post)).backward()
    got = _by_key(dec, keys, {"beta_weights": dec._beta_table.grad if dec._beta_table is not None else None,
                              "alpha_weights": dec._alpha_table.grad if dec._alpha_table is not None else None})
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("kind", ["n2d2", "n2d1", "nnms"])
def test_three_pass_check_backward_equals_general_form_at_full_size(built_lib, monkeypatch, kind):
    """The (16200,7200)-shaped code, early stop on, 256 frames at 6 dB and 256 at 2 dB (frames stop at different
    iterations, many run to the end): weight gradients of the three-pass check-side kernel against the general four-pass form taken for
    every check (LDPC_TRAIN_GENERAL=1), which is the form the golden gradients of the live reference pinned first."""
    L = built_lib
    T = 8
    code = L.codes.dvbs2_shaped(max_iterations=T)
    llr = torch.cat([L.awgn_llr(code.n, 256, 6.0, seed=3, llr_sign=1), L.awgn_llr(code.n, 256, 2.0, seed=4, llr_sign=1)])
    grads = []
    for general in ("0", "1"):
        monkeypatch.setenv("LDPC_TRAIN_GENERAL", general)
        torch.manual_seed(1)
        dec = L.NeuralMinSumDecoder(code, T) if kind == "nnms" else L.Neural2DMinSumDecoder(code, int(kind[-1]), T)
        with torch.no_grad():
            for p in dec.parameters():
                p.uniform_(0.55, 0.95)
        dec.differentiable = True
        _, post, iters = dec(llr)
        F.binary_cross_entropy_with_logits(-post, torch.zeros_like(post)).backward()
        grads.append([None if p.grad is None else p.grad.detach().cpu().numpy().copy() for p in dec.parameters()])
        if general == "0":
            assert len(set(iters.cpu().tolist())) > 1
        dec._engine(0).close()
    assert any(g is not None and np.abs(g).max() > 1e-6 for g in grads[0])
    for a, b in zip(*grads):
        assert (a is None) == (b is None)
        if a is not None:
            np.testing.assert_allclose(a, b, rtol=2e-4, atol=1e-7 + 2e-4 * np.abs(b).max())
