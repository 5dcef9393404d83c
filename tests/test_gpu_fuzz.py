"""Seeded random sweep over graph shapes, batch sizes, iteration counts and decoder families against the
oracle: catches interactions between the degree-specialised kernels (<= 8 | 9..64 | > 64 on both node
sides), frame compaction, frozen / mask-free message stores and the float64 path that the targeted tests
exercise one at a time."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _random_graph(rng):
    """Irregular H with a few heavy rows / columns, possibly empty ones, in one of three size classes."""
    cls = rng.integers(0, 3)
    if cls == 0:
        m, n = int(rng.integers(3, 12)), int(rng.integers(6, 30))
    elif cls == 1:
        m, n = int(rng.integers(20, 60)), int(rng.integers(60, 160))
    else:
        m, n = int(rng.integers(70, 100)), int(rng.integers(100, 140))
    H = np.zeros((m, n), dtype=np.int64)
    for i in range(m):
        d = int(rng.integers(1, min(n, 9)))
        H[i, rng.choice(n, d, replace=False)] = 1
    for _ in range(int(rng.integers(0, 4))):          # heavy checks
        d = int(rng.choice([9, 16, 31, 33, 64, 65, min(n, 80)]))
        H[rng.integers(0, m), rng.choice(n, min(d, n), replace=False)] = 1
    for _ in range(int(rng.integers(0, 4))):          # heavy variables
        d = int(rng.choice([9, 12, 16, 17, 40, 64, 66]))
        H[rng.choice(m, min(d, m), replace=False), rng.integers(0, n)] = 1
    if rng.random() < 0.3:
        H[rng.integers(0, m), :] = 0                  # empty check
    if rng.random() < 0.3:
        H[:, rng.integers(0, n)] = 0                  # degree-0 variable
    return H


@pytest.mark.parametrize("seed", range(36))
def test_random_configurations_vs_oracle(built_lib, monkeypatch, seed):
    from oracle import capi as O
    from oracle.restatement import MODE_NMS, MODE_OFFSET, MODE_RCQ, MODE_WRCQ, SparseGraph, quantizer_schedule
    L = built_lib
    rng = np.random.default_rng(9000 + seed)
    monkeypatch.setenv("LDPC_COMPACT_MIN_FRAMES", "128")
    monkeypatch.setenv("LDPC_GRAPHS", str(seed % 2))    # odd: graph replay; even: checkpoints + frame compaction
    H = _random_graph(rng)
    m, n = H.shape
    T = int(rng.integers(1, 14))
    B = int(rng.choice([1, 2, 127, 128, 129, 300, 900]))
    code = L.LDPCCode(n, max(1, n - m), H, max_iterations=T)
    og = SparseGraph.from_dense(H)
    scale = rng.choice([0.5, 2.0, 6.0])
    llr = scale * (1.0 + 1.2 * rng.standard_normal((B, n)))
    llr[rng.random((B, n)) < 0.02] = 0.0              # exact zeros (three-valued sign in the reference)
    llr32 = llr.astype(np.float32)
    kind = ["n2d", "nnms", "basic", "rcq", "wrcq", "oms"][seed % 6]
    torch.manual_seed(seed)
    x = torch.from_numpy(llr32).cuda()
    if kind == "n2d":
        wt = int(rng.integers(1, 5))
        dec = L.Neural2DMinSumDecoder(code, wt, T)
        with torch.no_grad():
            if dec._beta_table is not None:
                dec._beta_table.uniform_(0.2, 0.9)
            if dec._alpha_table is not None:
                dec._alpha_table.uniform_(0.3, 1.0)
        beta = (dec._beta_table.detach().numpy()[:, dec._beta_index] if dec._beta_table is not None
                else np.full((T, og.E), np.float32(0.7)))
        alpha = dec._alpha_table.detach().numpy()[:, dec._alpha_index] if dec._alpha_table is not None else None
        ref = O.decode(og, llr32, T=T, mode=MODE_NMS, beta=beta, alpha=alpha, nthreads=8)
        b, p, i = dec(x)
        assert np.array_equal(p.cpu().numpy(), ref.posterior)
        _, _, i2, s2 = dec._engine(0).decode_device(x)          # decode without posterior: mask-free kernels
        assert np.array_equal(i2.cpu().numpy(), ref.iterations) and np.array_equal(s2.cpu().numpy().astype(bool), ref.success)
    elif kind == "nnms":
        dec = L.NeuralMinSumDecoder(code, max_iterations=T)
        with torch.no_grad():
            dec._beta_table.uniform_(0.2, 0.9)
        ref = O.decode(og, llr32, T=T, mode=MODE_NMS, beta=dec._beta_table.detach().numpy(), nthreads=8)
        b, p, i = dec(x)
        assert np.array_equal(p.cpu().numpy(), ref.posterior)
    elif kind == "oms":                                   # offset rule: three-valued sign, alpha at the check node
        if rng.random() < 0.5:
            dec = L.Neural2DOffsetMinSumDecoder(code, int(rng.integers(1, 5)), T)
        else:
            dec = L.NeuralOffsetMinSumDecoder(code, T)
        with torch.no_grad():
            if dec._beta_table is not None:
                dec._beta_table.uniform_(-0.1, 0.6)
            if dec._alpha_table is not None:
                dec._alpha_table.uniform_(-0.05, 0.2)
        bt, at = dec._tables()
        beta = bt[:, dec._beta_index] if bt is not None else None
        alpha = at[:, dec._alpha_index] if at is not None else None
        ref = O.decode(og, llr32, T=T, mode=MODE_OFFSET, beta=beta, alpha=alpha, nthreads=8)
        b, p, i = dec(x)
        assert np.array_equal(p.cpu().numpy(), ref.posterior)
    elif kind == "basic":
        f = float(rng.choice([0.3, 0.7, 1.0]))
        ref = O.decode(og, llr, T=T, dtype=np.float64, beta=np.full((T, og.E), f), nthreads=8)
        b, s, i = L.BasicMinSumDecoder(code, f).decode(llr)
        assert np.array_equal(np.asarray(s), ref.success)
        b, i = torch.from_numpy(np.asarray(b)), torch.from_numpy(np.asarray(i))
    else:
        bc = int(rng.choice([2, 3, 4, 6]))
        qp = [(float(rng.uniform(2, 8)), float(rng.uniform(0.8, 1.5))) for _ in range(int(rng.integers(1, 4)))]
        if kind == "rcq":
            dec = L.RCQMinSumDecoder(code, bc, 8, qp, max_iterations=T)
        else:
            dec = L.WeightedRCQDecoder(code, bc, 8, qp, weight_sharing_type=int(rng.integers(1, 5)), max_iterations=T)
            with torch.no_grad():
                if dec._beta_table is not None:
                    dec._beta_table.uniform_(0.3, 1.0)
                if dec._alpha_table is not None:
                    dec._alpha_table.uniform_(0.5, 1.0)
        thr = np.array([q.thresholds for q in dec.quantizers], dtype=np.float64).astype(np.float32)
        kw = dict(T=T, bc=bc, thresholds=thr, quantizer_of_iter=quantizer_schedule(T, len(qp)), nthreads=8)
        if kind == "rcq":
            ref = O.decode(og, llr32, mode=MODE_RCQ, **kw)
            b, s, i = dec.decode(x)
            assert np.array_equal(s.cpu().numpy().astype(bool), ref.success)
        else:
            beta = (dec._beta_table.detach().numpy()[:, dec._beta_index] if dec._beta_table is not None
                    else np.full((T, og.E), np.float32(0.7)))
            alpha = (dec._alpha_table.detach().numpy()[:, dec._alpha_index] if dec._alpha_table is not None
                     else np.ones((T, n), np.float32))
            ref = O.decode(og, llr32, mode=MODE_WRCQ, beta=beta, alpha=alpha, **kw)
            b, p, i = dec(x)
            assert np.array_equal(p.cpu().numpy(), ref.posterior)
    assert np.array_equal(b.cpu().numpy().astype(np.uint8), ref.bits)
    assert np.array_equal(i.cpu().numpy(), ref.iterations)
