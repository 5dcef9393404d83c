"""Checks of degree 9..64 run in the bulk-async row-ring kernel (cn_wide_kernel).  Parity against the
oracle over: every boundary degree (9, 32, 33, 64) next to the classes on either side (8, 65), work items
with a partial last slab, batches that end inside a CTA's 512-frame block, per-edge weights (N-NMS),
per-check weights (type 2), RCQ codes, the float64 path, frames that stop at different iterations (frozen
messages must survive the masked stores) and the iteration-0 gather through the slot->variable map."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _code(L, rng, degrees, n=160, T=8, extra_small=6):
    m = len(degrees) + extra_small
    H = np.zeros((m, n), dtype=np.int64)
    for i, d in enumerate(degrees):
        H[i, rng.choice(n, d, replace=False)] = 1
    for i in range(len(degrees), m):
        H[i, rng.choice(n, int(rng.integers(2, 9)), replace=False)] = 1
    return L.LDPCCode(n, n - m, H, max_iterations=T), H


def _llr(rng, B, n):
    # a mix of confident, marginal and adversarial frames so that stopping iterations differ
    a = 2.0 + 2.0 * rng.standard_normal((B, n))
    a[: B // 3] = 6.0 + 1.5 * rng.standard_normal((B // 3, n))
    a[B // 3: B // 2] *= -1.0
    return a


@pytest.mark.parametrize("B", [1, 130, 513, 1400])
def test_wide_ring_all_decoders_vs_oracle(built_lib, B):
    from oracle import capi as O
    from oracle.restatement import MODE_RCQ, MODE_WRCQ, SparseGraph, quantizer_schedule
    L = built_lib
    rng = np.random.default_rng(100 + B)
    T = 8
    # 11 checks of degree 9 -> two work items (8 + 3 checks: 72 and 27 rows, both with a partial slab)
    degrees = [9] * 11 + [8, 8, 10, 17, 31, 32, 32, 33, 47, 64, 64, 65, 12, 12, 12]
    code, H = _code(L, rng, degrees, T=T)
    og = SparseGraph.from_dense(H)
    llr = _llr(rng, B, code.n)
    llr32 = llr.astype(np.float32)

    torch.manual_seed(B)
    nn = L.NeuralMinSumDecoder(code, max_iterations=T)      # per-edge beta
    with torch.no_grad():
        nn._beta_table.uniform_(0.4, 1.0)
    b, p, i = nn(torch.from_numpy(llr32).cuda())
    ref = O.decode(og, llr32, T=T, beta=nn._beta_table.detach().numpy(), nthreads=8)
    assert np.array_equal(b.cpu().numpy().reshape(ref.bits.shape), ref.bits)
    assert np.array_equal(np.atleast_1d(np.asarray(i.cpu() if torch.is_tensor(i) else i)), ref.iterations)
    assert np.array_equal(p.cpu().numpy().reshape(ref.posterior.shape), ref.posterior)
    if B > 100:
        assert len(set(ref.iterations.tolist())) > 1

    d2 = L.Neural2DMinSumDecoder(code, 2, T)                # per-check beta, per-variable alpha
    with torch.no_grad():
        d2._beta_table.uniform_(0.5, 1.0)
        d2._alpha_table.uniform_(0.8, 1.05)
    b, p, i = d2(torch.from_numpy(llr32).cuda())
    ref = O.decode(og, llr32, T=T, beta=d2._beta_table.detach().numpy()[:, d2._beta_index],
                   alpha=d2._alpha_table.detach().numpy()[:, d2._alpha_index], nthreads=8)
    assert np.array_equal(b.cpu().numpy().reshape(ref.bits.shape), ref.bits)
    assert np.array_equal(p.cpu().numpy().reshape(ref.posterior.shape), ref.posterior)

    bb, ss, ii = L.BasicMinSumDecoder(code, 0.75).decode(llr if B > 1 else llr[0])   # float64, V = 2
    ref = O.decode(og, llr, T=T, dtype=np.float64, beta=np.full((T, og.E), 0.75), nthreads=8)
    assert np.array_equal(np.asarray(bb).reshape(ref.bits.shape), ref.bits)
    assert np.array_equal(np.atleast_1d(ii), ref.iterations) and np.array_equal(np.atleast_1d(ss), ref.success)

    for bc, qp in ((3, [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]), (6, [(9.0, 1.1)])):
        rcq = L.RCQMinSumDecoder(code, bc, 8, qp, max_iterations=T)
        thr = np.array([q.thresholds for q in rcq.quantizers]).astype(np.float32)
        qoi = quantizer_schedule(T, len(qp))
        b, s, i = rcq.decode(torch.from_numpy(llr32).cuda())
        ref = O.decode(og, llr32, T=T, mode=MODE_RCQ, bc=bc, thresholds=thr, quantizer_of_iter=qoi, nthreads=8)
        assert np.array_equal(b.cpu().numpy().reshape(ref.bits.shape), ref.bits)
        assert np.array_equal(np.atleast_1d(np.asarray(i.cpu() if torch.is_tensor(i) else i)), ref.iterations)
        w = L.WeightedRCQDecoder(code, bc, 8, qp, weight_sharing_type=1, max_iterations=T)   # mixed dv: per-edge beta
        with torch.no_grad():
            w._beta_table.uniform_(0.6, 1.0)
        b, p, i = w(torch.from_numpy(llr32).cuda())
        ref = O.decode(og, llr32, T=T, mode=MODE_WRCQ, bc=bc, thresholds=thr, quantizer_of_iter=qoi,
                       beta=w._beta_table.detach().numpy()[:, w._beta_index], alpha=np.ones((T, code.n), np.float32),
                       nthreads=8)
        assert np.array_equal(b.cpu().numpy().reshape(ref.bits.shape), ref.bits)
        assert np.array_equal(p.cpu().numpy().reshape(ref.posterior.shape), ref.posterior)


def test_wide_ring_matches_register_streaming_kernel(built_lib, monkeypatch):
    """A/B: the row-ring kernel and the register-streaming kernel it replaces give identical outputs on the
    (9472,8192)-shaped QC code (dc = 29/30) at a batch the oracle would take minutes for."""
    L = built_lib
    code = L.codes.qc_shaped(max_iterations=10)
    B = 4096 + 300
    q = B // 4
    llr = torch.cat([L.awgn_llr(code.n, q if k < 3 else B - 3 * q, snr, seed=5 + k, frame0=0, llr_sign=1, device=0)
                     for k, snr in enumerate((3.0, 5.0, 6.5, 9.0))])

    def run():
        torch.manual_seed(1)
        dec = L.Neural2DMinSumDecoder(code, 2, 10)
        with torch.no_grad():
            dec._beta_table.uniform_(0.6, 0.9)
        return dec(llr)

    monkeypatch.setenv("LDPC_WIDE_RING", "0")
    b0, p0, i0 = run()
    monkeypatch.setenv("LDPC_WIDE_RING", "1")
    b1, p1, i1 = run()
    assert torch.equal(b0, b1) and torch.equal(p0, p1) and torch.equal(i0, i1)
    assert len(set(i1.tolist())) > 1   # frames stop at different iterations (frozen messages, masked stores)


@pytest.mark.parametrize("B", [3, 200, 700])
def test_wide_variables_stage_kernel_vs_oracle(built_lib, B):
    """Variables of degree 9..64 run in vn_wide_kernel (inputs staged once in shared memory, library summation
    orders evaluated from the stage): every boundary degree (8 | 9, 64 | 65), all decoder families, posterior
    output (frozen messages), early stop, float64."""
    from oracle import capi as O
    from oracle.restatement import MODE_RCQ, MODE_WRCQ, SparseGraph, quantizer_schedule
    L = built_lib
    rng = np.random.default_rng(7 + B)
    T = 8
    m, degs = 90, [8, 9, 9, 10, 12, 13, 15, 16, 17, 24, 31, 32, 33, 40, 63, 64, 65, 70] + [2, 3, 4] * 14
    n = len(degs)
    H = np.zeros((m, n), dtype=np.int64)
    for j, d in enumerate(degs):
        H[rng.choice(m, d, replace=False), j] = 1
    code = L.LDPCCode(n, max(1, n - m), H, max_iterations=T)
    og = SparseGraph.from_dense(H)
    llr = 1.5 + 2.5 * rng.standard_normal((B, n))
    llr[: B // 2] = 5.0 + 1.0 * rng.standard_normal((B // 2, n))
    llr32 = llr.astype(np.float32)

    torch.manual_seed(B)
    for wtype in (1, 3):                                  # type 1: beta(dc, dv); type 3: alpha(dv)
        d2 = L.Neural2DMinSumDecoder(code, wtype, T)
        with torch.no_grad():
            if d2._beta_table is not None:
                d2._beta_table.uniform_(0.2, 0.6)         # many inputs per variable: keep the messages small
            if d2._alpha_table is not None:
                d2._alpha_table.uniform_(0.2, 0.5)
        b, p, i = d2(torch.from_numpy(llr32).cuda())       # forward(): posterior wanted -> FREEZE variants
        beta = (d2._beta_table.detach().numpy()[:, d2._beta_index] if d2._beta_table is not None
                else np.full((T, og.E), np.float32(0.7)))
        alpha = d2._alpha_table.detach().numpy()[:, d2._alpha_index] if d2._alpha_table is not None else None
        ref = O.decode(og, llr32, T=T, beta=beta, alpha=alpha, nthreads=8)
        assert np.array_equal(b.cpu().numpy(), ref.bits) and np.array_equal(i.cpu().numpy(), ref.iterations)
        assert np.array_equal(p.cpu().numpy(), ref.posterior)
        _, _, it2, su2 = d2._engine(0).decode_device(torch.from_numpy(llr32).cuda())   # no posterior: mask-free variants
        assert np.array_equal(it2.cpu().numpy(), ref.iterations) and np.array_equal(su2.cpu().numpy().astype(bool), ref.success)

    bb, ss, ii = L.BasicMinSumDecoder(code, 0.3).decode(llr)
    ref = O.decode(og, llr, T=T, dtype=np.float64, beta=np.full((T, og.E), 0.3), nthreads=8)
    assert np.array_equal(bb, ref.bits) and np.array_equal(ii, ref.iterations) and np.array_equal(ss, ref.success)

    for bc, qp in ((3, [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]), (5, [(6.0, 1.1), (9.0, 1.0)])):
        rcq = L.RCQMinSumDecoder(code, bc, 8, qp, max_iterations=T)
        thr = np.array([q.thresholds for q in rcq.quantizers]).astype(np.float32)
        qoi = quantizer_schedule(T, len(qp))
        b, s, i = rcq.decode(torch.from_numpy(llr32).cuda())
        ref = O.decode(og, llr32, T=T, mode=MODE_RCQ, bc=bc, thresholds=thr, quantizer_of_iter=qoi, nthreads=8)
        assert np.array_equal(b.cpu().numpy(), ref.bits) and np.array_equal(i.cpu().numpy(), ref.iterations)
        w = L.WeightedRCQDecoder(code, bc, 8, qp, weight_sharing_type=2, max_iterations=T)
        with torch.no_grad():
            w._beta_table.uniform_(0.2, 0.5)
        b, p, i = w(torch.from_numpy(llr32).cuda())
        ref = O.decode(og, llr32, T=T, mode=MODE_WRCQ, bc=bc, thresholds=thr, quantizer_of_iter=qoi,
                       beta=w._beta_table.detach().numpy()[:, w._beta_index],
                       alpha=(w._alpha_table.detach().numpy()[:, w._alpha_index] if w._alpha_table is not None
                              else np.ones((T, n), np.float32)), nthreads=8)
        assert np.array_equal(b.cpu().numpy(), ref.bits) and np.array_equal(i.cpu().numpy(), ref.iterations)
        assert np.array_equal(p.cpu().numpy(), ref.posterior)
