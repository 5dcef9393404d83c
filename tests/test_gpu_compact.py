"""Frame compaction (early stop at scale): at checkpoints the frames still running are gathered into a
smaller dense batch that carries on from the same iteration, and a batch whose frames have all stopped ends
at once.  Every frame must see exactly the arithmetic of the plain schedule: results are compared with the
oracle, with the uncompacted CUDA path (LDPC_COMPACT=0), through the host-buffer pipeline and through the
Monte-Carlo round."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mixed_llr(rng, B, n, snrs, sign=1.0):
    parts = []
    per = B // len(snrs)
    for k, snr in enumerate(snrs):
        cnt = per if k < len(snrs) - 1 else B - per * (len(snrs) - 1)
        s2 = 10 ** (-snr / 10)
        parts.append(2 * (sign + np.sqrt(s2) * rng.standard_normal((cnt, n))) / s2)
    llr = np.concatenate(parts)
    rng.shuffle(llr, axis=0)     # running frames end up scattered over the lanes
    return llr


def _oracle_graph(g):
    from oracle.restatement import SparseGraph
    return SparseGraph.from_coo(g.n, g.m, g.edge_check, g.check_var)


@pytest.mark.parametrize("B", [700, 2500])
def test_compacted_decode_vs_oracle(built_lib, monkeypatch, B):
    from oracle import capi as O
    from oracle.restatement import MODE_RCQ, quantizer_schedule
    L = built_lib
    monkeypatch.setenv("LDPC_COMPACT_MIN_FRAMES", "128")
    T = 30
    code = L.codes.dvbs2_shaped(max_iterations=T, scale=20)
    g = code.graph
    og = _oracle_graph(g)
    rng = np.random.default_rng(B)
    llr = _mixed_llr(rng, B, g.n, (0.5, 2.0, 2.6, 3.2, 4.5, 6.0))
    llr32 = llr.astype(np.float32)

    torch.manual_seed(3)
    dec = L.Neural2DMinSumDecoder(code, weight_sharing_type=2, max_iterations=T)
    with torch.no_grad():
        dec._beta_table.uniform_(0.6, 0.95)
        dec._alpha_table.uniform_(0.9, 1.0)
    bits, post, iters = dec(torch.from_numpy(llr32).cuda())
    prof = dec._engine(0).profile_read()
    assert prof["compactions"] >= 2, prof
    ref = O.decode(og, llr32, T=T, beta=dec._beta_table.detach().numpy()[:, dec._beta_index],
                   alpha=dec._alpha_table.detach().numpy()[:, dec._alpha_index], nthreads=8)
    assert np.array_equal(iters.cpu().numpy(), ref.iterations)
    assert np.array_equal(bits.cpu().numpy(), ref.bits)
    assert np.array_equal(post.cpu().numpy(), ref.posterior)
    assert len(set(ref.iterations.tolist())) > 6 and (ref.iterations == T).any()

    # same frames through the host-buffer pipeline (its own root workspace, shared child levels)
    bh, ph, ih = dec(torch.from_numpy(llr32))
    assert torch.equal(bh, bits.cpu()) and torch.equal(ph, post.cpu()) and torch.equal(ih, iters.cpu())

    qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
    rcq = L.RCQMinSumDecoder(code, bc=3, bv=8, quantizer_params=qp, max_iterations=T)
    b, s, i = rcq.decode(torch.from_numpy(llr32).cuda())
    thr = np.array([q.thresholds for q in rcq.quantizers], dtype=np.float64).astype(np.float32)
    ref = O.decode(og, llr32, T=T, mode=MODE_RCQ, bc=3, thresholds=thr, quantizer_of_iter=quantizer_schedule(T, 3), nthreads=8)
    assert np.array_equal(b.cpu().numpy(), ref.bits) and np.array_equal(i.cpu().numpy(), ref.iterations)
    assert np.array_equal(s.cpu().numpy(), ref.success)
    assert rcq._engine(0).profile_read()["compactions"] >= 1

    # W-RCQ forward(): posterior of stopped frames from their frozen codes with the quantiser of THEIR last iteration
    from oracle.restatement import MODE_WRCQ
    w = L.WeightedRCQDecoder(code, bc=4, bv=8, quantizer_params=qp, weight_sharing_type=2, max_iterations=T)
    with torch.no_grad():
        w._beta_table.uniform_(0.8, 1.0)
    b, p, i = w(torch.from_numpy(llr32).cuda())
    thr4 = np.array([q.thresholds for q in w.quantizers], dtype=np.float64).astype(np.float32)
    ref = O.decode(og, llr32, T=T, mode=MODE_WRCQ, bc=4, thresholds=thr4, quantizer_of_iter=quantizer_schedule(T, 3),
                   beta=w._beta_table.detach().numpy()[:, w._beta_index],
                   alpha=(w._alpha_table.detach().numpy()[:, w._alpha_index] if w._alpha_table is not None
                          else np.ones((T, g.n), np.float32)), nthreads=8)
    assert np.array_equal(i.cpu().numpy(), ref.iterations) and np.array_equal(b.cpu().numpy(), ref.bits)
    assert np.array_equal(p.cpu().numpy(), ref.posterior)

    basic = L.BasicMinSumDecoder(code, factor=0.8)
    bb, ss, ii = basic.decode(llr)
    ref = O.decode(og, llr, T=T, dtype=np.float64, beta=np.full((T, g.E), 0.8), nthreads=8)
    assert np.array_equal(bb, ref.bits) and np.array_equal(ii, ref.iterations) and np.array_equal(ss, ref.success)


def test_all_frames_stop_early_exit(built_lib):
    """Confident frames: once every frame has stopped the decode returns at the next checkpoint instead of
    launching the remaining iterations."""
    L = built_lib
    T = 50
    code = L.codes.dvbs2_shaped(max_iterations=T, scale=20)
    llr = L.awgn_llr(code.n, 1000, 12.0, seed=3, llr_sign=1, device=0)
    dec = L.Neural2DMinSumDecoder(code, 4, T)
    bits, post, iters = dec(llr)
    prof = dec._engine(0).profile_read()
    worst = int(iters.max())
    assert worst < T - 10, "test needs frames that all stop early"
    assert prof["early_exits"] == 1, prof
    assert worst <= prof["cn_launches"] <= worst + 8, (worst, prof)   # next checkpoint is at most 8 iterations on


def test_compaction_ab_full_size(built_lib, monkeypatch):
    """Compaction on/off give identical outputs on the full (16200,7200)-shaped code, decode and Monte-Carlo."""
    L = built_lib
    T = 40
    code = L.codes.dvbs2_shaped(max_iterations=T)
    B = 6000
    llr = torch.cat([L.awgn_llr(code.n, B // 3, snr, seed=40 + k, llr_sign=1, device=0) for k, snr in enumerate((1.2, 2.0, 3.0))])
    llr = llr[torch.randperm(llr.shape[0], generator=torch.Generator().manual_seed(0)).cuda()]

    def run():
        torch.manual_seed(2)
        dec = L.Neural2DMinSumDecoder(code, 2, T)
        with torch.no_grad():
            dec._beta_table.uniform_(0.7, 0.9)
            dec._alpha_table.uniform_(0.95, 1.0)
        out = dec(llr)
        eng = dec._engine(0)
        prof = eng.profile_read()
        counters = torch.zeros(4, dtype=torch.int64, device="cuda")
        fbe = torch.zeros(4096, dtype=torch.int32, device="cuda")
        fit = torch.zeros(4096, dtype=torch.int32, device="cuda")
        eng.mc_round(2.0, 4096, seed=9, frame0=123, llr_sign=1, counters=counters, frame_bit_errors=fbe, frame_iterations=fit)
        return out, prof, counters.cpu(), fbe.cpu(), fit.cpu()

    monkeypatch.setenv("LDPC_COMPACT", "0")
    (b0, p0, i0), prof0, c0, fbe0, fit0 = run()
    monkeypatch.setenv("LDPC_COMPACT", "1")
    (b1, p1, i1), prof1, c1, fbe1, fit1 = run()
    assert prof0["compactions"] == 0 and prof1["compactions"] >= 1
    assert torch.equal(i0, i1) and torch.equal(b0, b1) and torch.equal(p0, p1)
    assert torch.equal(c0, c1) and torch.equal(fbe0, fbe1) and torch.equal(fit0, fit1)
    assert int(c1[3]) == 4096 and int(c1[2]) == int(fit1.sum())
    # fewer check-node bytes were streamed: the compacted run launches on smaller batches
    assert prof1["cn_launches"] > 0
