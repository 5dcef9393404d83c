"""Round-2 behaviours of the decode path, all against the oracle:
  * forward()'s posterior under both bookkeeping modes (refreshed every iteration / written when a frame stops) and
    with the speculative second span in flight -- identical results in every combination;
  * weight edits that bypass autograd's version counter (``param.data``) reach the device;
  * attributes the reference reads at call time (max_iterations, quantisers) select the matching engine."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
QP = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]


def _graph(g):
    from oracle.restatement import SparseGraph
    return SparseGraph.from_coo(g.n, g.m, g.edge_check, g.check_var)


def _mixed_llr(rng, B, n, snrs):
    parts = []
    per = B // len(snrs)
    for k, snr in enumerate(snrs):
        cnt = per if k < len(snrs) - 1 else B - per * (len(snrs) - 1)
        s2 = 10 ** (-snr / 10)
        parts.append(2 * (1.0 + np.sqrt(s2) * rng.standard_normal((cnt, n))) / s2)
    llr = np.concatenate(parts)
    rng.shuffle(llr, axis=0)
    return llr.astype(np.float32)


@pytest.mark.parametrize("post_mode,speculate,compact", [("1", "0", "1"), ("2", "0", "1"), ("0", "1", "1"), ("2", "1", "1"),
                                                          ("1", "1", "1"), ("2", "0", "0"), ("0", "0", "1"), ("0", "-1", "1"),
                                                          ("2", "-1", "1")])
def test_posterior_modes_and_speculation_are_result_neutral(built_lib, monkeypatch, post_mode, speculate, compact):
    from oracle import capi as O
    from oracle.restatement import MODE_WRCQ, quantizer_schedule
    L = built_lib
    monkeypatch.setenv("LDPC_POST_MODE", post_mode)
    monkeypatch.setenv("LDPC_SPECULATE", speculate)
    monkeypatch.setenv("LDPC_COMPACT", compact)
    monkeypatch.setenv("LDPC_COMPACT_MIN_FRAMES", "128")
    monkeypatch.setenv("LDPC_GRAPHS", "0")           # the span schedule, not the captured replay
    T = 30
    code = L.codes.dvbs2_shaped(max_iterations=T, scale=20)
    g = code.graph
    og = _graph(g)
    rng = np.random.default_rng(17)
    B = 1900
    llr = _mixed_llr(rng, B, g.n, (0.5, 2.0, 2.6, 3.2, 4.5, 6.0))
    torch.manual_seed(3)
    dec = L.Neural2DMinSumDecoder(code, weight_sharing_type=2, max_iterations=T)
    with torch.no_grad():
        dec._beta_table.uniform_(0.6, 0.95)
        dec._alpha_table.uniform_(0.9, 1.0)
    ref = O.decode(og, llr, T=T, beta=dec._beta_table.detach().numpy()[:, dec._beta_index],
                   alpha=dec._alpha_table.detach().numpy()[:, dec._alpha_index], nthreads=8)
    assert len(set(ref.iterations.tolist())) > 6 and (ref.iterations == T).any()
    for where in ("device", "host"):
        x = torch.from_numpy(llr)
        bits, post, iters = dec(x.cuda() if where == "device" else x)
        assert np.array_equal(iters.cpu().numpy(), ref.iterations), where
        assert np.array_equal(bits.cpu().numpy(), ref.bits), where
        assert np.array_equal(post.cpu().numpy(), ref.posterior), where
    if compact == "1":
        assert dec._engine(0).profile_read()["compactions"] >= 1
    # quantised: the posterior of a stopped frame decodes its codes with the quantiser of ITS last iteration
    w = L.WeightedRCQDecoder(code, bc=4, bv=8, quantizer_params=QP, weight_sharing_type=2, max_iterations=T)
    with torch.no_grad():
        w._beta_table.uniform_(0.8, 1.0)
    thr = np.array([q.thresholds for q in w.quantizers], dtype=np.float64).astype(np.float32)
    ref = O.decode(og, llr, T=T, mode=MODE_WRCQ, bc=4, thresholds=thr, quantizer_of_iter=quantizer_schedule(T, 3),
                   beta=w._beta_table.detach().numpy()[:, w._beta_index],
                   alpha=w._alpha_table.detach().numpy()[:, w._alpha_index], nthreads=8)
    b, p, i = w(torch.from_numpy(llr).cuda())
    assert np.array_equal(i.cpu().numpy(), ref.iterations) and np.array_equal(b.cpu().numpy(), ref.bits)
    assert np.array_equal(p.cpu().numpy(), ref.posterior)
    # float64 (two frames per lane), decisions only and with posteriors through the engine
    basic = L.BasicMinSumDecoder(code, factor=0.8)
    llr64 = llr.astype(np.float64)
    ref = O.decode(og, llr64, T=T, dtype=np.float64, beta=np.full((T, g.E), 0.8), nthreads=8)
    eb, ep, ei, es = basic._engine(0).decode_device(torch.from_numpy(llr64).cuda(), want_posterior=True)
    assert np.array_equal(eb.cpu().numpy(), ref.bits) and np.array_equal(ei.cpu().numpy(), ref.iterations)
    assert np.array_equal(ep.cpu().numpy(), ref.posterior) and np.array_equal(es.cpu().numpy(), ref.success)


def test_posterior_on_stop_with_wide_variables(built_lib, monkeypatch):
    """The on-stop pass through vn_wide_kernel (degree 9..64) and the generic path (> 64)."""
    from oracle import capi as O
    L = built_lib
    monkeypatch.setenv("LDPC_POST_MODE", "2")
    monkeypatch.setenv("LDPC_GRAPHS", "0")
    rng = np.random.default_rng(5)
    m, n = 90, 140
    H = np.zeros((m, n), dtype=np.int64)
    degs = [12] * 20 + [70] * 2 + [3] * 60 + [2] * 58
    for j, dv in enumerate(degs):
        H[rng.choice(m, size=dv, replace=False), j] = 1
    T = 12
    code = L.LDPCCode(n, n - m, H, max_iterations=T)
    g = code.graph
    og = _graph(g)
    llr = _mixed_llr(rng, 600, n, (3.0, 5.0, 7.0, 9.0))
    dec = L.Neural2DMinSumDecoder(code, weight_sharing_type=2, max_iterations=T)
    with torch.no_grad():
        dec._beta_table.uniform_(0.3, 0.6)
        dec._alpha_table.uniform_(0.9, 1.0)
    ref = O.decode(og, llr, T=T, beta=dec._beta_table.detach().numpy()[:, dec._beta_index],
                   alpha=dec._alpha_table.detach().numpy()[:, dec._alpha_index], nthreads=8)
    bits, post, iters = dec(torch.from_numpy(llr).cuda())
    assert len(set(ref.iterations.tolist())) >= 3
    assert np.array_equal(iters.cpu().numpy(), ref.iterations) and np.array_equal(bits.cpu().numpy(), ref.bits)
    assert np.array_equal(post.cpu().numpy(), ref.posterior)


def test_weight_edits_through_param_data_reach_the_device(built_lib):
    """``p.data.fill_()`` does not bump ``p._version``; the tables are compared by content."""
    from oracle import capi as O
    L = built_lib
    T = 8
    code = L.codes.dvbs2_shaped(max_iterations=T, scale=20)
    g = code.graph
    og = _graph(g)
    rng = np.random.default_rng(2)
    llr = _mixed_llr(rng, 300, g.n, (1.5, 2.5, 3.5))
    dec = L.Neural2DMinSumDecoder(code, weight_sharing_type=2, max_iterations=T)
    with torch.no_grad():
        dec._beta_table.fill_(0.9)
        dec._alpha_table.fill_(1.0)
    x = torch.from_numpy(llr).cuda()
    first = dec(x)
    v0 = dec._beta_table._version
    dec._beta_table.data.fill_(0.45)                       # no version bump
    dec._alpha_table.data.mul_(0.8)
    assert dec._beta_table._version == v0
    bits, post, iters = dec(x)
    ref = O.decode(og, llr, T=T, beta=np.full((T, g.E), np.float32(0.45)), alpha=np.full((T, g.n), np.float32(0.8)), nthreads=4)
    assert np.array_equal(bits.cpu().numpy(), ref.bits) and np.array_equal(iters.cpu().numpy(), ref.iterations)
    assert np.array_equal(post.cpu().numpy(), ref.posterior)
    assert not torch.equal(post, first[1])
    # a numpy alias of the table
    dec._beta_table.detach().numpy()[:] = 0.7
    bits, post, iters = dec(x)
    ref = O.decode(og, llr, T=T, beta=np.full((T, g.E), np.float32(0.7)), alpha=np.full((T, g.n), np.float32(0.8)), nthreads=4)
    assert np.array_equal(post.cpu().numpy(), ref.posterior)
    # edits through the reference-keyed view
    with torch.no_grad():
        for k in dec.beta_weights:
            dec.beta_weights[k].fill_(0.6)
    bits, post, iters = dec(x)
    ref = O.decode(og, llr, T=T, beta=np.full((T, g.E), np.float32(0.6)), alpha=np.full((T, g.n), np.float32(0.8)), nthreads=4)
    assert np.array_equal(post.cpu().numpy(), ref.posterior)


def test_call_time_attributes_select_the_engine(built_lib):
    """The reference reads max_iterations / bc / quantisers when decode() is called (rcq_decoder.py:169-208,
    neural_2d_decoder.py:159): changing them after a first decode must take effect."""
    from oracle import capi as O
    from oracle.restatement import MODE_RCQ, quantizer_schedule
    L = built_lib
    code = L.codes.dvbs2_shaped(max_iterations=10, scale=20)
    g = code.graph
    og = _graph(g)
    rng = np.random.default_rng(8)
    llr = _mixed_llr(rng, 256, g.n, (1.0, 2.0, 3.0))
    x = torch.from_numpy(llr).cuda()
    rcq = L.RCQMinSumDecoder(code, bc=3, bv=8, quantizer_params=QP, max_iterations=10)
    rcq.decode(x)
    rcq.max_iterations = 6
    rcq.quantizers = [L.NonUniformQuantizer(3, 4.0, 1.1), L.NonUniformQuantizer(3, 6.0, 1.2)]
    b, s, i = rcq.decode(x)
    thr = np.array([q.thresholds for q in rcq.quantizers], dtype=np.float64).astype(np.float32)
    ref = O.decode(og, llr, T=6, mode=MODE_RCQ, bc=3, thresholds=thr, quantizer_of_iter=quantizer_schedule(6, 2), nthreads=4)
    assert np.array_equal(b.cpu().numpy(), ref.bits) and np.array_equal(i.cpu().numpy(), ref.iterations)
    assert len(rcq._engines) == 1
    dec = L.Neural2DMinSumDecoder(code, weight_sharing_type=3, max_iterations=10)
    with torch.no_grad():
        dec._beta_table.uniform_(0.6, 0.9)
    dec(x)
    dec.max_iterations = 4
    bits, post, iters = dec(x)
    ref = O.decode(og, llr, T=4, beta=dec._beta_table.detach().numpy()[:4, dec._beta_index], nthreads=4)
    assert np.array_equal(bits.cpu().numpy(), ref.bits) and np.array_equal(post.cpu().numpy(), ref.posterior)
    assert int(iters.max()) <= 4
    dec.max_iterations = 11
    with pytest.raises(KeyError):
        dec(x)
    basic = L.BasicMinSumDecoder(code, factor=0.7)
    basic.decode(llr[:4].astype(np.float64))
    basic.factor = 0.9
    code.max_iterations = 5
    bb, ss, ii = basic.decode(llr.astype(np.float64))
    ref = O.decode(og, llr.astype(np.float64), T=5, dtype=np.float64, beta=np.full((5, g.E), 0.9), nthreads=4)
    assert np.array_equal(bb, ref.bits) and np.array_equal(ii, ref.iterations)


def test_decode_host_validates_caller_buffers(built_lib):
    L = built_lib
    code = L.create_test_ldpc_code()
    eng = L.BasicMinSumDecoder(code)._engine(0)
    llr = np.ones((5, 7))
    for bad in ({"bits": np.empty((5, 7), np.int32)}, {"bits": np.empty((4, 7), np.uint8)},
                {"iterations": np.empty(5, np.int64)}, {"bits": np.empty((7, 5), np.uint8).T},
                {"success": np.empty((5, 1), np.uint8)}):
        with pytest.raises(ValueError):
            eng.decode_host(llr, out=bad)
    out = {"bits": np.empty((5, 7), np.uint8), "iterations": np.empty(5, np.int32), "success": np.empty(5, np.uint8)}
    b, _, i, s = eng.decode_host(llr, out=out)
    assert b is out["bits"] and i is out["iterations"] and s is out["success"] and not b.any() and s.all()


@pytest.mark.parametrize("which", ["dvbs2_s20", "h74", "wide"])
def test_packed_decision_rows_equal_the_byte_rows(built_lib, monkeypatch, which):
    """ldpc_decode_{device,host}_packed: bit (j & 31) of word (j >> 5) of row f = decision of variable j of frame f,
    through the per-iteration path with compaction (frame maps), the on-chip small-code path and ragged n."""
    L = built_lib
    monkeypatch.setenv("LDPC_COMPACT_MIN_FRAMES", "128")
    rng = np.random.default_rng(11)
    if which == "dvbs2_s20":
        code, T, B = L.codes.dvbs2_shaped(max_iterations=30, scale=20), 30, 1500
    elif which == "h74":
        code, T, B = L.create_test_ldpc_code(), 10, 777
    else:
        H = (rng.random((40, 101)) < 0.08).astype(np.int64)
        code, T, B = L.LDPCCode(101, 61, H, max_iterations=12), 12, 300
    g = code.graph
    llr = _mixed_llr(rng, B, g.n, (1.0, 2.5, 4.0, 6.0))
    dec = L.Neural2DMinSumDecoder(code, weight_sharing_type=2, max_iterations=T)
    with torch.no_grad():
        dec._beta_table.uniform_(0.6, 0.9)
        dec._alpha_table.uniform_(0.9, 1.0)
    eng = dec._engine(0)
    x = torch.from_numpy(llr).cuda()
    bits, post, iters, succ = eng.decode_device(x, want_posterior=True)
    pk, post2, iters2, succ2 = eng.decode_device(x, want_posterior=True, packed_bits=True)
    assert pk.shape == (B, eng.row_words) and pk.dtype == torch.int32
    assert np.array_equal(eng.unpack_rows(pk.cpu().numpy().view(np.uint32)), bits.cpu().numpy())
    assert torch.equal(post, post2) and torch.equal(iters, iters2) and torch.equal(succ, succ2)
    hb, hp, hi, hs = eng.decode_host(llr, want_posterior=True, packed_bits=True)
    assert hb.dtype == np.uint32 and np.array_equal(eng.unpack_rows(hb), bits.cpu().numpy())
    assert np.array_equal(hp, post.cpu().numpy()) and np.array_equal(hi, iters.cpu().numpy())
    # bits beyond n in the last word are zero
    if g.n % 32:
        assert not (hb[:, -1] >> (g.n % 32)).any()
    with pytest.raises(ValueError):
        eng.decode_host(llr, out={"bits": np.empty((B, g.n), np.uint8)}, packed_bits=True)
