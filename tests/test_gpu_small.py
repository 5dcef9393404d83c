"""On-chip decode for small codes (csrc/ldpc_small.cu: one launch for all T iterations, messages in shared memory)
against the oracle and against the per-iteration path (LDPC_SMALL=0): every decoder family, ragged batches, degree-0
variables, degree-1 and empty checks, variable degrees beyond the unrolled sums, T = 1, early stop off, posteriors,
the host pipeline and the Monte-Carlo round."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tiny_graph(rng, f64):
    """Random H whose per-frame state fits the on-chip budget (896 bytes: 2E + n values + decision words)."""
    budget = 108 if f64 else 218     # (2E + n) values + the decision words within 896 bytes
    while True:
        m, n = int(rng.integers(2, 9)), int(rng.integers(4, 22))
        H = np.zeros((m, n), dtype=np.int64)
        for i in range(m):
            H[i, rng.choice(n, int(rng.integers(1, min(n, 7) + 1)), replace=False)] = 1
        if rng.random() < 0.4:
            H[rng.choice(m, min(m, int(rng.integers(2, 9))), replace=False), rng.integers(0, n)] = 1   # a heavier variable
        if m >= 3 and rng.random() < 0.3:
            H[rng.integers(0, m), :] = 0                                          # empty check
        if rng.random() < 0.3:
            H[:, rng.integers(0, n)] = 0                                          # degree-0 variable
        E = int(H.sum())
        if 0 < E and 2 * E + n <= budget and E * (8 if f64 else 4) >= n:   # (decisions are staged in the v2c columns)
            return H


def _decode_both(L, kind, code, og, llr, T, rng, seed):
    """Our decoder and the oracle on the same inputs -> ((bits, post|None, iters, success), ref, engine)."""
    from oracle import capi as O
    from oracle.restatement import MODE_NMS, MODE_OFFSET, MODE_RCQ, MODE_WRCQ, quantizer_schedule
    n = code.n
    llr32 = llr.astype(np.float32)
    x = torch.from_numpy(llr32).cuda()
    torch.manual_seed(seed)
    if kind == "basic":
        f = float(rng.choice([0.3, 0.7, 1.0]))
        dec = L.BasicMinSumDecoder(code, f)
        ref = O.decode(og, llr, T=T, dtype=np.float64, beta=np.full((T, og.E), f), nthreads=4)
        eng = dec._engine(0)
        got = eng.decode_device(torch.from_numpy(llr).cuda(), want_posterior=True)
        return got, ref, eng
    if kind in ("n2d", "nnms"):
        dec = L.Neural2DMinSumDecoder(code, int(rng.integers(1, 5)), T) if kind == "n2d" else L.NeuralMinSumDecoder(code, T)
        with torch.no_grad():
            if dec._beta_table is not None:
                dec._beta_table.uniform_(-0.2, 0.9)
            if dec._alpha_table is not None:
                dec._alpha_table.uniform_(0.3, 1.0)
        bt, at = dec._tables()
        beta = bt[:, dec._beta_index] if dec._beta_table is not None else np.full((T, og.E), np.float32(0.7))
        alpha = at[:, dec._alpha_index] if at is not None else None
        ref = O.decode(og, llr32, T=T, mode=MODE_NMS, beta=beta, alpha=alpha, nthreads=4)
    elif kind == "oms":
        dec = (L.Neural2DOffsetMinSumDecoder(code, int(rng.integers(1, 5)), T) if rng.random() < 0.5
               else L.NeuralOffsetMinSumDecoder(code, T))
        with torch.no_grad():
            if dec._beta_table is not None:
                dec._beta_table.uniform_(-0.1, 0.6)
            if dec._alpha_table is not None:
                dec._alpha_table.uniform_(-0.05, 0.2)
        bt, at = dec._tables()
        ref = O.decode(og, llr32, T=T, mode=MODE_OFFSET, beta=bt[:, dec._beta_index] if bt is not None else None,
                       alpha=at[:, dec._alpha_index] if at is not None else None, nthreads=4)
    else:
        bc = int(rng.choice([2, 3, 4, 6, 8]))
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
        kw = dict(T=T, bc=bc, thresholds=thr, quantizer_of_iter=quantizer_schedule(T, len(qp)), nthreads=4)
        if kind == "rcq":
            ref = O.decode(og, llr32, mode=MODE_RCQ, **kw)
        else:
            bt, at = dec._tables()
            beta = bt[:, dec._beta_index] if dec._beta_table is not None else np.full((T, og.E), np.float32(0.7))
            alpha = at[:, dec._alpha_index] if at is not None else np.ones((T, n), np.float32)
            ref = O.decode(og, llr32, mode=MODE_WRCQ, beta=beta, alpha=alpha, **kw)
    eng = dec._engine(0)
    return eng.decode_device(x, want_posterior=True), ref, eng


@pytest.mark.parametrize("seed", range(30))
def test_on_chip_decode_vs_oracle(built_lib, seed):
    from oracle.restatement import SparseGraph
    L = built_lib
    rng = np.random.default_rng(4100 + seed)
    kind = ["n2d", "nnms", "basic", "rcq", "wrcq", "oms"][seed % 6]
    H = _tiny_graph(rng, kind == "basic")
    m, n = H.shape
    T = int(rng.choice([1, 2, 5, 10, 13]))
    B = int(rng.choice([1, 37, 128, 129, 1000]))
    code = L.LDPCCode(n, max(1, n - m), H, max_iterations=T)
    og = SparseGraph.from_dense(H)
    llr = float(rng.choice([0.5, 2.0, 6.0])) * (1.0 + 1.2 * rng.standard_normal((B, n)))
    llr[rng.random((B, n)) < 0.03] = 0.0
    if kind != "basic":
        llr = llr.astype(np.float32).astype(np.float64)
    (bits, post, iters, succ), ref, eng = _decode_both(L, kind, code, og, llr, T, rng, seed)
    assert eng.profile_read()["small_decodes"] == 1
    assert np.array_equal(bits.cpu().numpy(), ref.bits)
    assert np.array_equal(iters.cpu().numpy(), ref.iterations)
    assert np.array_equal(succ.cpu().numpy().astype(bool), ref.success)
    assert np.array_equal(post.cpu().numpy(), ref.posterior)


@pytest.mark.parametrize("kind", ["n2d", "rcq", "basic", "oms"])
def test_on_chip_equals_per_iteration_path_and_host_pipeline(built_lib, monkeypatch, kind):
    from oracle.restatement import SparseGraph
    L = built_lib
    rng = np.random.default_rng({"n2d": 1, "rcq": 2, "basic": 3, "oms": 4}[kind])
    code = L.create_test_ldpc_code()
    T = 10
    og = SparseGraph.from_dense(np.asarray(code.H))
    B = 5000
    s2 = 10 ** (-2.0 / 10)
    llr = (2 * (1 + np.sqrt(s2) * rng.standard_normal((B, 7))) / s2).astype(np.float32).astype(np.float64)
    state = rng.bit_generator.state
    (b1, p1, i1, s1), ref, eng = _decode_both(L, kind, code, og, llr, T, rng, 7)
    assert eng.profile_read()["small_decodes"] == 1 and len(set(ref.iterations.tolist())) >= 3
    # the same frames through the host-buffer pipeline (chunks of the on-chip decode)
    hb, hp, hi, hs = eng.decode_host(llr.astype(eng.dtype), want_posterior=True)
    assert np.array_equal(hb, b1.cpu().numpy()) and np.array_equal(hi, i1.cpu().numpy()) and np.array_equal(hp, p1.cpu().numpy())
    monkeypatch.setenv("LDPC_SMALL", "0")
    rng.bit_generator.state = state
    (b0, p0, i0, s0), _, eng0 = _decode_both(L, kind, code, og, llr, T, rng, 7)
    assert eng0.profile_read()["small_decodes"] == 0
    for a, b in ((b0, b1), (p0, p1), (i0, i1), (s0, s1)):
        assert torch.equal(a, b)
    assert np.array_equal(b1.cpu().numpy(), ref.bits) and np.array_equal(p1.cpu().numpy(), ref.posterior)


def test_on_chip_early_stop_off_and_monte_carlo_round(built_lib, monkeypatch):
    L = built_lib
    code = L.create_test_ldpc_code()
    dec = L.Neural2DMinSumDecoder(code, 2, 10)
    with torch.no_grad():
        dec._beta_table.fill_(0.8)
        dec._alpha_table.fill_(0.95)
    res = {}
    for small in ("1", "0"):
        monkeypatch.setenv("LDPC_SMALL", small)
        d = L.Neural2DMinSumDecoder(code, 2, 10)
        d.load_state_dict(dec.state_dict())
        eng = d._engine(0)
        c = torch.zeros(4, dtype=torch.int64, device="cuda")
        fbe = torch.zeros(3000, dtype=torch.int32, device="cuda")
        fit = torch.zeros(3000, dtype=torch.int32, device="cuda")
        eng.mc_round(2.0, 3000, seed=5, frame0=123, llr_sign=1, counters=c, frame_bit_errors=fbe, frame_iterations=fit)
        res[small] = (c.tolist(), fbe.cpu(), fit.cpu(), eng.profile_read()["small_decodes"])
        # early stop off: every frame runs T iterations, success only from the last syndrome
        from ldpc_b200.engine import Engine
        b, a = d._tables()
        e2 = Engine(code.graph, max_iterations=10, early_stop=False, beta=b, beta_index=d._beta_index, alpha=a,
                    alpha_index=d._alpha_index, device=0)
        llr = L.awgn_llr(7, 500, 3.0, seed=1, llr_sign=1)
        res[small] += tuple(t.cpu() for t in e2.decode_device(llr, want_posterior=True))
    assert res["1"][3] == 1 and res["0"][3] == 0
    assert res["1"][0] == res["0"][0] and res["1"][0][3] == 3000 and 0 < res["1"][0][0] < 3000
    for k in (1, 2, 4, 5, 6, 7):
        assert torch.equal(res["1"][k], res["0"][k]), k
    assert int(res["1"][6].min()) == 10 and 0 < int(res["1"][7].sum()) < 500
