"""The oracle itself (CPU): both restatements against the golden vectors produced by the LIVE reference
(tests/golden/make_golden.py), the known-answer vectors of SURVEY.md appendix B, and the reduction-order
models against the live torch / numpy on this host."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, Golden, golden_cases
from oracle import capi as OC
from oracle import restatement as R


def oracle_run(impl, g: Golden):
    H, kind, T = g["H"], g.kind, int(g["T"])
    G = R.SparseGraph.from_dense(H)
    llr = g["llr"]
    if kind == "basic":
        beta = np.full((T, G.E), float(g["factor"]), dtype=np.float64)
        return impl.decode(G, llr, T=T, mode=R.MODE_NMS, dtype=np.float64, beta=beta)
    if kind in ("nnms", "n2d"):
        alpha = g["alpha_var"] if kind == "n2d" else None
        return impl.decode(G, llr, T=T, mode=R.MODE_NMS, beta=g["beta_edge"], alpha=alpha)
    th = g["thresholds"].astype(np.float32)
    qoi = g["quantizer_of_iter"]
    assert np.array_equal(R.quantizer_schedule(T, th.shape[0]), qoi)
    if kind == "rcq":
        return impl.decode(G, llr, T=T, mode=R.MODE_RCQ, bc=int(g["bc"]), thresholds=th, quantizer_of_iter=qoi)
    return impl.decode(G, llr, T=T, mode=R.MODE_WRCQ, bc=int(g["bc"]), thresholds=th, quantizer_of_iter=qoi,
                       beta=g["beta_edge"], alpha=g["alpha_var"])


@pytest.mark.parametrize("stem,case", golden_cases())
@pytest.mark.parametrize("impl", [R, OC], ids=["python", "c"])
def test_restatement_matches_live_reference(stem, case, impl):
    g = Golden(stem, case)
    res = oracle_run(impl, g)
    assert np.array_equal(res.bits, g["bits"])
    assert np.array_equal(res.iterations, g["iterations"])
    if "posterior" in g:
        assert np.array_equal(res.posterior, g["posterior"])   # float32, bit for bit
    if "success" in g:
        assert np.array_equal(res.success, g["success"])


def test_golden_covers_early_stop_and_failure():
    its = np.concatenate([Golden(s, c)["iterations"] for s, c in golden_cases()])
    assert (its == 1).any() and (its > 1).any() and len(set(its.tolist())) >= 6


# ---- SURVEY.md appendix B known answers, (7,4) code, T = 10 ----
H74 = np.array([[1, 1, 0, 1, 0, 0, 0], [0, 1, 1, 0, 1, 0, 0], [1, 0, 1, 0, 0, 1, 0], [1, 1, 1, 0, 0, 0, 1]])
LLR_A = np.array([1.25, -0.5, 2.75, -3.125, 0.625, -1.875, 4.5])
LLR_B = np.array([-0.75, 0.375, -2.5, 1.125, -0.25, 3.0, -1.5])
QP = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]


def kat_weights(G, wtype, T=10):
    return R.expand_2d_weights(G, wtype, T, lambda t, dc, dv: 0.75 + 0.015625 * t, lambda t, dv: 1 - 0.03125 * t)


@pytest.mark.parametrize("impl", [R, OC], ids=["python", "c"])
def test_appendix_b_known_answers(impl):
    G = R.SparseGraph.from_dense(H74)
    T = 10
    basic = lambda llr: impl.decode(G, llr, T=T, dtype=np.float64, beta=np.full((T, G.E), 0.7))
    r = basic(LLR_A)
    assert r.bits[0].tolist() == [1, 0, 0, 1, 0, 1, 0] and not r.success[0] and r.iterations[0] == 10
    r = basic(LLR_B)
    assert r.bits[0].tolist() == [1, 1, 1, 0, 0, 0, 1] and r.success[0] and r.iterations[0] == 2
    thr = np.array([R.quantizer_thresholds(3, C, gm) for C, gm in QP]).astype(np.float32)
    qoi = R.quantizer_schedule(T, 3)
    assert qoi.tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2, 2]
    r = impl.decode(G, LLR_A.astype(np.float32), T=T, mode=R.MODE_RCQ, bc=3, thresholds=thr, quantizer_of_iter=qoi)
    assert r.bits[0].tolist() == [1, 1, 0, 1, 0, 1, 0] and not r.success[0] and r.iterations[0] == 10
    r = impl.decode(G, LLR_B.astype(np.float32), T=T, mode=R.MODE_RCQ, bc=3, thresholds=thr, quantizer_of_iter=qoi)
    assert r.bits[0].tolist() == [1, 1, 1, 0, 0, 0, 1] and r.success[0] and r.iterations[0] == 2
    # N-2D type 2, LLR A
    beta, alpha = kat_weights(G, 2)
    r = impl.decode(G, LLR_A.astype(np.float32), T=T, beta=beta, alpha=alpha)
    want = np.array([-0.5322954654693604, 0.16371989250183105, 1.154575228691101, -3.1212146282196045,
                     0.21977722644805908, -0.7577279806137085, 4.427070140838623], dtype=np.float32)
    assert r.bits[0].tolist() == [1, 0, 0, 1, 0, 1, 0] and r.iterations[0] == 10
    assert np.array_equal(r.posterior[0], want)
    # N-2D type 4, LLR B
    beta, alpha = kat_weights(G, 4)
    r = impl.decode(G, LLR_B.astype(np.float32), T=T, beta=beta, alpha=alpha)
    want = np.array([-2.403749942779541, -1.2874999046325684, -2.4649999141693115, 1.1074999570846558,
                     0.22249996662139893, 3.1575000286102295, -1.4824999570846558], dtype=np.float32)
    assert r.iterations[0] == 2 and np.array_equal(r.posterior[0], want)
    # W-RCQ type 2, LLR B: 3 iterations
    beta, alpha = kat_weights(G, 2)
    r = impl.decode(G, LLR_B.astype(np.float32), T=T, mode=R.MODE_WRCQ, bc=3, thresholds=thr, quantizer_of_iter=qoi,
                    beta=beta, alpha=alpha)
    want = np.array([-2.52093505859375, -1.0634461641311646, -2.5, 1.125, 0.4692230820655823, 3.0, -1.5], dtype=np.float32)
    assert r.iterations[0] == 3 and np.array_equal(r.posterior[0], want)
    # W-RCQ type 1, LLR A: every C2V quantises to 0 -> posterior == LLR
    beta, alpha = kat_weights(G, 1)
    r = impl.decode(G, LLR_A.astype(np.float32), T=T, mode=R.MODE_WRCQ, bc=3, thresholds=thr, quantizer_of_iter=qoi,
                    beta=beta, alpha=alpha)
    assert r.iterations[0] == 10 and np.array_equal(r.posterior[0], LLR_A.astype(np.float32))
    assert r.bits[0].tolist() == [0, 1, 0, 1, 0, 1, 0]


def test_quantizer_known_answers():
    z = np.load(f"{GOLDEN_DIR}/quantizer_kat.npz")
    thr = z["thresholds"]
    assert np.allclose(thr, [0, 0.9622504486493761, 2.721655269759087, 5])
    assert np.array_equal(np.array(R.quantizer_thresholds(3, 5.0, 1.5)), thr)
    th32 = thr.astype(np.float32)
    codes = R.quantize(z["x"], th32, 3)
    assert np.array_equal(codes, z["codes"])
    assert codes[:12].tolist() == [6, 5, 0, 2, 2, 0, 0, 3, 3, 0, 3, 4]       # appendix B
    vals = R.dequantize(codes, th32, 3)
    assert np.array_equal(vals, z["values"])
    lib = OC.lib()
    import ctypes as C
    for x, c, v in zip(z["x"], z["codes"], z["values"]):
        assert lib.oracle_quantize_f32(C.c_float(float(x)), th32.ctypes.data, 4, 3) == int(c)
        assert lib.oracle_dequantize_f32(int(c), th32.ctypes.data, 3) == float(v)
    # C3 thresholds quoted in appendix A3
    t3 = np.array([R.quantizer_thresholds(3, C, g) for C, g in QP]).astype(np.float32)
    assert t3[0].tolist() == [0.0, 0.7192230820655823, 1.7709349393844604, 3.0]
    assert t3[2].tolist() == [0.0, 1.6781872510910034, 4.132181644439697, 7.0]


def test_reduction_order_models_match_this_hosts_libraries():
    """torch.sum (float32) / np.sum (float64) on contiguous vectors -- the orders the hard decisions hinge
    on; re-checked on whatever host runs the tests (the GPU box included)."""
    rng = np.random.default_rng(0)
    lib = OC.lib()
    for k in range(1, 66):
        for _ in range(40):
            x = (rng.standard_normal(k) * 10.0 ** rng.uniform(-3, 3, k)).astype(np.float32)
            want = torch.sum(torch.tensor(x)).item()
            assert float(R.torch_sum_f32([np.float32(v) for v in x])) == want
            assert lib.oracle_torch_sum_f32(x.ctypes.data, k) == want
            xd = rng.standard_normal(k) * 10.0 ** rng.uniform(-3, 3, k)
            want = float(np.sum(xd))
            assert float(R.np_sum_f64([np.float64(v) for v in xd])) == want
            assert lib.oracle_np_sum_f64(xd.ctypes.data, k) == want


def test_python_and_c_oracles_agree_on_irregular_random_code():
    rng = np.random.default_rng(5)
    m, n = 30, 64
    H = (rng.random((m, n)) < 0.12).astype(np.int64)
    H[:, 0] = 0
    H[0, :] = 0
    H[1, 3] = 1
    G = R.SparseGraph.from_dense(H)
    T = 6
    llr = (rng.standard_normal((9, n)) * 3).astype(np.float32)
    beta = rng.uniform(0.4, 1.0, (T, G.E)).astype(np.float32)
    alpha = rng.uniform(0.7, 1.1, (T, n)).astype(np.float32)
    a = R.decode(G, llr, T=T, beta=beta, alpha=alpha)
    b = OC.decode(G, llr, T=T, beta=beta, alpha=alpha, nthreads=2)
    assert np.array_equal(a.bits, b.bits) and np.array_equal(a.posterior, b.posterior)
    assert np.array_equal(a.iterations, b.iterations) and np.array_equal(a.success, b.success)
    a = R.decode(G, llr, T=T, beta=beta, alpha=alpha, early_stop=False)
    b = OC.decode(G, llr, T=T, beta=beta, alpha=alpha, early_stop=False)
    assert np.array_equal(a.bits, b.bits) and (b.iterations == T).all() and np.array_equal(a.success, b.success)


# ---- SURVEY 8f "next" rows: offset min-sum and layered RCQ (tests/golden/*_next.npz) ----
def next_cases():
    import glob
    import os
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*_next.npz"))):
        stem = os.path.splitext(os.path.basename(f))[0]
        z = np.load(f)
        out += [(stem, c) for c in sorted(set(k.split("/")[0] for k in z.files))]
    return out


def oracle_run_next(impl, g: Golden):
    G = R.SparseGraph.from_dense(g["H"])
    T = int(g["T"])
    if g.kind == "rcq_layered":
        th = g["thresholds"].astype(np.float32)
        return impl.decode_layered_rcq(G, g["llr"], T=T, bc=int(g["bc"]), thresholds=th,
                                       quantizer_of_iter=R.quantizer_schedule(T, th.shape[0]))
    alpha = g["alpha_var"] if g.kind == "n2doms" else None
    return impl.decode(G, g["llr"], T=T, mode=R.MODE_OFFSET, beta=g["beta_edge"], alpha=alpha)


@pytest.mark.parametrize("stem,case", next_cases())
@pytest.mark.parametrize("impl", [R, OC], ids=["python", "c"])
def test_next_rows_match_live_reference(stem, case, impl):
    g = Golden(stem, case)
    res = oracle_run_next(impl, g)
    assert np.array_equal(res.bits, g["bits"])
    assert np.array_equal(res.iterations, g["iterations"])
    if "posterior" in g:
        assert np.array_equal(res.posterior, g["posterior"])
    if "success" in g:
        assert np.array_equal(res.success, g["success"])


@pytest.mark.parametrize("case", ["n2d2_dvbs2", "rcq_dvbs2", "wrcq1_qc", "rcq_layered_dvbs2_s4", "rcq_layered_qc"])
def test_oracle_reproduces_fullsize_reference_vectors(case):
    """BASELINE's full code sizes: frames decoded by the LIVE reference (minutes per frame,
    tests/golden/make_golden_fullsize.py) against the C port -- bits, iterations, success, posteriors bit for bit."""
    from conftest import fullsize_tables, load_fullsize
    z, code = load_fullsize(case)
    import ldpc_b200 as L
    from oracle import capi as O
    from oracle.restatement import MODE_NMS, MODE_RCQ, MODE_WRCQ, SparseGraph, quantizer_schedule
    g = code.graph
    og = SparseGraph.from_coo(g.n, g.m, g.edge_check, g.check_var)
    T = int(z["T"])
    x = z["llr"]
    if case.startswith("n2d2"):
        dec = L.Neural2DMinSumDecoder(code, 2, T)
        beta, alpha = fullsize_tables(z, dec)
        ref = O.decode(og, x, T=T, mode=MODE_NMS, beta=beta, alpha=alpha, nthreads=4)
    else:
        thr = z["thresholds"].astype(np.float32)
        kw = dict(T=T, bc=3, thresholds=thr, quantizer_of_iter=quantizer_schedule(T, 3), nthreads=4)
        if case.startswith("rcq_layered"):
            ref = O.decode_layered_rcq(og, x, **kw)
        elif case.startswith("rcq"):
            ref = O.decode(og, x, mode=MODE_RCQ, **kw)
        else:
            dec = L.WeightedRCQDecoder(code, 3, 8, [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)], weight_sharing_type=1, max_iterations=T)
            beta, alpha = fullsize_tables(z, dec)
            ref = O.decode(og, x, mode=MODE_WRCQ, beta=beta,
                           alpha=alpha if alpha is not None else np.ones((T, g.n), np.float32), **kw)
    assert np.array_equal(ref.bits, z["bits"]) and np.array_equal(ref.iterations, z["iterations"])
    if "success" in z.files:
        assert np.array_equal(ref.success, z["success"])
    if "posterior" in z.files:
        assert np.array_equal(ref.posterior, z["posterior"])
