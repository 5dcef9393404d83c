"""Parity of the CUDA path against (a) the committed golden vectors produced by the live reference and
(b) the oracle on seeded inputs.  Bars (BASELINE.json north_star): hard decisions, success flags and
iteration counts bit-exact; RCQ integer path bit-exact; float posteriors within 1e-5 relative."""
import numpy as np
import pytest
import torch

from conftest import Golden, golden_cases

pytestmark = pytest.mark.gpu

POST_RTOL = 1e-5   # north_star tolerance for float posteriors
POST_ATOL = 1e-6


def make_decoder(L, g: Golden):
    H = g["H"].astype(np.int64)
    T = int(g["T"])
    code = L.LDPCCode(n=int(g["n"]), k=int(g["k"]), H=H, max_iterations=T)
    kind = g.kind
    if kind == "basic":
        return L.BasicMinSumDecoder(code, factor=float(g["factor"]))
    if kind == "nnms":
        dec = L.NeuralMinSumDecoder(code, max_iterations=T)
    elif kind == "n2d":
        dec = L.Neural2DMinSumDecoder(code, weight_sharing_type=int(g["wtype"]), max_iterations=T)
    elif kind == "rcq":
        return L.RCQMinSumDecoder(code, bc=int(g["bc"]), bv=8, quantizer_params=[tuple(p) for p in g["qp"]],
                                  max_iterations=T)
    elif kind == "wrcq":
        dec = L.WeightedRCQDecoder(code, bc=int(g["bc"]), bv=8, quantizer_params=[tuple(p) for p in g["qp"]],
                                   weight_sharing_type=int(g["wtype"]), max_iterations=T)
    else:
        raise ValueError(kind)
    state = {str(k): torch.tensor([float(v)]) for k, v in zip(g["weight_keys"], g["weight_vals"])}
    dec.load_reference_state_dict(state)
    return dec


def run_batch(dec, kind, llr, device):
    """Returns bits, posterior|None, iterations, success|None as numpy."""
    if kind == "basic":
        x = torch.from_numpy(llr).cuda() if device == "cuda" else llr
        b, s, i = dec.decode(x)
        if device == "cuda":
            b, s, i = b.cpu().numpy(), s.cpu().numpy(), i.cpu().numpy()
        return b, None, i, s
    x = torch.from_numpy(llr)
    if device == "cuda":
        x = x.cuda()
    if kind == "rcq":
        b, s, i = dec.decode(x)
        return b.cpu().numpy(), None, i.cpu().numpy(), s.cpu().numpy()
    b, p, i = dec(x)
    return b.cpu().numpy(), p.cpu().numpy(), i.cpu().numpy(), None


@pytest.mark.parametrize("stem,case", golden_cases())
@pytest.mark.parametrize("device", ["cuda", "host"])
def test_golden_batched(built_lib, stem, case, device):
    """Every frame of a golden case in ONE batched call, through the reference-shaped classes."""
    g = Golden(stem, case)
    dec = make_decoder(built_lib, g)
    bits, post, iters, succ = run_batch(dec, g.kind, g["llr"], device)
    assert np.array_equal(bits.astype(np.uint8), g["bits"])
    assert np.array_equal(iters, g["iterations"])
    if "success" in g:
        assert np.array_equal(succ.astype(bool), g["success"])
    if "posterior" in g:
        np.testing.assert_allclose(post, g["posterior"], rtol=POST_RTOL, atol=POST_ATOL)
        assert np.array_equal(post, g["posterior"]), "posterior is expected to be bit-identical (same add order)"


@pytest.mark.parametrize("stem,case", [c for c in golden_cases() if c[0] == "hamming74"])
def test_golden_single_frame_api(built_lib, stem, case):
    """One frame per call: return types and values of the reference's decode()/forward()."""
    g = Golden(stem, case)
    dec = make_decoder(built_lib, g)
    llr = g["llr"]
    for f in range(min(4, llr.shape[0])):
        if g.kind == "basic":
            b, s, i = dec.decode(llr[f])
            assert isinstance(b, np.ndarray) and b.dtype == np.int64 and b.shape == (7,)
            assert isinstance(s, bool) and isinstance(i, int)
            assert s == bool(g["success"][f])
        elif g.kind == "rcq":
            b, s, i = dec.decode(torch.from_numpy(llr[f]))
            assert b.dtype == torch.int32 and isinstance(s, bool) and isinstance(i, int)
            assert s == bool(g["success"][f])
            b = b.numpy()
        else:
            b, p, i = dec(torch.from_numpy(llr[f]))
            assert b.dtype == torch.int32 and p.dtype == torch.float32 and isinstance(i, int)
            assert np.array_equal(p.numpy(), g["posterior"][f])
            b = b.numpy()
        assert np.array_equal(np.asarray(b).astype(np.uint8), g["bits"][f])
        assert i == int(g["iterations"][f])


def _awgn(rng, B, n, snr_db, sign=1.0):
    s2 = 10 ** (-snr_db / 10)
    return (2 * (sign + np.sqrt(s2) * rng.standard_normal((B, n))) / s2)


def _oracle_graph(g):
    from oracle.restatement import SparseGraph
    return SparseGraph.from_coo(g.n, g.m, g.edge_check, g.check_var)


CODES = {
    "dvbs2_s20": lambda L: L.codes.dvbs2_shaped(max_iterations=10, scale=20),
    "qc_z16": lambda L: L.codes.qc_shaped(max_iterations=10, Z=16),
    "reg_3_6": lambda L: L.codes.regular_code(384, 3, 6, max_iterations=10, seed=5),
}
# operating points (dB) around each code's waterfall so that frames stop at many different iterations
SNRS = {"dvbs2_s20": (1.5, 3.0), "qc_z16": (4.5, 6.5), "reg_3_6": (2.0, 3.5)}


@pytest.mark.parametrize("cname", list(CODES))
@pytest.mark.parametrize("wtype", [1, 2, 3, 4])
def test_n2d_vs_oracle(built_lib, cname, wtype):
    from oracle import capi as O
    L = built_lib
    code = CODES[cname](L)
    g = code.graph
    rng = np.random.default_rng(1000 * wtype + list(CODES).index(cname))
    B, T = 300, 10   # not a multiple of 128: exercises the pad frames
    lo, hi = SNRS[cname]
    llr = np.concatenate([_awgn(rng, B // 3, g.n, lo), _awgn(rng, B // 3, g.n, hi),
                          _awgn(rng, B - 2 * (B // 3), g.n, hi, -1.0)]).astype(np.float32)
    torch.manual_seed(wtype)
    dec = L.Neural2DMinSumDecoder(code, weight_sharing_type=wtype, max_iterations=T)
    with torch.no_grad():
        if dec._beta_table is not None:
            dec._beta_table.uniform_(0.5, 1.0)
        if dec._alpha_table is not None:
            dec._alpha_table.uniform_(0.8, 1.1)
    bits, post, iters = dec(torch.from_numpy(llr).cuda())
    beta = (dec._beta_table.detach().numpy()[:, dec._beta_index] if dec._beta_table is not None
            else np.full((T, g.E), np.float32(0.7)))
    alpha = dec._alpha_table.detach().numpy()[:, dec._alpha_index] if dec._alpha_table is not None else None
    ref = O.decode(_oracle_graph(g), llr, T=T, beta=beta, alpha=alpha, nthreads=8)
    assert np.array_equal(bits.cpu().numpy(), ref.bits)
    assert np.array_equal(iters.cpu().numpy(), ref.iterations)
    np.testing.assert_allclose(post.cpu().numpy(), ref.posterior, rtol=POST_RTOL, atol=POST_ATOL)
    assert len(set(ref.iterations.tolist())) > 2, "test should cover several stopping iterations"
    assert (ref.iterations == T).any() and ref.success.any()


@pytest.mark.parametrize("cname", list(CODES))
def test_nnms_and_basic_vs_oracle(built_lib, cname):
    from oracle import capi as O
    L = built_lib
    code = CODES[cname](L)
    g = code.graph
    og = _oracle_graph(g)
    rng = np.random.default_rng(11)
    B, T = 130, 10
    llr = np.concatenate([_awgn(rng, B // 2, g.n, SNRS[cname][0]), _awgn(rng, B - B // 2, g.n, SNRS[cname][1])])
    torch.manual_seed(3)
    dec = L.NeuralMinSumDecoder(code, max_iterations=T)
    bits, post, iters = dec(torch.from_numpy(llr.astype(np.float32)).cuda())
    ref = O.decode(og, llr.astype(np.float32), T=T, beta=dec._beta_table.detach().numpy(), nthreads=8)
    assert np.array_equal(bits.cpu().numpy(), ref.bits) and np.array_equal(iters.cpu().numpy(), ref.iterations)
    np.testing.assert_allclose(post.cpu().numpy(), ref.posterior, rtol=POST_RTOL, atol=POST_ATOL)
    basic = L.BasicMinSumDecoder(code, factor=0.7)
    b, s, i = basic.decode(llr)
    ref = O.decode(og, llr, T=T, dtype=np.float64, beta=np.full((T, g.E), 0.7), nthreads=8)
    assert np.array_equal(b, ref.bits) and np.array_equal(i, ref.iterations) and np.array_equal(s, ref.success)


@pytest.mark.parametrize("cname", list(CODES))
@pytest.mark.parametrize("bc,qp", [(3, [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]), (4, [(6.0, 1.0)]), (8, [(12.0, 1.2), (20.0, 0.9)])])
def test_rcq_vs_oracle_bit_exact(built_lib, cname, bc, qp):
    from oracle import capi as O
    from oracle.restatement import MODE_RCQ, MODE_WRCQ, quantizer_schedule
    L = built_lib
    code = CODES[cname](L)
    g = code.graph
    og = _oracle_graph(g)
    rng = np.random.default_rng(bc)
    B, T = 200, 10
    llr = np.concatenate([_awgn(rng, B // 2, g.n, SNRS[cname][0] + 1.0),
                          _awgn(rng, B - B // 2, g.n, SNRS[cname][1] + 1.0)]).astype(np.float32)
    rcq = L.RCQMinSumDecoder(code, bc=bc, bv=8, quantizer_params=qp, max_iterations=T)
    thr = np.array([q.thresholds for q in rcq.quantizers], dtype=np.float64).astype(np.float32)
    qoi = quantizer_schedule(T, len(qp))
    b, s, i = rcq.decode(torch.from_numpy(llr).cuda())
    ref = O.decode(og, llr, T=T, mode=MODE_RCQ, bc=bc, thresholds=thr, quantizer_of_iter=qoi, nthreads=8)
    assert np.array_equal(b.cpu().numpy(), ref.bits)
    assert np.array_equal(i.cpu().numpy(), ref.iterations)
    assert np.array_equal(s.cpu().numpy(), ref.success)
    torch.manual_seed(bc)
    w = L.WeightedRCQDecoder(code, bc=bc, bv=8, quantizer_params=qp, weight_sharing_type=1, max_iterations=T)
    with torch.no_grad():
        w._beta_table.uniform_(0.6, 1.0)
    b, p, i = w(torch.from_numpy(llr).cuda())
    beta = w._beta_table.detach().numpy()[:, w._beta_index]
    ref = O.decode(og, llr, T=T, mode=MODE_WRCQ, bc=bc, thresholds=thr, quantizer_of_iter=qoi, beta=beta,
                   alpha=np.ones((T, g.n), np.float32), nthreads=8)
    assert np.array_equal(b.cpu().numpy(), ref.bits) and np.array_equal(i.cpu().numpy(), ref.iterations)
    # the posterior is a sum of reconstruction levels: bit-exact, not just close
    assert np.array_equal(p.cpu().numpy(), ref.posterior)


def test_batch_equals_single_frames_and_host_equals_device(built_lib):
    L = built_lib
    code = CODES["dvbs2_s20"](L)
    rng = np.random.default_rng(2)
    llr = _awgn(rng, 40, code.n, 2.0).astype(np.float32)
    torch.manual_seed(0)
    dec = L.Neural2DMinSumDecoder(code, 2, 10)
    with torch.no_grad():
        dec._beta_table.uniform_(0.6, 0.9)
        dec._alpha_table.uniform_(0.9, 1.0)
    B, P, I = dec(torch.from_numpy(llr).cuda())
    Bh, Ph, Ih = dec(torch.from_numpy(llr))
    assert torch.equal(B.cpu(), Bh) and torch.equal(P.cpu(), Ph) and torch.equal(I.cpu(), Ih)
    for f in range(0, 40, 7):
        b, p, i = dec(torch.from_numpy(llr[f]).cuda())
        assert torch.equal(b, B[f]) and torch.equal(p, P[f]) and i == int(I[f])


@pytest.mark.parametrize("case", ["n2d2_dvbs2", "rcq_dvbs2", "wrcq1_qc", "rcq_layered_dvbs2_s4", "rcq_layered_qc"])
def test_fullsize_reference_vectors(built_lib, case):
    """The CUDA path against frames decoded by the LIVE reference at BASELINE's full code sizes."""
    from conftest import fullsize_tables, load_fullsize
    z, code = load_fullsize(case)
    L = built_lib
    T = int(z["T"])
    x = torch.from_numpy(z["llr"]).cuda()
    if case.startswith("n2d2"):
        dec = L.Neural2DMinSumDecoder(code, 2, T)
        fullsize_tables(z, dec)
        b, p, i = dec(x)
        assert np.array_equal(p.cpu().numpy(), z["posterior"])
    elif case.startswith("rcq"):
        dec = L.RCQMinSumDecoder(code, 3, 8, [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)], max_iterations=T,
                                 layered=case.startswith("rcq_layered"))   # layered: the software-pipelined walk on the chain
                                                                           # code, the level-parallel kernel on the QC code
        b, s, i = dec.decode(x)
        assert np.array_equal(s.cpu().numpy().astype(bool), z["success"])
    else:
        dec = L.WeightedRCQDecoder(code, 3, 8, [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)], weight_sharing_type=1, max_iterations=T)
        fullsize_tables(z, dec)
        b, p, i = dec(x)
        assert np.array_equal(p.cpu().numpy(), z["posterior"])
    assert np.array_equal(b.cpu().numpy().astype(np.uint8), z["bits"])
    assert np.array_equal(i.cpu().numpy(), z["iterations"])
    # single-frame call, as the reference is driven
    out = dec.decode(x[0]) if case.startswith("rcq") else dec(x[0])
    assert np.array_equal(out[0].cpu().numpy().astype(np.uint8), z["bits"][0]) and out[2] == int(z["iterations"][0])
