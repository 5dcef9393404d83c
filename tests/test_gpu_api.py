"""C-ABI behaviour on the device: graph re-layout, generic-degree paths, error codes, big-batch
properties the domain offers (codewords stay codewords, linearity of the syndrome)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_graph_relayout_round_trip(built_lib):
    L = built_lib
    from ldpc_b200 import _lib
    code = L.codes.dvbs2_shaped(scale=20)
    g = code.graph
    slot = g.slot_of_edge(0)
    assert sorted(slot.tolist()) == list(range(g.E)), "slots are a permutation of the edges"
    dc_of_edge = g.check_degree[g.edge_check]
    order = np.argsort(slot)
    assert (np.diff(dc_of_edge[order]) >= 0).all(), "slots are sorted by check degree"
    # inside a check the slots are consecutive and keep ascending variable order
    for i in (0, 7, g.m - 1):
        s = slot[g.check_ptr[i]:g.check_ptr[i + 1]]
        assert (np.diff(s) == 1).all()
    assert g.query(0, _lib.GRAPH_E) == g.E and g.query(0, _lib.GRAPH_CHECK_CLASSES) == 4
    assert g.query(0, _lib.GRAPH_VAR_CLASSES) == 4 and g.query(0, _lib.GRAPH_MAX_DC) == 7
    assert g.query(0, _lib.GRAPH_MAX_DV) == 8


def test_error_codes(built_lib):
    from ldpc_b200 import _lib
    lib = _lib.load()
    out = C.c_void_p()
    ptr = np.array([0, 2, 3], dtype=np.int64)
    bad_order = np.array([1, 0, 2], dtype=np.int32)
    assert lib.ldpc_graph_create(0, 3, 2, ptr.ctypes.data, bad_order.ctypes.data, C.byref(out)) == _lib.LDPC_ERR_INVALID
    assert b"ascending" in lib.ldpc_last_error()
    oob = np.array([0, 5, 2], dtype=np.int32)
    assert lib.ldpc_graph_create(0, 3, 2, ptr.ctypes.data, oob.ctypes.data, C.byref(out)) == _lib.LDPC_ERR_INVALID
    good = np.array([0, 1, 2], dtype=np.int32)
    assert lib.ldpc_graph_create(0, 3, 2, ptr.ctypes.data, good.ctypes.data, C.byref(out)) == 0
    cfg = _lib.DecoderConfig()
    cfg.struct_size = C.sizeof(_lib.DecoderConfig)
    cfg.max_iterations = 0
    dec = C.c_void_p()
    assert lib.ldpc_decoder_create(out, C.byref(cfg), C.byref(dec)) == _lib.LDPC_ERR_INVALID
    cfg.max_iterations = 3
    cfg.bc = 3
    cfg.dtype = _lib.LDPC_F64
    assert lib.ldpc_decoder_create(out, C.byref(cfg), C.byref(dec)) == _lib.LDPC_ERR_UNSUPPORTED
    cfg.bc = 0
    assert lib.ldpc_decoder_create(out, C.byref(cfg), C.byref(dec)) == 0
    assert lib.ldpc_decode_host(dec, None, 1, None, None, None, None) == _lib.LDPC_ERR_INVALID
    assert lib.ldpc_decoder_destroy(dec) == 0 and lib.ldpc_graph_destroy(out) == 0
    with pytest.raises(IndexError):
        import ldpc_b200 as L
        L.BasicMinSumDecoder(L.create_test_ldpc_code()).decode(np.zeros(9))


def test_wide_checks_and_wide_variables_generic_paths(built_lib):
    """dc = 70 (> 64: sign re-read path), dc = 40 (mask path), dv = 20 and dv = 0 (generic sums)."""
    from oracle import capi as O
    from oracle.restatement import MODE_RCQ, SparseGraph, quantizer_schedule
    L = built_lib
    rng = np.random.default_rng(8)
    m, n = 24, 96
    H = np.zeros((m, n), dtype=np.int64)
    H[0, rng.choice(n, 70, replace=False)] = 1
    H[1, rng.choice(n, 40, replace=False)] = 1
    for i in range(2, m):
        H[i, rng.choice(n, rng.integers(2, 12), replace=False)] = 1
    H[:, 3] = 0                      # dv = 0
    H[rng.choice(m, 20, replace=False), 4] = 1   # dv >= 20
    code = L.LDPCCode(n, n - m, H, max_iterations=7)
    og = SparseGraph.from_dense(H)
    llr = (rng.standard_normal((70, n)) * 2.5 + 1.0)
    T = 7
    torch.manual_seed(2)
    dec = L.Neural2DMinSumDecoder(code, 1, T)
    with torch.no_grad():
        dec._beta_table.uniform_(0.5, 1.0)
    b, p, i = dec(torch.from_numpy(llr.astype(np.float32)).cuda())
    ref = O.decode(og, llr.astype(np.float32), T=T, beta=dec._beta_table.detach().numpy()[:, dec._beta_index])
    assert np.array_equal(b.cpu().numpy(), ref.bits) and np.array_equal(i.cpu().numpy(), ref.iterations)
    assert np.array_equal(p.cpu().numpy(), ref.posterior)
    bb, ss, ii = L.BasicMinSumDecoder(code, 0.8).decode(llr)
    ref = O.decode(og, llr, T=T, dtype=np.float64, beta=np.full((T, og.E), 0.8))
    assert np.array_equal(bb, ref.bits) and np.array_equal(ii, ref.iterations) and np.array_equal(ss, ref.success)
    rcq = L.RCQMinSumDecoder(code, 4, 8, [(4.0, 1.2), (8.0, 1.0)], max_iterations=T)
    b, s, i = rcq.decode(torch.from_numpy(llr.astype(np.float32)).cuda())
    thr = np.array([q.thresholds for q in rcq.quantizers]).astype(np.float32)
    ref = O.decode(og, llr.astype(np.float32), T=T, mode=MODE_RCQ, bc=4, thresholds=thr,
                   quantizer_of_iter=quantizer_schedule(T, 2))
    assert np.array_equal(b.cpu().numpy(), ref.bits) and np.array_equal(i.cpu().numpy(), ref.iterations)


def test_large_batch_properties_full_size_code(built_lib):
    """Size-independent properties on the full (16200,7200)-shaped code at a batch the oracle could not
    finish: (1) confident LLRs of the all-zero codeword decode in one iteration with zero syndrome,
    (2) success <=> zero syndrome of the returned bits, (3) iterations == T exactly when unsuccessful,
    (4) a random sample of frames agrees with the oracle bit for bit."""
    from oracle import capi as O
    from oracle.restatement import SparseGraph
    L = built_lib
    code = L.codes.dvbs2_shaped(max_iterations=10)
    g = code.graph
    B = 4096
    torch.manual_seed(0)
    dec = L.Neural2DMinSumDecoder(code, 2, 10)
    with torch.no_grad():
        dec._beta_table.fill_(0.8)
        dec._alpha_table.fill_(0.95)
    llr = L.awgn_llr(g.n, B, 2.6, seed=11, llr_sign=1)   # ~2/3 of the frames converge within 10 iterations
    bits, post, iters = dec(llr)
    eng = dec._engine(0)
    _, _, it2, succ = eng.decode_device(llr)
    assert torch.equal(iters, it2)
    syn = g.syndrome(bits.cpu().numpy())
    ok = (syn.sum(axis=1) == 0)
    assert np.array_equal(ok, succ.cpu().numpy().astype(bool))
    assert ((iters.cpu().numpy() == 10) | ok).all() and (iters.cpu().numpy()[~ok] == 10).all()
    assert 0.02 < ok.mean() < 0.999, "operating point should mix converged and failed frames"
    assert torch.equal(bits.bool(), post < 0)
    pick = np.random.default_rng(0).choice(B, 24, replace=False)
    og = SparseGraph.from_coo(g.n, g.m, g.edge_check, g.check_var)
    ref = O.decode(og, llr[pick].cpu().numpy(), T=10, beta=np.full((10, g.E), np.float32(0.8)),
                   alpha=np.full((10, g.n), np.float32(0.95)), nthreads=8)
    assert np.array_equal(bits[pick].cpu().numpy(), ref.bits)
    assert np.array_equal(iters[pick].cpu().numpy(), ref.iterations)
    assert np.array_equal(post[pick].cpu().numpy(), ref.posterior)
    clean = torch.full((256, g.n), 9.0, device="cuda")
    b, _, i = dec(clean)
    assert int(b.sum()) == 0 and (i == 1).all()


def test_concurrent_decoders_from_threads(built_lib):
    """One decoder object per host thread, as the reference's thread pool drives them
    (simulation_framework.py:192-198)."""
    import threading
    L = built_lib
    code = L.codes.dvbs2_shaped(max_iterations=6, scale=20)
    rng = np.random.default_rng(1)
    llr = (rng.standard_normal((64, code.n)) * 2 + 2).astype(np.float32)
    decs = [L.RCQMinSumDecoder(code, 3, 8, [(3.0 + k, 1.3)], max_iterations=6) for k in range(4)]
    want = [d.decode(torch.from_numpy(llr))[0] for d in decs]
    got = [None] * 4

    def work(k):
        for _ in range(5):
            got[k] = decs[k].decode(torch.from_numpy(llr))[0]
    ts = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert all(torch.equal(a, b) for a, b in zip(want, got))


def test_integration_md_binding_stub_runs(built_lib):
    """The ctypes stub shown in INTEGRATION.md (binding the C ABI underneath a reference-style module) is
    executed verbatim against the built library and must agree with the shipped class."""
    import os
    import re
    import types
    from conftest import ROOT
    L = built_lib
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code_block = re.search(r"```python\n(# ldpc_b200_binding\.py.*?)```", md, re.S).group(1)
    code_block = code_block.replace('C.CDLL("libldpc_b200.so")', f'C.CDLL({L._lib_mod.LIB_PATH!r})')
    ns = {}
    exec(compile(code_block, "INTEGRATION.md", "exec"), ns)

    code = L.create_test_ldpc_code()
    T = 6
    torch.manual_seed(4)
    ours = L.Neural2DMinSumDecoder(code, weight_sharing_type=2, max_iterations=T)
    with torch.no_grad():
        ours._beta_table.uniform_(0.5, 1.0)
        ours._alpha_table.uniform_(0.8, 1.0)
    beta_full = ours._beta_table.detach().numpy()[:, ours._beta_index]     # [T, E] in check-major edge order
    alpha_full = ours._alpha_table.detach().numpy()[:, ours._alpha_index]  # [T, n]
    rows, cols = np.nonzero(np.asarray(code.H) == 1)
    edge_of = {(int(i), int(j)): e for e, (i, j) in enumerate(zip(rows, cols))}
    # a stand-in with the reference module's surface the stub touches (neural_2d_decoder.py:84-131)
    fake = types.SimpleNamespace(
        code=types.SimpleNamespace(H=np.asarray(code.H)), max_iterations=T,
        _get_beta_weight=lambda t, i, j: torch.tensor(beta_full[t, edge_of[(i, j)]]),
        _get_alpha_weight=lambda t, i, j: torch.tensor(alpha_full[t, j]))
    ns["attach"](fake, device=0)
    rng = np.random.default_rng(0)
    for _ in range(5):
        llr = torch.from_numpy((rng.standard_normal(7) * 2 + 1).astype(np.float32))
        b0, p0, i0 = ours(llr)
        b1, p1, i1 = ns["forward"](fake, llr)
        assert torch.equal(b0, b1) and torch.equal(p0, p1) and i0 == i1


@pytest.mark.parametrize("kind,cname,snr", [("rcq", "dvbs2", 3.0), ("wrcq1", "qc", 5.8), ("n2d2", "qc", 5.8),
                                             ("wrcq1", "dvbs2", 3.0)])
def test_full_size_named_configs_properties_and_oracle_sample(built_lib, kind, cname, snr):
    """BASELINE configs 3 and 4 at their full code sizes (RCQ bc=3 on the (16200,7200) shape, W-RCQ type 1 on the
    (9472,8192)-shaped QC code, both through the row-ring / SWAR kernels): success <=> zero syndrome,
    iterations == T for every failed frame, and a random sample of frames bit-exact against the oracle."""
    from oracle import capi as O
    from oracle.restatement import MODE_NMS, MODE_RCQ, MODE_WRCQ, SparseGraph, quantizer_schedule
    L = built_lib
    T = 12
    code = L.codes.dvbs2_shaped(max_iterations=T) if cname == "dvbs2" else L.codes.qc_shaped(max_iterations=T)
    g = code.graph
    B = 3000
    qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
    torch.manual_seed(5)
    if kind == "rcq":
        dec = L.RCQMinSumDecoder(code, bc=3, bv=8, quantizer_params=qp, max_iterations=T)
    elif kind == "wrcq1":
        dec = L.WeightedRCQDecoder(code, bc=3, bv=8, quantizer_params=qp, weight_sharing_type=1, max_iterations=T)
        with torch.no_grad():
            dec._beta_table.uniform_(0.8, 1.0)
    else:
        dec = L.Neural2DMinSumDecoder(code, 2, T)
        with torch.no_grad():
            dec._beta_table.uniform_(0.7, 0.9)
            dec._alpha_table.fill_(1.0)
    llr = torch.cat([L.awgn_llr(g.n, B // 2, snr, seed=21, llr_sign=1), L.awgn_llr(g.n, B - B // 2, snr + 1.0, seed=22, llr_sign=1)])
    eng = dec._engine(0)
    bits, post, iters, succ = eng.decode_device(llr, want_posterior=(kind != "rcq"))
    bits_h, iters_h, succ_h = bits.cpu().numpy(), iters.cpu().numpy(), succ.cpu().numpy().astype(bool)
    ok = g.syndrome(bits_h).sum(axis=1) == 0
    assert np.array_equal(ok, succ_h)
    assert (iters_h[~ok] == T).all() and (iters_h >= 1).all() and (iters_h <= T).all()
    assert len(set(iters_h.tolist())) > 2, "operating point should spread the stopping iterations"
    pick = np.random.default_rng(1).choice(B, 16, replace=False)
    og = SparseGraph.from_coo(g.n, g.m, g.edge_check, g.check_var)
    x = llr[pick].cpu().numpy()
    if kind == "n2d2":
        ref = O.decode(og, x, T=T, mode=MODE_NMS, beta=dec._beta_table.detach().numpy()[:, dec._beta_index],
                       alpha=dec._alpha_table.detach().numpy()[:, dec._alpha_index], nthreads=8)
    else:
        thr = np.array([q.thresholds for q in dec.quantizers], dtype=np.float64).astype(np.float32)
        kw = dict(T=T, bc=3, thresholds=thr, quantizer_of_iter=quantizer_schedule(T, 3), nthreads=8)
        if kind == "rcq":
            ref = O.decode(og, x, mode=MODE_RCQ, **kw)
        else:
            ref = O.decode(og, x, mode=MODE_WRCQ, beta=dec._beta_table.detach().numpy()[:, dec._beta_index],
                           alpha=np.ones((T, g.n), np.float32), **kw)
    assert np.array_equal(bits_h[pick], ref.bits) and np.array_equal(iters_h[pick], ref.iterations)
    assert np.array_equal(succ_h[pick], ref.success.astype(bool))
    if post is not None:
        assert np.array_equal(post[pick].cpu().numpy(), ref.posterior)


@pytest.mark.parametrize("H", [np.zeros((3, 5), dtype=np.int64), np.array([[1]]), np.array([[1, 0, 0], [0, 0, 0]]),
                               np.zeros((0, 4), dtype=np.int64), np.ones((2, 1), dtype=np.int64)],
                         ids=["no-edges", "1x1", "one-edge", "no-checks", "one-variable"])
def test_degenerate_graphs(built_lib, H):
    """Empty / trivial Tanner graphs: no edges at all (every frame 'succeeds' after one iteration with
    bits = llr < 0), a single edge, zero checks, a single variable -- against the oracle."""
    from oracle import capi as O
    from oracle.restatement import MODE_RCQ, SparseGraph, quantizer_schedule
    L = built_lib
    m, n = H.shape
    T = 3
    code = L.LDPCCode(n, max(n - m, 0), H, max_iterations=T)
    og = SparseGraph.from_dense(H)
    rng = np.random.default_rng(m * 10 + n)
    llr = (rng.standard_normal((5, n)) * 2).astype(np.float32)
    llr[0, 0] = 0.0
    dec = L.Neural2DMinSumDecoder(code, 4, T)
    b, p, i = dec(torch.from_numpy(llr).cuda())
    ref = O.decode(og, llr, T=T, beta=np.full((T, og.E), np.float32(0.7)))
    assert np.array_equal(b.cpu().numpy(), ref.bits) and np.array_equal(i.cpu().numpy(), ref.iterations)
    assert np.array_equal(p.cpu().numpy(), ref.posterior)
    bb, ss, ii = L.BasicMinSumDecoder(code, 0.7).decode(llr.astype(np.float64))
    ref = O.decode(og, llr.astype(np.float64), T=T, dtype=np.float64, beta=np.full((T, og.E), 0.7))
    assert np.array_equal(bb, ref.bits) and np.array_equal(ii, ref.iterations) and np.array_equal(ss, ref.success)
    qp = [(3.0, 1.3)]
    rcq = L.RCQMinSumDecoder(code, 3, 8, qp, max_iterations=T)
    b, s, i = rcq.decode(torch.from_numpy(llr).cuda())
    thr = np.array([q.thresholds for q in rcq.quantizers]).astype(np.float32)
    ref = O.decode(og, llr, T=T, mode=MODE_RCQ, bc=3, thresholds=thr, quantizer_of_iter=quantizer_schedule(T, 1))
    assert np.array_equal(b.cpu().numpy(), ref.bits) and np.array_equal(i.cpu().numpy(), ref.iterations)
    assert np.array_equal(s.cpu().numpy().astype(bool), ref.success)


def test_launch_bound_decodes_replay_a_cuda_graph(built_lib, monkeypatch):
    """Tiny codes / tiny batches are launch-bound: the T-iteration launch sequence is captured once and replayed
    (profile.graph_replays), with identical results to the plain launch path (LDPC_GRAPHS=0), also after the
    weights change and after the workspace is re-allocated for a larger batch."""
    L = built_lib
    code = L.create_test_ldpc_code()
    rng = np.random.default_rng(3)
    llr = torch.from_numpy((rng.standard_normal((300, 7)) * 2 + 0.7).astype(np.float32)).cuda()

    def make():
        torch.manual_seed(1)
        d = L.Neural2DMinSumDecoder(code, 2, 10)
        with torch.no_grad():
            d._beta_table.uniform_(0.5, 1.0)
            d._alpha_table.uniform_(0.8, 1.0)
        return d

    monkeypatch.setenv("LDPC_SMALL", "0")     # (this code would otherwise take the one-launch on-chip decode,
    monkeypatch.setenv("LDPC_RESIDENT", "0")  #  batches of a few frames the CTA-resident one)
    monkeypatch.setenv("LDPC_GRAPHS", "0")
    plain = make()
    ref = [plain(llr), plain(llr[:5]), plain(llr[7])]
    assert plain._engine(0).profile_read()["graph_replays"] == 0
    monkeypatch.setenv("LDPC_GRAPHS", "1")
    dec = make()
    for _ in range(3):
        out = [dec(llr), dec(llr[:5]), dec(llr[7])]
        for a, b in zip(ref, out):
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
            assert (a[2] == b[2]) if isinstance(a[2], int) else torch.equal(a[2], b[2])
    prof = dec._engine(0).profile_read()
    assert prof["graph_replays"] == 9 and prof["cn_launches"] == 90, prof
    with torch.no_grad():                      # new weights: same graph, new table contents
        dec._beta_table.mul_(0.9)
        plain._beta_table.mul_(0.9)
    a, b = plain(llr), dec(llr)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    big = torch.from_numpy((rng.standard_normal((5000, 7)) * 2 + 0.7).astype(np.float32)).cuda()
    a, b = plain(big), dec(big)                # workspace grows: cached graphs are dropped and re-captured
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    a, b = plain(llr), dec(llr)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert dec._engine(0).profile_read()["graph_replays"] == 3


@pytest.mark.parametrize("kind", ["n2d2", "rcq", "basic"])
def test_encode_channel_symmetry_round_trip_full_size(built_lib, kind):
    """Size-independent property at the full (16200,7200) shape: encode random information bits (the code is
    IRA: parity = running XOR of the information syndromes), send the codeword c instead of the all-zero word
    over the same noise (llr_c = llr_0 * (1 - 2c)), and the decoder must return bits_0 XOR c with the same
    iteration counts and the sign-flipped posteriors, bit for bit (min-sum is symmetric; negation is exact)."""
    L = built_lib
    T = 10
    code = L.codes.dvbs2_shaped(max_iterations=T)
    g = code.graph
    n, k, m = g.n, code.k, g.m
    B = 1536
    rng = np.random.default_rng(12)
    H = code.H.tocsr()
    info = rng.integers(0, 2, size=(B, k)).astype(np.int64)
    s = (H[:, :k] @ info.T) % 2                       # [m, B] information part of every check
    par = np.cumsum(s, axis=0) % 2                    # p_i = p_{i-1} xor s_i (dual-diagonal parity part)
    cw = np.concatenate([info, par.T], axis=1).astype(np.uint8)
    assert not g.syndrome(cw[:64]).any() and cw.any()
    llr0 = torch.cat([L.awgn_llr(n, B // 2, 2.2, seed=3, llr_sign=1), L.awgn_llr(n, B - B // 2, 3.0, seed=4, llr_sign=1)])
    flip = torch.from_numpy(1.0 - 2.0 * cw.astype(np.float32)).cuda()
    llrc = llr0 * flip
    if kind == "n2d2":
        dec = L.Neural2DMinSumDecoder(code, 2, T)
        with torch.no_grad():
            dec._beta_table.fill_(0.8)
            dec._alpha_table.fill_(0.97)
        b0, p0, i0 = dec(llr0)
        bc, pc, ic = dec(llrc)
        assert torch.equal(pc, p0 * flip)
    elif kind == "rcq":
        dec = L.RCQMinSumDecoder(code, 3, 8, [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)], max_iterations=T)
        b0, s0, i0 = dec.decode(llr0)
        bc, sc, ic = dec.decode(llrc)
        assert torch.equal(s0, sc)
    else:
        dec = L.BasicMinSumDecoder(code, 0.8)
        b0, s0, i0 = dec.decode(llr0.double())
        bc, sc, ic = dec.decode(llrc.double())
        assert torch.equal(s0, sc)
    cwt = torch.from_numpy(cw).to(b0.device)
    assert torch.equal(i0, ic) and len(set(i0.tolist())) > 2
    # (a posterior of exactly 0 would decide bit 0 on both sides; with continuous noise that does not occur)
    assert torch.equal(bc.to(torch.uint8), b0.to(torch.uint8) ^ cwt)


def test_one_decoder_many_calls_of_changing_shape(built_lib, monkeypatch):
    """One handle, 60 calls with changing batch sizes, entry points and output sets (workspace growth, graph
    cache eviction and invalidation, compaction levels of changing size, host pipeline staging re-allocation):
    every call must equal a fresh handle's answer for the same frames."""
    L = built_lib
    monkeypatch.setenv("LDPC_COMPACT_MIN_FRAMES", "128")
    monkeypatch.setenv("LDPC_HOST_CHUNK", "512")
    T = 16
    code = L.codes.dvbs2_shaped(max_iterations=T, scale=20)
    rng = np.random.default_rng(77)

    def make():
        torch.manual_seed(5)
        d = L.Neural2DMinSumDecoder(code, 2, T)
        with torch.no_grad():
            d._beta_table.uniform_(0.7, 0.9)
            d._alpha_table.uniform_(0.95, 1.0)
        return d

    dec = make()
    pool = torch.cat([L.awgn_llr(code.n, 1500, snr, seed=30 + k, llr_sign=1) for k, snr in enumerate((1.0, 2.5, 3.5, 6.0))])
    pool = pool[torch.randperm(pool.shape[0], generator=torch.Generator().manual_seed(1)).cuda()]
    sizes = [1, 3000, 7, 128, 129, 2, 5000, 640, 1, 333] + [int(x) for x in rng.integers(1, 4000, size=50)]
    for k, B in enumerate(sizes):
        off = int(rng.integers(0, pool.shape[0] - B + 1))
        x = pool[off:off + B]
        mode = k % 3
        fresh = make()
        if mode == 0:                                   # device, with posterior
            a, b = dec(x), fresh(x)
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]), (k, B)
        elif mode == 1:                                 # device, decisions only (mask-free kernels)
            a = dec._engine(0).decode_device(x)
            b = fresh._engine(0).decode_device(x)
            assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3]), (k, B)
        else:                                           # host pipeline (chunks of 512 frames)
            a, b = dec(x.cpu()), fresh(x.cpu())
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]), (k, B)
        del fresh
    prof = dec._engine(0).profile_read()
    assert prof["graph_replays"] > 0 and prof["compactions"] > 0, prof
