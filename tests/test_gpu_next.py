"""SURVEY section 8f rows 1-2 on the GPU: offset min-sum decoders and the layered RCQ schedule, against
the golden vectors produced by the live reference and against the oracle on seeded batches."""
import numpy as np
import pytest
import torch

from conftest import Golden
from test_oracle import next_cases

pytestmark = pytest.mark.gpu


def make_next_decoder(L, g: Golden):
    T = int(g["T"])
    code = L.LDPCCode(n=int(g["n"]), k=int(g["k"]), H=g["H"].astype(np.int64), max_iterations=T)
    if g.kind == "rcq_layered":
        return L.RCQMinSumDecoder(code, bc=int(g["bc"]), bv=8, quantizer_params=[tuple(p) for p in g["qp"]],
                                  max_iterations=T, layered=True)
    if g.kind == "noms":
        dec = L.NeuralOffsetMinSumDecoder(code, max_iterations=T)
    else:
        dec = L.Neural2DOffsetMinSumDecoder(code, weight_sharing_type=int(g["wtype"]), max_iterations=T)
    dec.load_reference_state_dict({str(k): torch.tensor([float(v)]) for k, v in zip(g["weight_keys"], g["weight_vals"])})
    return dec


@pytest.mark.parametrize("stem,case", next_cases())
@pytest.mark.parametrize("device", ["cuda", "host"])
def test_next_rows_golden(built_lib, stem, case, device):
    L = built_lib
    g = Golden(stem, case)
    if stem.startswith("onecheck"):
        # the only graph family where the reference's "subtract previous C2V" is live: rejected, loudly
        dec = make_next_decoder(L, g)
        with pytest.raises(L.LdpcError):
            dec.decode(torch.from_numpy(g["llr"]))
        return
    dec = make_next_decoder(L, g)
    x = torch.from_numpy(g["llr"])
    if device == "cuda":
        x = x.cuda()
    if g.kind == "rcq_layered":
        b, s, i = dec.decode(x)
        assert np.array_equal(s.cpu().numpy(), g["success"])
    else:
        b, p, i = dec(x)
        assert np.array_equal(p.cpu().numpy(), g["posterior"])
    assert np.array_equal(b.cpu().numpy().astype(np.uint8), g["bits"])
    assert np.array_equal(i.cpu().numpy(), g["iterations"])


def test_weight_tables_of_offset_decoders(built_lib):
    L = built_lib
    for stem, case in next_cases():
        g = Golden(stem, case)
        if g.kind not in ("noms", "n2doms"):
            continue
        dec = make_next_decoder(L, g)
        T, graph = int(g["T"]), dec.code.graph
        b, a = dec._tables()
        beta = b[:, dec._beta_index] if b is not None else np.zeros((T, graph.E), np.float32)
        alpha = a[:, dec._alpha_index] if a is not None else np.zeros((T, graph.n), np.float32)
        assert np.array_equal(beta, g["beta_edge"]) and np.array_equal(alpha, g["alpha_var"])


def _awgn(rng, B, n, snr_db, sign=1.0):
    s2 = 10 ** (-snr_db / 10)
    return (2 * (sign + np.sqrt(s2) * rng.standard_normal((B, n))) / s2).astype(np.float32)


@pytest.mark.parametrize("cname,snrs", [("dvbs2_s20", (2.0, 3.5)), ("qc_z16", (5.0, 7.0))])
def test_offset_and_layered_vs_oracle(built_lib, cname, snrs):
    from oracle import capi as O
    from oracle.restatement import MODE_OFFSET, SparseGraph, quantizer_schedule
    L = built_lib
    code = (L.codes.dvbs2_shaped(max_iterations=10, scale=20) if cname == "dvbs2_s20"
            else L.codes.qc_shaped(max_iterations=10, Z=16))
    g = code.graph
    og = SparseGraph.from_coo(g.n, g.m, g.edge_check, g.check_var)
    rng = np.random.default_rng(17)
    B, T = 150, 10
    llr = np.concatenate([_awgn(rng, B // 2, g.n, snrs[0]), _awgn(rng, B - B // 2, g.n, snrs[1])])
    llr[3, ::9] = 0.0          # exact zeros: the three-valued sign matters for the offset rule
    llr[4, 5::11] = -0.0
    for wt in (1, 2, 3, 4):
        torch.manual_seed(wt)
        dec = L.Neural2DOffsetMinSumDecoder(code, wt, T)
        with torch.no_grad():
            if dec._beta_table is not None:
                dec._beta_table.uniform_(-0.1, 0.5)
            if dec._alpha_table is not None:
                dec._alpha_table.uniform_(-0.05, 0.15)
        b, p, i = dec(torch.from_numpy(llr).cuda())
        bt, at = dec._tables()
        beta = bt[:, dec._beta_index] if bt is not None else None
        alpha = at[:, dec._alpha_index] if at is not None else None
        ref = O.decode(og, llr, T=T, mode=MODE_OFFSET, beta=beta, alpha=alpha, nthreads=8)
        assert np.array_equal(b.cpu().numpy(), ref.bits), wt
        assert np.array_equal(i.cpu().numpy(), ref.iterations), wt
        assert np.array_equal(p.cpu().numpy(), ref.posterior), wt
    torch.manual_seed(9)
    dec = L.NeuralOffsetMinSumDecoder(code, T)
    b, p, i = dec(torch.from_numpy(llr).cuda())
    ref = O.decode(og, llr, T=T, mode=MODE_OFFSET, beta=dec._beta_table.detach().numpy(), nthreads=8)
    assert np.array_equal(b.cpu().numpy(), ref.bits) and np.array_equal(i.cpu().numpy(), ref.iterations)
    assert np.array_equal(p.cpu().numpy(), ref.posterior)
    qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
    lay = L.RCQMinSumDecoder(code, 3, 8, qp, max_iterations=T, layered=True)
    b, s, i = lay.decode(torch.from_numpy(llr).cuda())
    thr = np.array([q.thresholds for q in lay.quantizers]).astype(np.float32)
    ref = O.decode_layered_rcq(og, llr, T=T, bc=3, thresholds=thr, quantizer_of_iter=quantizer_schedule(T, 3), nthreads=8)
    assert np.array_equal(b.cpu().numpy(), ref.bits) and np.array_equal(i.cpu().numpy(), ref.iterations)
    assert np.array_equal(s.cpu().numpy(), ref.success)
    assert len(set(ref.iterations.tolist())) > 2


def test_create_test_decoders_builds_all_ten(built_lib):
    L = built_lib
    code = L.create_test_ldpc_code()
    decs = L.create_test_decoders(code)
    assert list(decs) == ['Basic MinSum', 'N-NMS', 'N-OMS', 'N-2D-NMS Type 1', 'N-2D-NMS Type 2', 'N-2D-NMS Type 3',
                          'N-2D-NMS Type 4', 'N-2D-OMS Type 2', 'RCQ MinSum', 'W-RCQ Type 2']
    llr = torch.tensor([1.25, -0.5, 2.75, -3.125, 0.625, -1.875, 4.5])
    for name, d in decs.items():
        out = d.decode(llr.double().numpy()) if name == 'Basic MinSum' else (d.decode(llr) if name == 'RCQ MinSum' else d(llr))
        assert len(out) == 3 and len(out[0]) == 7 and isinstance(out[2], int)


def test_layered_level_parallel_equals_sequential(built_lib, monkeypatch):
    """Quasi-cyclic codes have few dependency levels (one per block row): the level-parallel layered kernel must
    give exactly the sequential kernel's result (full (9472,8192)-shaped code, frames stopping at different
    iterations), and a chain-structured code must keep using the sequential kernel."""
    L = built_lib
    T = 10
    qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
    code = L.codes.qc_shaped(max_iterations=T)
    llr = torch.cat([L.awgn_llr(code.n, 5600, snr, seed=60 + k, llr_sign=1) for k, snr in enumerate((5.0, 6.5, 8.0))])   # >= 16 384 frames: staged
    outs = []
    for flag, stage in (("0", "1"), ("1", "1"), ("1", "0"), ("1", "2")):   # sequential / level-parallel staged (4 frames per lane) / unstaged / staged (2 frames per lane)
        monkeypatch.setenv("LDPC_LAYERED_LEVELS", flag)
        monkeypatch.setenv("LDPC_LAYERED_STAGE", stage)
        dec = L.RCQMinSumDecoder(code, 3, 8, qp, max_iterations=T, layered=True)
        b, s, i = dec.decode(llr)
        outs.append((b, s, i, dec._engine(0).profile_read()["cn_launches"]))
    monkeypatch.delenv("LDPC_LAYERED_STAGE")
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    for k in (2, 3):
        assert torch.equal(outs[k][0], outs[1][0]) and torch.equal(outs[k][1], outs[1][1]) and torch.equal(outs[k][2], outs[1][2])
    assert len(set(outs[1][2].tolist())) > 2
    assert outs[1][3] > outs[0][3] >= 1          # one launch per level instead of one per iteration
    chain = L.codes.dvbs2_shaped(max_iterations=4, scale=20)      # dual-diagonal parity: as many levels as checks
    dec = L.RCQMinSumDecoder(chain, 3, 8, qp, max_iterations=4, layered=True)
    dec.decode(L.awgn_llr(chain.n, 64, 0.0, seed=1, llr_sign=-1))
    assert dec._engine(0).profile_read()["cn_launches"] == 4


def test_layered_pipelined_walk_equals_plain_walk(built_lib, monkeypatch):
    """Chain-structured code at full size ((16200,7200) shape, dual-diagonal parity: 9000 dependency levels): the
    software-pipelined sequential walk (inputs copied ahead / forwarded through shared memory) must give exactly
    the plain sequential kernel's decisions, flags, iteration counts -- frames stopping at different iterations,
    a batch that is not a multiple of the CTA size."""
    L = built_lib
    T = 10
    qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
    code = L.codes.dvbs2_shaped(max_iterations=T)
    assert code.graph.query(0, 9) == 1 and code.graph.query(0, 8) > 8000
    llr = torch.cat([L.awgn_llr(code.n, 333, snr, seed=80 + k, llr_sign=1) for k, snr in enumerate((2.0, 3.5, 5.0))])
    outs = []
    for pipe, v2_from in (("0", "0"), ("1", "0"), ("1", "128")):   # plain walk | pipelined | pipelined, two frames per thread
        monkeypatch.setenv("LDPC_LAYERED_PIPE", pipe)
        monkeypatch.setenv("LDPC_LAYERED_V2_FRAMES", v2_from)
        dec = L.RCQMinSumDecoder(code, 3, 8, qp, max_iterations=T, layered=True)
        outs.append(dec.decode(llr))
    for k in (1, 2):
        for a, b in zip(outs[0], outs[k]):
            assert torch.equal(a, b)
    assert len(set(outs[1][2].tolist())) > 2


@pytest.mark.parametrize("seed", range(6))
def test_layered_pipelined_walk_dense_overlaps_vs_oracle(built_lib, monkeypatch, seed):
    """Small dense graphs: most variables are shared by checks fewer than a ring depth apart (forwarding at every
    distance 1..7, several readers in a row, variables first read late in the walk), degree-1 and empty checks."""
    from oracle import capi as O
    from oracle.restatement import SparseGraph, quantizer_schedule
    L = built_lib
    rng = np.random.default_rng(500 + seed)
    m, n = int(rng.integers(12, 60)), int(rng.integers(9, 40))
    H = np.zeros((m, n), dtype=np.int64)
    for i in range(m):
        H[i, rng.choice(n, int(rng.integers(1, 9)), replace=False)] = 1
    H[rng.integers(1, m - 1), :] = 0
    T = int(rng.integers(2, 9))
    code = L.LDPCCode(n, max(1, n - m), H, max_iterations=T)
    assert code.graph.query(0, 9) == 1
    llr = (3.0 * (1.0 + 1.1 * rng.standard_normal((257, n)))).astype(np.float32)
    llr[rng.random(llr.shape) < 0.02] = 0.0
    qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
    monkeypatch.setenv("LDPC_LAYERED_V2_FRAMES", "128" if seed % 2 else "0")    # odd seeds: two frames per thread
    dec = L.RCQMinSumDecoder(code, 3, 8, qp, max_iterations=T, layered=True)
    b, s, i = dec.decode(torch.from_numpy(llr).cuda())
    thr = np.array([q.thresholds for q in dec.quantizers]).astype(np.float32)
    ref = O.decode_layered_rcq(SparseGraph.from_dense(H), llr, T=T, bc=3, thresholds=thr,
                               quantizer_of_iter=quantizer_schedule(T, 3), nthreads=4)
    assert np.array_equal(b.cpu().numpy(), ref.bits) and np.array_equal(i.cpu().numpy(), ref.iterations)
    assert np.array_equal(s.cpu().numpy(), ref.success)
    # ragged batches (partly filled warps / CTAs, stopped lanes walking along) and repeated calls (the rings are
    # filled asynchronously: a race would show as a run-to-run difference)
    for B in (1, 31, 129):
        for _ in range(3):
            b2, s2, i2 = dec.decode(torch.from_numpy(llr[:B]).cuda())
            assert np.array_equal(b2.cpu().numpy(), ref.bits[:B]) and np.array_equal(i2.cpu().numpy(), ref.iterations[:B])


@pytest.mark.parametrize("m", [2, 3, 7, 8, 9, 15, 16, 17, 23, 24, 25, 31, 32, 33, 57, 100])
def test_layered_pipelined_walk_ring_boundaries(built_lib, monkeypatch, m):
    """Chains of m checks around the sizes of the kernel's rings (8 check slots, records staged 24 checks ahead into
    a 32-record ring): dual-diagonal parity part plus random information columns, against the oracle."""
    from oracle import capi as O
    from oracle.restatement import SparseGraph, quantizer_schedule
    L = built_lib
    rng = np.random.default_rng(700 + m)
    k = max(3, m)                                   # information variables
    n = k + m
    H = np.zeros((m, n), dtype=np.int64)
    for i in range(m):
        H[i, k + i] = 1                             # parity chain: check i touches p_i and p_{i-1}
        if i:
            H[i, k + i - 1] = 1
        H[i, rng.choice(k, int(rng.integers(1, min(k, 6) + 1)), replace=False)] = 1
    T = 6
    code = L.LDPCCode(n, k, H, max_iterations=T)
    assert code.graph.query(0, 9) == 1
    llr = (4.0 * (1.0 + 0.9 * rng.standard_normal((200, n)))).astype(np.float32)
    qp = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]
    ref = None
    for v2_from in ("0", "128"):          # one frame per thread | two frames per thread (normally from 65 536 frames on)
        monkeypatch.setenv("LDPC_LAYERED_V2_FRAMES", v2_from)
        dec = L.RCQMinSumDecoder(code, 3, 8, qp, max_iterations=T, layered=True)
        b, s, i = dec.decode(torch.from_numpy(llr).cuda())
        if ref is None:
            thr = np.array([q.thresholds for q in dec.quantizers]).astype(np.float32)
            ref = O.decode_layered_rcq(SparseGraph.from_dense(H), llr, T=T, bc=3, thresholds=thr,
                                       quantizer_of_iter=quantizer_schedule(T, 3), nthreads=4)
        assert np.array_equal(b.cpu().numpy(), ref.bits) and np.array_equal(i.cpu().numpy(), ref.iterations), v2_from
        assert np.array_equal(s.cpu().numpy(), ref.success), v2_from
