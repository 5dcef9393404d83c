"""CTA-resident decode (csrc/ldpc_resident.cu: one thread block per frame, the frame's messages in shared memory for
all T iterations) against the oracle and against the per-iteration path: every float32 decoder family, irregular
graphs (degree-0 variables, empty and degree-1 checks, check degrees beyond 8, variable degrees beyond 8), ragged
batches, early stop on / off, posteriors, packed rows and the host pipeline."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
QP = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]


def _graph(rng, which):
    if which == "small":
        m, n = int(rng.integers(3, 12)), int(rng.integers(6, 30))
    else:
        m, n = int(rng.integers(30, 90)), int(rng.integers(90, 200))
    H = np.zeros((m, n), dtype=np.int64)
    for i in range(m):
        H[i, rng.choice(n, int(rng.integers(1, min(n, 9))), replace=False)] = 1
    for _ in range(int(rng.integers(0, 3))):          # heavy checks (run-time degree path)
        H[rng.integers(0, m), rng.choice(n, min(int(rng.choice([9, 17, 33])), n), replace=False)] = 1
    for _ in range(int(rng.integers(0, 3))):          # heavy variables (local-array path)
        H[rng.choice(m, min(int(rng.choice([9, 12, 20])), m), replace=False), rng.integers(0, n)] = 1
    if rng.random() < 0.3:
        H[rng.integers(0, m), :] = 0
    if rng.random() < 0.3:
        H[:, rng.integers(0, n)] = 0
    return H


@pytest.mark.parametrize("seed", range(24))
def test_resident_decode_vs_oracle(built_lib, monkeypatch, seed):
    from oracle import capi as O
    from oracle.restatement import MODE_NMS, MODE_OFFSET, MODE_RCQ, MODE_WRCQ, SparseGraph, quantizer_schedule
    L = built_lib
    monkeypatch.setenv("LDPC_RESIDENT", "1")
    monkeypatch.setenv("LDPC_SMALL", "0")
    rng = np.random.default_rng(7300 + seed)
    H = _graph(rng, "small" if seed % 3 == 0 else "mid")
    m, n = H.shape
    T = int(rng.choice([1, 2, 6, 12]))
    B = int(rng.choice([1, 3, 150, 700]))
    code = L.LDPCCode(n, max(1, n - m), H, max_iterations=T)
    og = SparseGraph.from_dense(H)
    llr = (float(rng.choice([0.5, 2.0, 6.0])) * (1.0 + 1.2 * rng.standard_normal((B, n)))).astype(np.float32)
    llr[rng.random((B, n)) < 0.02] = 0.0
    x = torch.from_numpy(llr).cuda()
    kind = ["n2d", "nnms", "rcq", "wrcq", "oms", "n2d"][seed % 6]
    torch.manual_seed(seed)
    if kind in ("n2d", "nnms"):
        dec = L.Neural2DMinSumDecoder(code, int(rng.integers(1, 5)), T) if kind == "n2d" else L.NeuralMinSumDecoder(code, T)
        with torch.no_grad():
            if dec._beta_table is not None:
                dec._beta_table.uniform_(-0.2, 0.9)
            if dec._alpha_table is not None:
                dec._alpha_table.uniform_(0.3, 1.0)
        bt, at = dec._tables()
        ref = O.decode(og, llr, T=T, mode=MODE_NMS,
                       beta=bt[:, dec._beta_index] if dec._beta_table is not None else np.full((T, og.E), np.float32(0.7)),
                       alpha=at[:, dec._alpha_index] if at is not None else None, nthreads=8)
    elif kind == "oms":
        dec = L.Neural2DOffsetMinSumDecoder(code, int(rng.integers(1, 5)), T) if rng.random() < 0.5 else L.NeuralOffsetMinSumDecoder(code, T)
        with torch.no_grad():
            if dec._beta_table is not None:
                dec._beta_table.uniform_(-0.1, 0.6)
            if dec._alpha_table is not None:
                dec._alpha_table.uniform_(-0.05, 0.2)
        bt, at = dec._tables()
        ref = O.decode(og, llr, T=T, mode=MODE_OFFSET, beta=bt[:, dec._beta_index] if bt is not None else None,
                       alpha=at[:, dec._alpha_index] if at is not None else None, nthreads=8)
    else:
        bc = int(rng.choice([2, 3, 4, 8]))
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
            ref = O.decode(og, llr, mode=MODE_RCQ, **kw)
        else:
            bt, at = dec._tables()
            ref = O.decode(og, llr, mode=MODE_WRCQ,
                           beta=bt[:, dec._beta_index] if dec._beta_table is not None else np.full((T, og.E), np.float32(0.7)),
                           alpha=at[:, dec._alpha_index] if at is not None else np.ones((T, n), np.float32), **kw)
    eng = dec._engine(0)
    bits, post, iters, succ = eng.decode_device(x, want_posterior=True)
    assert eng.profile_read()["resident_decodes"] == 1
    assert np.array_equal(bits.cpu().numpy(), ref.bits)
    assert np.array_equal(iters.cpu().numpy(), ref.iterations)
    assert np.array_equal(succ.cpu().numpy().astype(bool), ref.success)
    assert np.array_equal(post.cpu().numpy(), ref.posterior)
    pk, _, i2, _ = eng.decode_device(x, packed_bits=True)
    assert np.array_equal(eng.unpack_rows(pk.cpu().numpy().view(np.uint32)), ref.bits) and torch.equal(i2, iters)
    hb, hp, hi, hs = eng.decode_host(llr, want_posterior=True)
    assert np.array_equal(hb, ref.bits) and np.array_equal(hp, ref.posterior) and np.array_equal(hi, ref.iterations)


def test_resident_equals_per_iteration_path_on_a_full_size_shape(built_lib, monkeypatch):
    """(16200,7200)-shaped code, E = 48 599: 194 KB of messages per frame, one block per SM."""
    L = built_lib
    T = 10
    code = L.codes.dvbs2_shaped(max_iterations=T)
    g = code.graph
    s2 = 10 ** (-2.6 / 10)
    rng = np.random.default_rng(4)
    llr = (2 * (1 + np.sqrt(s2) * rng.standard_normal((600, g.n))) / s2).astype(np.float32)
    llr[::7] *= -1            # frames that never converge
    x = torch.from_numpy(llr).cuda()
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("LDPC_RESIDENT", mode)
        dec = L.Neural2DMinSumDecoder(code, 2, T)
        with torch.no_grad():
            dec._beta_table.fill_(0.8)
            dec._alpha_table.fill_(0.95)
        eng = dec._engine(0)
        out[mode] = eng.decode_device(x, want_posterior=True)
        assert eng.profile_read()["resident_decodes"] == (1 if mode == "1" else 0)
        rcq = L.RCQMinSumDecoder(code, 3, 8, QP, max_iterations=T)
        out[mode] += rcq._engine(0).decode_device(x)
        eng0 = dec._engine(0)
        # early stop off
        from ldpc_b200.engine import Engine
        b, a = dec._tables()
        e2 = Engine(g, max_iterations=T, early_stop=False, beta=b, beta_index=dec._beta_index, alpha=a,
                    alpha_index=dec._alpha_index, device=0)
        out[mode] += e2.decode_device(x[:200], want_posterior=True)
    its = out["0"][2].cpu().numpy()
    assert len(set(its.tolist())) >= 3 and (its == T).any()
    for a, b in zip(out["0"], out["1"]):
        if a is None:
            assert b is None
        else:
            assert torch.equal(a, b)


def test_policy_takes_the_resident_decode_for_a_few_waves(built_lib, monkeypatch):
    """Without LDPC_RESIDENT the library decides (ldpc_api.cu fill_resident): a batch of at most a few waves of thread
    blocks (4 with one block per SM, 2 where an SM holds several frames) decodes CTA-resident, a larger one with the
    per-iteration kernels; results are the same either way."""
    L = built_lib
    monkeypatch.delenv("LDPC_RESIDENT", raising=False)
    monkeypatch.delenv("LDPC_RESIDENT_MAX_FRAMES", raising=False)
    T = 8
    small = L.codes.dvbs2_shaped(max_iterations=T, scale=20)        # ~10 KB per frame: many blocks per SM
    big = L.codes.dvbs2_shaped(max_iterations=T)                    # 194 KB per frame: one block per SM
    for code, cases in ((small, ((1, 1), (300, 1), (60000, 0))), (big, ((1, 1), (400, 1), (2000, 0)))):
        dec = L.Neural2DMinSumDecoder(code, 2, T)
        with torch.no_grad():
            dec._beta_table.fill_(0.8)
            dec._alpha_table.fill_(0.95)
        eng = dec._engine(0)
        for B, want in cases:
            x = L.awgn_llr(code.n, B, 3.0, seed=B, llr_sign=1)
            eng.profile_read(reset=True)
            got = eng.decode_device(x, want_posterior=True)
            assert eng.profile_read()["resident_decodes"] == want, (code.n, B)
            monkeypatch.setenv("LDPC_RESIDENT", str(1 - want))
            other = L.Neural2DMinSumDecoder(code, 2, T)
            with torch.no_grad():
                other._beta_table.fill_(0.8)
                other._alpha_table.fill_(0.95)
            e2 = other._engine(0)
            ref = e2.decode_device(x, want_posterior=True)
            assert e2.profile_read()["resident_decodes"] == 1 - want
            monkeypatch.delenv("LDPC_RESIDENT")
            for a, b in zip(got, ref):
                assert torch.equal(a, b)
            e2.close()
