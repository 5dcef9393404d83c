"""Monte-Carlo leg on the device: Philox AWGN generator, error counting, the fused round and the
LDPSimulator loop."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

M32 = 0xFFFFFFFF


def philox4x32_10(c, k):
    """Reference Philox4x32-10 (Salmon et al.) on python ints."""
    c = list(c)
    k = list(k)
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & M32, p1 & M32, ((p0 >> 32) ^ c[3] ^ k[1]) & M32, p0 & M32]
        k = [(k[0] + 0x9E3779B9) & M32, (k[1] + 0xBB67AE85) & M32]
    return c


def model_llr(n, frames, frame0, seed, snr_db, sign, codeword=None):
    out = np.zeros((frames, n), dtype=np.float64)
    s2 = 1.0 / 10 ** (snr_db / 10)
    sigma = np.float32(np.sqrt(s2))
    kk = np.float32(2.0 / s2)
    for f in range(frames):
        gf = frame0 + f
        for jg in range((n + 3) // 4):
            r = philox4x32_10([jg, gf & M32, gf >> 32, 0x4C445043], [seed & M32, seed >> 32])
            u = [np.float32(x) * np.float32(2.3283064365386963e-10) + np.float32(1.1641532182693481e-10) for x in r]
            z = []
            for a, b in ((u[0], u[1]), (u[2], u[3])):
                rad = np.sqrt(-2.0 * np.log(np.float64(a)))
                z += [rad * np.cos(2 * np.pi * np.float64(b)), rad * np.sin(2 * np.pi * np.float64(b))]
            for i in range(4):
                j = 4 * jg + i
                if j < n:
                    cw = 0.0 if codeword is None else float(codeword[j])
                    sym = sign * (1.0 - 2.0 * cw)
                    out[f, j] = (sym + float(sigma) * z[i]) * float(kk)
    return out


def test_philox_known_answer():
    # Random123 known-answer test vectors for philox4x32-10
    assert philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert philox4x32_10([M32] * 4, [M32, M32]) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]


def test_awgn_matches_counter_based_model(built_lib):
    L = built_lib
    n, frames, frame0, seed = 23, 9, (1 << 32) - 3, 0x1234567890ABCDEF
    cw = np.array([j % 3 == 0 for j in range(n)], dtype=np.uint8)
    for sign, cwd in ((1, None), (-1, cw)):
        got = L.awgn_llr(n, frames, 2.5, seed=seed, frame0=frame0, llr_sign=sign,
                         codeword=None if cwd is None else torch.from_numpy(cwd).cuda()).cpu().numpy()
        want = model_llr(n, frames, frame0, seed, 2.5, sign, cwd)
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5)
    # statistics: mean 2/s2, variance 4/s2
    big = L.awgn_llr(2000, 2000, 1.0, seed=5).double()
    s2 = 10 ** (-0.1)
    assert abs(big.mean().item() - 2 / s2) < 0.01 and abs(big.var().item() - 4 / s2) < 0.03
    # frames are independent of batch boundaries
    a = L.awgn_llr(64, 300, 3.0, seed=9, frame0=100)
    b = L.awgn_llr(64, 100, 3.0, seed=9, frame0=250)
    assert torch.equal(a[150:250], b)


@pytest.mark.parametrize("kind", ["n2d", "rcq", "basic"])
def test_mc_round_equals_decode_plus_count(built_lib, kind):
    L = built_lib
    code = L.codes.dvbs2_shaped(max_iterations=8, scale=20)
    if kind == "n2d":
        torch.manual_seed(1)
        dec = L.Neural2DMinSumDecoder(code, 2, 8)
        with torch.no_grad():
            dec._beta_table.uniform_(0.6, 0.9)
            dec._alpha_table.uniform_(0.9, 1.0)
    elif kind == "rcq":
        dec = L.RCQMinSumDecoder(code, 3, 8, [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)], max_iterations=8)
    else:
        dec = L.BasicMinSumDecoder(code, 0.7)
    eng = dec._engine(0)
    B, snr, seed, frame0 = 333, 1.6, 77, 1000
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    fbe = torch.zeros(B, dtype=torch.int32, device="cuda")
    fit = torch.zeros(B, dtype=torch.int32, device="cuda")
    eng.mc_round(snr, B, seed=seed, frame0=frame0, llr_sign=1, counters=counters, frame_bit_errors=fbe, frame_iterations=fit)
    llr = L.awgn_llr(code.n, B, snr, seed=seed, frame0=frame0, llr_sign=1)
    if kind == "basic":
        llr = llr.double()
    bits, _, iters, succ = eng.decode_device(llr)
    be = bits.sum(dim=1).int()
    assert torch.equal(fbe, be) and torch.equal(fit, iters)
    want = [int((be > 0).sum()), int(be.sum()), int(iters.sum()), B]
    assert counters.tolist() == want
    c2 = L.count_errors(bits, iterations=iters)
    assert c2.tolist() == want
    assert 0 < want[0] < B, "operating point should give both frame errors and successes"
    # non-zero codeword bookkeeping
    cw = torch.zeros(code.n, dtype=torch.uint8, device="cuda")
    cw[::7] = 1
    c3 = L.count_errors(bits, codeword=cw, iterations=iters)
    assert c3[1].item() == int((bits != cw[None]).sum()) and c3[3].item() == B


def test_simulator_loop_and_stop_rule(built_lib, tmp_path):
    L = built_lib
    from ldpc_b200.simulation_framework import LDPSimulator, SimulationConfig
    code = L.codes.dvbs2_shaped(max_iterations=10, scale=20)
    torch.manual_seed(0)
    dec = L.Neural2DMinSumDecoder(code, 2, 10)
    with torch.no_grad():
        dec._beta_table.fill_(0.8)
        dec._alpha_table.fill_(1.0)
    outs = {}
    for batch in (256, 1000):
        sim = LDPSimulator(SimulationConfig(batch_frames=batch, seed=3, save_results=False))
        outs[batch] = sim.simulate_single_snr(dec, code, 1.8, max_frames=4000, max_errors=40)
    a, b = outs[256], outs[1000]
    assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2] and a[4:] == b[4:], "result must not depend on the batch size"
    fer, ber, avg_it, secs, frames, errs = a
    assert errs == 40 and frames < 4000 and 0 < fer < 1 and 0 < ber < fer and 1 <= avg_it <= 10
    # reference sign convention: FER = 1, stops after max_errors frames exactly
    sim = LDPSimulator(SimulationConfig(batch_frames=512, reference_convention=True, save_results=False))
    fer, ber, avg_it, _, frames, errs = sim.simulate_single_snr(dec, code, 2.0, max_frames=10000, max_errors=100)
    assert (fer, frames, errs, avg_it) == (1.0, 100, 100, 10.0)
    # max_frames bound
    sim = LDPSimulator(SimulationConfig(batch_frames=300, save_results=False))
    fer, _, _, _, frames, errs = sim.simulate_single_snr(dec, code, 6.0, max_frames=700, max_errors=100)
    assert frames == 700 and errs < 100
    # sweep + JSON round trip
    sim = LDPSimulator(SimulationConfig(snr_range=(1.0, 2.0), snr_step=0.5, max_frames=600, max_errors=50,
                                        batch_frames=300, results_dir=str(tmp_path)))
    res = sim.simulate_multiple_decoders({"n2d": dec, "rcq": L.RCQMinSumDecoder(code, 3, 8, [(5.0, 1.3)], 10)}, code)
    assert res["n2d"].snr_values == [1.0, 1.5, 2.0] and len(res["rcq"].frame_error_rates) == 3
    assert res["n2d"].frame_error_rates[0] >= res["n2d"].frame_error_rates[2]
    sim.save_results(res, "r.json")
    assert sim.load_results("r.json")["n2d"].total_frames == res["n2d"].total_frames


def test_simulate_awgn_channel_reference_convention(built_lib):
    L = built_lib
    np.random.seed(4)
    cw = np.array([0, 1, 1, 0, 1, 0, 1])
    a = L.simulate_awgn_channel(cw, 30.0)
    assert a.dtype == np.float64 and a.shape == (7,)
    assert np.array_equal(a > 0, cw.astype(bool))        # bit 1 -> positive LLR (ldpc_decoder.py:289)
    np.random.seed(4)
    assert np.array_equal(a, L.simulate_awgn_channel(cw, 30.0))
