"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol, the
reference-shaped classes build the right weight tables / index maps, code generators, quantiser class,
Monte-Carlo bookkeeping."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, ROOT, Golden, golden_cases


def test_library_loads_and_exports_every_declared_symbol(built_lib):
    from ldpc_b200 import _lib
    header = open(os.path.join(ROOT, "include", "ldpc_b200.h")).read()
    declared = set(re.findall(r"\b(ldpc_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert getattr(lib, name) is not None
    assert _lib.load().ldpc_version() == 104
    assert ctypes.sizeof(_lib.DecoderConfig) == 88 and ctypes.sizeof(_lib.Profile) == 96


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device failure mode")
def test_no_cpu_fallback_without_a_device(built_lib):
    L = built_lib
    code = L.create_test_ldpc_code()
    with pytest.raises(L.LdpcError):
        L.BasicMinSumDecoder(code).decode(np.zeros(7))
    with pytest.raises(L.LdpcError):
        L.Neural2DMinSumDecoder(code, 2, 5)(torch.zeros(7))


def test_package_has_no_oracle_import():
    pkg = os.path.join(ROOT, __import__("__graft_entry__").PKG)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "").lower() or f == "build.py", f


def test_ldpc_code_and_test_code(built_lib):
    L = built_lib
    code = L.create_test_ldpc_code()
    assert (code.n, code.k, code.max_iterations) == (7, 4, 10) and abs(code.rate - 4 / 7) < 1e-12
    assert code.check_node_degrees == {0: 3, 1: 3, 2: 3, 3: 4}
    assert code.variable_node_degrees == {0: 3, 1: 3, 2: 3, 3: 1, 4: 1, 5: 1, 6: 1}
    g = code.graph
    assert g.E == 13 and g.check_var.tolist() == [0, 1, 3, 1, 2, 4, 0, 2, 5, 0, 1, 2, 6]
    H2 = np.array([[1, 2, 0], [0, 1, 1]])
    g2 = L.LDPCCode(3, 1, H2).graph            # the entry 2 is not an edge (H == 1 test)
    assert g2.E == 3 and g2.check_var.tolist() == [0, 1, 2]
    assert g.syndrome(np.array([1, 1, 1, 0, 0, 0, 1])).tolist() == [[0, 0, 0, 0]]


def test_code_shapes_match_survey_appendix_d(built_lib):
    L = built_lib
    c = L.codes.dvbs2_shaped()
    g = c.graph
    assert (c.n, c.k, g.m, g.E) == (16200, 7200, 9000, 48599)
    assert dict(zip(*np.unique(g.var_degree, return_counts=True))) == {1: 1, 2: 8999, 3: 5400, 8: 1800}
    assert dict(zip(*np.unique(g.check_degree, return_counts=True))) == {4: 1441, 5: 3239, 6: 3600, 7: 720}
    c = L.codes.qc_shaped()
    g = c.graph
    assert (c.n, c.k, g.m, g.E) == (9472, 8192, 1280, 37888)
    assert set(g.var_degree.tolist()) == {4}
    assert dict(zip(*np.unique(g.check_degree, return_counts=True))) == {29: 512, 30: 768}
    assert np.array_equal(L.codes.dvbs2_shaped(seed=3).graph.check_var, L.codes.dvbs2_shaped(seed=3).graph.check_var)


def _make(L, g):
    import test_gpu_parity
    return test_gpu_parity.make_decoder(L, g)


@pytest.mark.parametrize("stem,case", [c for c in golden_cases() if Golden(*c).kind in ("nnms", "n2d", "wrcq")])
def test_weight_tables_reproduce_reference_lookups(built_lib, stem, case):
    """Import the reference's ParameterDict by key, then expand table[t, index[...]] and compare with the
    values the reference's own _get_beta_weight/_get_alpha_weight returned (recorded in the golden file)."""
    g = Golden(stem, case)
    dec = _make(built_lib, g)
    T = int(g["T"])
    b, a = dec._tables()
    graph = dec.code.graph
    if dec._beta_table is not None:
        beta = b[:, dec._beta_index]
    else:
        beta = np.broadcast_to(b, (T, graph.E)) if b is not None else np.ones((T, graph.E), np.float32)
    assert np.array_equal(beta.astype(np.float32), g["beta_edge"])
    alpha = a[:, dec._alpha_index] if a is not None else np.ones((T, graph.n), np.float32)
    assert np.array_equal(alpha.astype(np.float32), g["alpha_var"])
    # key views and state-dict round trip
    sd = dec.reference_state_dict()
    assert set(sd) == set(str(k) for k in g["weight_keys"])
    for k, v in zip(g["weight_keys"], g["weight_vals"]):
        assert float(sd[str(k)]) == float(v)


def test_parameter_counts_match_survey(built_lib):
    L = built_lib
    code = L.create_test_ldpc_code()
    want = {1: (40, 0), 2: (20, 20), 3: (20, 0), 4: (0, 20)}
    for wt, (nb, na) in want.items():
        d = L.Neural2DMinSumDecoder(code, wt, 10)
        assert (len(d.beta_weights), len(d.alpha_weights)) == (nb, na)
        assert sum(p.numel() for p in d.parameters()) == nb + na
    assert len(L.NeuralMinSumDecoder(code, 10).beta_weights) == 130
    with pytest.raises(ValueError):
        L.Neural2DMinSumDecoder(code, 5, 10)
    w = L.WeightedRCQDecoder(code, 3, 8, [(3.0, 1.3)], weight_sharing_type=9, max_iterations=4)  # no validation
    with pytest.raises(StopIteration):
        w(torch.zeros(7))


def test_initialisation_matches_reference_under_same_seed(built_lib):
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("live reference only exists in the build container")
    ref = ref_shim.load()
    L = built_lib
    code = L.create_test_ldpc_code()
    rcode = ref.ldpc_decoder.create_test_ldpc_code()
    for wt in (1, 2, 3, 4):
        torch.manual_seed(123)
        r = ref.neural_2d_decoder.Neural2DMinSumDecoder(rcode, wt, 6)
        torch.manual_seed(123)
        d = L.Neural2DMinSumDecoder(code, wt, 6)
        rs = {k: float(v) for k, v in r.state_dict().items()}
        ds = {k: float(v) for k, v in d.reference_state_dict().items()}
        assert rs == ds
        assert d.check_node_degrees == r.check_node_degrees and d.variable_node_degrees == r.variable_node_degrees
    torch.manual_seed(7)
    r = ref.neural_minsum_decoder.NeuralMinSumDecoder(rcode, 5)
    torch.manual_seed(7)
    d = L.NeuralMinSumDecoder(code, 5)
    assert {k: float(v) for k, v in r.state_dict().items()} == {k: float(v) for k, v in d.reference_state_dict().items()}
    # a reference checkpoint loads into ours
    d2 = L.NeuralMinSumDecoder(code, 5)
    d2.load_reference_state_dict(r.state_dict())
    assert torch.equal(d2._beta_table, d._beta_table)


def test_quantizer_class_matches_reference_known_answers(built_lib):
    L = built_lib
    z = np.load(f"{GOLDEN_DIR}/quantizer_kat.npz")
    q = L.NonUniformQuantizer(3, 5.0, 1.5)
    assert q.thresholds == z["thresholds"].tolist()
    codes = q.quantize(torch.tensor(z["x"]))
    assert codes.dtype == torch.int64 and np.array_equal(codes.numpy(), z["codes"])
    vals = q.dequantize(codes)
    assert vals.dtype == torch.float32 and np.array_equal(vals.numpy(), z["values"])
    r = L.RCQMinSumDecoder(L.create_test_ldpc_code(), 3, 8, [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)], max_iterations=10)
    assert [r.quantizers.index(r._get_quantizer(t)) for t in range(10)] == [0, 0, 0, 1, 1, 1, 2, 2, 2, 2]
    assert r.bv == 8 and len(r.quantizers) == 3
    assert L.RCQMinSumDecoder(L.create_test_ldpc_code(), 3, 8, [(3.0, 1.3)], layered=True).layered


def test_simulation_bookkeeping(built_lib, tmp_path):
    from ldpc_b200.simulation_framework import (LDPSimulator, SimulationConfig, SimulationResult, split_round,
                                                truncate_in_frame_order)
    shares = [split_round(10, 4, r) for r in range(4)]
    assert shares == [(0, 3), (3, 3), (6, 2), (8, 2)]
    be = np.array([0, 2, 0, 1, 3, 0])
    it = np.array([1, 2, 3, 4, 5, 6])
    assert truncate_in_frame_order(be, it, 1, 3) == (2, 3, 10, 4)
    assert truncate_in_frame_order(be, it, 0, 99) == (3, 6, 21, 6)
    res = SimulationResult("x", [0.0, 0.5, 1.0])
    res.add_result(2, 0.5, 0.1, 3.0, 1.5, 100, 50)
    assert res.frame_error_rates == [0.0, 0.0, 0.5] and res.total_frames == [0, 0, 100]
    sim = LDPSimulator(SimulationConfig(results_dir=str(tmp_path / "r")))
    sim.save_results({"x": res}, "a.json")
    back = sim.load_results("a.json")["x"]
    assert back.total_errors == [0, 0, 50] and back.snr_values == [0.0, 0.5, 1.0]
    import json
    blob = json.load(open(tmp_path / "r" / "a.json"))
    assert set(blob["x"]) == {"decoder_name", "snr_values", "frame_error_rates", "bit_error_rates",
                              "average_iterations", "simulation_times", "total_frames", "total_errors"}


def test_host_pipeline_chunk_plan(built_lib):
    """The chunk plan of ldpc_decode_host (test hook, no device needed): covers the batch exactly, no chunk above
    the limit, geometric ramp at the head, equal block-aligned parts after it, no small trailing chunk."""
    from ldpc_b200 import _lib
    lib = _lib.load()

    def plan(B, chunk, V=4):
        out = np.zeros(4096, dtype=np.int64)
        nc = ctypes.c_int32(0)
        _lib.check(lib.ldpc_host_chunk_plan(B, chunk, V, out.ctypes.data, out.size, ctypes.byref(nc)))
        return out[:nc.value].tolist()

    assert plan(65536, 8192) == [1024, 2048, 3072, 5120, 8192, 8192, 8192, 8192, 7168, 7168, 7168]
    assert plan(1, 8192) == [1] and plan(100, 8192) == [100] and plan(8192, 8192) == [8192]
    assert plan(65536, 0) == plan(65536, 4096)                       # default chunk
    rng = np.random.default_rng(0)
    for _ in range(300):
        B = int(rng.integers(1, 300000))
        chunk = int(rng.choice([0, 100, 512, 1000, 4096, 8192, 16384]))
        V = int(rng.choice([2, 4]))
        p = plan(B, chunk, V)
        lim = min(B, chunk if chunk > 0 else 4096)
        assert sum(p) == B and all(0 < c <= lim for c in p), (B, chunk, p)
        if B > 4 * lim and lim >= 1024:
            assert p[-1] >= lim // 2, (B, chunk, p)                   # no small un-overlapped tail
            assert p[0] <= max(lim // 8, 128 * 8), (B, chunk, p)        # the kernels start early


def test_figures_go_through_matplotlib_when_it_is_there(built_lib, monkeypatch):
    """The plot methods of the simulator / trainer / analyzer (simulation_framework.py:218-336, training_framework.py:266-295,
    354-377) import matplotlib on demand: a recording stand-in sees one curve per decoder and panel, the file name and
    show(); without matplotlib the call is an ImportError, never a silent no-op."""
    import sys
    import types
    L = built_lib
    from ldpc_b200.training_framework import GradientExplosionAnalyzer, PosteriorJointTrainer, TrainingConfig
    calls = []

    class Axis:
        def __getattr__(self, name):
            return lambda *a, **k: calls.append((name, a, k))

    plt = types.ModuleType("matplotlib.pyplot")

    def subplots(rows=1, cols=1, **kw):
        calls.append(("subplots", (rows, cols), kw))
        grid = [[Axis() for _ in range(cols)] for _ in range(rows)]
        return object(), (grid[0][0] if rows * cols == 1 else grid[0] if rows == 1 else grid)

    plt.subplots = subplots
    for fn in ("tight_layout", "savefig", "show"):
        setattr(plt, fn, (lambda n: lambda *a, **k: calls.append((n, a, k)))(fn))
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)

    res = {}
    for name in ("A", "B"):
        r = L.SimulationResult(name, [0.0, 1.0, 2.0])
        for i in range(3):
            r.add_result(i, 0.1 / (i + 1), 0.01 / (i + 1), 5.0 - i, 0.2, 100, 10 - i)
        res[name] = r
    sim = L.LDPSimulator(L.SimulationConfig(save_results=False))
    sim.plot_fer_curves(res, save_path="fer.png")
    assert [c[0] for c in calls].count("semilogy") == 2 and ("savefig", ("fer.png",), {"dpi": 300, "bbox_inches": "tight"}) in calls
    assert calls[-1][0] == "show"
    calls.clear()
    sim.plot_ber_curves(res, log_scale=False)
    sim.plot_iteration_curves(res)
    assert [c[0] for c in calls].count("plot") == 4 and not any(c[0] in ("semilogy", "savefig") for c in calls)
    calls.clear()
    sim.plot_comprehensive_comparison(res)
    names = [c[0] for c in calls]
    assert names.count("semilogy") == 4 and names.count("plot") == 4 and names.count("set_title") == 4
    assert {c[1][0] for c in calls if c[0] == "set_ylabel"} == {"Frame Error Rate (FER)", "Bit Error Rate (BER)", "Average Iterations",
                                                                "Simulation Time (s)"}
    calls.clear()
    tr = PosteriorJointTrainer.__new__(PosteriorJointTrainer)
    tr.train_losses, tr.train_accuracies, tr.gradient_norms = [1.0, 0.5], [0.6, 0.8], [2.0, 1.0]
    tr.plot_training_history()
    assert [c[1][0] for c in calls if c[0] == "set_title"] == ["Training Loss", "Training Accuracy", "Gradient Norm"]
    calls.clear()
    GradientExplosionAnalyzer.plot_gradient_analysis(None, {"gradient_magnitudes": [1.0, 2.0], "iteration_counts": [3, 4]}, save_path="g.png")
    assert [c[0] for c in calls].count("hist") == 1 and [c[0] for c in calls].count("scatter") == 1
    # no matplotlib: a clear error
    monkeypatch.setitem(sys.modules, "matplotlib", None)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", None)
    with pytest.raises(ImportError):
        sim.plot_fer_curves(res)
