"""Checkpoint interchange with the reference's nn.ParameterDict layout (neural_2d_decoder.py:46-82,
neural_minsum_decoder.py:47-53, rcq_decoder.py:398-431): ``state_dict()`` / ``load_state_dict()`` of our modules
speak the reference's keys in both directions.  The live-reference legs run in the build container only; the
key-format legs run everywhere (no GPU needed: no decode is called)."""
import numpy as np
import pytest
import torch

QP = [(3.0, 1.3), (5.0, 1.3), (7.0, 1.3)]


def _ours(L, kind, code, T):
    if kind == "nnms":
        return L.NeuralMinSumDecoder(code, T)
    if kind == "noms":
        return L.NeuralOffsetMinSumDecoder(code, T)
    if kind.startswith("n2d"):
        return L.Neural2DMinSumDecoder(code, int(kind[-1]), T)
    if kind.startswith("oms"):
        return L.Neural2DOffsetMinSumDecoder(code, int(kind[-1]), T)
    return L.WeightedRCQDecoder(code, 3, 8, QP, weight_sharing_type=int(kind[-1]), max_iterations=T)


def _theirs(ref, kind, code, T):
    if kind == "nnms":
        return ref.neural_minsum_decoder.NeuralMinSumDecoder(code, T)
    if kind == "noms":
        return ref.neural_minsum_decoder.NeuralOffsetMinSumDecoder(code, T)
    if kind.startswith("n2d"):
        return ref.neural_2d_decoder.Neural2DMinSumDecoder(code, int(kind[-1]), T)
    if kind.startswith("oms"):
        return ref.neural_2d_decoder.Neural2DOffsetMinSumDecoder(code, int(kind[-1]), T)
    return ref.rcq_decoder.WeightedRCQDecoder(code, 3, 8, QP, weight_sharing_type=int(kind[-1]), max_iterations=T)


KINDS = ["nnms", "noms", "n2d1", "n2d2", "n2d3", "n2d4", "oms2", "wrcq1", "wrcq2", "wrcq3", "wrcq4"]


@pytest.mark.parametrize("kind", KINDS)
def test_state_dict_round_trips_with_the_live_reference(built_lib, kind):
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("live reference only exists in the build container")
    ref = ref_shim.load()
    L = built_lib
    T = 4
    code, rcode = L.create_test_ldpc_code(), ref.ldpc_decoder.create_test_ldpc_code()
    torch.manual_seed(11)
    theirs = _theirs(ref, kind, rcode, T)
    torch.manual_seed(99)
    ours = _ours(L, kind, code, T)
    rs = theirs.state_dict()
    # same key set in the same order as the reference's own state_dict
    assert list(ours.state_dict().keys()) == list(rs.keys())
    assert all(v.shape == rs[k].shape and v.dtype == rs[k].dtype for k, v in ours.state_dict().items())
    # reference -> ours
    res = ours.load_state_dict(rs)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in ours.state_dict().items():
        assert torch.equal(v, rs[k]), k
    for name in ("beta_weights", "alpha_weights"):
        for k in getattr(ours, name, ()):
            assert float(getattr(ours, name)[k].detach()) == float(getattr(theirs, name)[k].detach())
    # ours -> reference (fresh values, so the load is visible)
    with torch.no_grad():
        for p in ours.parameters():
            p.copy_(torch.arange(p.numel(), dtype=torch.float32).reshape(p.shape) / 16 - 1)
    res = theirs.load_state_dict(ours.state_dict())
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in theirs.state_dict().items():
        assert torch.equal(v, ours.state_dict()[k]), k
    # named_parameters of both cover the same number of scalars
    assert sum(p.numel() for p in ours.parameters()) == sum(p.numel() for p in theirs.parameters())


def test_load_state_dict_reports_missing_and_unexpected_keys(built_lib):
    L = built_lib
    code = L.create_test_ldpc_code()
    d = L.Neural2DMinSumDecoder(code, 2, 3)
    sd = d.state_dict()
    assert "beta_weights.iter_0_dc3" in sd and "alpha_weights.iter_2_dv1" in sd and "_beta_table" not in sd
    partial = {k: v for k, v in sd.items() if k != "beta_weights.iter_1_dc4"}
    partial["beta_weights.iter_9_dc3"] = torch.zeros(1)
    with pytest.raises(RuntimeError) as e:
        d.load_state_dict(partial)
    assert "iter_1_dc4" in str(e.value) and "iter_9_dc3" in str(e.value)
    res = d.load_state_dict(partial, strict=False)
    assert res.missing_keys == ["beta_weights.iter_1_dc4"] and res.unexpected_keys == ["beta_weights.iter_9_dc3"]
    with pytest.raises(RuntimeError):
        d.load_state_dict({**sd, "beta_weights.iter_0_dc3": torch.zeros(2)})
    # the native dense layout still loads (checkpoints written before the hooks existed)
    d2 = L.Neural2DMinSumDecoder(code, 2, 3)
    d2.load_state_dict({"_beta_table": d._beta_table.detach().clone() + 1, "_alpha_table": d._alpha_table.detach().clone()})
    assert torch.equal(d2._beta_table, d._beta_table + 1)
    # works under a parent module's prefix and through deepcopy / .to()
    import copy
    parent = torch.nn.ModuleDict({"dec": d})
    assert "dec.beta_weights.iter_0_dc3" in parent.state_dict()
    parent2 = torch.nn.ModuleDict({"dec": L.Neural2DMinSumDecoder(code, 2, 3)})
    parent2.load_state_dict(parent.state_dict())
    assert torch.equal(parent2["dec"]._alpha_table, d._alpha_table)
    c = copy.deepcopy(d)
    assert list(c.state_dict().keys()) == list(sd.keys())


def test_weight_tables_follow_max_iterations_like_the_reference_keys(built_lib):
    L = built_lib
    code = L.create_test_ldpc_code()
    d = L.Neural2DMinSumDecoder(code, 2, 6)
    b, a = d._tables()
    assert b.shape == (6, 2) and a.shape == (6, 2)
    d.max_iterations = 4          # the reference reads it at call time: the first four iterations' keys
    b4, a4 = d._tables()
    assert np.array_equal(b4, b[:4]) and np.array_equal(a4, a[:4])
    d.max_iterations = 7          # the reference fails on the missing key iter_6_...
    with pytest.raises(KeyError):
        d._tables()


def test_analyze_weight_patterns_and_legacy_class(built_lib):
    L = built_lib
    from ldpc_b200 import ldpc_decoder as ours_ld
    from ldpc_b200.neural_minsum_decoder import analyze_weight_patterns
    code = L.create_test_ldpc_code()
    torch.manual_seed(5)
    legacy = ours_ld.NeuralMinSumDecoder(code, 3)
    assert len(legacy.beta_weights) == 39 and len(legacy.alpha_weights) == 0
    assert abs(float(legacy._beta_table.detach().mean())) < 0.1          # 0.1 * randn, no 0.7 shift (ldpc_decoder.py:173)
    torch.manual_seed(5)
    d = L.NeuralMinSumDecoder(code, 3)
    got = analyze_weight_patterns(d, code)
    assert set(got) == {"weight_statistics", "iteration_patterns", "node_degree_correlations"}
    assert sorted(got["iteration_patterns"]) == [0, 1, 2]
    assert sorted(got["node_degree_correlations"]) == ["check_degree_3", "check_degree_4"]
    assert got["node_degree_correlations"]["check_degree_3"]["count"] == 9
    from oracle import ref_shim
    if not ref_shim.available():
        return
    ref = ref_shim.load()
    rcode = ref.ldpc_decoder.create_test_ldpc_code()
    torch.manual_seed(5)
    rl = ref.ldpc_decoder.NeuralMinSumDecoder(rcode, 3)
    assert {k: float(v) for k, v in rl.state_dict().items()} == {k: float(v) for k, v in legacy.state_dict().items()}
    torch.manual_seed(5)
    r = ref.neural_minsum_decoder.NeuralMinSumDecoder(rcode, 3)
    want = ref.neural_minsum_decoder.analyze_weight_patterns(r, rcode)
    for t in range(3):
        for k in ("mean", "std", "min", "max"):
            assert got["iteration_patterns"][t][k] == pytest.approx(want["iteration_patterns"][t][k], rel=1e-12, abs=1e-15)
    for key, rec in want["node_degree_correlations"].items():
        assert got["node_degree_correlations"][key]["count"] == rec["count"]
        assert got["node_degree_correlations"][key]["mean"] == pytest.approx(rec["mean"], rel=1e-12)
        assert got["node_degree_correlations"][key]["std"] == pytest.approx(rec["std"], rel=1e-9)


def test_ldpccode_notices_a_new_H(built_lib):
    L = built_lib
    code = L.create_test_ldpc_code()
    assert code.graph.E == 13 and code.check_node_degrees[3] == 4
    H2 = np.array(code.H).copy()
    H2[3, 3] = 1
    code.H = H2
    assert code.graph.E == 14 and code.check_node_degrees[3] == 5
    code.H[0, 6] = 1               # in-place edit: needs invalidate()
    assert code.graph.E == 14
    code.invalidate()
    assert code.graph.E == 15 and code.variable_node_degrees[6] == 2


def test_the_recorded_repairs_make_the_reference_trainer_run(built_lib):
    """oracle/reference_training_repairs.patch (the oracle agreed for training): applied to the reference's
    training_framework.py in memory, one epoch of its PosteriorJointTrainer runs on the (7,4) code."""
    import os
    import sys
    import types
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("live reference only exists in the build container")
    ref = ref_shim.load()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(ref_shim.REFERENCE_ROOT, "training_framework.py"), newline="").read().split("\n")
    patch = open(os.path.join(root, "oracle", "reference_training_repairs.patch"), newline="").read().split("\n")
    out, pos, i = [], 0, 0
    while i < len(patch):
        line = patch[i]
        if line.startswith("@@"):
            start = int(line.split()[1].split(",")[0][1:]) - 1
            out += src[pos:start]
            pos = start
            i += 1
            while i < len(patch) and not patch[i].startswith("@@"):
                h = patch[i]
                if h.startswith("+"):
                    out.append(h[1:])
                elif h.startswith("-"):
                    assert src[pos] == h[1:], (src[pos], h)
                    pos += 1
                elif h.startswith(" "):
                    assert src[pos] == h[1:], (src[pos], h)
                    out.append(src[pos])
                    pos += 1
                i += 1
        else:
            i += 1
    out += src[pos:]
    mod = types.ModuleType("ref_training_repaired")
    saved = {k: sys.modules.get(k) for k in ("ldpc_decoder", "neural_2d_decoder", "rcq_decoder")}
    sys.modules.update(ldpc_decoder=ref.ldpc_decoder, neural_2d_decoder=ref.neural_2d_decoder, rcq_decoder=ref.rcq_decoder)
    try:
        exec(compile("\n".join(out).replace("\r", ""), "training_framework_repaired.py", "exec"), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    code = ref.ldpc_decoder.create_test_ldpc_code()
    torch.manual_seed(0)
    np.random.seed(0)
    model = ref.neural_2d_decoder.Neural2DMinSumDecoder(code, 2, 3)
    tr = mod.PosteriorJointTrainer(model, mod.TrainingConfig(batch_size=4, num_epochs=1))
    hist = tr.train(code, num_train_samples=8, num_val_samples=4)
    assert len(hist["train_losses"]) == 1 and np.isfinite(hist["train_losses"][0]) and hist["gradient_norms"][0] > 0
