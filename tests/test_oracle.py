"""The oracle (oracle/spectral_oracle.py) against the committed outputs of the
REAL reference (tests/golden/, made by oracle/gen_golden.py) and against the
reference's own known-answer tests (/root/reference/tests/test_metrics.py)."""

import json
import math
from pathlib import Path

import numpy as np
import pytest
import spectral_oracle as orc
import torch
from _inputs import build_case, checksum, golden_case_names
from _vit_stub import StubViT, WrappedViT

GOLD = Path(__file__).parent / "golden"
RECORDS = json.loads((GOLD / "metrics_golden.json").read_text())
SVS = np.load(GOLD / "sv_golden.npz")
MODELS = json.loads((GOLD / "model_golden.json").read_text())


def same(a, b, rtol=1e-12, atol=1e-14):
    if a is None or b is None:
        return a is b
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    return abs(a - b) <= atol + rtol * abs(b)


@pytest.mark.parametrize("name", golden_case_names())
def test_case_matches_reference(name):
    rec = RECORDS[name]
    w = build_case(name)
    assert checksum(w) == rec["input_crc32"], "input generator drifted; regenerate golden"
    got = orc.get_spectral_metrics(w)
    assert list(got) == list(orc.METRIC_KEYS)
    for k in orc.METRIC_KEYS:
        assert same(got[k], rec["metrics"][k]), (k, got[k], rec["metrics"][k])
    assert orc.integer_outputs(w) == rec["ints"]
    if rec["sv_len"] >= 0:
        s = orc.singular_values(w)
        np.testing.assert_allclose(s, SVS[name], rtol=1e-12, atol=1e-300)
    if "alpha_fit_range_2_12" in rec:
        assert same(orc.alpha_exponent(np.asarray(w, np.float64), fit_range=(2, 12)), rec["alpha_fit_range_2_12"])
        assert same(orc.power_law_alpha_hill(np.asarray(w, np.float64), k=7), rec["hill_k7"])


def test_aggregate_matches_reference():
    g = json.loads((GOLD / "aggregate_golden.json").read_text())
    out = orc.aggregate_spectral_metrics(g["input"])
    assert list(out) == list(g["output"])
    for k, v in g["output"].items():
        assert same(out[k], v)
    assert orc.aggregate_spectral_metrics([]) == g["empty"] == {}


def _build(tag):
    if tag == "E_seed42":
        return StubViT(embed_dim=32, depth=1, seed=42)
    if tag == "C_seed142":
        return StubViT(embed_dim=96, depth=3, seed=142)
    if tag == "E_wrapped_seed7":
        return WrappedViT(embed_dim=32, depth=2, seed=7)
    if tag == "E_sepqkv_seed3":
        return StubViT(embed_dim=32, depth=1, seed=3, separate_qkv=True)
    raise KeyError(tag)


@pytest.mark.parametrize("tag", list(MODELS))
def test_model_level_matches_reference(tag):
    gold = MODELS[tag]
    model = _build(tag)
    crc = checksum(np.concatenate([p.detach().numpy().ravel() for p in model.parameters()]))
    assert crc == gold["params_crc32"], "torch CPU generator drifted; regenerate golden"
    res = orc.extract_and_analyze_weights(model, torch.device("cpu"))
    assert list(res["per_layer_metrics"]) == list(gold["analysis"]["per_layer_metrics"])
    for name, m in gold["analysis"]["per_layer_metrics"].items():
        for k, v in m.items():
            assert same(res["per_layer_metrics"][name][k], v)
    for k, v in gold["analysis"]["aggregated_metrics"].items():
        assert same(res["aggregated_metrics"][k], v)
    for name, s in gold["analysis"]["singular_values"].items():
        np.testing.assert_allclose(res["singular_values"][name], s, rtol=1e-12)
    tm = orc.compute_spectral_metrics_trainer(model, ["blocks.0"], True, True, True)
    assert list(tm) == list(gold["trainer_metrics"])
    for k, v in gold["trainer_metrics"].items():
        assert same(tm[k], v)


# ---- the reference's own known-answer tests, restated (tests/test_metrics.py) ----
def test_identity_entropy_and_rank():  # :12-25, :75-84
    assert np.isclose(orc.spectral_entropy(np.eye(10)), np.log(10), rtol=1e-4)
    assert np.isclose(orc.stable_rank(np.eye(10)), 10.0, rtol=1e-4)


def test_rank_one():  # :27-39, :86-97
    u = np.random.default_rng(0).standard_normal((10, 1))
    assert orc.spectral_entropy(u @ u.T) < 0.5
    assert np.isclose(orc.stable_rank(u @ u.T), 1.0, rtol=1e-4)


def test_bounds_and_powerlaw():  # :99-107, :113-146
    w = np.random.default_rng(1).standard_normal((30, 50))
    assert 1.0 <= orc.stable_rank(w) <= 30
    assert abs(orc.alpha_exponent(build_case("powerlaw:100:2.0:f64")) - 2.0) < 1.0
    assert abs(orc.alpha_exponent(np.eye(50))) < 1.0


def test_hill_and_small():  # :148-183
    np.random.seed(42)
    a = orc.power_law_alpha_hill(np.random.randn(100, 100))
    assert np.isfinite(a) and a > 0
    assert np.isnan(orc.alpha_exponent(np.random.randn(4, 4)))
    assert np.isnan(orc.power_law_alpha_hill(np.random.randn(5, 5)))


def test_contract():  # :189-220, :63-69, :347-353
    m = orc.get_spectral_metrics(torch.randn(32, 32))
    assert set(m) == set(orc.METRIC_KEYS) and np.isfinite(m["spectral_entropy"])
    assert np.isnan(orc.spectral_entropy(np.random.randn(10)))
    assert orc.get_spectral_distribution(np.random.randn(10)) is None


def test_distribution_invariants():  # :321-345
    np.random.seed(42)
    d = orc.get_spectral_distribution(np.random.randn(64, 64), name="t", matrix_type="x")
    assert np.all(np.diff(d.singular_values) <= 0) and np.all(d.normalized_sv <= 1.0)
    assert np.all(np.diff(d.cumulative_variance) >= 0) and np.isclose(d.cumulative_variance[-1], 1.0)


def test_lowrank_restatements_match_the_reference_goldens():
    """oracle restatements of tail_truncation.py:63-152 and gradient_alignment.py:48-70 against outputs of the real
    reference (tests/golden/lowrank_golden.npz, oracle/gen_golden_lowrank.py)."""
    from pathlib import Path

    from _inputs import build_case

    gold = np.load(Path(__file__).parent / "golden" / "lowrank_golden.npz")
    names = sorted({k.split("/")[0] for k in gold.files})
    assert len(names) == 6
    for key in names:
        case = {"vit_C_0_q": "vit:C:0:q", "vit_C_0_mlp_up": "vit:C:0:mlp_up", "vit_E_0_mlp_down": "vit:E:0:mlp_down",
                "randn_30x50_f64": "randn:30x50:f64", "sgd_96x384": "sgd:96x384", "powerlaw_64_0.5_f32": "powerlaw:64:0.5:f32"}[key]
        w = build_case(case)
        t90, i90 = orc.truncate_weight_matrix(w, 0.9)
        t50, i50 = orc.truncate_weight_matrix(w, 0.5)
        te, ie = orc.truncate_by_energy(w, 0.95)
        np.testing.assert_allclose(t90, gold[f"{key}/trunc90"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(t50, gold[f"{key}/trunc50"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(te, gold[f"{key}/energy95"], rtol=0, atol=1e-12)
        info = gold[f"{key}/info"]
        assert (i90["original_rank"], i90["truncated_rank"], i50["truncated_rank"], ie["truncated_rank"]) == tuple(int(v) for v in info[[0, 1, 3, 5]])
        np.testing.assert_allclose([i90["energy_retained"], i50["energy_retained"], ie["energy_retained"]], info[[2, 4, 6]], rtol=1e-12)
        np.testing.assert_allclose(orc.compute_rank_reducing_gradient(w), gold[f"{key}/polar"], rtol=0, atol=1e-10)
