"""Host-side logic of the drop-in layer that needs no GPU: extraction rules, aggregate,
distribution arrays, tracker (de)serialisation, MLflow logging contract, artifact JSON,
arena layouts and work partitioning."""

import json
import math
from pathlib import Path

import numpy as np
import pytest
import spectral_oracle as orc
import torch
from _vit_stub import StubViT, WrappedViT

from vision_spectra_b200.experiments.run_spectral_analysis import log_spectral_metrics, write_spectral_artifacts
from vision_spectra_b200.metrics import (
    SpectralDistribution,
    SpectralTracker,
    aggregate_spectral_metrics,
    extract_all_weights,
    extract_attention_weights,
    extract_mlp_weights,
    extract_patch_embed_weights,
    extract_qkv_weights,
    get_spectral_distribution,
    get_spectral_metrics,
)
from vision_spectra_b200.metrics.spectral import EpochSpectralSnapshot, distribution_from_sv
from vision_spectra_b200.sweep import CheckpointLayout, matrix_cost, partition_lpt, shard_checkpoints

GOLD = Path(__file__).parent / "golden"
MODELS = json.loads((GOLD / "model_golden.json").read_text())


def _same(a, b):
    return (math.isnan(a) and math.isnan(b)) or abs(a - b) <= 1e-12 * max(1.0, abs(b))


@pytest.mark.parametrize("make", [
    lambda: StubViT(embed_dim=32, depth=2, seed=1),
    lambda: WrappedViT(embed_dim=32, depth=2, seed=2),
    lambda: StubViT(embed_dim=32, depth=1, seed=3, separate_qkv=True),
])
def test_extraction_matches_reference_rules(make):
    """Same names, order, types, layer indices and values as the oracle's restatement of
    extraction.py (itself pinned to the real reference through model_golden.json)."""
    model = make()
    for ours, theirs in (
        (extract_qkv_weights(model), orc.extract_qkv_weights(model)),
        (extract_attention_weights(model), orc.extract_attention_weights(model)),
        (extract_mlp_weights(model), orc.extract_mlp_weights(model)),
        (extract_patch_embed_weights(model), orc.extract_patch_embed_weights(model)),
        (extract_all_weights(model, ["blocks.0"], True, True, True, True), orc.extract_all_weights(model, ["blocks.0"], True, True, True, True)),
        (extract_all_weights(model), orc.extract_all_weights(model)),
    ):
        assert [(w.name, w.layer_idx, w.matrix_type, tuple(w.shape)) for w in ours] == [
            (w.name, w.layer_idx, w.matrix_type, tuple(w.shape)) for w in theirs
        ]
        for a, b in zip(ours, theirs):
            np.testing.assert_array_equal(a.numpy(), b.weight)


def test_qkv_are_views_of_the_fused_buffer():
    model = StubViT(embed_dim=32, depth=1, seed=0)
    q, k, v = extract_qkv_weights(model)
    base = model.blocks[0].attn.qkv.weight
    assert q.weight.data_ptr() == base.data_ptr()
    assert k.weight.data_ptr() == base.data_ptr() + 32 * 32 * 4
    assert v.weight.data_ptr() == base.data_ptr() + 2 * 32 * 32 * 4
    assert q.name == "blocks.0.attn.qkv.q" and (q.matrix_type, k.matrix_type, v.matrix_type) == ("q", "k", "v")


def test_layer_pattern_is_substring_match():
    model = StubViT(embed_dim=32, depth=11, seed=0)
    names = {w.name for w in extract_qkv_weights(model, ["blocks.1"])}
    assert any(n.startswith("blocks.10.") for n in names) and any(n.startswith("blocks.1.") for n in names)
    assert not any(n.startswith("blocks.2.") for n in names)


def test_aggregate_matches_reference():
    g = json.loads((GOLD / "aggregate_golden.json").read_text())
    out = aggregate_spectral_metrics(g["input"])
    assert list(out) == list(g["output"])
    for k, v in g["output"].items():
        assert _same(out[k], v)
    assert aggregate_spectral_metrics([]) == {}


def test_distribution_arrays_match_reference():
    g = MODELS["C_seed142"]["dist0"]
    full_sv = np.array(MODELS["C_seed142"]["analysis"]["singular_values"][g["name"]])
    d = distribution_from_sv(full_sv, g["metrics"], g["name"], g["matrix_type"])
    k = len(g["singular_values"])  # the tracker truncated to max_singular_values
    np.testing.assert_allclose(d.singular_values[:k], g["singular_values"], rtol=1e-13)
    np.testing.assert_allclose(d.eigenvalues[:k], g["eigenvalues"], rtol=1e-13)
    np.testing.assert_allclose(d.normalized_sv[:k], g["normalized_sv"], rtol=1e-13)
    np.testing.assert_allclose(d.cumulative_variance[:k], g["cumulative_variance"], rtol=1e-13)
    assert distribution_from_sv(None, {}) is None
    assert get_spectral_distribution(np.zeros(7)) is None  # non-2-D -> None (spectral.py:532)
    assert all(math.isnan(v) for v in get_spectral_metrics(np.zeros(7)).values())


def test_tracker_roundtrip_and_history(tmp_path):
    t = SpectralTracker(layer_patterns=["blocks.0"], include_mlp=True, max_singular_values=5)
    for epoch, scale in ((0, 1.0), (5, 2.0)):
        sv = np.array([3.0, 2.0, 1.0]) * scale
        dist = distribution_from_sv(sv, {"stable_rank": 14 / 9, "alpha_exponent": float("nan")}, "blocks.0.attn.qkv.q", "q")
        t.history.append(EpochSpectralSnapshot(epoch, [dist], {"stable_rank_mean": 14 / 9, "alpha_exponent_mean": float("nan")}))
    assert t.get_metric_history("stable_rank_mean") == ([0, 5], [14 / 9, 14 / 9])
    assert t.get_metric_history("alpha_exponent_mean") == ([], [])  # NaN filtered (spectral.py:724)
    assert t.get_all_layer_names() == ["blocks.0.attn.qkv.q"]
    ep, svs = t.get_layer_sv_history("blocks.0.attn.qkv.q")
    assert ep == [0, 5] and np.allclose(svs[1], [6, 4, 2])
    d = t.to_dict()
    assert list(d) == ["layer_patterns", "include_qkv", "include_mlp", "include_patch_embed", "max_singular_values", "history"]
    assert list(d["history"][0]) == ["epoch", "timestamp", "aggregated_metrics", "distributions"]
    assert list(d["history"][0]["distributions"][0]) == ["name", "matrix_type", "singular_values", "metrics"]
    t.save(tmp_path / "x" / "tracker.json")
    t2 = SpectralTracker.load(tmp_path / "x" / "tracker.json")
    assert t2.max_singular_values == 5 and t2.include_mlp and len(t2.history) == 2
    np.testing.assert_allclose(t2.history[1].distributions[0].singular_values, [6, 4, 2])
    np.testing.assert_allclose(t2.history[1].distributions[0].eigenvalues, [36, 16, 4])


class _Recorder:
    def __init__(self):
        self.metrics, self.artifacts = [], []

    def log_metric(self, key, value, step=None):
        self.metrics.append((key, value, step))

    def log_artifact(self, path, artifact_path=None):
        self.artifacts.append((Path(path).name, artifact_path))


def test_mlflow_contract_and_artifacts(tmp_path):
    """Keys `spectral/{metric}_{mean,std}`, step = epoch, finite values only
    (run_spectral_analysis.py:511-513); artifacts under spectral/epoch_{N} with NaN -> null."""
    analysis = {
        "per_layer_metrics": {"blocks.0.attn.qkv.q": {"spectral_entropy": 1.5, "alpha_exponent": float("nan")}},
        "aggregated_metrics": {"spectral_entropy_mean": 1.5, "spectral_entropy_std": 0.0, "alpha_exponent_mean": float("nan"), "alpha_exponent_std": float("nan")},
        "singular_values": {"blocks.0.attn.qkv.q": [2.0, 1.0]},
    }
    rec = _Recorder()
    assert log_spectral_metrics(rec, analysis, epoch=29) == 2
    assert rec.metrics == [("spectral/spectral_entropy_mean", 1.5, 29), ("spectral/spectral_entropy_std", 0.0, 29)]
    d = write_spectral_artifacts(analysis, 29, tmp_path, rec)
    assert d.name == "epoch_29"
    assert json.loads((d / "singular_values.json").read_text()) == analysis["singular_values"]
    assert json.loads((d / "layer_metrics.json").read_text()) == {"blocks.0.attn.qkv.q": {"spectral_entropy": 1.5, "alpha_exponent": None}}
    assert rec.artifacts == [("singular_values.json", "spectral/epoch_29"), ("layer_metrics.json", "spectral/epoch_29")]


def test_layout_order_matches_driver_extraction_order():
    """CheckpointLayout.vit lists slots in the order extract_and_analyze_weights visits them."""
    model = StubViT(embed_dim=32, depth=3, seed=0)
    ours = [(s.name, s.matrix_type, s.layer_idx, (s.rows, s.cols)) for s in CheckpointLayout.vit(32, 3).slots]
    ref = orc.extract_qkv_weights(model) + orc.extract_attention_weights(model) + orc.extract_mlp_weights(model)
    assert ours == [(w.name, w.matrix_type, w.layer_idx, tuple(w.shape)) for w in ref]
    lay = CheckpointLayout.vit(192, 6)
    assert lay.matrices == 36 and lay.bytes == 10_616_832
    arena = torch.arange(lay.arena_elems, dtype=torch.float32)
    views = lay.views(arena)
    assert views[0].data_ptr() == arena.data_ptr() and views[1].data_ptr() == arena.data_ptr() + 192 * 192 * 4
    assert all(v.shape == (s.rows, s.cols) for v, s in zip(views, lay.slots))


def test_partition_lpt_and_round_robin():
    shapes = [(192, 192)] * 24 + [(768, 192)] * 6 + [(192, 768)] * 6
    costs = [matrix_cost(r, c) for r, c in shapes]
    groups = [i // 3 if i < 18 else 100 + i for i in range(36)]  # q/k/v triples stay together
    for world in (1, 2, 4, 8):
        shards = partition_lpt(costs, world, groups)
        assert sorted(i for s in shards for i in s) == list(range(36))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) <= 1.35 * (sum(costs) / world)
        for s in shards:
            for g in range(6):
                members = [i for i in range(18) if i // 3 == g]
                assert all(i in s for i in members) or not any(i in s for i in members)
    assert shard_checkpoints(10, 4, 1) == [1, 5, 9]
    assert sorted(sum((shard_checkpoints(93, 8, r) for r in range(8)), [])) == list(range(93))


@pytest.mark.parametrize("make", [
    lambda: StubViT(embed_dim=32, depth=2, seed=1),
    lambda: WrappedViT(embed_dim=32, depth=2, seed=2),
    lambda: StubViT(embed_dim=32, depth=1, seed=3, separate_qkv=True),
])
def test_state_dict_selection_matches_module_extraction(make, tmp_path):
    """checkpoint.select_matrices on a saved state dict == the reference's extraction rules on
    the live model (names, order, types, layer indices, values), incl. the trainer's union."""
    from vision_spectra_b200.checkpoint import load_checkpoint, select_matrices

    model = make()
    path = tmp_path / "ckpt.pt"
    torch.save({"epoch": 3, "model_state_dict": model.state_dict(), "best_val_metric": 0.5}, path)  # base.py:576-594
    sd = load_checkpoint(path)
    ours = select_matrices(sd)
    ref = orc.extract_qkv_weights(model) + orc.extract_attention_weights(model) + orc.extract_mlp_weights(model)
    assert [(w.name, w.layer_idx, w.matrix_type, tuple(w.shape)) for w in ours] == [
        (w.name, w.layer_idx, w.matrix_type, tuple(w.shape)) for w in ref
    ]
    for a, b in zip(ours, ref):
        np.testing.assert_array_equal(a.numpy(), b.weight)
    ours2 = select_matrices(sd, layer_patterns=["blocks.0"], include_mlp=False, include_patch_embed=True)
    ref2 = orc.extract_all_weights(model, ["blocks.0"], True, True, False, True)
    assert [(w.name, w.matrix_type, tuple(w.shape)) for w in ours2] == [(w.name, w.matrix_type, tuple(w.shape)) for w in ref2]


def test_trainer_epoch_artifact_json_layout(tmp_path):
    """training/base.py:453-511 (JSON half): file name, directory, keys and nesting of spectral_epoch_%04d.json, and
    the mlflow.log_artifact call with artifact_path 'spectral/json'."""
    import json

    import numpy as np
    from vision_spectra_b200.metrics.spectral import EpochSpectralSnapshot, SpectralDistribution
    from vision_spectra_b200.training.base import save_epoch_spectral_artifacts

    d = SpectralDistribution("blocks.0.attn.qkv.q", "q", np.array([3.0, 2.0, 1.0]), np.array([9.0, 4.0, 1.0]),
                             np.array([1.0, 2 / 3, 1 / 3]), np.array([9 / 14, 13 / 14, 1.0]),
                             {"spectral_entropy": 0.9, "stable_rank": 1.5, "alpha_exponent": 1.0, "pl_alpha_hill": 2.0})
    snap = EpochSpectralSnapshot(epoch=7, distributions=[d], aggregated_metrics={"stable_rank_mean": 1.5})
    calls = []

    class FakeMlflow:
        @staticmethod
        def log_artifact(path, artifact_path=None):
            calls.append((path, artifact_path))

    out = save_epoch_spectral_artifacts(snap, 7, tmp_path, FakeMlflow)
    assert out == tmp_path / "spectral" / "json" / "spectral_epoch_0007.json"
    data = json.loads(out.read_text())
    assert list(data) == ["epoch", "timestamp", "aggregated_metrics", "distributions"] and data["epoch"] == 7
    assert list(data["distributions"][0]) == ["name", "matrix_type", "singular_values", "metrics"]
    assert data["distributions"][0]["singular_values"] == [3.0, 2.0, 1.0]
    assert calls == [(str(out), "spectral/json")]
