"""Pin the numpy oracle against outputs of the reference itself (tests/golden/make_golden.py)."""

from __future__ import annotations

import numpy as np

from oracle import modernbert_numpy as onp
from oracle import postprocess_numpy as opp
from open_provence_b200 import synthetic as syn


def _unpadded(golden):
    return [golden["input_ids"][b, :n].tolist() for b, n in enumerate(golden["lengths"].tolist())]


def test_oracle_fp64_matches_reference_fp64(forward_golden, tiny_weights, tiny_config):
    cfg = tiny_config["base_model_config"]
    w64 = onp.cast_weights(tiny_weights, np.float64)
    rank, prunes = onp.forward_batch(_unpadded(forward_golden), w64, cfg)
    # HF forces fp32 for the RoPE angles (HF:162-172); everything else is fp64 on both sides.
    assert np.abs(rank - forward_golden["ranking_logits_f64"]).max() < 1e-9
    for b, pr in enumerate(prunes):
        n = pr.shape[0]
        assert np.abs(pr - forward_golden["pruning_logits_f64"][b, :n]).max() < 1e-9


def test_oracle_fp32_matches_reference_fp32(forward_golden, tiny_weights, tiny_config):
    cfg = tiny_config["base_model_config"]
    rank, prunes = onp.forward_batch(_unpadded(forward_golden), onp.cast_weights(tiny_weights, np.float32), cfg)
    assert rank.dtype == np.float32
    # tolerance 2e-5: two fp32 evaluation orders of a 4-layer model (reference fp32 is itself 6e-6 off fp64)
    assert np.abs(rank - forward_golden["ranking_logits_f32"]).max() < 2e-5
    for b, pr in enumerate(prunes):
        n = pr.shape[0]
        assert np.abs(pr - forward_golden["pruning_logits_f32"][b, :n]).max() < 2e-5


def test_unpadded_equals_padded_on_valid_tokens(forward_golden):
    """The fixture was produced from a right-padded batch; row 0 has a single token."""
    assert forward_golden["lengths"].min() == 1
    assert forward_golden["attention_mask"].sum() == forward_golden["lengths"].sum()


def test_score_conversion_matches_recorded_blocks(process_golden):
    """sigmoid / 2-way softmax restatement vs what the reference produced for str_str."""
    case = next(c for c in process_golden["cases"] if c["name"] == "str_str")
    block = case["blocks"][0]
    score = opp.ranking_score_from_logits(np.asarray(block["rank_logits"], dtype=np.float32))
    assert abs(score - case["result"]["reranking_score"]) < 1e-7
    probs = opp.keep_probs_from_logits(np.asarray(block["prune_logits"], dtype=np.float32))
    assert probs.dtype == np.float32 and 0.0 <= probs.min() and probs.max() <= 1.0


def test_package_flops_formula_equals_oracle():
    for name, S in [("base-130M", 2048), ("xsmall-30M", 512), ("large-310M", 4096), ("en-gte-149M", 8192)]:
        cfg = syn.backbone_config(name)
        assert syn.algorithmic_flops_per_pair(cfg, S) == onp.algorithmic_flops_per_pair(cfg, S)


def test_flops_formula_matches_baseline_table():
    base = dict(hidden_size=512, num_hidden_layers=19, intermediate_size=2048, num_labels=1, local_attention=128)
    assert abs(onp.algorithmic_flops_per_pair(base, 2048) / 1e9 - 392.94) < 0.01
    xsmall = dict(hidden_size=256, num_hidden_layers=10, intermediate_size=1024, num_labels=1, local_attention=128)
    assert abs(onp.algorithmic_flops_per_pair(xsmall, 512) / 1e9 - 12.19) < 0.01
    large = dict(hidden_size=768, num_hidden_layers=25, intermediate_size=3072, num_labels=1, local_attention=128)
    assert abs(onp.algorithmic_flops_per_pair(large, 4096) / 1e9 - 2422.37) < 0.01


def test_oracle_mean_pooling_matches_reference_fp64(tiny_weights, tiny_config):
    """classifier_pooling = "mean" (HF:623-630) against the reference run with that backbone config
    (tests/golden/make_golden_mean.py)."""
    from pathlib import Path

    data = np.load(Path(__file__).resolve().parent / "golden" / "forward_tiny_mean.npz")
    cfg = dict(tiny_config["base_model_config"], classifier_pooling="mean")
    cu = np.concatenate([[0], np.cumsum(data["lengths"])])
    seqs = [data["input_ids"][cu[i] : cu[i + 1]].tolist() for i in range(len(data["lengths"]))]
    rank, prunes = onp.forward_batch(seqs, onp.cast_weights(tiny_weights, np.float64), cfg)
    assert np.abs(rank - data["ranking_logits_f64"]).max() < 1e-9
    # positions up to 1023: HF's fp32 RoPE angles (HF:162-172) leave 2e-9 on the token logits
    assert np.abs(np.concatenate(prunes) - data["pruning_logits_f64"]).max() < 1e-8
    cls_rank, _ = onp.forward_batch(seqs[-2:], onp.cast_weights(tiny_weights, np.float64), tiny_config["base_model_config"])
    assert np.abs(cls_rank - data["ranking_logits_f64"][-2:]).max() > 1e-3  # the fixture does distinguish the poolings
