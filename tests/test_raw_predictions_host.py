"""``get_raw_predictions[_batch]`` / ``predict_with_thresholds`` (standalone:1742-1881) against fixtures recorded from
the reference's own methods (tests/golden/make_golden_raw.py).  The forward is replaced by the logits the reference
forward produced for the same rows, as the reference's tests swap ``forward``."""

from __future__ import annotations

import json

import numpy as np
import pytest
import torch

from open_provence_b200.config import OpenProvenceConfig
from open_provence_b200.modeling import OpenProvenceModel, OpenProvenceOutput


@pytest.fixture(scope="module")
def golden(tiny_ckpt_dir):
    return json.loads((tiny_ckpt_dir.parent / "raw_tiny.json").read_text())


@pytest.fixture(scope="module")
def model(golden, tiny_ckpt_dir):
    from transformers import AutoTokenizer

    tok = AutoTokenizer.from_pretrained(str(tiny_ckpt_dir))
    m = OpenProvenceModel(OpenProvenceConfig.from_pretrained(tiny_ckpt_dir), None, tok, scorer=object())
    m.max_length = golden["max_length"]
    by_ids = {tuple(b["ids"]): b for b in golden["blocks"]}

    def recorded_forward(input_ids=None, attention_mask=None, **_kw):
        B, S = input_ids.shape
        rank = torch.zeros(B, 1)
        prune = torch.zeros(B, S, 2)
        for b in range(B):
            n = int(attention_mask[b].sum())
            rec = by_ids[tuple(int(t) for t in input_ids[b, :n])]  # KeyError = tokenisation differs from the reference's
            rank[b] = torch.tensor(rec["rank_logits"])
            prune[b, :n] = torch.tensor(rec["prune_logits"])
        return OpenProvenceOutput(ranking_logits=rank, pruning_logits=prune, logits=rank)

    m.forward = recorded_forward
    return m


def test_get_raw_predictions_batch(golden, model):
    raws = model.get_raw_predictions_batch(golden["batch_queries"], golden["batch_contexts"])
    assert len(raws) == len(golden["raw_batch"])
    sep = model.tokenizer.sep_token
    for got, want, q, ctx in zip(raws, golden["raw_batch"], golden["batch_queries"], golden["batch_contexts"]):
        assert got.ranking_score == pytest.approx(want["ranking_score"], abs=1e-6)
        assert [list(r) for r in got.context_ranges] == want["context_ranges"]
        # the reference returns the whole padded row; positions past the sequence are never read (the ranges stop
        # at the last valid token) and hold whatever its padded forward produced, so only valid tokens are compared
        n = len(model.tokenizer(q + sep + "".join(ctx), truncation=True, max_length=golden["max_length"])["input_ids"])
        assert got.pruning_probs.shape[0] == len(want["pruning_probs"])
        np.testing.assert_allclose(got.pruning_probs[:n], want["pruning_probs"][:n], atol=1e-6)
    single = model.get_raw_predictions(golden["query"], golden["contexts"])
    assert single.ranking_score == pytest.approx(golden["raw_single"]["ranking_score"], abs=1e-6)
    assert [list(r) for r in single.context_ranges] == golden["raw_single"]["context_ranges"]


@pytest.mark.parametrize("which", [0, 1])
def test_predict_with_thresholds(golden, model, which):
    want = golden["thresholds"][which]
    got = model.predict_with_thresholds(golden["query"], golden["contexts"], [0.05, 0.1, 0.5], use_majority=want["use_majority"])
    assert {str(k): v for k, v in got["predictions"].items()} == want["predictions"]
    assert [list(r) for r in got["context_ranges"]] == want["context_ranges"]
    assert got["ranking_score"] == pytest.approx(want["ranking_score"], abs=1e-6)


def test_mismatched_query_count_raises(model):
    with pytest.raises(ValueError, match="must match contexts_batch"):
        model.get_raw_predictions_batch(["a", "b"], [["x"]])
