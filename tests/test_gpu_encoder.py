"""encoder.py APIs end to end on the GPU (tokenizer -> engine -> token-level pruning) against the fixtures
recorded from the reference's ``OpenProvenceEncoder``, plus the save_pretrained / from_pretrained round trip."""

from __future__ import annotations

import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from open_provence_b200.encoder import OpenProvenceEncoder  # noqa: E402
from open_provence_b200.modeling import OpenProvenceModel  # noqa: E402


@pytest.fixture(scope="module")
def golden(tiny_ckpt_dir):
    return json.loads((tiny_ckpt_dir.parent / "encoder_tiny.json").read_text())


@pytest.fixture(scope="module")
def encoder(tiny_ckpt_dir, golden):
    enc = OpenProvenceEncoder.from_pretrained(tiny_ckpt_dir, device="cuda", dtype="fp32")
    enc.max_length = enc._model.max_length = golden["max_length"]
    return enc


def _pairs(golden):
    return [tuple(p) for p in golden["pairs"]]


def test_predict_and_token_scores_match_reference_fp32(golden, encoder):
    np.testing.assert_allclose(encoder.predict(_pairs(golden), batch_size=2), golden["predict"], atol=1e-5)
    ref = golden["predict_context"]
    chunks = [[tuple(c) for c in ch] for ch in ref["chunks"]]
    outs = encoder.predict_context(_pairs(golden), chunks, batch_size=3)
    for got, want in zip(outs, ref["outputs"]):
        np.testing.assert_allclose(got.token_scores, want["token_scores"], atol=1e-5)
        np.testing.assert_allclose(got.chunk_scores, want["chunk_scores"], atol=1e-5)
        assert np.asarray(got.chunk_predictions).astype(int).tolist() == want["chunk_predictions"]


@pytest.mark.parametrize("which", [0, 1])
def test_token_masks_and_documents_identical(golden, encoder, which):
    case = golden["predict_with_pruning"][which]
    token_scores = [np.asarray(o["token_scores"]) for o in golden["predict_context"]["outputs"]]
    outs = encoder.predict_with_pruning(_pairs(golden), pruning_threshold=case["threshold"], return_documents=True)
    for got, ref, probs in zip(outs, case["outputs"], token_scores):
        # a token whose reference probability sits within 1e-5 of the threshold may legitimately flip
        assert probs.size == 0 or np.abs(probs - case["threshold"]).min() > 1e-5, "fixture too close to the threshold"
        assert np.asarray(got.pruning_masks).astype(int).tolist() == ref["pruning_masks"]
        assert got.pruned_documents == ref["pruned_documents"]
        assert got.sentences == ref["tokens"]


def test_save_pretrained_round_trip(tmp_path, golden, encoder, tiny_ckpt_dir):
    encoder.save_pretrained(tmp_path / "ckpt")
    again = OpenProvenceEncoder.from_pretrained(tmp_path / "ckpt", device="cuda", dtype="fp32")
    assert again.max_length == golden["max_length"]  # written into config.json
    a = encoder.predict_with_pruning(_pairs(golden), return_documents=True)
    b = again.predict_with_pruning(_pairs(golden), return_documents=True)
    for x, y in zip(a, b):
        assert np.array_equal(x.pruning_masks, y.pruning_masks) and x.pruned_documents == y.pruned_documents
        assert np.array_equal(x.ranking_scores, y.ranking_scores)
    # the same directory is a valid checkpoint for the process() class as well
    model = OpenProvenceModel.from_pretrained(tmp_path / "ckpt", device="cuda", dtype="fp32")
    from open_provence_b200.host_text import simple_sentence_splitter

    out = model.process(question="What are bananas?", context=golden["pairs"][1][1], threshold=0.1, show_progress=False,
                        sentence_splitter=simple_sentence_splitter)
    assert set(out) >= {"pruned_context", "reranking_score", "compression_rate"}


def test_bf16_encoder_close(golden, tiny_ckpt_dir):
    enc = OpenProvenceEncoder.from_pretrained(tiny_ckpt_dir, device="cuda")  # bf16 engine (default)
    enc.max_length = enc._model.max_length = golden["max_length"]
    scores = enc.predict(_pairs(golden))
    np.testing.assert_allclose(scores, golden["predict"], atol=2e-2)
