"""Host side of the encoder.py APIs (predict / predict_with_pruning / predict_context / prune / prune_texts)
against fixtures produced by the reference's own ``OpenProvenceEncoder`` (tests/golden/make_golden_encoder.py).

The device stage is replaced at the ``packed_forward`` seam by the logits the reference forward recorded for
each pair; token probabilities come from the CPU oracle.  This pins tokenisation arguments, document-span
resolution, the token threshold, offset-based reconstruction and the chunk vote."""

from __future__ import annotations

import json

import numpy as np
import pytest

from open_provence_b200.config import OpenProvenceConfig
from open_provence_b200.encoder import OpenProvenceEncoder, evaluate_chunks, rebuild_document, resolve_document_span
from open_provence_b200.modeling import OpenProvenceModel
from oracle import postprocess_numpy as opp


@pytest.fixture(scope="module")
def golden(tiny_ckpt_dir):
    return json.loads((tiny_ckpt_dir.parent / "encoder_tiny.json").read_text())


@pytest.fixture(scope="module")
def encoder(golden, tiny_ckpt_dir):
    from transformers import AutoTokenizer

    by_ids = {tuple(e["ids"]): e for e in golden["examples"]}

    def recorded_forward(id_lists):
        ranks, probs = [], []
        for ids in id_lists:
            rec = by_ids[tuple(int(t) for t in ids)]  # KeyError = tokenisation differs from the reference's
            ranks.append(np.asarray(rec["rank_logits"], dtype=np.float32))
            probs.append(opp.keep_probs_from_logits(np.asarray(rec["prune_logits"], dtype=np.float32)))
        return np.stack(ranks), probs

    tok = AutoTokenizer.from_pretrained(str(tiny_ckpt_dir))
    model = OpenProvenceModel(OpenProvenceConfig.from_pretrained(tiny_ckpt_dir), None, tok, scorer=object())
    model.max_length = golden["max_length"]
    return OpenProvenceEncoder(model, packed_forward=recorded_forward)


def _pairs(golden):
    return [tuple(p) for p in golden["pairs"]]


def test_predict_scores(golden, encoder):
    scores = encoder.predict(_pairs(golden), batch_size=2)
    assert isinstance(scores, np.ndarray)
    np.testing.assert_allclose(scores, golden["predict"], rtol=0, atol=1e-6)
    single = encoder.predict(_pairs(golden)[0])
    np.testing.assert_allclose(single, golden["predict_single"], rtol=0, atol=1e-6)
    assert isinstance(encoder.predict(_pairs(golden), convert_to_numpy=False), list)


@pytest.mark.parametrize("which", [0, 1])
def test_predict_with_pruning_matches_reference(golden, encoder, which):
    case = golden["predict_with_pruning"][which]
    outs = encoder.predict_with_pruning(_pairs(golden), batch_size=2, pruning_threshold=case["threshold"], return_documents=True)
    assert len(outs) == len(case["outputs"])
    for got, ref in zip(outs, case["outputs"]):
        assert np.asarray(got.pruning_masks).astype(int).tolist() == ref["pruning_masks"]
        assert got.sentences == ref["tokens"]
        assert got.pruned_documents == ref["pruned_documents"]
        assert got.num_pruned_sentences == ref["num_pruned_sentences"]
        assert got.compression_ratio == pytest.approx(ref["compression_ratio"], abs=1e-12)
        np.testing.assert_allclose(got.ranking_scores, ref["ranking_scores"], atol=1e-6)


def test_single_pair_returns_one_output(golden, encoder):
    out = encoder.predict_with_pruning(_pairs(golden)[1], pruning_threshold=0.5, return_documents=True)
    ref = golden["predict_with_pruning_single"]
    assert out.pruned_documents == ref["pruned_documents"]
    assert out.compression_ratio == pytest.approx(ref["compression_ratio"], abs=1e-12)


def test_predict_context_matches_reference(golden, encoder):
    ref = golden["predict_context"]
    chunks = [[tuple(c) for c in ch] for ch in ref["chunks"]]
    outs = encoder.predict_context(_pairs(golden), chunks, batch_size=3, token_threshold=0.5, chunk_threshold=0.5)
    for got, want in zip(outs, ref["outputs"]):
        assert np.asarray(got.chunk_predictions).astype(int).tolist() == want["chunk_predictions"]
        np.testing.assert_allclose(got.chunk_scores, want["chunk_scores"], atol=1e-7)
        np.testing.assert_allclose(got.token_scores, want["token_scores"], atol=1e-7)
        assert got.compression_ratio == pytest.approx(want["compression_ratio"], abs=1e-12)
        assert got.ranking_scores == pytest.approx(want["ranking_scores"], abs=1e-6)


def test_prune_and_prune_texts(golden, encoder):
    q, d = _pairs(golden)[0]
    assert encoder.prune(q, d, threshold=0.5) == golden["prune"]["plain"]
    detail = encoder.prune(q, d, threshold=0.5, return_sentences=True)
    for key in ("pruned_document", "sentences", "pruning_masks", "num_pruned_sentences"):
        assert detail[key] == golden["prune"]["detail"][key]
    assert detail["compression_ratio"] == pytest.approx(golden["prune"]["detail"]["compression_ratio"], abs=1e-12)
    res = encoder.prune_texts([p[0] for p in _pairs(golden)], [p[1] for p in _pairs(golden)], threshold=0.5, batch_size=2)
    for got, want in zip(res, golden["prune_texts"]):
        assert got["pruned_text"] == want["pruned_text"]
        assert got["kept_ratio"] == pytest.approx(want["kept_ratio"], abs=1e-12)
    with_mask = encoder.prune_texts([q], [d], return_tokens=True)
    assert with_mask[0]["pruning_mask"].dtype == bool


def test_span_resolution_fallbacks():
    offsets = [(0, 0), (0, 3), (0, 0), (0, 2), (2, 5), (0, 0)]
    ids = [1, 10, 2, 11, 12, 2]
    assert resolve_document_span(ids, offsets, [0, 0, 0, 1, 1, 1], [1, 0, 1, 0, 0, 1], [2]) == (3, 5)  # segment ids
    assert resolve_document_span(ids, offsets, None, [1, 0, 1, 0, 0, 1], [2]) == (3, 5)               # separators
    assert resolve_document_span(ids, offsets, None, None, []) == (1, 5)                               # non-special run
    assert resolve_document_span([1, 2], [(0, 0), (0, 0)], None, None, []) is None


def test_chunk_vote_and_reconstruction_edge_cases():
    offsets = np.array([[0, 3], [3, 6], [7, 9], [0, 0]])
    probs = np.array([0.9, 0.6, 0.1, 0.99], dtype=np.float32)
    scores, preds = evaluate_chunks([(0, 6), (6, 9), (20, 30)], probs, offsets, 0.5, 0.5)
    # the token starting at offset 0 is skipped by the reference's `start != 0 and end != 0` test
    assert scores[0] == pytest.approx(float(np.float32(0.6))) and preds.tolist() == [1, 0, 0] and scores[2] == 0.0
    assert rebuild_document("abcdefghij", np.array([1, 1, 0, 1], bool), offsets) == "abcdef"
    assert rebuild_document("abcdefghij", np.array([1, 0, 1, 0], bool), offsets) == "abc hi"
    assert rebuild_document("abc", np.zeros(4, bool), offsets) == ""
