"""Host side of ``process()`` against fixtures produced by the reference's own ``process()``.

The device stage is replaced at the scorer seam by the logits the reference recorded for each block
(the reference's tests swap ``OpenProvenceModel.forward`` the same way,
/root/reference/tests/test_modeling_open_provence.py:934-1063), evaluated with the CPU oracle.  This pins
input normalisation, titles, sentence normalisation, fragmentising, block assembly, the range tables and
the string post-processing; the CUDA kernels are pinned separately by the ``-m gpu`` tests.
"""

from __future__ import annotations

import numpy as np
import pytest

from open_provence_b200.config import OpenProvenceConfig
from open_provence_b200.host_text import simple_sentence_splitter
from open_provence_b200.modeling import OpenProvenceModel
from oracle import postprocess_numpy as opp

CASES = [
    "str_str", "str_list", "aligned", "nested", "presplit_sentences", "explicit_titles", "first_line_title",
    "title_none", "multi_block", "multi_block_respect", "overlong_sentence", "strip_sentences", "reorder_topk",
    "no_best_score", "japanese", "empty_context", "batch_size_1",
]


class RecordedLogitScorer:
    """Looks every block up in the logits the reference recorded, then applies the oracle's arithmetic."""

    def __init__(self, recorded_blocks):
        self.by_ids = {tuple(b["ids"]): b for b in recorded_blocks}
        self.seen: list[tuple[int, ...]] = []

    def run(self, table, threshold):
        rank_score = np.zeros(table.n_blocks, dtype=np.float32)
        probs = []
        for b, ids in enumerate(table.block_ids):
            key = tuple(int(t) for t in ids)
            self.seen.append(key)
            assert key in self.by_ids, f"block {b} was never fed to the reference forward: {key[:12]}..."
            rec = self.by_ids[key]
            rank_score[b] = opp.ranking_score_from_logits(np.asarray(rec["rank_logits"], dtype=np.float32))
            probs.append(opp.keep_probs_from_logits(np.asarray(rec["prune_logits"], dtype=np.float32)))
        frag_mean = []
        for blk, (start, end) in zip(table.frag_block, table.frag_local):
            frag_mean.append(1.0 if end <= start else float(probs[blk][start:end].mean()))
        sent_prob, keep = [], []
        for s in range(table.n_sentences):
            members = table.sent_frag_index[table.sent_offsets[s] : table.sent_offsets[s + 1]]
            vals = [frag_mean[k] for k in members]
            p = max(0.0, min(float(np.mean(vals)) if vals else 0.0, 1.0))
            sent_prob.append(p)
            keep.append(p > threshold)
        return {"rank_score": rank_score, "sent_prob": np.asarray(sent_prob), "keep": np.asarray(keep, dtype=bool)}


@pytest.fixture(scope="module")
def tiny_tokenizer(tiny_ckpt_dir):
    from transformers import AutoTokenizer

    return AutoTokenizer.from_pretrained(str(tiny_ckpt_dir))


def _model(tiny_ckpt_dir, tokenizer, case):
    config = OpenProvenceConfig.from_pretrained(tiny_ckpt_dir)
    scorer = RecordedLogitScorer(case["blocks"])
    model = OpenProvenceModel(config, None, tokenizer, scorer=scorer)
    model.max_length = case["max_length"]
    return model, scorer


def _approx_nested(a, b, tol):
    if isinstance(a, (list, tuple)):
        assert isinstance(b, (list, tuple)) and len(a) == len(b), (a, b)
        for x, y in zip(a, b):
            _approx_nested(x, y, tol)
    elif a is None or b is None:
        assert a is None and b is None
    else:
        assert abs(float(a) - float(b)) <= tol, (a, b)


@pytest.mark.parametrize("name", CASES)
def test_process_matches_reference(name, process_golden, tiny_ckpt_dir, tiny_tokenizer):
    case = next(c for c in process_golden["cases"] if c["name"] == name)
    model, scorer = _model(tiny_ckpt_dir, tiny_tokenizer, case)
    kwargs = dict(case["kwargs"])
    kwargs["sentence_splitter"] = simple_sentence_splitter
    result = model.process(**kwargs)
    golden = case["result"]

    # the blocks handed to the device are exactly the sequences the reference fed to its forward
    assert sorted(scorer.seen) == sorted(tuple(b["ids"]) for b in case["blocks"])
    assert result["pruned_context"] == golden["pruned_context"]
    assert result["kept_sentences"] == golden["kept_sentences"]
    assert result["removed_sentences"] == golden["removed_sentences"]
    assert result["title"] == golden["title"]
    _approx_nested(result["compression_rate"], golden["compression_rate"], 1e-9)
    _approx_nested(result["reranking_score"], golden["reranking_score"], 1e-6)
    _approx_nested(result["sentence_probabilities"], golden["sentence_probabilities"], 1e-6)
    assert list(result.keys())[:6] == ["pruned_context", "reranking_score", "compression_rate", "title", "timing",
                                       "performance_trace"]
    assert set(result["timing"]) == set(result["performance_trace"].as_dict())


@pytest.mark.parametrize("name", ["str_list", "nested", "explicit_titles", "multi_block", "reorder_topk", "empty_context"])
def test_chunked_host_pipeline_gives_the_same_result(name, process_golden, tiny_ckpt_dir, tiny_tokenizer):
    """One context per host-preparation chunk (worker thread prepares chunk k+1 while chunk k is scored) must
    reproduce the reference result exactly, like the single-chunk run above."""
    case = next(c for c in process_golden["cases"] if c["name"] == name)
    model, scorer = _model(tiny_ckpt_dir, tiny_tokenizer, case)
    kwargs = dict(case["kwargs"])
    kwargs["sentence_splitter"] = simple_sentence_splitter
    kwargs["preprocess_batch_size"] = 1
    result = model.process(**kwargs)
    golden = case["result"]
    assert sorted(scorer.seen) == sorted(tuple(b["ids"]) for b in case["blocks"])
    for key in ("pruned_context", "kept_sentences", "removed_sentences", "title"):
        assert result[key] == golden[key]
    _approx_nested(result["compression_rate"], golden["compression_rate"], 1e-9)
    _approx_nested(result["reranking_score"], golden["reranking_score"], 1e-6)
    _approx_nested(result["sentence_probabilities"], golden["sentence_probabilities"], 1e-6)


def test_process_default_result_keys(process_golden, tiny_ckpt_dir, tiny_tokenizer):
    case = next(c for c in process_golden["cases"] if c["name"] == "str_str")
    model, _ = _model(tiny_ckpt_dir, tiny_tokenizer, case)
    res = model.process(question=case["kwargs"]["question"], context=case["kwargs"]["context"],
                        threshold=case["kwargs"]["threshold"], sentence_splitter=simple_sentence_splitter)
    assert set(res) == {"pruned_context", "reranking_score", "compression_rate", "title", "timing", "performance_trace"}


def test_input_validation_messages(tiny_ckpt_dir, tiny_tokenizer, process_golden):
    case = process_golden["cases"][0]
    model, _ = _model(tiny_ckpt_dir, tiny_tokenizer, case)
    with pytest.raises(ValueError, match="Number of contexts must match number of queries"):
        model.process(question=["a", "b"], context=["only one"], sentence_splitter=simple_sentence_splitter)
    with pytest.raises(ValueError, match="Unsupported context format"):
        model.process(question="a", context=3, sentence_splitter=simple_sentence_splitter)
    with pytest.raises(ValueError, match="first_line_as_title=True cannot be combined"):
        model.process(question="a", context="x", title="T", first_line_as_title=True,
                      sentence_splitter=simple_sentence_splitter)
    with pytest.raises(ValueError, match="language must be provided"):
        model.process(question="a", context="x", sentence_splitter={"ja": simple_sentence_splitter})
    with pytest.raises(TypeError, match="debug_messages"):
        model.process(question="a", context="x", sentence_splitter=simple_sentence_splitter, debug_messages=3)


def test_threshold_resolution(tiny_ckpt_dir):
    config = OpenProvenceConfig.from_pretrained(tiny_ckpt_dir)
    assert config.default_threadshold == pytest.approx(0.1)
    model = OpenProvenceModel(config, None, None, scorer=object())
    assert model._resolve_process_threshold(None) == pytest.approx(0.1)
    assert model._resolve_process_threshold(0.37) == pytest.approx(0.37)
    with pytest.warns(RuntimeWarning, match="default_threshold"):
        cfg = OpenProvenceConfig(default_threshold=0.3)
    assert cfg.default_threadshold == pytest.approx(0.3)
    assert OpenProvenceConfig().resolve_default_threshold() == pytest.approx(0.1)


@pytest.mark.parametrize("name", ["str_list", "multi_block", "overlong_sentence"])
def test_process_ignores_tokenizer_state_left_by_earlier_calls(name, process_golden, tiny_ckpt_dir, tiny_tokenizer):
    """HF fast tokenizers keep the truncation / padding of their last call on the Rust backend.  The reference's
    ``tokenizer(list, add_special_tokens=False)`` resets it; the Rust ``encode_batch`` fast path must do the same, or
    every sentence is padded to the longest one / cut at max_length (host_text.tokenize_batch)."""
    case = next(c for c in process_golden["cases"] if c["name"] == name)
    model, scorer = _model(tiny_ckpt_dir, tiny_tokenizer, case)
    kwargs = dict(case["kwargs"])
    kwargs["sentence_splitter"] = simple_sentence_splitter
    model.process(**kwargs)  # first call: caches the separator length (a plain tokenizer call that resets the state)
    scorer.seen.clear()
    try:
        tiny_tokenizer(["a b c d e f g h i j k l m n o p", "a"], padding=True, truncation=True, max_length=6)
        assert tiny_tokenizer.backend_tokenizer.padding is not None
        assert tiny_tokenizer.backend_tokenizer.truncation is not None
        result = model.process(**kwargs)
    finally:
        tiny_tokenizer.backend_tokenizer.no_padding()
        tiny_tokenizer.backend_tokenizer.no_truncation()
    golden = case["result"]
    assert sorted(scorer.seen) == sorted(tuple(b["ids"]) for b in case["blocks"])
    assert result["pruned_context"] == golden["pruned_context"]
    assert result["kept_sentences"] == golden["kept_sentences"]


def test_context_ranges_leave_no_truncation_state(tiny_ckpt_dir, tiny_tokenizer, process_golden):
    case = process_golden["cases"][0]
    model, _ = _model(tiny_ckpt_dir, tiny_tokenizer, case)
    model.max_length = 32
    ranges = model._context_ranges_from_contexts("what is it?", ["Alpha beta gamma. ", "Delta epsilon. " * 20])
    assert ranges[0][1] == ranges[1][0] and ranges[-1][1] <= 32
    assert tiny_tokenizer.backend_tokenizer.truncation is None
