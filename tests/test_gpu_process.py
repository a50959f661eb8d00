"""End to end on the GPU: ``from_pretrained`` -> ``process()`` / ``forward()`` against the reference's results."""

from __future__ import annotations

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from open_provence_b200.host_text import simple_sentence_splitter  # noqa: E402
from open_provence_b200.modeling import OpenProvenceModel  # noqa: E402
from open_provence_b200.scoring import BlockTable, DeviceScorer  # noqa: E402
from oracle import postprocess_numpy as opp  # noqa: E402

CASES = [
    "str_str", "str_list", "aligned", "nested", "presplit_sentences", "explicit_titles", "first_line_title",
    "title_none", "multi_block", "multi_block_respect", "overlong_sentence", "strip_sentences", "reorder_topk",
    "no_best_score", "japanese", "empty_context", "batch_size_1",
]


@pytest.fixture(scope="module")
def model_fp32(tiny_ckpt_dir):
    return OpenProvenceModel.from_pretrained(tiny_ckpt_dir, device="cuda", dtype="fp32")


@pytest.fixture(scope="module")
def model_bf16(tiny_ckpt_dir):
    return OpenProvenceModel.from_pretrained(tiny_ckpt_dir, device="cuda")


def _flatten(x):
    if isinstance(x, (list, tuple)):
        out = []
        for v in x:
            out.extend(_flatten(v))
        return out
    return [x]


def _run(model, case):
    kwargs = dict(case["kwargs"])
    kwargs["sentence_splitter"] = simple_sentence_splitter
    model.max_length = case["max_length"]
    return model.process(**kwargs)


@pytest.mark.parametrize("name", CASES)
def test_process_fp32_identical_to_reference(name, process_golden, model_fp32):
    """fp32 engine: pruned text / kept sentence sets bit-exact, probabilities within 1e-5."""
    case = next(c for c in process_golden["cases"] if c["name"] == name)
    res = _run(model_fp32, case)
    gold = case["result"]
    assert res["pruned_context"] == gold["pruned_context"]
    assert res["kept_sentences"] == gold["kept_sentences"]
    assert res["removed_sentences"] == gold["removed_sentences"]
    assert res["title"] == gold["title"]
    a, b = _flatten(res["sentence_probabilities"]), _flatten(gold["sentence_probabilities"])
    assert len(a) == len(b)
    assert max((abs(x - y) for x, y in zip(a, b)), default=0.0) < 1e-5
    sa, sb = _flatten(res["reranking_score"]), _flatten(gold["reranking_score"])
    assert all((x is None) == (y is None) for x, y in zip(sa, sb))
    assert max((abs(x - y) for x, y in zip(sa, sb) if x is not None), default=0.0) < 1e-5


@pytest.mark.parametrize("name", CASES)
def test_process_bf16_close_to_reference(name, process_golden, model_bf16):
    """bf16 engine: probabilities within 1e-2; keep decisions identical wherever the reference's probability is
    further than that from the threshold."""
    case = next(c for c in process_golden["cases"] if c["name"] == name)
    res = _run(model_bf16, case)
    gold = case["result"]
    a, b = _flatten(res["sentence_probabilities"]), _flatten(gold["sentence_probabilities"])
    assert len(a) == len(b)
    err = max((abs(x - y) for x, y in zip(a, b)), default=0.0)
    assert err < 1e-2, f"sentence probability error {err:.3e}"
    thr = case["kwargs"]["threshold"]
    if all(abs(y - thr) > 1e-2 for y in b):
        assert res["kept_sentences"] == gold["kept_sentences"]
        assert res["pruned_context"] == gold["pruned_context"]
    sa, sb = _flatten(res["reranking_score"]), _flatten(gold["reranking_score"])
    assert max((abs(x - y) for x, y in zip(sa, sb) if x is not None and y is not None), default=0.0) < 5e-3


def test_forward_padded_interface(model_fp32, forward_golden):
    """``forward(input_ids, attention_mask)`` keeps the reference's [B, S] interface (standalone:1666-1739)."""
    ids = torch.from_numpy(forward_golden["input_ids"])
    mask = torch.from_numpy(forward_golden["attention_mask"])
    out = model_fp32.forward(input_ids=ids, attention_mask=mask, return_dict=True, token_type_ids=torch.zeros_like(ids))
    assert out.ranking_logits.shape == (ids.shape[0], 1) and out.pruning_logits.shape == (*ids.shape, 2)
    assert out["ranking_logits"] is out.logits
    rank = out.ranking_logits.cpu().double().numpy()
    prune = out.pruning_logits.cpu().double().numpy()
    m = forward_golden["attention_mask"][..., None]
    assert np.abs(rank - forward_golden["ranking_logits_f64"]).max() < 1e-5
    assert np.abs((prune - forward_golden["pruning_logits_f64"]) * m).max() < 1e-5
    tup = model_fp32.forward(input_ids=ids, attention_mask=mask, return_dict=False)
    assert isinstance(tup, tuple) and len(tup) == 2
    with pytest.raises(ValueError, match="input_ids must be provided"):
        model_fp32.forward(input_ids=None)
    # left padding (and any other mask): kept tokens are compacted, logits scattered back to their columns
    n = int(mask[1].sum())
    shifted_ids, shifted_mask = ids.clone(), mask.clone()
    pad = ids.shape[1] - n
    if pad > 0:
        shifted_ids[1] = torch.cat([ids[1, n:], ids[1, :n]])
        shifted_mask[1] = torch.cat([mask[1, n:], mask[1, :n]])
        left = model_fp32.forward(input_ids=shifted_ids, attention_mask=shifted_mask, return_dict=True)
        assert torch.equal(left.ranking_logits, out.ranking_logits)
        assert torch.equal(left.pruning_logits[1, pad:], out.pruning_logits[1, :n])
        assert torch.count_nonzero(left.pruning_logits[1, :pad]) == 0
    empty = mask.clone()
    empty[-1] = 0
    with pytest.raises(ValueError, match="at least one token"):
        model_fp32.forward(input_ids=ids, attention_mask=empty)


def test_prune_kernels_match_oracle(model_fp32):
    """fragment means / sentence prune on the device vs the numpy restatement, incl. empty ranges,
    sentences without fragments and the guard-band re-evaluation at an exact threshold tie."""
    eng = model_fp32.engine
    rng = np.random.default_rng(3)
    n_tokens = 5000
    logits = rng.normal(0, 3, size=(n_tokens, 2)).astype(np.float32)
    rank_logits = rng.normal(0, 2, size=(7, 1)).astype(np.float32)
    ranges, at = [], 0
    while at < n_tokens - 60:
        step = int(rng.integers(1, 60))
        ranges.append((at, at + step))
        at += step
    ranges += [(10, 10), (20, 5), (n_tokens - 3, n_tokens + 50)]  # empty, inverted, clipped
    ranges = np.asarray(ranges, dtype=np.int32)
    d_logits = torch.from_numpy(logits).cuda()
    frag_mean, score = eng.fragment_means(d_logits, torch.from_numpy(ranges).cuda(), torch.from_numpy(rank_logits).cuda())
    probs = opp.keep_probs_from_logits(logits)
    ref = []
    for s, e in ranges:
        s, e = max(0, min(int(s), n_tokens)), min(int(e), n_tokens)
        e = max(s, e)
        ref.append(1.0 if e <= s else float(probs[s:e].mean()))
    torch.cuda.synchronize()
    assert np.abs(frag_mean.cpu().numpy() - np.asarray(ref)).max() < 2e-6
    ref_score = [opp.ranking_score_from_logits(r) for r in rank_logits]
    assert np.abs(score.cpu().numpy() - np.asarray(ref_score)).max() < 1e-6

    # sentences: groups of 1..4 fragments, plus one sentence with no fragment at all
    offsets, index = [0], []
    f = 0
    while f < len(ranges):
        k = int(rng.integers(1, 5))
        index.extend(range(f, min(f + k, len(ranges))))
        offsets.append(len(index))
        f += k
    offsets.append(len(index))  # empty sentence
    fm = frag_mean.cpu().numpy()
    ref_prob = [max(0.0, min(float(np.mean([float(v) for v in fm[index[a:b]]])) if b > a else 0.0, 1.0)) for a, b in zip(offsets[:-1], offsets[1:])]
    thr = ref_prob[3]  # exact tie: reference says "not kept" (strict >)
    prob, keep, near = eng.sentence_prune(frag_mean, torch.tensor(offsets, dtype=torch.int32).cuda(),
                                          torch.tensor(index, dtype=torch.int32).cuda(), thr, 1e-5)
    torch.cuda.synchronize()
    assert np.abs(prob.cpu().numpy() - np.asarray(ref_prob)).max() < 1e-12
    assert near.cpu().numpy()[3] == 1
    far = np.abs(np.asarray(ref_prob) - thr) > 1e-5
    assert np.array_equal(keep.cpu().numpy().astype(bool)[far], (np.asarray(ref_prob) > thr)[far])
    assert prob.cpu().numpy()[-1] == 0.0 and keep.cpu().numpy()[-1] == 0


def test_device_scorer_guard_band_matches_exact_reference_procedure(model_fp32):
    """A sentence sitting exactly on the threshold is re-evaluated with torch-CPU softmax + numpy mean."""
    eng = model_fp32.engine
    rng = np.random.default_rng(5)
    table = BlockTable()
    for b in range(3):
        n = 40 + 10 * b
        ids = rng.integers(5, 260, size=n).astype(np.int32)
        ids[0] = 1
        table.block_ids.append(ids)
        for s in range(4):
            table.frag_block.append(b)
            table.frag_local.append((2 + 8 * s, 2 + 8 * s + 8))
            table.sent_frag_index.append(len(table.frag_block) - 1)
            table.sent_offsets.append(len(table.sent_frag_index))
    scorer = DeviceScorer(eng)
    first = scorer.run(table, 0.5)
    target = 5
    thr = float(first["sent_prob"][target])
    second = scorer.run(table, thr)
    assert bool(second["near"][target])
    # the re-evaluation uses the reference's exact CPU procedure on the engine's own fp32 logits
    from open_provence_b200.scoring import exact_fragment_mean, exact_sentence_probability

    blk = table.frag_block[target]
    ids = torch.from_numpy(table.block_ids[blk]).cuda()
    cu = torch.tensor([0, ids.numel()], dtype=torch.int32, device="cuda")
    prune, _ = eng.forward_packed(ids, cu, int(ids.numel()))
    a, b = table.frag_local[target]
    exact = exact_sentence_probability([exact_fragment_mean(prune[a:b].cpu().numpy())])
    assert second["sent_prob"][target] == exact
    assert bool(second["keep"][target]) == (exact > thr)
    assert abs(exact - thr) < 1e-6
    others = np.abs(first["sent_prob"] - thr) > 1e-5
    assert np.array_equal(second["keep"][others], (first["sent_prob"] > thr)[others])


def test_raw_predictions_and_thresholds_on_the_engine(model_fp32, tiny_ckpt_dir):
    """get_raw_predictions_batch / predict_with_thresholds end to end (fp32 engine) against the reference's recorded
    results (tests/golden/make_golden_raw.py)."""
    import json

    golden = json.loads((tiny_ckpt_dir.parent / "raw_tiny.json").read_text())
    saved = model_fp32.max_length
    model_fp32.max_length = golden["max_length"]
    try:
        raws = model_fp32.get_raw_predictions_batch(golden["batch_queries"], golden["batch_contexts"])
        sep = model_fp32.tokenizer.sep_token
        for got, want, q, ctx in zip(raws, golden["raw_batch"], golden["batch_queries"], golden["batch_contexts"]):
            n = len(model_fp32.tokenizer(q + sep + "".join(ctx), truncation=True, max_length=golden["max_length"])["input_ids"])
            assert abs(got.ranking_score - want["ranking_score"]) < 1e-5
            assert [list(r) for r in got.context_ranges] == want["context_ranges"]
            np.testing.assert_allclose(got.pruning_probs[:n], want["pruning_probs"][:n], atol=1e-5)
        for want in golden["thresholds"]:
            got = model_fp32.predict_with_thresholds(golden["query"], golden["contexts"], [0.05, 0.1, 0.5],
                                                     use_majority=want["use_majority"])
            assert {str(k): v for k, v in got["predictions"].items()} == want["predictions"]
    finally:
        model_fp32.max_length = saved


@pytest.mark.parametrize("which", ["fp32", "bf16"])
def test_single_launch_fast_path_equals_general_path(which, model_fp32, model_bf16):
    """One staged copy each way (DeviceScorer._run_single_launch) against the per-array copies of
    score_blocks() + prune(): same kernels on the same inputs, so every output is the same bits."""
    eng = (model_fp32 if which == "fp32" else model_bf16).engine
    rng = np.random.default_rng(9)
    table = BlockTable()
    for b in range(7):
        n = int(rng.integers(20, 300))
        ids = rng.integers(5, 260, size=n).astype(np.int32)
        ids[0] = 1
        table.block_ids.append(ids)
        at = 2
        while at < n - 1:
            end = min(n - 1, at + int(rng.integers(1, 40)))
            table.frag_block.append(b)
            table.frag_local.append((at, end if rng.random() > 0.1 else at))  # some empty ranges (mean := 1.0)
            if rng.random() < 0.7 or not table.sent_frag_index:
                table.sent_frag_index.append(len(table.frag_block) - 1)
                table.sent_offsets.append(len(table.sent_frag_index))
            else:  # fragment joins the previous sentence
                table.sent_frag_index.append(len(table.frag_block) - 1)
                table.sent_offsets[-1] = len(table.sent_frag_index)
            at = end
        table.sent_offsets.append(len(table.sent_frag_index))  # a sentence without fragments (probability 0)
    scorer = DeviceScorer(eng)
    fast = scorer.run(table, 0.3)
    assert getattr(scorer, "_stage", None) is not None, "the fast path did not run"
    scorer.single_launch_fast_path = False
    slow = scorer.run(table, 0.3)
    for key in ("rank_score", "frag_mean", "sent_prob", "keep", "near"):
        assert np.array_equal(fast[key], slow[key]), key
    # a table too large for one launch falls back to the general path
    scorer.single_launch_fast_path = True
    scorer.max_tokens = 256
    assert scorer._run_single_launch(table, 0.3) is None
    again = scorer.run(table, 0.3)
    assert np.array_equal(again["keep"], slow["keep"]) and np.allclose(again["sent_prob"], slow["sent_prob"], atol=1e-12)


def test_load_and_process_the_way_eval_mldr_does(tiny_ckpt_dir, process_golden):
    """scripts/eval_mldr.py:145-165 (`_load_process_fn`): ``OpenProvenceModel.from_pretrained(path, device=..., max_length=...,
    trust_remote_code=..., torch_dtype=...)`` then ``.eval()`` and the bound ``process`` called with the keyword set of
    ``build_records`` (385-418) after the signature filter the script applies (``log_timing`` is not a parameter here
    either way)."""
    import inspect

    model = OpenProvenceModel.from_pretrained(str(tiny_ckpt_dir), device="cuda", max_length=96, trust_remote_code=True,
                                              torch_dtype=torch.float32)
    assert model.eval() is model and model.max_length == 96
    process_fn = model.process
    kwargs = {"question": ["Which topic?", "What are bananas?"], "context": [["Sentence one is here. Sentence two follows!"],
              ["Bananas are berries. They grow in clusters."]], "title": [[None], ["Bananas"]], "threshold": 0.25,
              "batch_size": 4, "log_timing": False, "use_best_reranker_score": True, "show_progress": False,
              "return_sentence_texts": True}
    supported = set(inspect.signature(process_fn).parameters)
    assert {"question", "context", "title", "threshold", "batch_size", "use_best_reranker_score", "show_progress",
            "return_sentence_texts"} <= supported
    result = process_fn(**{k: v for k, v in kwargs.items() if k in supported}, sentence_splitter=simple_sentence_splitter)
    for key in ("pruned_context", "reranking_score", "compression_rate", "kept_sentences", "removed_sentences", "title"):
        assert key in result and len(result[key]) == 2
    assert all(isinstance(s, float) for row in result["reranking_score"] for s in row)
