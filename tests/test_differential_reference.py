"""Randomised differential test of the whole host path of ``process()`` against THE REFERENCE ITSELF.

Runs only where ``/root/reference`` exists (the build container; skipped on the GPU box): the unmodified reference
module is imported as in ``tests/golden/make_golden.py``, its ``forward`` is replaced by a deterministic function of
the token ids (the seam its own tests use), the same function feeds this repo's scorer seam, and ``process()`` of both
is called on randomly generated questions / contexts / options.  Everything the host does -- input normalisation,
titles, sentence handling, tokenisation, fragmentising, block assembly, range tables, sentence means, thresholds,
string joins, compression rates, reordering -- must come out identical."""

from __future__ import annotations

import random
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

REF_FILE = Path("/root/reference/open_provence/modeling_open_provence_standalone.py")
pytestmark = pytest.mark.skipif(not REF_FILE.exists(), reason="the reference tree is only present in the build container")

from open_provence_b200.config import OpenProvenceConfig  # noqa: E402
from open_provence_b200.host_text import simple_sentence_splitter  # noqa: E402
from open_provence_b200.modeling import OpenProvenceModel  # noqa: E402
from oracle import postprocess_numpy as opp  # noqa: E402


def fake_logits(ids: list[int]) -> tuple[np.ndarray, np.ndarray]:
    """Deterministic (rank logit [1], prune logits [n, 2]) from the token ids of one block."""
    arr = np.asarray(ids, dtype=np.int64)
    pos = np.arange(arr.size, dtype=np.int64)
    keep = (((arr * 2654435761 + pos * 40503) % 1000).astype(np.float32) / np.float32(1000.0)) * np.float32(8.0) - np.float32(4.0)
    prune = np.stack([np.zeros_like(keep), keep], axis=1).astype(np.float32)
    rank = np.asarray([np.float32((int(arr.sum()) % 97) / 97.0 * 6.0 - 3.0)], dtype=np.float32)
    return rank, prune


class FakeScorer:
    """This repo's device stage, replaced by ``fake_logits`` + the oracle's arithmetic (cf. tests/test_process_host.py)."""

    def run(self, table, threshold):
        rank_score = np.zeros(table.n_blocks, dtype=np.float32)
        probs = []
        for b, ids in enumerate(table.block_ids):
            rank, prune = fake_logits([int(t) for t in ids])
            rank_score[b] = opp.ranking_score_from_logits(rank)
            probs.append(opp.keep_probs_from_logits(prune))
        frag_mean = [1.0 if e <= s else float(probs[blk][s:e].mean()) for blk, (s, e) in zip(table.frag_block, table.frag_local)]
        sent_prob, keep = [], []
        for s in range(table.n_sentences):
            vals = [frag_mean[k] for k in table.sent_frag_index[table.sent_offsets[s] : table.sent_offsets[s + 1]]]
            p = max(0.0, min(float(np.mean(vals)) if vals else 0.0, 1.0))
            sent_prob.append(p)
            keep.append(p > threshold)
        return {"rank_score": rank_score, "sent_prob": np.asarray(sent_prob), "keep": np.asarray(keep, dtype=bool)}


@pytest.fixture(scope="module")
def both(tiny_ckpt_dir):
    sys.path.insert(0, str(tiny_ckpt_dir.parent))
    import make_golden as mg

    ref = mg.load_reference_module()
    from transformers import AutoTokenizer

    fast = AutoTokenizer.from_pretrained(str(tiny_ckpt_dir))
    ref.AutoTokenizer.from_pretrained = staticmethod(lambda *_a, **_k: mg.Tokenizer457Shim(fast))
    import json

    cfg = json.loads((tiny_ckpt_dir / "config.json").read_text())
    ref_model = ref.OpenProvenceModel(ref.OpenProvenceConfig(
        base_model_config=cfg["base_model_config"], tokenizer_name_or_path="tiny_ckpt", pruning_config=cfg["pruning_config"],
        max_length=512, default_threadshold=0.1))
    ref_model.eval()

    def ref_forward(self, input_ids=None, attention_mask=None, **_kw):
        B, S = input_ids.shape
        rank = torch.zeros(B, 1)
        prune = torch.zeros(B, S, 2)
        for b in range(B):
            n = int(attention_mask[b].sum()) if attention_mask is not None else S
            r, p = fake_logits(input_ids[b, :n].tolist())
            rank[b] = torch.from_numpy(r)
            prune[b, :n] = torch.from_numpy(p)
        return {"ranking_logits": rank, "pruning_logits": prune}

    ref.OpenProvenceModel.forward = ref_forward
    ours = OpenProvenceModel(OpenProvenceConfig.from_pretrained(tiny_ckpt_dir), None, fast, scorer=FakeScorer())
    return ref_model, ours


WORDS = ["alpha", "beta", "gamma", "delta", "tower", "banana", "river", "東京", "タワー", "question", "answer", "x", "pruning",
         "context", "sentence", "the", "of", "is"]
ENDS = [". ", "! ", "? ", "。", ".\n", "\n\n", "  "]


def random_text(rng: random.Random, n_sent: int) -> str:
    return "".join(" ".join(rng.choice(WORDS) for _ in range(rng.randint(1, 14))) + rng.choice(ENDS) for _ in range(n_sent))


def random_case(rng: random.Random) -> dict:
    n_q = rng.randint(1, 3)
    questions = ["what " + " ".join(rng.choice(WORDS) for _ in range(rng.randint(1, 6))) + "?" for _ in range(n_q)]
    shape = rng.choice(["str_str", "str_list", "aligned", "nested", "presplit"])
    kw: dict = {}
    if shape == "str_str":
        kw.update(question=questions[0], context=random_text(rng, rng.randint(0, 12)))
    elif shape == "str_list":
        kw.update(question=questions[0], context=[random_text(rng, rng.randint(0, 10)) for _ in range(rng.randint(1, 4))])
    elif shape == "aligned":
        kw.update(question=questions, context=[random_text(rng, rng.randint(1, 10)) for _ in questions])
    elif shape == "nested":
        kw.update(question=questions, context=[[random_text(rng, rng.randint(0, 9)) for _ in range(rng.randint(1, 3))] for _ in questions])
    else:
        kw.update(question=questions,
                  context=[[[random_text(rng, 1) for _ in range(rng.randint(1, 8))] for _ in range(rng.randint(1, 2))] for _ in questions])
    kw["threshold"] = rng.choice([0.05, 0.3, 0.5, 0.7])
    kw["strip_sentences"] = rng.random() < 0.3
    kw["respect_sentence_boundaries"] = rng.random() < 0.3
    kw["always_select_title"] = rng.random() < 0.4
    kw["use_best_reranker_score"] = rng.random() < 0.7
    kw["zero_score_when_empty"] = rng.random() < 0.7
    kw["first_line_as_title"] = rng.random() < 0.2
    title_mode = rng.choice(["default", "none", "explicit"])
    if title_mode == "none" or kw["first_line_as_title"]:
        kw["title"] = None
    elif title_mode == "explicit" and shape in ("str_list", "nested"):
        def title():
            return rng.choice(["", "Title " + rng.choice(WORDS), rng.choice(WORDS) + "\n"])
        if shape == "str_list":
            kw["title"] = [title() for _ in kw["context"]]
        else:
            kw["title"] = [[title() for _ in docs] for docs in kw["context"]]
    if rng.random() < 0.3:
        kw["reorder"] = True
        kw["top_k"] = rng.choice([None, 1, 2])
    kw.update(sentence_splitter=simple_sentence_splitter, return_sentence_metrics=True, return_sentence_texts=True, show_progress=False)
    return kw


def _close(a, b, path="result"):
    if isinstance(a, (list, tuple)):
        assert isinstance(b, (list, tuple)) and len(a) == len(b), f"{path}: {a!r} vs {b!r}"
        for i, (x, y) in enumerate(zip(a, b)):
            _close(x, y, f"{path}[{i}]")
    elif isinstance(a, float) or isinstance(b, float):
        assert a is not None and b is not None and abs(float(a) - float(b)) <= 1e-6, f"{path}: {a!r} vs {b!r}"
    else:
        assert a == b, f"{path}: {a!r} vs {b!r}"


@pytest.mark.parametrize("seed", range(120))
def test_process_equals_reference_on_random_inputs(both, seed):
    ref_model, ours = both
    rng = random.Random(seed)
    kw = random_case(rng)
    max_length = rng.choice([48, 96, 160, 512])
    ref_model.max_length = ours.max_length = max_length
    ref_model.config.max_length = max_length
    want = ref_model.process(**kw)
    got = ours.process(**kw, preprocess_batch_size=rng.choice([None, 1, 2]))
    for key in ("pruned_context", "reranking_score", "compression_rate", "title", "kept_sentences", "removed_sentences",
                "sentence_probabilities"):
        _close(got[key], want[key], key)
