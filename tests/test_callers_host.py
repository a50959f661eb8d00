"""The reference's own callers of ``process()`` as an acceptance harness:
``scripts/eval_datasets.py:331-346``, ``scripts/eval_mldr.py:385-418`` and the five input shapes of
``scripts/hf_utils/hf_model_process_check.py:42-64``, called with exactly the keyword arguments those
scripts pass and consumed the way they consume the result.  Device stage = a constant scorer (host test)."""

from __future__ import annotations

import inspect

import numpy as np
import pytest

from open_provence_b200.config import OpenProvenceConfig
from open_provence_b200.host_text import simple_sentence_splitter
from open_provence_b200.modeling import OpenProvenceModel

QUESTION = "How tall is Tokyo Tower?"
CONTEXT = (
    "Tokyo Tower is a communications and observation tower in Minato.\n"
    "It was completed in 1958. At 332.9 meters it is the second-tallest structure in Japan.\n"
    "Over 150 million people have visited the tower.\n"
)


class RampScorer:
    """Sentence probabilities 0, 1/(n-1), ..., 1 and rank score 0.5 for every block."""

    def run(self, table, threshold):
        n = table.n_sentences
        prob = np.linspace(0.0, 1.0, n) if n else np.zeros(0)
        return {"rank_score": np.full(table.n_blocks, 0.5, np.float32), "sent_prob": prob, "keep": prob > threshold}


@pytest.fixture(scope="module")
def model(tiny_ckpt_dir):
    from transformers import AutoTokenizer

    tok = AutoTokenizer.from_pretrained(str(tiny_ckpt_dir))
    return OpenProvenceModel(OpenProvenceConfig.from_pretrained(tiny_ckpt_dir), None, tok, scorer=RampScorer())


def test_eval_datasets_call(model):
    """eval_datasets.py: question=list, context=list[list[list[str]]] (pre-split sentences), title=None."""
    sentences = [line for line in CONTEXT.splitlines(True) if line.strip()]
    questions = [QUESTION, "second question?"]
    contexts_nested = [[sentences, sentences[:2]], [sentences]]
    outputs = model.process(**{
        "question": questions, "context": contexts_nested, "title": None, "batch_size": 8, "threshold": 0.1,
        "sentence_splitter": None, "show_progress": False, "debug_messages": False, "return_sentence_metrics": True,
        "show_inference_progress": False,
    })
    pruned, rates = outputs["pruned_context"], outputs["compression_rate"]
    probs = outputs.get("sentence_probabilities") or []
    assert [len(p) for p in pruned] == [2, 1] and [len(r) for r in rates] == [2, 1]
    assert [len(p) for p in probs] == [2, 1] and len(probs[0][0]) == len(sentences)
    timing = outputs.get("timing") or {}
    for key in ("preprocess_seconds", "assembly_seconds", "inference_seconds", "postprocess_seconds", "total_seconds",
                "sentence_collect_seconds", "sentence_normalize_seconds", "tokenize_seconds", "fragment_split_seconds",
                "fragment_decode_seconds"):
        assert isinstance(float(timing[key]), float)
    assert hasattr(outputs.get("performance_trace"), "as_dict")


def test_eval_mldr_call(model):
    """eval_mldr.py filters its kwargs by ``inspect.signature(process)`` and then reads five result lists."""
    kwargs = {
        "question": [QUESTION, "q2"], "context": [[CONTEXT, CONTEXT], [CONTEXT]], "title": [["Tokyo Tower", ""], [""]],
        "threshold": 0.1, "batch_size": 16, "log_timing": True, "use_best_reranker_score": True, "show_progress": False,
        "return_sentence_texts": True, "sentence_splitter": simple_sentence_splitter,
    }
    supported = set(inspect.signature(model.process).parameters)
    assert "question" in supported and "log_timing" not in supported  # same as the reference signature
    result = model.process(**{k: v for k, v in kwargs.items() if k in supported})
    for key in ("pruned_context", "reranking_score", "compression_rate", "kept_sentences", "removed_sentences", "title"):
        assert key in result and [len(x) for x in result[key]] == [2, 1], key
    assert all(isinstance(s, str) for s in result["kept_sentences"][0][0])


def test_hf_model_process_check_shapes(model):
    """The five (question, context) structures of hf_model_process_check.py::build_cases."""
    sentences = [line for line in CONTEXT.splitlines(True) if line.strip()]
    cases = [
        (QUESTION, CONTEXT, str),
        ([QUESTION, QUESTION], [CONTEXT, CONTEXT], list),
        (QUESTION, sentences, list),
        (QUESTION, [sentences], list),
        ([QUESTION, QUESTION], [[sentences], [sentences]], list),
    ]
    for question, context, kind in cases:
        result = model.process(question=question, context=context, threshold=0.1, show_progress=False,
                               sentence_splitter=simple_sentence_splitter)
        pruned, score, rate = result["pruned_context"], result["reranking_score"], result["compression_rate"]
        assert isinstance(pruned, kind)
        flat = [pruned] if isinstance(pruned, str) else pruned
        while flat and isinstance(flat[0], list):
            flat = [x for sub in flat for x in sub]
        assert all(isinstance(p, str) for p in flat)
        assert type(score) is type(rate) or isinstance(score, (float, type(None)))
