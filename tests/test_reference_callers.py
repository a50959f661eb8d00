"""The reference's OWN caller code executed against this repo's ``process()``.

Runs only where ``/root/reference`` exists (the build container).  The unmodified bodies of
``scripts/hf_utils/hf_model_process_check.py::run_cases`` (42-64 + 100-128), ``scripts/eval_datasets.py::evaluate_dataset``
(247-486, call at 331-346) and ``scripts/eval_mldr.py::build_records`` (238-524, call at 385-418) are imported from the
reference tree and run twice: once with the reference's ``OpenProvenceModel`` and once with this repo's, both with the same
deterministic stand-in for the forward (tests/test_differential_reference.py).  Whatever those callers compute from the
result -- table rows, span accuracy / precision / recall / ROC payloads, per-passage records and score statistics -- must
be identical.  The device kernels behind the same seam are pinned by the ``gpu`` tests."""

from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import pytest

REF_ROOT = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REF_ROOT / "scripts" / "eval_datasets.py").exists(),
                                reason="the reference tree is only present in the build container")

from open_provence_b200.host_text import simple_sentence_splitter  # noqa: E402
from test_differential_reference import both  # noqa: E402,F401  (fixture: reference model + this repo's model)


def _load_script(name: str, path: Path, ref_module) -> types.ModuleType:
    """Import a reference script with ``open_provence.modeling_open_provence_standalone`` bound to the already loaded
    reference module (the package __init__ would pull in the training stack) and absent third parties stubbed."""
    pkg = types.ModuleType("open_provence")
    pkg.__path__ = []  # a package
    pkg.modeling_open_provence_standalone = ref_module
    sys.modules.setdefault("open_provence", pkg)
    sys.modules.setdefault("open_provence.modeling_open_provence_standalone", ref_module)
    if "litellm" not in sys.modules:
        try:
            import litellm  # noqa: F401
        except ImportError:
            sys.modules["litellm"] = types.ModuleType("litellm")  # only used by the LLM-judge half of eval_mldr.py
    spec = importlib.util.spec_from_file_location(name, path)
    module = importlib.util.module_from_spec(spec)
    sys.modules[name] = module
    spec.loader.exec_module(module)
    return module


class _DefaultSplitter:
    """What differs between the two containers is only which sentence splitter 'auto' resolves to (nltk / fast-bunkai are
    not installed here); the callers that do not pass one get the same explicit splitter on both sides."""

    def __init__(self, model):
        self._model = model

    def process(self, question, context, title="first_sentence", threshold=None, batch_size=32,
                use_best_reranker_score=True, show_progress=False, return_sentence_texts=False, **kwargs):
        # named parameters: eval_mldr.build_records drops every keyword that inspect.signature() does not show
        kwargs.update(question=question, context=context, title=title, threshold=threshold, batch_size=batch_size,
                      use_best_reranker_score=use_best_reranker_score, show_progress=show_progress,
                      return_sentence_texts=return_sentence_texts)
        if kwargs.get("sentence_splitter") is None and not _is_presplit(context):
            kwargs["sentence_splitter"] = simple_sentence_splitter
        return self._model.process(**kwargs)


def _is_presplit(context) -> bool:
    return (isinstance(context, list) and context and isinstance(context[0], list) and context[0]
            and isinstance(context[0][0], list))


@pytest.fixture(scope="module")
def ref_module(both):  # noqa: F811
    return sys.modules["ref_standalone"]


def test_hf_model_process_check_run_cases(both, ref_module):  # noqa: F811
    ref_model, ours = both
    script = _load_script("ref_hf_model_process_check", REF_ROOT / "scripts" / "hf_utils" / "hf_model_process_check.py",
                          ref_module)
    ref_model.max_length = ours.max_length = 512
    ref_model.config.max_length = 512
    want = script.run_cases(_DefaultSplitter(ref_model), 0.1, True)
    got = script.run_cases(_DefaultSplitter(ours), 0.1, True)
    assert len(got) == len(want) == 10  # 1 + 2 + 4 (one "document" per sentence) + 1 + 2 samples
    for a, b in zip(got, want):
        assert (a.case, a.sample, a.pruned) == (b.case, b.sample, b.pruned)
        assert (a.score is None) == (b.score is None) and (a.score is None or abs(a.score - b.score) < 1e-6)
        assert abs(a.compression - b.compression) < 1e-9
    assert script._format_table(got) == script._format_table(want)


def test_eval_datasets_evaluate_dataset(both, ref_module):  # noqa: F811
    from datasets import Dataset

    ref_model, ours = both
    script = _load_script("ref_eval_datasets", REF_ROOT / "scripts" / "eval_datasets.py", ref_module)
    text_a = "Tokyo Tower is a tower in Minato. It was completed in 1958. Many people visit the tower every year."
    text_b = "Bananas are yellow. Rivers flow to the sea."
    spans = lambda t: [[m, n] for m, n in _sentence_spans(t)]  # noqa: E731
    rows = [
        {"query": "How tall is Tokyo Tower?", "texts": [text_a, text_b], "context_spans": [spans(text_a), spans(text_b)],
         "context_spans_relevance": [[1, 0, 1], [0, 0]]},
        {"query": "what is a banana?", "texts": [text_b], "context_spans": [spans(text_b)], "context_spans_relevance": [[1, 0]]},
    ]
    dataset = Dataset.from_list(rows)
    kw = dict(threshold=0.3, batch_size=8, dataset_label="tiny", show_progress=False, debug_messages=False,
              print_timing_summary=False, silent=True)
    ref_model.max_length = ours.max_length = 96
    ref_model.config.max_length = 96
    want = script.evaluate_dataset(ref_model, dataset, **kw)
    got = script.evaluate_dataset(ours, dataset, **kw)
    for key in ("span_total", "span_correct", "span_accuracy", "span_skipped", "contexts", "mean_compression", "precision",
                "recall", "f2", "confusion_matrix"):
        assert got[key] == want[key], key
    assert got["span_total"] == 7 and got["contexts"] == 3
    assert got["roc_data"]["labels"] == want["roc_data"]["labels"]
    assert got["roc_data"]["predictions"] == want["roc_data"]["predictions"]
    assert all(abs(a - b) < 1e-6 for a, b in zip(got["roc_data"]["scores"], want["roc_data"]["scores"]))
    assert set(got["timing"]) >= {"preprocess_seconds", "assembly_seconds", "inference_seconds", "postprocess_seconds",
                                  "total_seconds"}


def _sentence_spans(text: str):
    at = 0
    for part in text.split(". "):
        end = min(len(text), at + len(part) + 2)
        yield at, end
        at = end


def test_eval_mldr_build_records(both, ref_module):  # noqa: F811
    from datasets import Dataset

    ref_model, ours = both
    script = _load_script("ref_eval_mldr", REF_ROOT / "scripts" / "eval_mldr.py", ref_module)
    rows = [
        {"query_id": "q1", "query": "How tall is Tokyo Tower?",
         "positive_passages": [{"docid": "d1", "title": "Tokyo Tower", "text": "Tokyo Tower is tall. It is 332.9 meters tall. " * 6}],
         "negative_passages": [{"docid": "d2", "title": "", "text": "Bananas are yellow. Rivers flow to the sea."},
                               {"docid": "d3", "title": "Rivers", "text": "The river is long.\nIt flows to the sea."}]},
        {"query_id": "q2", "query": "what is a banana?",
         "positive_passages": [{"docid": "d4", "title": None, "text": "A banana is a fruit. It is yellow!"}],
         "negative_passages": []},
    ]
    dataset = Dataset.from_list(rows)
    kw = dict(threshold=0.3, batch_size=4, log_timing=False, use_best_reranker_score=True, show_progress=False)
    ref_model.max_length = ours.max_length = 64  # several blocks per long passage
    ref_model.config.max_length = 64
    want_records, want_stats, want_n = script.build_records(_DefaultSplitter(ref_model).process, dataset, **kw)
    got_records, got_stats, got_n = script.build_records(_DefaultSplitter(ours).process, dataset, **kw)
    assert got_n == want_n == 2 and len(got_records) == len(want_records) == 4
    for a, b in zip(got_records, want_records):
        for key in ("query_id", "docid", "label", "title", "original_text", "pruned_text", "kept_sentences", "removed_sentences"):
            assert a[key] == b[key], key
        assert abs(a["reranking_score"] - b["reranking_score"]) < 1e-6
        assert abs(a["compression_rate"] - b["compression_rate"]) < 1e-9
    for key in want_stats:
        assert len(got_stats[key]) == len(want_stats[key])
        assert all(abs(x - y) < 1e-6 for x, y in zip(got_stats[key], want_stats[key]))
