"""Native block assembly (``opv_pack_build``, csrc/host_pack.cu) against the Python restatement of the same
reference code (``OpenProvenceModel._fragmentize`` + ``_build_table``, pinned to the reference's ``process()`` by
tests/test_process_host.py and tests/test_differential_reference.py).  Host code only: no GPU needed, but the
C-ABI library has to be built.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from open_provence_b200 import _native, host_pack
from open_provence_b200.config import OpenProvenceConfig
from open_provence_b200.host_text import simple_sentence_splitter
from open_provence_b200.modeling import OpenProvenceModel


class _Capture:
    """Scorer stand-in that keeps the table and returns deterministic pseudo-random sentence probabilities."""

    def __init__(self):
        self.tables = []

    def run(self, table, threshold):
        self.tables.append(table)
        n = table.n_sentences
        prob = (np.arange(n) * 0.37) % 1.0
        return {"rank_score": np.linspace(0.1, 0.9, table.n_blocks).astype(np.float32), "sent_prob": prob,
                "keep": prob > threshold}


@pytest.fixture(scope="module")
def tokenizer(tiny_ckpt_dir):
    from transformers import AutoTokenizer

    return AutoTokenizer.from_pretrained(str(tiny_ckpt_dir))


def _model(tiny_ckpt_dir, tokenizer, mode, max_length):
    model = OpenProvenceModel(OpenProvenceConfig.from_pretrained(tiny_ckpt_dir), None, tokenizer, scorer=_Capture())
    model.max_length = max_length
    model.host_pack_mode = mode
    return model


def _table_lists(table):
    return {
        "block_ids": [[int(t) for t in b] for b in table.block_ids],
        "frag_block": [int(b) for b in table.frag_block],
        "frag_local": [(int(s), int(e)) for s, e in np.asarray(table.frag_local).reshape(-1, 2)],
        "sent_offsets": [int(v) for v in table.sent_offsets],
        "sent_frag_index": [int(v) for v in table.sent_frag_index],
    }


WORDS = ["alpha", "beta", "gamma", "delta", "pruning", "context", "question", "answer", "tokyo", "river", "東京",
         "は", "日本の", "首都です"]
ENDS = [". ", "! ", "? ", "。", "\n", ".\n\n", " \n", "\n \n"]


def _random_call(rng):
    def sentence():
        if rng.random() < 0.1:
            return str(rng.choice(["\n", " \n", "  ", "\n\n"]))  # whitespace-only: the empty-fragment filter
        return " ".join(rng.choice(WORDS, size=int(rng.integers(1, 40)))) + str(rng.choice(ENDS))

    def context():
        return "".join(sentence() for _ in range(int(rng.integers(0, 14))))

    n_q = int(rng.integers(1, 4))
    questions = [" ".join(rng.choice(WORDS, size=int(rng.integers(1, 30)))) + "?" for _ in range(n_q)]
    shape = int(rng.integers(0, 3))
    if shape == 0:
        contexts = [[context() for _ in range(int(rng.integers(1, 4)))] for _ in range(n_q)]
    elif shape == 1:  # pre-split sentences
        contexts = [[[sentence() for _ in range(int(rng.integers(1, 8)))] for _ in range(int(rng.integers(1, 3)))]
                    for _ in range(n_q)]
    else:
        contexts = [context() for _ in range(n_q)]
    kw = dict(question=questions, context=contexts, threshold=float(rng.choice([0.1, 0.5])),
              sentence_splitter=simple_sentence_splitter, show_progress=False, return_sentence_metrics=True,
              return_sentence_texts=True, strip_sentences=bool(rng.integers(0, 2)),
              respect_sentence_boundaries=bool(rng.integers(0, 2)), preprocess_batch_size=int(rng.choice([1, 2, 64])))
    title_mode = int(rng.integers(0, 4))
    if title_mode == 0:
        kw["title"] = None
    elif title_mode == 1 and shape != 1:
        kw["first_line_as_title"] = True
    elif title_mode == 2 and shape == 2:
        kw["title"] = ["Title %d\n" % i for i in range(n_q)]
    if rng.random() < 0.3:
        kw["always_select_title"] = True
    return kw


@pytest.mark.parametrize("seed", range(40))
def test_native_pack_equals_python_pack(tiny_ckpt_dir, tokenizer, seed):
    rng = np.random.default_rng(seed)
    kw = _random_call(rng)
    max_length = int(rng.choice([24, 48, 96, 512]))
    results = {}
    for mode in ("native", "python"):
        model = _model(tiny_ckpt_dir, tokenizer, mode, max_length)
        out = model.process(**kw)
        tables = [_table_lists(t) for t in model._scorer.tables]
        results[mode] = (tables, {k: v for k, v in out.items() if k not in ("timing", "performance_trace")})
    assert results["native"][0] == results["python"][0]
    assert results["native"][1] == results["python"][1]


def test_native_path_is_taken_and_skips_the_decode(tiny_ckpt_dir, tokenizer, monkeypatch):
    calls = {"pack": 0, "decoded": 0}
    real = host_pack.pack_blocks

    def spy(*args, **kwargs):
        packed = real(*args, **kwargs)
        calls["pack"] += 1
        calls["decoded"] += packed.decoded_fragments
        return packed

    monkeypatch.setattr(host_pack, "pack_blocks", spy)
    model = _model(tiny_ckpt_dir, tokenizer, "native", 64)
    model.process(question="what is alpha?", context="alpha beta gamma. \n delta river tokyo! \n\n answer.",
                  sentence_splitter=simple_sentence_splitter, show_progress=False)
    assert calls["pack"] == 1
    # only the whitespace-only sentence(s) needed the tokenizer's decode
    assert calls["decoded"] <= 2
    table = model._scorer.tables[0]
    assert table.packed_ids is not None and int(table.block_offsets[-1]) == table.packed_ids.shape[0]


def test_token_visibility_is_conservative(tokenizer):
    """Every token marked visible really decodes to visible text, alone and next to other tokens."""
    visible = host_pack.token_visibility(tokenizer)
    assert visible is not None and visible.dtype == np.uint8
    backend = tokenizer.backend_tokenizer
    vocab = visible.shape[0]
    for i in np.nonzero(visible)[0]:
        assert backend.decode([int(i)], skip_special_tokens=True).strip()
    rng = np.random.default_rng(0)
    for _ in range(200):
        ids = [int(t) for t in rng.integers(0, vocab, size=int(rng.integers(1, 6)))]
        if any(visible[t] for t in ids):
            assert backend.decode(ids, skip_special_tokens=True).strip()


def test_special_token_template(tokenizer):
    assert host_pack.special_token_template(tokenizer, True, 1, 2) == ([1], [2], [2])
    assert host_pack.special_token_template(tokenizer, True, None, 2) == ([], [2], [2])

    class Bert:
        all_special_ids = [101, 102]

        def build_inputs_with_special_tokens(self, a, b):
            return [101] + list(a) + [102] + list(b) + [102]

    assert host_pack.special_token_template(Bert(), False, None, None) == ([101], [102], [102])

    class Roberta(Bert):
        def build_inputs_with_special_tokens(self, a, b):
            return [0] + list(a) + [2, 2] + list(b) + [2]

    assert host_pack.special_token_template(Roberta(), False, None, None) == ([0], [2, 2], [2])

    class Odd(Bert):  # not head + q + mid + ctx + tail: the Python path keeps such tokenizers
        def build_inputs_with_special_tokens(self, a, b):
            return [101] + list(b) + [102] + list(a)

    assert host_pack.special_token_template(Odd(), False, None, None) is None


def test_pack_rejects_bad_arguments():
    lib = _native.load()
    handle = C.c_void_p()
    with pytest.raises(ValueError):
        _native.check(lib.opv_pack_build(None, C.byref(handle)), "opv_pack_build")
    inp = _native.OpvPackInput(abi_version=_native.OPV_ABI_VERSION + 1)
    with pytest.raises(ValueError, match="abi_version"):
        _native.check(lib.opv_pack_build(C.byref(inp), C.byref(handle)), "opv_pack_build")
    ctx_sent = np.array([0, 1], dtype=np.int64)
    sent = np.array([0, 0], dtype=np.int64)
    ctx_query = np.array([3], dtype=np.int32)  # only one query exists
    ctx_prefix = np.array([0], dtype=np.int32)
    q_off = np.array([0, 0], dtype=np.int64)
    inp = _native.OpvPackInput(
        abi_version=_native.OPV_ABI_VERSION, max_length=16, max_fragment_tokens=8, n_contexts=1, n_queries=1,
        h_sent_offsets=sent.ctypes.data, h_ctx_sent_offsets=ctx_sent.ctypes.data, h_ctx_query=ctx_query.ctypes.data,
        h_ctx_prefix=ctx_prefix.ctypes.data, h_query_offsets=q_off.ctypes.data)
    with pytest.raises(ValueError, match="names query 3"):
        _native.check(lib.opv_pack_build(C.byref(inp), C.byref(handle)), "opv_pack_build")
    assert lib.opv_pack_destroy(None) == 0


def test_pack_without_visibility_table_asks_for_every_fragment():
    """h_token_visible = NULL (slow tokenizers): the first call lists the windows, the second builds."""
    lib = _native.load()
    tokens = np.arange(10, 30, dtype=np.int32)
    sent = np.array([0, 7, 7, 20], dtype=np.int64)  # an empty sentence in the middle
    ctx_sent = np.array([0, 3], dtype=np.int64)
    ctx_query = np.zeros(1, dtype=np.int32)
    ctx_prefix = np.zeros(1, dtype=np.int32)
    q_tok = np.array([5, 6], dtype=np.int32)
    q_off = np.array([0, 2], dtype=np.int64)
    head, mid = np.array([1], dtype=np.int32), np.array([2], dtype=np.int32)
    inp = _native.OpvPackInput(
        abi_version=_native.OPV_ABI_VERSION, max_length=16, max_fragment_tokens=8, sep_len=1, n_contexts=1,
        n_queries=1, n_head=1, n_mid=1, n_tail=1, h_head=head.ctypes.data, h_mid=mid.ctypes.data,
        h_tail=mid.ctypes.data, h_tokens=tokens.ctypes.data, h_sent_offsets=sent.ctypes.data,
        h_ctx_sent_offsets=ctx_sent.ctypes.data, h_ctx_query=ctx_query.ctypes.data,
        h_ctx_prefix=ctx_prefix.ctypes.data, h_query_tokens=q_tok.ctypes.data, h_query_offsets=q_off.ctypes.data)
    handle, view = C.c_void_p(), _native.OpvPackView()
    _native.check(lib.opv_pack_build(C.byref(inp), C.byref(handle)))
    _native.check(lib.opv_pack_view_get(handle, C.byref(view)))
    assert view.needs_decode == 1 and view.n_raw_fragments == 3 and view.n_uncertain == 3  # 7 | 8 + 5
    lib.opv_pack_destroy(handle)
    drop = np.array([0, 1, 0], dtype=np.uint8)  # the middle window "decodes to nothing"
    inp.h_frag_drop = drop.ctypes.data
    _native.check(lib.opv_pack_build(C.byref(inp), C.byref(handle)))
    _native.check(lib.opv_pack_view_get(handle, C.byref(view)))
    assert view.needs_decode == 0 and view.n_slots == 2 and view.n_sentences == 3
    ids = host_pack._copy(view.h_ids, int(view.n_tokens), np.int32)
    offs = host_pack._copy(view.h_block_offsets, int(view.n_blocks) + 1, np.int64)
    # capacity = 16 - 2 = 14; base = 2 + 1: the 7-token window fills block 0, the 5-token one fits too (3+7+5 > 14?)
    blocks = [ids[offs[b] : offs[b + 1]].tolist() for b in range(int(view.n_blocks))]
    assert blocks == [[1, 5, 6, 2] + list(range(10, 17)) + [2], [1, 5, 6, 2] + list(range(25, 30)) + [2]]
    csr = host_pack._copy(view.h_sent_slot_offsets, 4, np.int32).tolist()
    assert csr == [0, 1, 1, 2]
    lib.opv_pack_destroy(handle)
