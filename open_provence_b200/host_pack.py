"""Native block assembly for ``process()``: the ``opv_pack_*`` entry points of libopv_sm100.so.

Everything between the tokenizer and the device -- fragment windows, the empty-fragment filter, greedy block
packing, ``[CLS] q [SEP] ctx [SEP]`` ids, fragment ranges, the sentence -> fragment CSR -- runs in C++ over
flat int32 arrays (``csrc/host_pack.cu``; reference: standalone:686-713, 846-894, 2104-2259, 3076-3099) and
comes back as the packed :class:`~open_provence_b200.scoring.BlockTable` the scorer uploads as is.

``standalone:N`` = /root/reference/open_provence/modeling_open_provence_standalone.py:N
"""

from __future__ import annotations

import ctypes as C
import weakref
from itertools import chain
from typing import Any, Sequence

import numpy as np

from . import _native
from .host_text import _rust_backend, decode_batch
from .scoring import BlockTable

_VISIBLE: "weakref.WeakKeyDictionary[Any, np.ndarray]" = weakref.WeakKeyDictionary()


def token_visibility(tokenizer: Any) -> np.ndarray | None:
    """uint8 [vocab]: 1 where the token, decoded on its own with special tokens skipped, shows a character that
    is neither whitespace nor U+FFFD.  A fragment holding such a token cannot decode to an empty (or
    whitespace-only) string, so the reference's filter (standalone:874-886) keeps it whatever else it holds;
    only fragments without one are decoded.  None for tokenizers without a Rust backend (every fragment is
    decoded then, as the reference does)."""
    backend = _rust_backend(tokenizer)
    if backend is None:
        return None
    try:
        cached = _VISIBLE.get(backend)
    except TypeError:  # backend not weak-referenceable
        cached = getattr(tokenizer, "_opv_token_visible", None)
    if cached is not None:
        return cached
    vocab = int(backend.get_vocab_size(with_added_tokens=True))
    texts = backend.decode_batch([[i] for i in range(vocab)], skip_special_tokens=True)
    table = np.fromiter((1 if t.replace("\ufffd", "").strip() else 0 for t in texts), dtype=np.uint8, count=vocab)
    try:
        _VISIBLE[backend] = table
    except TypeError:
        try:
            tokenizer._opv_token_visible = table
        except Exception:
            pass
    return table


def special_token_template(tokenizer: Any, manual: bool, cls_id: int | None, sep_id: int | None):
    """``(head, mid, tail)`` with block ids = head + query + mid + context + tail, or None when the tokenizer's
    ``build_inputs_with_special_tokens`` is not of that form (standalone:2104-2143).  The manual path
    (ModernBERT-EN tokenizers, standalone:2123-2135) is that form by construction."""
    if manual:
        head = [cls_id] if cls_id is not None else []
        mid = [sep_id] if sep_id is not None else []
        return head, mid, list(mid)
    build = getattr(tokenizer, "build_inputs_with_special_tokens", None)
    if not callable(build):
        return None
    special = set(int(t) for t in (getattr(tokenizer, "all_special_ids", None) or []))
    probe = [t for t in range(0, 4096) if t not in special][-7:]
    if len(probe) < 7:
        return None
    found = None
    for q, c in ((probe[:2], probe[2:3]), (probe[3:4], probe[4:7])):
        try:
            ids = [int(t) for t in build(list(q), list(c))]
        except Exception:
            return None
        n = len(ids) - len(q) - len(c)
        if n < 0:
            return None
        parts = None
        for i in range(0, n + 1):  # head length
            if ids[i : i + len(q)] != q:
                continue
            for j in range(i + len(q), len(ids) - len(c) + 1):
                if ids[j : j + len(c)] == c:
                    parts = (ids[:i], ids[i + len(q) : j], ids[j + len(c) :])
                    break
            if parts:
                break
        if parts is None or (found is not None and parts != found):
            return None
        found = parts
    if found is None or any(t in probe for part in found for t in part):
        return None
    return found


def _as_ptr(array: np.ndarray | None) -> int | None:
    return None if array is None else array.ctypes.data


def _copy(ptr: int | None, n: int, dtype: Any) -> np.ndarray:
    if not ptr or n <= 0:
        return np.zeros(0, dtype=dtype)
    out = np.empty(n, dtype=dtype)
    C.memmove(out.ctypes.data, ptr, out.nbytes)
    return out


class PackedBlocks:
    """Result of :func:`pack_blocks`: the table plus, per context, its block range."""

    __slots__ = ("table", "ctx_block_offsets", "ctx_sentence_base", "decoded_fragments")

    def __init__(self, table: BlockTable, ctx_block_offsets: np.ndarray, ctx_sentence_base: np.ndarray, decoded: int):
        self.table = table
        self.ctx_block_offsets = ctx_block_offsets
        self.ctx_sentence_base = ctx_sentence_base
        self.decoded_fragments = decoded


def pack_blocks(
    tokenizer: Any,
    token_lists: Sequence[Sequence[int]],
    ctx_sentence_counts: Sequence[int],
    ctx_query: Sequence[int],
    ctx_prefix: Sequence[int],
    query_tokens: Sequence[Sequence[int]],
    *,
    template: tuple[Sequence[int], Sequence[int], Sequence[int]],
    max_length: int,
    max_fragment_tokens: int,
    keep_sentence_boundaries: bool,
    sep_len: int,
    strip_sentences: bool,
) -> PackedBlocks:
    """One ``opv_pack_build`` call for a chunk of contexts.  ``token_lists`` are the tokenised sentences of
    all contexts in order; ``ctx_sentence_counts[c]`` of them belong to context ``c``."""
    lib = _native.load()
    n_ctx = len(ctx_sentence_counts)
    sent_len = np.fromiter(map(len, token_lists), dtype=np.int64, count=len(token_lists))
    sent_offsets = np.zeros(len(token_lists) + 1, dtype=np.int64)
    np.cumsum(sent_len, out=sent_offsets[1:])
    tokens = np.fromiter(chain.from_iterable(token_lists), dtype=np.int32, count=int(sent_offsets[-1]))
    ctx_sent_offsets = np.zeros(n_ctx + 1, dtype=np.int64)
    np.cumsum(np.asarray(ctx_sentence_counts, dtype=np.int64), out=ctx_sent_offsets[1:])
    if int(ctx_sent_offsets[-1]) != len(token_lists):
        raise ValueError("ctx_sentence_counts does not add up to the number of token lists")
    q_len = np.fromiter(map(len, query_tokens), dtype=np.int64, count=len(query_tokens))
    q_offsets = np.zeros(len(query_tokens) + 1, dtype=np.int64)
    np.cumsum(q_len, out=q_offsets[1:])
    q_tokens = np.fromiter(chain.from_iterable(query_tokens), dtype=np.int32, count=int(q_offsets[-1]))
    ctx_query_a = np.ascontiguousarray(ctx_query, dtype=np.int32)
    ctx_prefix_a = np.ascontiguousarray(ctx_prefix, dtype=np.int32)
    head, mid, tail = (np.ascontiguousarray(part, dtype=np.int32) for part in template)
    visible = token_visibility(tokenizer)

    inp = _native.OpvPackInput(
        abi_version=_native.OPV_ABI_VERSION, max_length=int(max_length), max_fragment_tokens=int(max_fragment_tokens),
        keep_sentence_boundaries=int(bool(keep_sentence_boundaries)), sep_len=int(sep_len), n_contexts=n_ctx,
        n_queries=len(query_tokens), vocab_size=0 if visible is None else int(visible.shape[0]),
        n_head=int(head.shape[0]), n_mid=int(mid.shape[0]), n_tail=int(tail.shape[0]),
        h_head=_as_ptr(head), h_mid=_as_ptr(mid), h_tail=_as_ptr(tail), h_tokens=_as_ptr(tokens),
        h_sent_offsets=_as_ptr(sent_offsets), h_ctx_sent_offsets=_as_ptr(ctx_sent_offsets),
        h_ctx_query=_as_ptr(ctx_query_a), h_ctx_prefix=_as_ptr(ctx_prefix_a), h_query_tokens=_as_ptr(q_tokens),
        h_query_offsets=_as_ptr(q_offsets), h_token_visible=_as_ptr(visible), h_frag_drop=None)
    handle = C.c_void_p()
    view = _native.OpvPackView()
    decoded = 0
    _native.check(lib.opv_pack_build(C.byref(inp), C.byref(handle)), "opv_pack_build")
    try:
        _native.check(lib.opv_pack_view_get(handle, C.byref(view)), "opv_pack_view_get")
        if view.needs_decode:
            # fragments no visible token settles: ask the tokenizer, exactly as standalone:864-886 does
            n_raw = int(view.n_raw_fragments)
            uncertain = np.nonzero(_copy(view.h_raw_uncertain, n_raw, np.uint8))[0]
            start = _copy(view.h_raw_start, n_raw, np.int64)
            length = _copy(view.h_raw_len, n_raw, np.int32)
            texts = decode_batch(tokenizer, [tokens[start[i] : start[i] + length[i]].tolist() for i in uncertain])
            drop = np.zeros(n_raw, dtype=np.uint8)
            for i, text in zip(uncertain, texts):
                drop[i] = 0 if (text.strip() if strip_sentences else text) else 1
            decoded = int(uncertain.shape[0])
            lib.opv_pack_destroy(handle)
            handle = C.c_void_p()
            inp.h_frag_drop = _as_ptr(drop)
            _native.check(lib.opv_pack_build(C.byref(inp), C.byref(handle)), "opv_pack_build")
            _native.check(lib.opv_pack_view_get(handle, C.byref(view)), "opv_pack_view_get")
        n_blocks, n_slots, n_sent = int(view.n_blocks), int(view.n_slots), int(view.n_sentences)
        ids = _copy(view.h_ids, int(view.n_tokens), np.int32)
        block_offsets = _copy(view.h_block_offsets, n_blocks + 1, np.int64)
        table = BlockTable(
            block_ids=[ids[block_offsets[b] : block_offsets[b + 1]] for b in range(n_blocks)],
            frag_block=_copy(view.h_frag_block, n_slots, np.int32),
            frag_local=_copy(view.h_frag_local, 2 * n_slots, np.int32).reshape(-1, 2),
            sent_offsets=_copy(view.h_sent_slot_offsets, n_sent + 1, np.int32),
            sent_frag_index=_copy(view.h_sent_slot_index, n_slots, np.int32),
        )
        table.packed_ids = ids
        table.block_offsets = block_offsets
        ctx_block_offsets = _copy(view.h_ctx_block_offsets, n_ctx + 1, np.int64)
    finally:
        lib.opv_pack_destroy(handle)
    return PackedBlocks(table, ctx_block_offsets, ctx_sent_offsets[:-1].copy(), decoded)
