"""Host-side text handling for ``process()``: sentence splitting, sentence normalisation and the
token-level fragmentiser.  Pure Python / tokenizer work -- nothing here touches the GPU.

Behavioural contract (what must match the reference, checked by tests/test_process_host.py against
fixtures produced by the reference):
  * splitters                     standalone:1002-1143
  * sentence normalisation        standalone:582-661
  * fragment split / filter       standalone:686-713, 846-894
``standalone:N`` = /root/reference/open_provence/modeling_open_provence_standalone.py:N
"""

from __future__ import annotations

import math
import re
from dataclasses import dataclass
from typing import Any, Callable, Iterable, Mapping, Sequence

SentenceSplitter = Callable[[str], list[str]]

ENGLISH_SENTENCE_MAX_CHARS = 1200
_SIMPLE_SENTENCE_RE = re.compile(r".+?(?:。|！|？|!|\?|\n|$)", re.S)
_BULLET_RE = re.compile(r"^\s*(?:[\-\*••]+|\d{1,4}[:.)]|[A-Za-z]{1}[:.)])\s+", re.UNICODE)
_SENTENCE_END_CHARS = ".?!"


# ------------------------------------------------------------------------------------------------
# language detection + splitters
# ------------------------------------------------------------------------------------------------
def _is_kana(cp: int) -> bool:
    return (0x3041 <= cp <= 0x3096) or (0x30A1 <= cp <= 0x30FA) or (0x31F0 <= cp <= 0x31FF) or (0xFF71 <= cp <= 0xFF9D)


def is_japanese_fast(text: str, window: int = 500, min_kana_per_window: int = 1) -> bool:
    """Kana-density heuristic (standalone:135-155): Japanese iff >= ceil(len/window)*min kana letters."""
    if not text or text.isascii():
        return False
    needed = math.ceil(len(text) / window) * min_kana_per_window
    if needed <= 0:
        return False
    seen = 0
    for ch in text:
        cp = ord(ch)
        if cp > 0x7F and _is_kana(cp):
            seen += 1
            if seen >= needed:
                return True
    return False


def simple_sentence_splitter(text: str) -> list[str]:
    """Regex splitter on 。！？!? and newlines, delimiters kept (standalone:1018-1029)."""
    if not text:
        return []
    parts = [m for m in _SIMPLE_SENTENCE_RE.findall(text) if m]
    return parts or [text]


_FAST_BUNKAI = None


def fast_bunkai_sentence_splitter(text: str) -> list[str]:
    """Japanese splitter backed by fast-bunkai (optional dependency, standalone:1002-1015)."""
    global _FAST_BUNKAI
    if _FAST_BUNKAI is None:
        try:
            from fast_bunkai import FastBunkai
        except ImportError as exc:
            raise RuntimeError(
                "fast-bunkai is not installed. Install `fast-bunkai` or provide a custom sentence_splitter "
                "(e.g. `simple_sentence_splitter`)."
            ) from exc
        _FAST_BUNKAI = FastBunkai()
    parts = [s for s in _FAST_BUNKAI(text) if s]
    if parts:
        return parts
    return [text] if text else []


def _hard_wrap(sentence: str, max_chars: int, keep_whitespace: bool) -> list[str]:
    """Cut a sentence longer than max_chars, preferring newline then punctuation boundaries
    (standalone:532-579)."""
    work = sentence if keep_whitespace else sentence.strip()
    if not work:
        return []
    if len(work) <= max_chars:
        return [work]
    out: list[str] = []
    pos, total = 0, len(work)
    while pos < total:
        limit = min(pos + max_chars, total)
        cut = None
        nl = work.rfind("\n", pos + 1, limit)
        if nl != -1:
            cut = nl + 1
        if cut is None or cut <= pos:
            for j in range(limit, pos, -1):
                if work[j - 1] in ".?!;:\n":
                    cut = j
                    break
        if cut is None or cut <= pos:
            cut = limit
        piece = work[pos:cut]
        if not keep_whitespace:
            piece = piece.strip()
        if piece:
            out.append(piece)
        pos = cut
    return out or [work]


def _bullet_blocks(text: str) -> Iterable[tuple[str, int, int]]:
    """Group lines into blocks, starting a new block at every bullet-like line (standalone:485-529)."""
    if not text:
        return
    lines = text.splitlines(keepends=True)
    if not lines:
        yield text, 0, len(text)
        return
    consumed = 0
    parts: list[str] = []
    start = 0
    for line in lines:
        line_start = consumed
        consumed += len(line)
        if _BULLET_RE.match(line.rstrip("\r\n")) and parts:
            block = "".join(parts)
            if block:
                yield block, start, start + len(block)
            parts, start = [line], line_start
        else:
            if not parts:
                start = line_start
            parts.append(line)
    if parts:
        block = "".join(parts)
        if block:
            yield block, start, start + len(block)
    if consumed < len(text) and text[consumed:]:
        yield text[consumed:], consumed, len(text)


_PUNKT = None


def _punkt():
    global _PUNKT
    if _PUNKT is None:
        try:
            import nltk
        except ImportError as exc:
            raise RuntimeError(
                "nltk is not installed: the English sentence splitter needs NLTK punkt. Install `nltk` or pass "
                "a custom sentence_splitter (e.g. `simple_sentence_splitter`)."
            ) from exc
        try:
            _PUNKT = nltk.data.load("tokenizers/punkt/english.pickle")
        except LookupError as exc:
            raise LookupError("Missing NLTK punkt tokenizer data. Run `python -m nltk.downloader punkt`.") from exc
    return _PUNKT


def create_english_sentence_splitter(max_chars: int = ENGLISH_SENTENCE_MAX_CHARS) -> SentenceSplitter:
    """Punkt-based splitter that preserves whitespace/newlines and wraps overlong sentences
    (standalone:1032-1117)."""
    if max_chars <= 0:
        raise ValueError("max_chars must be positive")

    def split(text: str) -> list[str]:
        if not text:
            return []
        punkt = _punkt()
        out: list[str] = []
        for block, b0, b1 in _bullet_blocks(text):
            if not block:
                continue
            spans = list(punkt.span_tokenize(block))
            if not spans:
                seg = text[b0:b1]
                if seg.strip():
                    out.extend(_hard_wrap(seg, max_chars, True))
                continue
            for s0, s1 in spans:
                g0, g1 = b0 + s0, b0 + s1
                while g1 < b1 and text[g1].isspace():
                    g1 += 1
                seg = text[g0:g1]
                if seg and seg.strip():
                    out.extend(_hard_wrap(seg, max_chars, True))
        if out:
            return out
        stripped = text.strip()
        return [stripped] if stripped else []

    return split


_DEFAULT_ENGLISH = create_english_sentence_splitter()


def english_sentence_splitter(text: str) -> list[str]:
    return _DEFAULT_ENGLISH(text)


def create_auto_sentence_splitter(
    *,
    japanese_splitter: SentenceSplitter = fast_bunkai_sentence_splitter,
    english_splitter: SentenceSplitter = english_sentence_splitter,
    kana_window: int = 500,
    min_kana_per_window: int = 1,
) -> SentenceSplitter:
    """Kana density picks the Japanese or the English splitter (standalone:1129-1143)."""

    def split(text: str) -> list[str]:
        if is_japanese_fast(text, window=kana_window, min_kana_per_window=min_kana_per_window):
            return japanese_splitter(text)
        return english_splitter(text)

    return split


def resolve_sentence_splitter(
    splitter: SentenceSplitter | Mapping[str, SentenceSplitter] | None,
    language: str | None,
    default_language: str | None = "auto",
) -> SentenceSplitter:
    """standalone:2007-2039 (same error messages)."""
    if isinstance(splitter, Mapping):
        if language is None:
            raise ValueError("language must be provided when sentence_splitter is a mapping")
        if language in splitter:
            return splitter[language]
        raise ValueError(f"No sentence splitter registered for language '{language}'")
    if callable(splitter):
        return splitter
    lang = language if language is not None else default_language
    if lang is None:
        lang = "auto"
    key = str(lang).lower()
    if key == "auto":
        return create_auto_sentence_splitter()
    if key == "ja":
        return fast_bunkai_sentence_splitter
    if key == "en":
        return english_sentence_splitter
    raise ValueError(
        f"Unsupported language code for sentence splitting: '{lang}'. Supported values are 'auto', 'en', and 'ja'."
    )


# ------------------------------------------------------------------------------------------------
# sentence normalisation
# ------------------------------------------------------------------------------------------------
def _split_on_lines(text: str, strip: bool) -> list[str]:
    """A multi-line "sentence" without enough end punctuation becomes one sentence per line
    (standalone:582-612)."""
    whole = [text.strip() if strip else text]
    if "\n" not in text:
        return whole
    lines = [seg for seg in text.splitlines(keepends=not strip) if seg.strip()]
    if len(lines) <= 1:
        return whole
    if sum(1 for ch in text if ch in _SENTENCE_END_CHARS) >= len(lines):
        return whole
    if any(len(seg.strip()) > ENGLISH_SENTENCE_MAX_CHARS for seg in lines):
        return whole
    pieces = [seg.strip() for seg in lines] if strip else lines
    pieces = [p for p in pieces if p]
    return pieces or whole


def normalize_sentences(raw: Sequence[str], context_text: str, strip: bool) -> list[str]:
    """standalone:640-661: drop empty entries, expand multi-line entries, fall back to the whole context."""
    out: list[str] = []
    for entry in raw:
        text = str(entry)
        if not text:
            continue
        out.extend(seg for seg in _split_on_lines(text, strip) if seg)
    if out:
        return out
    if not strip:
        return [context_text]
    return [context_text.strip() or context_text]


def fallback_sentence(context_text: str, strip: bool) -> str:
    if not strip:
        return context_text
    return context_text.strip() or context_text


# ------------------------------------------------------------------------------------------------
# fragments
# ------------------------------------------------------------------------------------------------
@dataclass
class Fragment:
    """A run of at most ``max_fragment_tokens`` tokens of one sentence (reference ``_FragmentRecord``,
    standalone:990-999; the decoded text is only needed to drop empty fragments)."""

    token_ids: list[int]
    sentence_index: int
    fragment_index: int
    global_index: int

    @property
    def token_length(self) -> int:
        return len(self.token_ids)


def split_token_lists(
    token_lists: Sequence[Sequence[int]], max_fragment_tokens: int, *, keep_sentence_boundaries: bool = False
) -> list[Fragment]:
    """standalone:686-713: every sentence is cut into consecutive windows of ``max_fragment_tokens``."""
    step = max(1, int(max_fragment_tokens))
    out: list[Fragment] = []
    for s_idx, ids in enumerate(token_lists):
        ids = list(ids)
        if not ids:
            continue
        if keep_sentence_boundaries and len(ids) <= max_fragment_tokens:
            out.append(Fragment(ids, s_idx, 0, len(out)))
            continue
        for f_idx, at in enumerate(range(0, len(ids), step)):
            out.append(Fragment(ids[at : at + step], s_idx, f_idx, len(out)))
    return out


def filter_decodable(fragments: Sequence[Fragment], texts: Sequence[str], strip: bool) -> list[Fragment]:
    """standalone:874-886: drop fragments whose decoded text is empty (whitespace-only when strip)."""
    kept = []
    for frag, text in zip(fragments, texts):
        if strip:
            if not text.strip():
                continue
        elif not text:
            continue
        kept.append(frag)
    return kept


def _rust_backend(tokenizer: Any) -> Any:
    """The ``tokenizers.Tokenizer`` behind a fast HF tokenizer, or None (slow / stub tokenizers)."""
    if not getattr(tokenizer, "is_fast", False):
        return None
    backend = getattr(tokenizer, "backend_tokenizer", None)
    return backend if hasattr(backend, "encode_batch") and hasattr(backend, "decode_batch") else None


def tokenize_batch(tokenizer: Any, sentences: Sequence[str]) -> list[list[int]]:
    """``tokenizer(list, add_special_tokens=False)`` (standalone:664-672) in one call.

    With a fast tokenizer the Rust ``encode_batch`` is called directly: same ids, without the per-row
    ``BatchEncoding`` bookkeeping that dominated host time at thousands of sentences per call."""
    if not sentences:
        return []
    backend = _rust_backend(tokenizer)
    if backend is not None:
        # HF fast tokenizers leave the truncation / padding of their LAST call on the Rust object and never restore
        # it; ``encode_batch`` would then silently pad every sentence to the longest one or cut it at max_length.
        # The reference's ``tokenizer(list, add_special_tokens=False)`` resets that state on every call -- so do we.
        try:
            if backend.truncation is not None:
                backend.no_truncation()
            if backend.padding is not None:
                backend.no_padding()
            return [enc.ids for enc in backend.encode_batch(list(sentences), add_special_tokens=False)]
        except Exception:  # a backend without these controls: fall back to the HF call
            pass
    enc = tokenizer(list(sentences), add_special_tokens=False, return_attention_mask=False)
    ids = enc.get("input_ids", []) if isinstance(enc, Mapping) or hasattr(enc, "get") else []
    return [[int(t) for t in row] for row in ids]


def decode_batch(tokenizer: Any, token_lists: Sequence[Sequence[int]]) -> list[str]:
    """``tokenizer.batch_decode(..., skip_special_tokens=True, clean_up_tokenization_spaces=False)``
    (standalone:846-870), through the Rust ``decode_batch`` when the tokenizer is a fast one."""
    if not token_lists:
        return []
    backend = _rust_backend(tokenizer)
    if backend is not None:
        try:
            return list(backend.decode_batch([list(t) for t in token_lists], skip_special_tokens=True))
        except Exception:
            pass
    return list(tokenizer.batch_decode([list(t) for t in token_lists], skip_special_tokens=True,
                                       clean_up_tokenization_spaces=False))
