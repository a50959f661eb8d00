"""Numpy restatement of the reference's score conversion and per-sentence prune.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Not imported by the product.

``standalone:N`` = /root/reference/open_provence/modeling_open_provence_standalone.py:N
"""

from __future__ import annotations

from collections import defaultdict
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np


def ranking_score_from_logits(rank_logits: np.ndarray) -> float:
    """standalone:2913-2916 -- sigmoid of the first (or only) ranking logit, fp32."""
    x = np.float32(np.asarray(rank_logits, dtype=np.float32).reshape(-1)[0])
    return float(np.float32(1.0) / (np.float32(1.0) + np.exp(-x, dtype=np.float32)))


def keep_probs_from_logits(prune_logits: np.ndarray) -> np.ndarray:
    """standalone:2918-2920 -- fp32 ``softmax(logits, -1)[:, 1]``."""
    lg = np.asarray(prune_logits, dtype=np.float32)
    m = lg.max(axis=-1, keepdims=True)
    e = np.exp(lg - m, dtype=np.float32)
    return (e[:, 1] / (e[:, 0] + e[:, 1])).astype(np.float32)


@dataclass
class OracleBlock:
    """One inference block of a context: the fragments packed into it."""

    keep_probs: np.ndarray  # fp32 [n_tokens] (standalone:2918)
    ranges: Sequence[tuple[int, int]]  # block-local token ranges (standalone:2172-2184)
    frag_global_index: Sequence[int]
    frag_sentence_index: Sequence[int]
    ranking_score: float | None


@dataclass
class OracleContext:
    """What ``contexts_info[(q, c)]`` carries in the reference (standalone:2736-2746)."""

    n_sentences: int
    prefix_length: int
    prefix_token_counts: Sequence[int]
    title_is_first_sentence: bool
    frag_table: Sequence[tuple[int, int]]  # every fragment: (global_index, sentence_index)
    blocks: list[OracleBlock] = field(default_factory=list)


def postprocess_context(
    ctx: OracleContext,
    *,
    threshold: float,
    always_select_title: bool = False,
    use_best_reranker_score: bool = True,
) -> dict:
    """standalone:3065-3136: fragment means -> sentence means -> keep flags, score = max over blocks."""
    fragment_scores: dict[int, list[float]] = defaultdict(list)
    ranking_score: float | None = None
    for block in ctx.blocks:
        probs = block.keep_probs
        for g_idx, s_idx, (start, end) in zip(block.frag_global_index, block.frag_sentence_index, block.ranges):
            offset = sum(ctx.prefix_token_counts[:s_idx])  # standalone:3076 (title quirk)
            start = max(0, start - offset)
            end = max(start, end - offset)
            end = min(end, len(probs))
            start = min(start, len(probs))
            mean_prob = 1.0 if end <= start else float(probs[start:end].mean())  # numpy fp32 pairwise mean
            fragment_scores[g_idx].append(mean_prob)
        if block.ranking_score is not None:
            if use_best_reranker_score:
                ranking_score = block.ranking_score if ranking_score is None else max(ranking_score, block.ranking_score)
            elif ranking_score is None:
                ranking_score = block.ranking_score

    sentence_scores: dict[int, list[float]] = defaultdict(list)
    for g_idx, s_idx in ctx.frag_table:  # standalone:3094-3099
        if g_idx in fragment_scores:
            sentence_scores[s_idx].extend(fragment_scores[g_idx])

    title_sentence_index = None
    if always_select_title:  # standalone:3108-3112
        if ctx.prefix_length > 0:
            title_sentence_index = 0
        elif ctx.title_is_first_sentence and ctx.n_sentences > ctx.prefix_length:
            title_sentence_index = ctx.prefix_length

    probs_out: list[float] = []
    any_above = False
    for s in range(ctx.n_sentences):  # standalone:3116-3122
        vals = sentence_scores.get(s)
        avg = float(np.mean(vals)) if vals else 0.0
        avg = max(0.0, min(avg, 1.0))
        probs_out.append(avg)
        any_above = any_above or avg > threshold
    force_title = title_sentence_index is not None and any_above
    keep = []
    for s in range(ctx.n_sentences):  # standalone:3128-3134
        flag = probs_out[s] > threshold
        if force_title and s == title_sentence_index:
            flag = True
        keep.append(bool(flag))
    return {"sentence_probabilities": probs_out, "keep": keep, "ranking_score": ranking_score}
