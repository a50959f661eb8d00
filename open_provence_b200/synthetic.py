"""Synthetic checkpoints and pre-tokenised workloads for the benchmark and the
full-size parity tests.  There is no network: weights are random-initialised with the distribution
HF's ``ModernBertPreTrainedModel._init_weights`` uses (HF:361-420), data is uniform random token ids.
"""

from __future__ import annotations

import math
from typing import Any

import numpy as np
import torch

# dims from the public model cards; the loader reads real checkpoints' dims from
# config.json -- these are only for synthetic runs.
MODEL_DIMS: dict[str, dict[str, int]] = {
    "tiny": dict(hidden_size=128, num_hidden_layers=4, intermediate_size=128, num_attention_heads=2, vocab_size=264),
    "xsmall-30M": dict(hidden_size=256, num_hidden_layers=10, intermediate_size=1024, num_attention_heads=4, vocab_size=102400),
    "base-130M": dict(hidden_size=512, num_hidden_layers=19, intermediate_size=2048, num_attention_heads=8, vocab_size=102400),
    "large-310M": dict(hidden_size=768, num_hidden_layers=25, intermediate_size=3072, num_attention_heads=12, vocab_size=102400),
    "en-gte-149M": dict(hidden_size=768, num_hidden_layers=22, intermediate_size=1152, num_attention_heads=12, vocab_size=50368),
}

CLS_ID, SEP_ID, PAD_ID = 1, 2, 0


def backbone_config(name: str) -> dict[str, Any]:
    if name not in MODEL_DIMS:
        raise KeyError(f"unknown synthetic model {name!r}; choose from {sorted(MODEL_DIMS)}")
    cfg: dict[str, Any] = dict(MODEL_DIMS[name])
    cfg.update(
        model_type="modernbert",
        max_position_embeddings=8192,
        local_attention=128,
        global_attn_every_n_layers=3,
        norm_eps=1e-5,
        pad_token_id=PAD_ID,
        bos_token_id=CLS_ID,
        cls_token_id=CLS_ID,
        eos_token_id=SEP_ID,
        sep_token_id=SEP_ID,
        classifier_pooling="cls",
        initializer_range=0.02,
        initializer_cutoff_factor=2.0,
    )
    return cfg


def _trunc_normal(shape, std: float, cutoff: float, gen: torch.Generator) -> torch.Tensor:
    t = torch.empty(shape, dtype=torch.float32)
    torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-cutoff * std, b=cutoff * std, generator=gen)
    return t


def random_state_dict(cfg: dict[str, Any], seed: int = 0, num_labels: int = 1) -> dict[str, torch.Tensor]:
    """fp32 CPU state dict with the reference's key names.

    Stds follow HF:379-384 (in = 0.02, out = 0.02/sqrt(2L), final_out = H^-0.5).  The pruning head is
    N(0, 1.06^2 / H) with bias (0, logit(0.05)) so per-sentence means straddle the 0.1 threshold (section 8d).
    LayerNorm gains get a small jitter so a dropped gain cannot go unnoticed.
    """
    gen = torch.Generator().manual_seed(seed)
    H, L, I, V = cfg["hidden_size"], cfg["num_hidden_layers"], cfg["intermediate_size"], cfg["vocab_size"]
    std_in = float(cfg.get("initializer_range", 0.02))
    std_out = std_in / math.sqrt(2.0 * L)
    cutoff = float(cfg.get("initializer_cutoff_factor", 2.0) or 3.0)
    p = "ranking_model."

    def gain(n):
        return 1.0 + 0.05 * torch.randn(n, generator=gen)

    sd: dict[str, torch.Tensor] = {}
    sd[p + "model.embeddings.tok_embeddings.weight"] = _trunc_normal((V, H), std_in, cutoff, gen)
    sd[p + "model.embeddings.norm.weight"] = gain(H)
    for l in range(L):
        lp = f"{p}model.layers.{l}."
        if l > 0:
            sd[lp + "attn_norm.weight"] = gain(H)
        sd[lp + "attn.Wqkv.weight"] = _trunc_normal((3 * H, H), std_in, cutoff, gen)
        sd[lp + "attn.Wo.weight"] = _trunc_normal((H, H), std_out, cutoff, gen)
        sd[lp + "mlp_norm.weight"] = gain(H)
        sd[lp + "mlp.Wi.weight"] = _trunc_normal((2 * I, H), std_in, cutoff, gen)
        sd[lp + "mlp.Wo.weight"] = _trunc_normal((H, I), std_out, cutoff, gen)
    sd[p + "model.final_norm.weight"] = gain(H)
    sd[p + "head.dense.weight"] = _trunc_normal((H, H), std_out, cutoff, gen)
    sd[p + "head.norm.weight"] = gain(H)
    sd[p + "classifier.weight"] = _trunc_normal((num_labels, H), H**-0.5, cutoff, gen)
    sd[p + "classifier.bias"] = torch.zeros(num_labels)
    # keep-logit margin ~ N(log(0.05 / 0.95), 1.5^2) per token: token keep-probabilities spread over (0, 0.5) and the
    # per-sentence means land on BOTH sides of the default 0.1 threshold (r1 used N(0, 0.5^2) weights: |logit| ~ 40,
    # saturated probabilities, 99.9 % of the sentences kept -- a weak check of the keep decisions)
    sd["pruning_head.classifier.weight"] = torch.randn(2, H, generator=gen) * (1.06 / math.sqrt(H))
    sd["pruning_head.classifier.bias"] = torch.tensor([0.0, math.log(0.05 / 0.95)])
    return sd


def make_workload(cfg: dict[str, Any], n_pairs: int, seq_len: int, *, mode: str = "dense", seed: int = 1234) -> dict[str, Any]:
    """Pre-tokenised (question, context) blocks: ``[CLS] q [SEP] ctx [SEP]``.

    dense: every block has exactly ``seq_len`` tokens (roofline runs); ragged: U{seq_len/2 .. seq_len}.
    The context is partitioned into sentences of U{8..48} tokens; each sentence is one fragment.
    Returns numpy arrays: ids int32 [T], cu_seqlens int32 [n+1], frag_ranges int32 [F, 2] (packed-token
    coordinates), sent_offsets int32 [F+1], sent_frag_index int32 [F], frag_block int32 [F].
    """
    rng = np.random.default_rng(seed)
    V = int(cfg["vocab_size"])
    ids_parts, lengths, ranges, frag_block = [], [], [], []
    offset = 0
    for b in range(n_pairs):
        n = seq_len if mode == "dense" else int(rng.integers(seq_len // 2, seq_len + 1))
        n = max(n, 8)
        q_len = int(min(rng.integers(16, 33), max(1, n - 4)))
        row = rng.integers(8, V, size=n, dtype=np.int64).astype(np.int32)
        row[0] = CLS_ID
        row[1 + q_len] = SEP_ID
        row[n - 1] = SEP_ID
        ctx_start, ctx_end = 2 + q_len, n - 1
        at = ctx_start
        while at < ctx_end:
            step = int(rng.integers(8, 49))
            end = min(ctx_end, at + step)
            ranges.append((offset + at, offset + end))
            frag_block.append(b)
            at = end
        ids_parts.append(row)
        lengths.append(n)
        offset += n
    n_frags = len(ranges)
    return {
        "ids": np.concatenate(ids_parts).astype(np.int32),
        "cu_seqlens": np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32),
        "lengths": np.asarray(lengths, dtype=np.int32),
        "max_seqlen": int(max(lengths)),
        "frag_ranges": np.asarray(ranges, dtype=np.int32).reshape(n_frags, 2),
        "sent_offsets": np.arange(n_frags + 1, dtype=np.int32),
        "sent_frag_index": np.arange(n_frags, dtype=np.int32),
        "frag_block": np.asarray(frag_block, dtype=np.int32),
        "n_pairs": n_pairs,
        "seq_len": seq_len,
        "mode": mode,
    }


def algorithmic_flops_per_pair(cfg: dict[str, Any], S: int, num_labels: int = 1) -> float:
    """SURVEY.md section 8(d) / BASELINE.md section 3: F(S) = S.L.(8H^2 + 6HI) + n_glob.4H.S^2 + n_loc.4H.P(S)
    + 2H^2 + 2H.num_labels + 4HS, with P(S) the exact band pair count; 1 MAC = 2 FLOP, valid tokens only."""
    H, L, I = int(cfg["hidden_size"]), int(cfg["num_hidden_layers"]), int(cfg["intermediate_size"])
    half = int(cfg.get("local_attention", 128)) // 2
    every = int(cfg.get("global_attn_every_n_layers", 3))
    layer_types = cfg.get("layer_types")
    n_glob = sum(1 for l in range(L) if (layer_types[l] == "full_attention" if layer_types else l % every == 0))
    i = np.arange(S)
    band = int((np.minimum(S - 1, i + half) - np.maximum(0, i - half) + 1).sum())
    return float(S * L * (8 * H * H + 6 * H * I) + n_glob * 4 * H * S * S + (L - n_glob) * 4 * H * band
                 + 2 * H * H + 2 * H * num_labels + S * 4 * H)


def slice_workload(wl: dict[str, Any], blocks) -> dict[str, Any]:
    """The sub-workload made of ``blocks`` (indices into ``wl``, kept in the given order), re-packed: same keys as
    :func:`make_workload`.  Used to shard a fixed global block list over ranks / launches (strong scaling)."""
    blocks = np.asarray(blocks, dtype=np.int64)
    cu = wl["cu_seqlens"].astype(np.int64)
    lengths = wl["lengths"][blocks]
    new_cu = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    ids = np.concatenate([wl["ids"][cu[b] : cu[b + 1]] for b in blocks]) if blocks.size else np.zeros(0, np.int32)
    ranges, frag_block = [], []
    order = np.argsort(wl["frag_block"], kind="stable")
    first = np.searchsorted(wl["frag_block"][order], np.arange(len(cu)))
    for k, b in enumerate(blocks):
        sel = order[first[b] : first[b + 1]]
        ranges.append(wl["frag_ranges"][sel].astype(np.int64) - cu[b] + new_cu[k])
        frag_block.append(np.full(sel.size, k, dtype=np.int32))
    ranges = np.concatenate(ranges) if ranges else np.zeros((0, 2), np.int64)
    n_frags = int(ranges.shape[0])
    return {
        "ids": ids.astype(np.int32),
        "cu_seqlens": new_cu.astype(np.int32),
        "lengths": lengths.astype(np.int32),
        "max_seqlen": int(lengths.max()) if blocks.size else 0,
        "frag_ranges": ranges.astype(np.int32).reshape(n_frags, 2),
        "sent_offsets": np.arange(n_frags + 1, dtype=np.int32),
        "sent_frag_index": np.arange(n_frags, dtype=np.int32),
        "frag_block": np.concatenate(frag_block).astype(np.int32) if frag_block else np.zeros(0, np.int32),
        "n_pairs": int(blocks.size),
        "seq_len": wl["seq_len"],
        "mode": wl["mode"],
        "global_blocks": blocks.astype(np.int64),
    }
