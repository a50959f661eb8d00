"""Numpy restatement of the reference forward (ModernBERT backbone + heads).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Not imported by the product.

Every function cites the code it restates:

* ``standalone:N`` = /root/reference/open_provence/modeling_open_provence_standalone.py:N
* ``encoder:N``    = /root/reference/open_provence/encoder.py:N
* ``HF:N``         = transformers 5.5.0 ``models/modernbert/modeling_modernbert.py:N``
  (third-party dependency of the reference, pinned 4.57.1 in its ``uv.lock:3640``;
  5.5.0 is what is installed and what the golden fixtures were generated with).

The arithmetic runs in the dtype of the weights handed in (float64 for the
"truth" comparand, float32 to mimic the reference CPU path).  Sequences are
processed unpadded and independently, which equals the padded HF result on
valid tokens (verified by ``tests/test_oracle_golden.py`` against fixtures that
were produced with right-padded batches).
"""

from __future__ import annotations

import math
from typing import Mapping, Sequence

import numpy as np

try:  # scipy is present in the image; erf is the only thing needed from it
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)


def layer_norm(x: np.ndarray, weight: np.ndarray, eps: float) -> np.ndarray:
    """``nn.LayerNorm(H, eps, bias=False)`` -- HF:60 (embeddings.norm), HF:321,323,432,499."""
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)  # biased variance
    return (x - mean) / np.sqrt(var + eps) * weight


def gelu_erf(x: np.ndarray) -> np.ndarray:
    """Exact GELU (``hidden_activation = classifier_activation = "gelu"``), HF:85,497."""
    return (0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))).astype(x.dtype)


def layer_is_global(cfg: Mapping, layer: int) -> bool:
    """configuration_modernbert.py:113-120: layer i is full attention iff i % every_n == 0."""
    layer_types = cfg.get("layer_types")
    if layer_types:
        return layer_types[layer] == "full_attention"
    every = int(cfg.get("global_attn_every_n_layers", 3))
    return layer % every == 0


def rope_thetas(cfg: Mapping) -> tuple[float, float]:
    """configuration_modernbert.py:77,141-148 (defaults 160000 global / 10000 local)."""
    rp = cfg.get("rope_parameters") or {}
    g = (rp.get("full_attention") or {}).get("rope_theta", cfg.get("global_rope_theta", 160000.0))
    l = (rp.get("sliding_attention") or {}).get("rope_theta", cfg.get("local_rope_theta", 10000.0))
    return float(g), float(l)


def rope_tables(n_pos: int, head_dim: int, theta: float) -> tuple[np.ndarray, np.ndarray]:
    """cos/sin tables exactly as HF:139-172 builds them: fp32 inv_freq, fp32 angles."""
    exponent = (np.arange(0, head_dim, 2, dtype=np.int64).astype(np.float32) / np.float32(head_dim))
    inv_freq = (np.float32(1.0) / np.power(np.float32(theta), exponent, dtype=np.float32)).astype(np.float32)
    pos = np.arange(n_pos, dtype=np.float32)
    freqs = (pos[:, None] * inv_freq[None, :]).astype(np.float32)  # HF:168 (fp32 matmul of outer product)
    emb = np.concatenate([freqs, freqs], axis=-1)
    return np.cos(emb).astype(np.float32), np.sin(emb).astype(np.float32)


def _rotate_half(x: np.ndarray) -> np.ndarray:
    """HF:197-201."""
    d = x.shape[-1] // 2
    return np.concatenate([-x[..., d:], x[..., :d]], axis=-1)


def forward_sequence(
    ids: Sequence[int],
    weights: Mapping[str, np.ndarray],
    cfg: Mapping,
    *,
    return_hidden: bool = False,
):
    """One unpadded sequence through backbone + both heads.

    Restates standalone:1666-1739 (``OpenProvenceModel.forward``; twin encoder:174-245)
    = HF:446-490 (model) + HF:313-342 (layer) + HF:232-310 (attention) + HF:74-91 (MLP)
    + HF:621-634,493-502 (sequence-classification head) + standalone:434-448 (pruning head).

    Returns ``(ranking_logits [num_labels], pruning_logits [n, 2])``.
    """
    W = weights
    p = "ranking_model."
    dt = W[p + "model.embeddings.tok_embeddings.weight"].dtype
    H = int(cfg["hidden_size"])
    nh = int(cfg["num_attention_heads"])
    d = H // nh
    L = int(cfg["num_hidden_layers"])
    eps = float(cfg.get("norm_eps", 1e-5))
    half_window = int(cfg.get("local_attention", 128)) // 2
    theta_g, theta_l = rope_thetas(cfg)
    ids = np.asarray(ids, dtype=np.int64)
    n = ids.shape[0]

    cos_g, sin_g = rope_tables(n, d, theta_g)
    cos_l, sin_l = rope_tables(n, d, theta_l)

    # HF:60-71 embeddings: LN(tok_emb[ids])
    h = layer_norm(W[p + "model.embeddings.tok_embeddings.weight"][ids], W[p + "model.embeddings.norm.weight"], eps)
    idx = np.arange(n)
    band = np.abs(idx[:, None] - idx[None, :]) <= half_window  # masking_utils.py:121-131

    for l in range(L):
        lp = f"{p}model.layers.{l}."
        x = h if l == 0 else layer_norm(h, W[lp + "attn_norm.weight"], eps)  # HF:318-321
        qkv = x @ W[lp + "attn.Wqkv.weight"].T  # HF:280
        qkv = qkv.reshape(n, 3, nh, d)  # HF:281-282 view(..., 3, heads, d).unbind(-3)
        q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
        is_global = layer_is_global(cfg, l)
        cos, sin = (cos_g, sin_g) if is_global else (cos_l, sin_l)
        cos_c = cos.astype(dt)[:, None, :]  # HF:172 cast to activation dtype
        sin_c = sin.astype(dt)[:, None, :]
        q = q * cos_c + _rotate_half(q) * sin_c  # HF:224-227
        k = k * cos_c + _rotate_half(k) * sin_c
        scores = np.einsum("ihd,jhd->hij", q, k) * (d ** -0.5)  # HF:185
        if not is_global:
            scores = np.where(band[None], scores, -np.inf)
        scores = scores - scores.max(axis=-1, keepdims=True)
        probs = np.exp(scores)
        probs = probs / probs.sum(axis=-1, keepdims=True)  # HF:189 softmax
        attn = np.einsum("hij,jhd->ihd", probs, v).reshape(n, H)
        h = h + attn @ W[lp + "attn.Wo.weight"].T  # HF:308-309,340
        y = layer_norm(h, W[lp + "mlp_norm.weight"], eps)
        u = y @ W[lp + "mlp.Wi.weight"].T
        I = u.shape[-1] // 2
        h = h + (gelu_erf(u[:, :I]) * u[:, I:]) @ W[lp + "mlp.Wo.weight"].T  # HF:90-91,341

    h = layer_norm(h, W[p + "model.final_norm.weight"], eps)  # HF:488 == hidden_states[-1]

    pooling = cfg.get("classifier_pooling", "cls")
    pooled = h[0] if pooling == "cls" else h.mean(axis=0)  # HF:621-630
    pooled = layer_norm(gelu_erf(pooled @ W[p + "head.dense.weight"].T), W[p + "head.norm.weight"], eps)  # HF:501
    rank = pooled @ W[p + "classifier.weight"].T + W[p + "classifier.bias"]  # HF:634
    prune = h @ W["pruning_head.classifier.weight"].T + W["pruning_head.classifier.bias"]  # standalone:446-447
    if return_hidden:
        return rank, prune, h
    return rank, prune


def forward_batch(
    sequences: Sequence[Sequence[int]],
    weights: Mapping[str, np.ndarray],
    cfg: Mapping,
) -> tuple[np.ndarray, list[np.ndarray]]:
    """Oracle for a batch of unpadded sequences -> (rank [B, num_labels], list of prune [n_i, 2])."""
    ranks, prunes = [], []
    for ids in sequences:
        r, pr = forward_sequence(ids, weights, cfg)
        ranks.append(r)
        prunes.append(pr)
    return np.stack(ranks), prunes


def cast_weights(weights: Mapping[str, np.ndarray], dtype) -> dict[str, np.ndarray]:
    return {k: np.asarray(v).astype(dtype) for k, v in weights.items()}


def algorithmic_flops_per_pair(cfg: Mapping, S: int) -> float:
    """SURVEY.md section 8(d) / BASELINE.md section 3: F(S), 1 MAC = 2 FLOP, valid tokens only."""
    H = int(cfg["hidden_size"])
    L = int(cfg["num_hidden_layers"])
    I = int(cfg["intermediate_size"])
    num_labels = int(cfg.get("num_labels", 1))
    half = int(cfg.get("local_attention", 128)) // 2
    n_glob = sum(1 for l in range(L) if layer_is_global(cfg, l))
    n_loc = L - n_glob
    i = np.arange(S)
    P = int((np.minimum(S - 1, i + half) - np.maximum(0, i - half) + 1).sum())
    return float(
        S * L * (8 * H * H + 6 * H * I)
        + n_glob * 4 * H * S * S
        + n_loc * 4 * H * P
        + 2 * H * H
        + 2 * H * num_labels
        + S * 4 * H
    )
