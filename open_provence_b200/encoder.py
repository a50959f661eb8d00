"""B200-native stand-in for the reference's ``OpenProvenceEncoder`` inference APIs
(``/root/reference/open_provence/encoder.py``: ``forward`` 174-245, ``predict`` 247-343,
``predict_with_pruning`` 345-529, ``predict_context`` 682-839, ``prune`` 901-938, ``prune_texts`` 940-991,
``save_pretrained`` / ``from_pretrained`` 1040-1202).

Same forward as ``OpenProvenceModel`` (one engine, packed sequences); what differs from ``process()`` is the
post-processing: pruning is decided per TOKEN (``keep_prob > threshold``) and the pruned document is
rebuilt from the tokenizer's character offsets, or evaluated against caller-supplied character chunks.
The keep probabilities are computed on the device (``opv_token_keep_probs``); thresholds, span
resolution and string work stay on the host as in the reference.  Training-time behaviour of the
reference class (it is an ``nn.Module`` that owns trainable weights) is out of scope.
"""

from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Sequence

import numpy as np
import torch

from .config import OpenProvenceConfig
from .modeling import OpenProvenceModel, _load_state_dict


@dataclass
class OpenProvenceOutput:
    """Chunk-level result of ``predict_context`` (reference ``data_structures.py:15-45``)."""

    ranking_scores: float | np.ndarray | None = None
    chunk_predictions: np.ndarray | None = None
    chunk_scores: np.ndarray | None = None
    token_scores: np.ndarray | None = None
    chunk_positions: list | None = None
    compression_ratio: float | None = None


@dataclass
class RerankingOpenProvenceOutput:
    """Token-level result of ``predict_with_pruning`` (reference ``data_structures.py:88-146``); as in the
    reference ``sentences`` holds the document TOKENS and ``num_pruned_sentences`` counts pruned tokens."""

    ranking_scores: np.ndarray | None = None
    ranking_logits: torch.Tensor | None = None
    pruning_masks: np.ndarray | None = None
    pruning_logits: torch.Tensor | None = None
    pruning_probs: np.ndarray | None = None
    sentences: list | None = None
    sentence_boundaries: list | None = None
    original_positions: list | None = None
    compression_ratio: float | None = None
    num_pruned_sentences: int | None = None
    pruned_documents: list | None = None


def _is_special(mask_value: int | None, offset: tuple[int, int]) -> bool:
    """encoder.py:559-569 -- flagged by the tokenizer, or an empty (0, 0) character span."""
    return (mask_value is not None and int(mask_value) == 1) or (offset[0] == 0 and offset[1] == 0)


def _trim(start: int, end: int, offsets: Sequence[tuple[int, int]], special: Sequence[int] | None):
    """Drop special tokens from both ends of [start, end) (encoder.py:571-598)."""
    n = len(offsets)
    start, end = max(0, min(start, n)), max(0, min(end, n))
    while start < end and _is_special(special[start] if special is not None else None, offsets[start]):
        start += 1
    while end > start and _is_special(special[end - 1] if special is not None else None, offsets[end - 1]):
        end -= 1
    return (start, end) if end > start else None


def resolve_document_span(
    token_ids: Sequence[int],
    offsets: Sequence[tuple[int, int]],
    token_type_ids: Sequence[int] | None,
    special_tokens_mask: Sequence[int] | None,
    separator_ids: Sequence[int],
) -> tuple[int, int] | None:
    """Token range [start, end) of the document half of a (query, document) pair, found the way the
    reference does (encoder.py:600-680): segment ids first, separator tokens second, non-special run last."""
    if token_type_ids is not None:
        doc = [i for i, t in enumerate(token_type_ids) if int(t) == 1]
        if doc:
            span = _trim(doc[0], doc[-1] + 1, offsets, special_tokens_mask)
            if span is not None:
                return span
    if separator_ids:
        seps = sorted({i for i, t in enumerate(token_ids) if int(t) in separator_ids})
        if len(seps) >= 2:
            span = _trim(seps[0] + 1, seps[-1], offsets, special_tokens_mask)
            if span is not None:
                return span
        elif seps:
            span = _trim(seps[0] + 1, len(offsets), offsets, special_tokens_mask)
            if span is not None:
                return span
    plain = [i for i, off in enumerate(offsets)
             if not _is_special(special_tokens_mask[i] if special_tokens_mask is not None else None, off)]
    if not plain or plain[-1] + 1 <= plain[0]:
        return None
    return plain[0], plain[-1] + 1


def evaluate_chunks(chunks: Sequence[tuple[int, int]], probs: np.ndarray, offsets: np.ndarray,
                    token_threshold: float, chunk_threshold: float) -> tuple[np.ndarray, np.ndarray]:
    """Chunk score = mean keep probability of the overlapping tokens; chunk kept when the share of tokens
    above ``token_threshold`` reaches ``chunk_threshold`` (encoder.py:841-899).  Tokens whose start OR end
    offset is 0 are skipped, exactly as the reference's ``token_start != 0 and token_end != 0`` test does
    (so the document's first token never votes)."""
    scores = np.zeros(len(chunks), dtype=np.float64)
    preds = np.zeros(len(chunks), dtype=np.int64)
    if len(offsets) == 0:
        return scores, preds
    starts, ends = offsets[:, 0], offsets[:, 1]
    usable = (starts != 0) & (ends != 0)
    for c, (c_start, c_end) in enumerate(chunks):
        hit = usable & (starts < c_end) & (ends > c_start)
        if not hit.any():
            continue
        # the reference averages a Python list of float32 ``.item()`` values with np.mean -> float64
        p = probs[hit].astype(np.float64)
        scores[c] = p.mean()
        preds[c] = 1 if (p > token_threshold).sum() / p.size >= chunk_threshold else 0
    return scores, preds


def rebuild_document(text: str, keep: np.ndarray, offsets: np.ndarray) -> str:
    """Kept character ranges, merged when they touch or overlap, joined with one space (encoder.py:482-512)."""
    ranges = sorted((int(s), int(e)) for k, (s, e) in zip(keep, offsets) if k and not (s == 0 and e == 0))
    if not ranges:
        return ""
    merged = [ranges[0]]
    for s, e in ranges[1:]:
        if s <= merged[-1][1]:
            merged[-1] = (merged[-1][0], max(merged[-1][1], e))
        else:
            merged.append((s, e))
    return " ".join(text[s:e] for s, e in merged)


class OpenProvenceEncoder:
    """Inference-side ``OpenProvenceEncoder`` on the sm_100a engine."""

    def __init__(self, model: OpenProvenceModel, state_dict: dict[str, torch.Tensor] | None = None, *,
                 packed_forward: Any = None) -> None:
        """``packed_forward(list of token-id lists) -> (rank logits [n, labels], [keep probs per pair])`` replaces
        the device stage; the host tests use it the way the reference's tests swap ``forward``."""
        if model.engine is None and packed_forward is None:
            raise RuntimeError("OpenProvenceEncoder needs an OpenProvenceModel with an engine")
        self._model = model
        self._packed_forward = packed_forward or self._engine_forward
        self._state_dict = state_dict  # kept (on the host) only so that save_pretrained can write it back
        self.tokenizer = model.tokenizer
        self.config = model.config
        self.mode = "reranking_pruning"
        self.num_labels = model.num_labels
        self.max_length = model.max_length
        self.device = str(model.device)
        self.model_name_or_path = model.config.base_model_name_or_path or ""

    # ------------------------------------------------------------------ loading / saving
    @classmethod
    def from_pretrained(cls, model_name_or_path: str | Path, device: str | None = None, **kwargs: Any) -> "OpenProvenceEncoder":
        """Reads the checkpoint layout ``save_pretrained`` writes (encoder.py:1040-1094): ``config.json`` with
        ``mode`` / ``max_length`` / ``pruning_config`` / ``base_model_config``, ``model.safetensors`` (or
        ``pytorch_model.bin``) with ``ranking_model.*`` and ``pruning_head.*`` keys, tokenizer files."""
        kwargs.pop("trust_remote_code", None)
        keep_weights = bool(kwargs.pop("keep_weights", True))
        path = Path(model_name_or_path)
        config = OpenProvenceConfig.from_pretrained(path)
        mode = getattr(config, "mode", "reranking_pruning")
        if mode != "reranking_pruning":
            raise ValueError(
                "Checkpoints saved in 'pruning_only' mode are no longer supported. "
                "Please export a reranking+pruning checkpoint."
            )
        state = dict(_load_state_dict(path))
        if not any(k.startswith("pruning_head.") for k in state):
            raise ValueError("No pruning head found in the model")
        model = OpenProvenceModel.from_pretrained(path, device=device or "cuda", **kwargs)
        return cls(model, state if keep_weights else None)

    def state_dict(self) -> dict[str, torch.Tensor]:
        if self._state_dict is None:
            raise RuntimeError("this encoder was loaded with keep_weights=False; there is nothing to save")
        return self._state_dict

    def save_pretrained(self, save_directory: str | Path) -> None:
        """``config.json`` + ``model.safetensors`` + tokenizer, loadable by this class, by
        ``OpenProvenceModel.from_pretrained`` and by the reference (same keys, encoder.py:1050-1088)."""
        from safetensors.torch import save_file

        out = Path(save_directory)
        out.mkdir(parents=True, exist_ok=True)
        save_file({k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()}, str(out / "model.safetensors"))
        cfg = self.config.to_dict()
        cfg["max_length"] = int(self.max_length)
        cfg["mode"] = "reranking_pruning"
        from .hf_auto import ARCHITECTURES, AUTO_MAP, with_auto_map

        # encoder.py:1078-1094: the reference writes ``auto_map`` AND copies modeling_open_provence_standalone.py next to
        # the weights, so that stock ``AutoModel.from_pretrained(dir, trust_remote_code=True)`` finds the class.  That
        # file belongs to the reference; it is carried over from the checkpoint this encoder was loaded from when it is
        # there, and ``auto_map`` is only written when the module it names is actually in the directory.
        shipped = False
        source = Path(getattr(self.config, "_name_or_path", "") or "")
        modules = {entry.split(".")[0] for entry in AUTO_MAP.values()}
        for module in modules:
            src = source / f"{module}.py"
            if source.is_dir() and src.is_file():
                if src.resolve() != (out / src.name).resolve():
                    (out / src.name).write_bytes(src.read_bytes())
                shipped = True
        shipped = shipped and all((out / f"{module}.py").is_file() for module in modules)
        cfg.pop("auto_map", None)
        cfg["architectures"] = list(ARCHITECTURES)
        if shipped:
            cfg = with_auto_map(cfg)
        (out / "config.json").write_text(json.dumps(cfg, indent=2, ensure_ascii=False, default=str))
        if self.tokenizer is not None:
            self.tokenizer.save_pretrained(str(out))

    def eval(self) -> "OpenProvenceEncoder":
        return self

    def to(self, *args: Any, **kwargs: Any) -> "OpenProvenceEncoder":
        self._model.to(*args, **kwargs)
        return self

    # ------------------------------------------------------------------ forward
    def forward(self, input_ids: torch.Tensor | None = None, attention_mask: torch.Tensor | None = None,
                sentence_boundaries: torch.Tensor | None = None, return_dict: bool = True, **kwargs: Any):
        """encoder.py:174-245: ``{"ranking_logits", "pruning_logits", "hidden_states"}`` (hidden states are not
        materialised by the engine and come back as None)."""
        del sentence_boundaries
        if input_ids is None and "sentence_features" in kwargs:
            features = kwargs.pop("sentence_features")
            if features:
                input_ids = features[0].get("input_ids")
                attention_mask = features[0].get("attention_mask")
        if input_ids is None:
            raise ValueError("input_ids must be provided")
        if attention_mask is None:
            raise ValueError("attention_mask must be provided")
        out = self._model.forward(input_ids=input_ids, attention_mask=attention_mask, return_dict=True)
        if return_dict:
            return {"ranking_logits": out["ranking_logits"], "pruning_logits": out["pruning_logits"], "hidden_states": None}
        return out["ranking_logits"], out["pruning_logits"]

    __call__ = forward

    # ------------------------------------------------------------------ scoring core
    def _engine_forward(self, id_lists: list[list[int]]) -> tuple[np.ndarray, list[np.ndarray]]:
        """One packed launch: forward + token keep probabilities on the device, results to the host."""
        engine = self._model.engine
        lengths = [len(ids) for ids in id_lists]
        flat = np.fromiter((t for ids in id_lists for t in ids), dtype=np.int32, count=sum(lengths))
        cu = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32)
        ids_dev = torch.from_numpy(flat).to(engine.device, non_blocking=True)
        cu_dev = torch.from_numpy(cu).to(engine.device, non_blocking=True)
        prune, rank = engine.forward_packed(ids_dev, cu_dev, max(lengths))
        prob = engine.token_keep_probs(prune).cpu().numpy()
        return rank.cpu().numpy(), [prob[cu[j] : cu[j + 1]] for j in range(len(lengths))]

    def _score_pairs(self, pairs: list[tuple[str, str]], batch_size: int, need_tokens: bool):
        """Tokenise (truncating to max_length), run the packed forward, and return per pair:
        rank logits [num_labels], keep probabilities [n_i] and the tokenizer's side tables."""
        max_tokens = max(131072, int(batch_size) * max(int(self.max_length), 1))
        results: list[dict[str, Any]] = []
        enc_kwargs: dict[str, Any] = dict(padding=False, truncation=True, max_length=self.max_length)
        if need_tokens:
            enc_kwargs.update(return_offsets_mapping=True, return_token_type_ids=True, return_special_tokens_mask=True)
        encoded = self.tokenizer([(str(q), str(d)) for q, d in pairs], **enc_kwargs) if pairs else {"input_ids": []}
        ids_all = encoded["input_ids"]
        at = 0
        while at < len(ids_all):
            end, tokens = at, 0
            while end < len(ids_all) and (end == at or tokens + len(ids_all[end]) <= max_tokens):
                tokens += len(ids_all[end])
                end += 1
            rank_host, probs = self._packed_forward([list(ids_all[i]) for i in range(at, end)])
            for j, i in enumerate(range(at, end)):
                rec: dict[str, Any] = {"ids": ids_all[i], "rank": rank_host[j], "prob": np.asarray(probs[j], dtype=np.float32)}
                if need_tokens:
                    rec["offsets"] = [(int(s), int(e)) for s, e in encoded["offset_mapping"][i]]
                    tt = encoded.get("token_type_ids")
                    sm = encoded.get("special_tokens_mask")
                    rec["token_type_ids"] = tt[i] if tt is not None else None
                    rec["special_tokens_mask"] = sm[i] if sm is not None else None
                results.append(rec)
            at = end
        return results

    @staticmethod
    def _rank_score(rank_logits: np.ndarray) -> float:
        """First (or only) ranking logit, raw (encoder.py:316-325)."""
        return float(np.asarray(rank_logits).reshape(-1)[0])

    def _separator_ids(self) -> list[int]:
        out = []
        for name in ("eos_token_id", "sep_token_id"):
            value = getattr(self.tokenizer, name, None)
            if value is not None:
                out.append(int(value))
        return out

    @staticmethod
    def _as_list(sentences):
        single = isinstance(sentences[0], str)
        return ([tuple(sentences)] if single else [tuple(p) for p in sentences]), single

    # ------------------------------------------------------------------ public APIs
    def predict(self, sentences, batch_size: int = 32, show_progress_bar: bool = False, convert_to_numpy: bool = True,
                convert_to_tensor: bool = False, apply_pruning: bool = False, pruning_threshold: float = 0.5,
                return_documents: bool = False):
        """Raw ranking scores of (query, document) pairs (encoder.py:247-343)."""
        if apply_pruning:
            return self.predict_with_pruning(sentences, batch_size=batch_size, pruning_threshold=pruning_threshold,
                                             return_documents=return_documents, show_progress_bar=show_progress_bar)
        pairs, _ = self._as_list(sentences)
        scores = [self._rank_score(r["rank"]) for r in self._score_pairs(pairs, batch_size, need_tokens=False)]
        if convert_to_tensor:
            return torch.tensor(scores)
        if convert_to_numpy:
            return np.array(scores)
        return scores

    def predict_with_pruning(self, sentences, batch_size: int = 32, pruning_threshold: float = 0.5,
                             return_documents: bool = False, show_progress_bar: bool = False):
        """Token-level pruning of each document (encoder.py:345-529)."""
        del show_progress_bar
        pairs, single = self._as_list(sentences)
        outputs = []
        for (_, document), rec in zip(pairs, self._score_pairs(pairs, batch_size, need_tokens=True)):
            score = np.array([self._rank_score(rec["rank"])])
            span = resolve_document_span(rec["ids"], rec["offsets"], rec["token_type_ids"], rec["special_tokens_mask"],
                                         self._separator_ids())
            if span is None:
                out = RerankingOpenProvenceOutput(ranking_scores=score, pruning_masks=np.array([[]]), sentences=[[]],
                                                  compression_ratio=0.0, num_pruned_sentences=0)
                if return_documents:
                    out.pruned_documents = [""]
                outputs.append(out)
                continue
            start, end = span
            keep = rec["prob"][start:end] > pruning_threshold
            total = end - start
            kept = int(keep.sum())
            out = RerankingOpenProvenceOutput(
                ranking_scores=score,
                pruning_masks=np.array([keep]),
                sentences=[self.tokenizer.convert_ids_to_tokens(rec["ids"][start:end])],
                compression_ratio=1.0 - (kept / total) if total > 0 else 0.0,
                num_pruned_sentences=int(total - kept),
            )
            if return_documents:
                offsets = np.asarray(rec["offsets"][start:end], dtype=np.int64).reshape(-1, 2)
                out.pruned_documents = [rebuild_document(str(document), keep, offsets)]
            outputs.append(out)
        return outputs[0] if single else outputs

    def predict_context(self, sentences, chunk_positions, batch_size: int = 32, token_threshold: float = 0.5,
                        chunk_threshold: float = 0.5, show_progress_bar: bool = False):
        """Chunk-level relevance from token keep probabilities (encoder.py:682-839)."""
        del show_progress_bar
        pairs, single = self._as_list(sentences)
        chunk_lists = [chunk_positions] if single else list(chunk_positions)
        outputs = []
        for chunks, rec in zip(chunk_lists, self._score_pairs(pairs, batch_size, need_tokens=True)):
            score = self._rank_score(rec["rank"])
            span = resolve_document_span(rec["ids"], rec["offsets"], rec["token_type_ids"], rec["special_tokens_mask"],
                                         self._separator_ids())
            if span is None:
                outputs.append(OpenProvenceOutput(ranking_scores=score, chunk_predictions=np.array([]), chunk_scores=np.array([]),
                                                  token_scores=np.array([]), chunk_positions=chunks, compression_ratio=0.0))
                continue
            start, end = span
            probs = rec["prob"][start:end]
            offsets = np.asarray(rec["offsets"][start:end], dtype=np.int64).reshape(-1, 2)
            flat = chunks if (isinstance(chunks, list) and len(chunks) > 0 and isinstance(chunks[0], tuple)) else chunks[0]
            chunk_scores, chunk_preds = evaluate_chunks([tuple(c) for c in flat], probs, offsets, token_threshold, chunk_threshold)
            n_chunks = len(chunks)
            outputs.append(OpenProvenceOutput(
                ranking_scores=score, chunk_predictions=chunk_preds, chunk_scores=chunk_scores, token_scores=probs.copy(),
                chunk_positions=chunks, compression_ratio=1.0 - (chunk_preds.sum() / n_chunks) if n_chunks > 0 else 0.0))
        return outputs[0] if single else outputs

    def prune(self, query: str, document: str, threshold: float = 0.5, min_sentences: int = 1, return_sentences: bool = False):
        """One document, token-level (encoder.py:901-938)."""
        del min_sentences
        out = self.predict_with_pruning((query, document), pruning_threshold=threshold, return_documents=True)
        if not return_sentences:
            return out.pruned_documents[0]
        return {
            "pruned_document": out.pruned_documents[0],
            "sentences": [],
            "pruning_masks": [],
            "ranking_score": float(np.asarray(out.ranking_scores).reshape(-1)[0]) if out.ranking_scores is not None else None,
            "compression_ratio": out.compression_ratio,
            "num_pruned_sentences": 0,
        }

    def prune_texts(self, queries: list[str], texts: list[str], threshold: float = 0.5, batch_size: int = 32,
                    return_tokens: bool = False, show_progress_bar: bool = False) -> list[dict[str, Any]]:
        """encoder.py:940-991.  ``return_tokens=True`` returns the token mask under ``pruning_mask`` (the reference
        reads a non-existent ``output.pruning_mask`` attribute there and raises; the mask is what it documents)."""
        outputs = self.predict_with_pruning(list(zip(queries, texts)), batch_size=batch_size, pruning_threshold=threshold,
                                            return_documents=True, show_progress_bar=show_progress_bar)
        results = []
        for text, out in zip(texts, outputs):
            item = {"pruned_text": out.pruned_documents[0] if out.pruned_documents else text,
                    "kept_ratio": 1.0 - out.compression_ratio}
            if return_tokens:
                item["pruning_mask"] = out.pruning_masks[0] if out.pruning_masks is not None else None
            results.append(item)
        return results
