"""Drop-in ``OpenProvenceModel`` for the scoring-and-pruning hot path on B200.

Mirrors the reference interface for this path only
(``/root/reference/open_provence/modeling_open_provence_standalone.py``, cited as ``standalone:N``):

* ``OpenProvenceModel.from_pretrained``  standalone:1557-1664  (HF ``config.json`` + safetensors)
* ``OpenProvenceModel.forward``          standalone:1666-1739  (same signature / outputs)
* ``get_raw_predictions[_batch]``        standalone:1741-1841
* ``OpenProvenceModel.process``          standalone:3314-3808  (same kwargs, same result dict)

The arithmetic (ModernBERT forward, score conversion, per-sentence prune) runs in ``libopv_sm100.so``;
this file is host-side planning only: splitting, tokenising, fragmentising, packing blocks, and turning
keep flags back into strings.  It is organised around flat tables (:class:`scoring.BlockTable`) instead
of the reference's per-block dictionaries so one ``process()`` call costs a handful of launches.
"""

from __future__ import annotations

import contextlib
import json
from collections.abc import Callable, Mapping, Sequence
from dataclasses import dataclass, field
from pathlib import Path
from time import perf_counter
from typing import Any

import numpy as np
import torch

from . import host_pack
from .config import DEFAULT_PROCESS_THRESHOLD, OpenProvenceConfig
from .host_text import (
    Fragment,
    SentenceSplitter,
    fallback_sentence,
    filter_decodable,
    normalize_sentences,
    resolve_sentence_splitter,
    split_token_lists,
    tokenize_batch,
    decode_batch,
)
from .scoring import BlockTable, DeviceScorer

DEFAULT_SPLITTER_LANGUAGE = "auto"


@dataclass(frozen=True)
class ProcessPerformanceTrace:
    """Same fields as the reference's trace (standalone:377-404)."""

    preprocess_seconds: float = 0.0
    assembly_seconds: float = 0.0
    inference_seconds: float = 0.0
    postprocess_seconds: float = 0.0
    total_seconds: float = 0.0
    sentence_collect_seconds: float = 0.0
    sentence_normalize_seconds: float = 0.0
    tokenize_seconds: float = 0.0
    fragment_split_seconds: float = 0.0
    fragment_decode_seconds: float = 0.0

    def as_dict(self) -> dict[str, float]:
        return {k: float(getattr(self, k)) for k in self.__dataclass_fields__}


@dataclass
class OpenProvenceRawPrediction:
    """standalone:451-459."""

    query: str
    contexts: list[str]
    ranking_score: float | None
    pruning_probs: np.ndarray
    context_ranges: list[tuple[int, int]]


class OpenProvenceOutput(dict):
    """``forward`` result: mapping *and* attribute access (the reference's callers use both,
    standalone:1540-1555)."""

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError as exc:
            raise AttributeError(name) from exc


@dataclass
class _ContextPlan:
    """One (query, context) pair on its way through ``process()`` (reference: ``contexts_info`` entry,
    standalone:2736-2746)."""

    query_idx: int
    context_idx: int
    entry: Any
    context_text: str
    sentences: list[str]
    prefix_sentences: list[str]
    prefix_token_counts: list[int]
    title_is_first_sentence: bool
    token_lists: list[list[int]] = field(default_factory=list)
    fragments: list[Fragment] = field(default_factory=list)
    blocks: list[list[Fragment]] = field(default_factory=list)
    block_slots: list[int] = field(default_factory=list)  # indices into BlockTable.block_ids
    sentence_base: int = 0  # first sentence slot in the BlockTable CSR
    n_fragments: int = 0  # after the empty-fragment filter
    n_blocks: int = 0


def _is_seq(value: Any) -> bool:
    return isinstance(value, Sequence) and not isinstance(value, (str, bytes, bytearray))


def resolve_inference_device(device: str | torch.device | None) -> torch.device:
    """Device resolution with the reference's error messages (standalone:305-339); CUDA only."""
    if isinstance(device, torch.device):
        cand = device
    elif device is None or str(device).strip().lower() in ("", "auto"):
        cand = torch.device("cuda")
    else:
        text = str(device).strip().lower()
        if text == "cpu" or text.startswith("mps"):
            raise ValueError(
                f"Unsupported device specification: {device!r} (the sm_100a engine runs on CUDA only; "
                "there is no CPU fallback)"
            )
        if not text.startswith("cuda"):
            raise ValueError(f"Unsupported device specification: {device!r}")
        cand = torch.device(text)
    if cand.type != "cuda":
        raise ValueError(f"Unsupported device specification: {device!r} (CUDA only)")
    if not torch.cuda.is_available():
        raise ValueError("CUDA device requested but CUDA is not available.")
    if cand.index is not None and not (0 <= cand.index < torch.cuda.device_count()):
        raise ValueError(f"CUDA device index {cand.index} out of range (count={torch.cuda.device_count()}).")
    return cand


def convert_legacy_state_dict(state: Mapping[str, torch.Tensor]) -> Mapping[str, torch.Tensor]:
    """Un-prefixed legacy checkpoints get ``ranking_model.`` (standalone:1452-1464)."""
    if any(k.startswith("ranking_model.") for k in state):
        return state
    return {(k if k.startswith("pruning_head.") else f"ranking_model.{k}"): v for k, v in state.items()}


def _load_state_dict(path: Path) -> Mapping[str, torch.Tensor]:
    st = path / "model.safetensors"
    if st.exists():
        from safetensors.torch import load_file

        return load_file(str(st))
    index = path / "model.safetensors.index.json"
    if index.exists():
        from safetensors.torch import load_file

        shards = sorted(set(json.loads(index.read_text())["weight_map"].values()))
        merged: dict[str, torch.Tensor] = {}
        for shard in shards:
            merged.update(load_file(str(path / shard)))
        return merged
    legacy = path / "pytorch_model.bin"
    if legacy.exists():
        return torch.load(str(legacy), map_location="cpu", weights_only=True)
    raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin under {path}")


class OpenProvenceModel:
    """B200-native stand-in for the reference's ``OpenProvenceModel`` (hot path only)."""

    def __init__(
        self,
        config: OpenProvenceConfig,
        state_dict: Mapping[str, torch.Tensor] | None = None,
        tokenizer: Any = None,
        *,
        device: str | torch.device | None = None,
        dtype: Any = None,
        fuse_epilogues: bool = True,
        scorer: Any = None,
        process_group: Any = None,
        data_parallel: bool = False,
    ) -> None:
        self.config = config
        self.max_length = int(config.max_length)
        self.num_labels = int(config.num_labels)
        self.num_pruning_labels = int(config.num_pruning_labels)
        self.default_splitter_language = DEFAULT_SPLITTER_LANGUAGE
        self.default_threshold = config.resolve_default_threshold()
        self.tokenizer = tokenizer
        self.engine = None
        self._scorer = scorer
        self._manual_special_tokens_required = False
        self._manual_cls_token_id: int | None = None
        self._manual_sep_token_id: int | None = None
        if state_dict is not None:
            from .engine import Engine

            resolved = resolve_inference_device(device)
            if not config.base_model_config:
                raise ValueError("OpenProvenceConfig must define base_model_config or base_model_name_or_path.")
            if "model_type" not in config.base_model_config:
                raise ValueError("base_model_config must include 'model_type' to rebuild the backbone.")
            self.engine = Engine(
                config.base_model_config,
                convert_legacy_state_dict(state_dict),
                device=resolved,
                dtype=dtype,
                num_labels=self.num_labels,
                fuse_epilogues=fuse_epilogues,
            )
            self._runtime_device = self.engine.device
            if scorer is None:  # an injected scorer (tests, custom sharding) is kept
                self._scorer = DeviceScorer(self.engine)
            if data_parallel or process_group is not None:
                self.enable_data_parallel(process_group)
        else:
            self._runtime_device = torch.device("cpu")
        if tokenizer is not None:
            self._update_tokenizer_runtime()
            self._update_runtime_defaults()

    # ------------------------------------------------------------------ multi-GPU
    def enable_data_parallel(self, process_group: Any = None) -> "OpenProvenceModel":
        """Shard the blocks of every ``process()`` call over the ranks of ``process_group`` (default: the world
        group of an initialised ``torch.distributed``; one process per GPU, NCCL).  Weights stay replicated; every
        rank must call ``process()`` with the same arguments and receives the same, complete result.  One
        all-gather of ``[rank scores | fragment means]`` per call (``sharding.ShardedScorer``)."""
        import torch.distributed as dist

        from .sharding import ShardedScorer

        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("enable_data_parallel() needs an initialised torch.distributed process group")
        if isinstance(self._scorer, ShardedScorer):
            self._scorer.group = process_group
            return self
        base = self.config.base_model_config or {}
        self._scorer = ShardedScorer(self._scorer, process_group, hidden=int(base.get("hidden_size", 512)),
                                     inter=int(base.get("intermediate_size", 2048)))
        return self

    # ------------------------------------------------------------------ loading
    @classmethod
    def from_pretrained(
        cls,
        pretrained_model_name_or_path: str | Path,
        *,
        device: str | torch.device | None = None,
        trust_remote_code: bool = True,
        max_length: int | None = None,
        torch_dtype: torch.dtype | str | None = None,
        **kwargs: Any,
    ) -> "OpenProvenceModel":
        """Load ``config.json`` + safetensors + tokenizer of a reference checkpoint (standalone:1557-1664).

        ``dtype`` / ``torch_dtype``: bf16 (default on CUDA, as in the reference, standalone:219-233) or
        fp32 (parity mode).  ``attn_implementation`` is accepted and ignored: attention is the engine's."""
        del trust_remote_code
        try:
            resolved_device = resolve_inference_device(device)
        except ValueError as exc:
            raise ValueError(f"Invalid device specification for OpenProvenceModel: {device!r}") from exc
        if "torch_dtype" in kwargs and "dtype" not in kwargs:
            kwargs["dtype"] = kwargs.pop("torch_dtype")
        dtype = kwargs.pop("dtype", None)
        if dtype is None:
            dtype = torch_dtype
        if dtype is None or (isinstance(dtype, str) and dtype.lower() == "auto"):
            dtype = torch.bfloat16
        kwargs.pop("attn_implementation", None)
        fuse = bool(kwargs.pop("fuse_epilogues", True))
        process_group = kwargs.pop("process_group", None)
        data_parallel = bool(kwargs.pop("data_parallel", False))

        path = Path(pretrained_model_name_or_path)
        if not path.is_dir():
            try:
                from huggingface_hub import snapshot_download

                path = Path(snapshot_download(str(pretrained_model_name_or_path)))
            except Exception as exc:
                raise OSError(
                    f"{pretrained_model_name_or_path!r} is not a local checkpoint directory and could not be "
                    "downloaded from the Hugging Face hub."
                ) from exc
        config = OpenProvenceConfig.from_pretrained(path)
        state = _load_state_dict(path)
        tokenizer = cls._init_tokenizer(config, path)
        model = cls(config, state, tokenizer, device=resolved_device, dtype=dtype, fuse_epilogues=fuse,
                    process_group=process_group, data_parallel=data_parallel)
        if max_length is not None:
            model.max_length = int(max_length)
            model.config.max_length = int(max_length)
        model._update_tokenizer_runtime(max_length_override=max_length)
        model._update_runtime_defaults()
        return model

    @staticmethod
    def _init_tokenizer(config: OpenProvenceConfig, ckpt_dir: Path) -> Any:
        from transformers import AutoTokenizer

        candidates = []
        ref = config.tokenizer_name_or_path
        if ref:
            for cand in (ckpt_dir / ref, ckpt_dir.parent / ref, Path(ref)):
                if cand.is_dir():
                    candidates.append(str(cand))
        if any((ckpt_dir / name).exists() for name in ("tokenizer.json", "tokenizer_config.json", "spiece.model")):
            candidates.insert(0, str(ckpt_dir))
        candidates.extend(c for c in (ref, config.base_model_name_or_path) if c)
        last: Exception | None = None
        for cand in candidates:
            try:
                return AutoTokenizer.from_pretrained(cand)
            except Exception as exc:  # try the next location
                last = exc
        raise RuntimeError(f"Failed to initialize tokenizer from '{ref or ckpt_dir}'.") from last

    # nn.Module-flavoured no-ops so callers written for the reference keep working
    def eval(self) -> "OpenProvenceModel":
        return self

    def to(self, *args: Any, **kwargs: Any) -> "OpenProvenceModel":
        target = kwargs.get("device", args[0] if args else None)
        if target is not None and self.engine is not None and torch.device(target).type != "cuda":
            raise ValueError("the sm_100a engine cannot be moved off CUDA (no CPU fallback)")
        return self

    @property
    def device(self) -> torch.device:
        return self._runtime_device

    # ------------------------------------------------------------------ tokenizer runtime
    def _update_tokenizer_runtime(self, max_length_override: int | None = None) -> None:
        """standalone:1391-1399: make sure the tokenizer never truncates/warns below max_length."""
        if self.tokenizer is None:
            return
        upper = max(getattr(self.tokenizer, "model_max_length", 0) or 0, 1_000_000)
        if max_length_override is not None and max_length_override > 0:
            upper = max(upper, int(max_length_override))
        elif self.max_length and self.max_length > 0:
            upper = max(upper, int(self.max_length))
        try:
            self.tokenizer.model_max_length = upper
        except Exception:
            pass

    def _special_id(self, *names: str) -> int | None:
        tok = self.tokenizer
        special_map = getattr(tok, "special_tokens_map", {}) or {}
        for name in names:
            for value in (getattr(tok, name, None), special_map.get(name) if isinstance(special_map, Mapping) else None):
                if isinstance(value, int):
                    return value
        return None

    def _requires_manual_special_tokens(self) -> bool:
        """standalone:1501-1538, plus: a tokenizer without ``build_inputs_with_special_tokens``
        (transformers >= 5 fast tokenizers) always takes the manual ``[CLS] q [SEP] ctx [SEP]`` path."""
        tok = self.tokenizer
        build = getattr(tok, "build_inputs_with_special_tokens", None)
        if not callable(build):
            return True
        try:
            q = tok.encode("open provence query", add_special_tokens=False)
            c = tok.encode("open provence document", add_special_tokens=False)
        except Exception:
            return False
        if not q or not c:
            return False
        built = [int(t) for t in build(q, c)]
        cls_ids = [v for v in (self._special_id("cls_token_id"), self._special_id("bos_token_id")) if v is not None]
        sep_ids = [v for v in (self._special_id("sep_token_id"), self._special_id("eos_token_id")) if v is not None]
        missing_cls = bool(cls_ids) and not any(t in cls_ids for t in built)
        missing_sep = bool(sep_ids) and not any(t in sep_ids for t in built)
        return missing_cls or missing_sep

    def _update_runtime_defaults(self) -> None:
        self._manual_special_tokens_required = self._requires_manual_special_tokens()
        if self._manual_special_tokens_required:
            self._manual_cls_token_id = self._special_id("cls_token_id", "bos_token_id")
            self._manual_sep_token_id = self._special_id("sep_token_id", "eos_token_id")
        else:
            self._manual_cls_token_id = None
            self._manual_sep_token_id = None

    def _resolve_process_threshold(self, threshold: float | None) -> float:
        """standalone:1482-1493."""
        if threshold is None:
            resolved = getattr(self, "default_threshold", DEFAULT_PROCESS_THRESHOLD)
            if resolved is None:
                resolved = DEFAULT_PROCESS_THRESHOLD
        else:
            resolved = threshold
        try:
            return float(resolved)
        except (TypeError, ValueError) as exc:
            raise TypeError("Resolved threshold must be numeric.") from exc

    # ------------------------------------------------------------------ forward
    def forward(
        self,
        input_ids: torch.Tensor | None = None,
        attention_mask: torch.Tensor | None = None,
        labels: torch.Tensor | None = None,
        return_dict: bool | None = None,
        **kwargs: Any,
    ) -> OpenProvenceOutput | tuple[torch.Tensor, ...]:
        """Same contract as standalone:1666-1739: ``[B, S]`` ids (+ mask) -> ``ranking_logits [B, num_labels]`` and
        ``pruning_logits [B, S, 2]`` (fp32; zeros at masked columns, where the reference returns the backbone's output
        for the [PAD] token -- no caller reads those: ``context_ranges`` never cover padding)."""
        if input_ids is None:
            raise ValueError("input_ids must be provided")
        if self.engine is None:
            raise RuntimeError("this OpenProvenceModel has no engine (constructed without weights)")
        del kwargs  # token_type_ids etc.: ModernBERT ignores them
        dev = self.engine.device
        ids = input_ids.to(dev)
        if ids.dim() != 2:
            raise ValueError("input_ids must have shape [batch, seq_len]")
        B, S = ids.shape
        if attention_mask is None:
            mask = torch.ones((B, S), dtype=torch.bool, device=dev)
        else:
            mask = attention_mask.to(dev).bool()
        # Any mask is accepted: the kept tokens of a row are compacted into one packed sequence and the logits are
        # scattered back to their columns (masked columns: logit 0).  Positions (RoPE, sliding window, CLS = first kept
        # token) count KEPT tokens, exactly as in the reference's unpadded flash-attention path; for right-padded
        # batches -- what every caller in the reference produces -- that is also what its eager / sdpa path computes.
        lengths = mask.sum(dim=1)
        if bool((lengths == 0).any()):
            raise ValueError("every row of attention_mask must keep at least one token")
        cu = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        cu[1:] = torch.cumsum(lengths, dim=0).to(torch.int32)
        packed = ids[mask].to(torch.int32).contiguous()
        prune_packed, rank = self.engine.forward_packed(packed, cu, int(lengths.max().item()))
        prune = torch.zeros((B, S, 2), dtype=torch.float32, device=dev)
        prune[mask] = prune_packed
        loss = None
        if labels is not None:
            labels = labels.to(dev)
            if self.num_labels == 1:
                loss = torch.nn.functional.binary_cross_entropy_with_logits(rank.view(-1), labels.float())
            else:
                loss = torch.nn.functional.cross_entropy(rank.view(-1, self.num_labels), labels.view(-1))
        if return_dict is not None and not return_dict:
            out: tuple[torch.Tensor, ...] = (rank, prune)
            return ((loss,) + out) if loss is not None else out
        return OpenProvenceOutput(loss=loss, logits=rank, ranking_logits=rank, pruning_logits=prune,
                                  hidden_states=None, attentions=None)

    __call__ = forward

    # ------------------------------------------------------------------ raw predictions
    @torch.no_grad()
    def get_raw_predictions(self, query: str, contexts: Sequence[str]) -> OpenProvenceRawPrediction:
        return self.get_raw_predictions_batch(query, [list(contexts)])[0]

    @torch.no_grad()
    def get_raw_predictions_batch(
        self, query: str | Sequence[str], contexts_batch: Sequence[Sequence[str]], batch_size: int | None = None
    ) -> list[OpenProvenceRawPrediction]:
        """standalone:1752-1841: ``query + sep + "".join(contexts)`` tokenised with specials, truncated."""
        if not contexts_batch:
            return []
        sep = getattr(self.tokenizer, "sep_token", None) or ""
        if batch_size is None or batch_size <= 0:
            batch_size = len(contexts_batch)
        if _is_seq(query):
            queries = [str(q) for q in query]
            if len(queries) != len(contexts_batch):
                raise ValueError("When providing multiple queries, their count must match contexts_batch.")
        else:
            queries = [str(query)] * len(contexts_batch)
        results: list[OpenProvenceRawPrediction] = []
        for at in range(0, len(contexts_batch), batch_size):
            chunk = contexts_batch[at : at + batch_size]
            chunk_q = queries[at : at + batch_size]
            texts = [q + sep + "".join(c) for q, c in zip(chunk_q, chunk)]
            enc = self.tokenizer(texts, padding=True, truncation=True, max_length=self.max_length, return_tensors="pt")
            out = self.forward(input_ids=enc["input_ids"], attention_mask=enc.get("attention_mask"), return_dict=True)
            rank = out.ranking_logits.detach().cpu().float()
            prune = out.pruning_logits.detach().cpu().float()
            for i, contexts in enumerate(chunk):
                if len(contexts) == 0:
                    continue
                score = float(torch.sigmoid(rank[i].flatten())[0])
                probs = torch.softmax(prune[i], dim=-1).numpy()[:, 1]
                results.append(OpenProvenceRawPrediction(
                    query=chunk_q[i], contexts=list(contexts), ranking_score=score, pruning_probs=probs,
                    context_ranges=self._context_ranges_from_contexts(chunk_q[i], contexts)))
        return results

    def _context_ranges_from_contexts(self, query: str, contexts: Sequence[str]) -> list[tuple[int, int]]:
        """Token span of each context via cumulative tokenisation (standalone:1926-1969)."""
        if not contexts:
            return []
        sep = getattr(self.tokenizer, "sep_token", None) or ""
        prefix = query + sep

        def n_tokens(text: str, truncate: bool) -> int:
            enc = self.tokenizer(text, padding=False, truncation=truncate,
                                 **({"max_length": self.max_length} if truncate else {}))
            return len(enc["input_ids"])

        # call order as in the reference: the truncating calls first, the untruncated prefix call LAST, so the
        # tokenizer's Rust backend is left without truncation state (fast tokenizers keep the last call's)
        ends, acc = [], prefix
        for ctx in contexts:
            acc += ctx
            ends.append(n_tokens(acc, True))
        ranges, prev = [], n_tokens(prefix, False)
        for end in ends:
            ranges.append((prev, end))
            prev = end
        return ranges

    def predict_with_thresholds(self, query: str, contexts: Sequence[str], thresholds: Sequence[float], *,
                                use_majority: bool = False) -> dict[str, Any]:
        """standalone:1843-1881."""
        raw = self.get_raw_predictions(query, contexts)
        predictions: dict[float, list[int]] = {}
        for thr in thresholds:
            flags = []
            for start, end in raw.context_ranges:
                seg = raw.pruning_probs[start:end]
                if seg.size == 0:
                    flags.append(1)
                elif use_majority:
                    flags.append(1 if np.count_nonzero(seg > thr) >= seg.size / 2 else 0)
                else:
                    flags.append(1 if float(seg.mean()) > thr else 0)
            predictions[thr] = flags
        return {"query": raw.query, "contexts": raw.contexts, "ranking_score": raw.ranking_score,
                "predictions": predictions, "context_ranges": raw.context_ranges, "pruning_probs": raw.pruning_probs}

    # ------------------------------------------------------------------ process(): input shapes
    @staticmethod
    def _normalize_inputs(question: str | Sequence[str], context: Any) -> tuple[list[str], list[list[Any]], str]:
        """str / list / aligned / nested (standalone:2261-2323, same error messages)."""
        queries = [question] if isinstance(question, str) else [str(q) for q in question]

        def norm(values: Sequence[Any]) -> list[Any]:
            return [[str(e) for e in item] if _is_seq(item) else str(item) for item in values]

        if isinstance(context, str):
            structure, contexts = "str", [[context]]
        elif not _is_seq(context):
            raise ValueError("Unsupported context format")
        elif len(queries) == 1:
            structure, contexts = "list", [norm(context)]
        else:
            items = list(context)
            if all(not _is_seq(e) for e in items):
                if len(items) != len(queries):
                    raise ValueError("Number of contexts must match number of queries")
                structure, contexts = "aligned", [[str(e)] for e in items]
            else:
                structure, contexts = "nested", []
                for e in items:
                    if not _is_seq(e):
                        raise ValueError("Number of context lists must match number of queries")
                    contexts.append(norm(e))
        if structure == "list" and len(queries) != 1:
            raise ValueError("Single list of contexts requires a single query")
        if structure == "nested" and len(contexts) != len(queries):
            raise ValueError("Number of context lists must match number of queries")
        if structure == "str" and len(queries) != 1:
            raise ValueError("Single context string requires a single query")
        return queries, contexts, structure

    @staticmethod
    def _prepare_titles(title: Any, queries: list[str], contexts: list[list[Any]]) -> list[Any]:
        """standalone:2325-2360."""
        n = len(queries)
        if title is None:
            return [None] * n
        if isinstance(title, str):
            if title == "first_sentence":
                return ["first_sentence"] * n
            return [[title for _ in ctxs] for ctxs in contexts]
        if _is_seq(title):
            items = [[str(v) for v in e] if _is_seq(e) else str(e) for e in title]
            if n == 1 and all(isinstance(i, str) for i in items):
                return [[str(i) for i in items]]
            if len(items) == n and all(isinstance(i, list) for i in items):
                return [list(map(str, i)) for i in items]
            if len(items) == n and all(isinstance(i, str) for i in items):
                return [[v for _ in contexts[k]] for k, v in enumerate(items)]
        raise ValueError("Unsupported title format")

    @staticmethod
    def _extract_first_line_titles(contexts: list[list[Any]]) -> tuple[list[list[Any]], list[list[str]]]:
        """First non-empty line (or pre-split segment) becomes the title (standalone:2362-2410)."""
        new_contexts, titles = [], []
        for group in contexts:
            g_ctx, g_titles = [], []
            for entry in group:
                if isinstance(entry, list):
                    segs = [str(v) for v in entry]
                    title, rest = "", segs
                    for i, seg in enumerate(segs):
                        if seg.strip():
                            title, rest = seg.rstrip("\r\n"), segs[i + 1 :]
                            break
                    g_titles.append(title)
                    g_ctx.append(rest)
                else:
                    text = str(entry)
                    title, rest_text = "", ""
                    if text:
                        lines = text.splitlines(keepends=True)
                        rest_lines = lines
                        for i, line in enumerate(lines):
                            if line.strip():
                                title, rest_lines = line.rstrip("\r\n"), lines[i + 1 :]
                                break
                        rest_text = "".join(rest_lines)
                    g_titles.append(title)
                    g_ctx.append(rest_text)
            new_contexts.append(g_ctx)
            titles.append(g_titles)
        return new_contexts, titles

    def _resolve_titles(self, queries, contexts, title, *, first_line_as_title: bool):
        if first_line_as_title:
            if title not in (None, "first_sentence"):
                raise ValueError("first_line_as_title=True cannot be combined with an explicit title override.")
            contexts, title = self._extract_first_line_titles(contexts)
        return contexts, self._prepare_titles(title, queries, contexts)

    @staticmethod
    def _resolve_prefix_sentences(title_spec: Any, context_idx: int) -> tuple[list[str], bool]:
        """standalone:1971-2005: explicit titles become prefix sentences; the last one ends with a newline."""
        prefix: list[str] = []
        first = False
        if title_spec == "first_sentence":
            first = True
        elif isinstance(title_spec, list):
            raw = title_spec[context_idx] if context_idx < len(title_spec) else None
            if title_spec and isinstance(title_spec[0], list):
                if raw:
                    prefix.extend(t.strip() for t in raw if isinstance(t, str) and t.strip())
            elif isinstance(raw, str) and raw.strip():
                prefix.append(raw.strip())
        elif isinstance(title_spec, str) and title_spec.strip():
            prefix.append(title_spec.strip())
        if prefix:
            prefix[-1] = prefix[-1].rstrip("\n") + "\n"
        return prefix, first

    # ------------------------------------------------------------------ process(): planning
    def _plan_contexts(self, queries, contexts, titles, splitter: SentenceSplitter, strip: bool, timing: dict,
                       pairs: Sequence[tuple[int, int]] | None = None, query_tokens: list | None = None):
        """Sentences for every (query, context) pair -- or for ``pairs`` only -- then ONE batched tokenizer call
        for all of them."""
        plans: list[_ContextPlan] = []
        t_collect = t_norm = 0.0
        if pairs is None:
            pairs = [(qi, ci) for qi in range(len(queries)) for ci in range(len(contexts[qi]))]
        for qi, ci in pairs:
            for entry in (contexts[qi][ci],):
                if isinstance(entry, list):
                    manual = [str(s) for s in entry if str(s).strip()]
                    text = "".join(manual)
                else:
                    manual, text = None, entry
                prefix, first = self._resolve_prefix_sentences(titles[qi], ci)
                t0 = perf_counter()
                raw = [str(s) for s in prefix if s is not None]
                raw.extend(str(s) for s in (manual if manual is not None else splitter(str(text))) if s is not None)
                t1 = perf_counter()
                sentences = normalize_sentences(raw, str(text), strip)
                t_collect += t1 - t0
                t_norm += perf_counter() - t1
                plans.append(_ContextPlan(qi, ci, entry, str(text), sentences, prefix, [], first))
        t0 = perf_counter()
        flat = [s for p in plans for s in p.sentences]
        flat_tokens = tokenize_batch(self.tokenizer, flat)
        at = 0
        for p in plans:
            p.token_lists = flat_tokens[at : at + len(p.sentences)]
            at += len(p.sentences)
            p.prefix_token_counts = [len(t) for t in p.token_lists[: len(p.prefix_sentences)]]
        if query_tokens is None:
            query_tokens = [[int(t) for t in self.tokenizer.encode(q, add_special_tokens=False)] for q in queries]
        timing["sentence_collect_seconds"] += t_collect
        timing["sentence_normalize_seconds"] += t_norm
        timing["tokenize_seconds"] += perf_counter() - t0
        return plans, query_tokens

    def _fragmentize(self, plans: list[_ContextPlan], max_fragment_tokens: int, strip: bool, respect: bool, timing: dict):
        """standalone:795-843 for every plan; decoding (needed only to drop empty fragments) is batched."""
        t0 = perf_counter()
        for p in plans:
            p.fragments = split_token_lists(p.token_lists, max_fragment_tokens, keep_sentence_boundaries=respect)
            if not p.fragments:
                tokens = self.tokenizer.encode(fallback_sentence(p.context_text, strip), add_special_tokens=False)
                p.fragments = [Fragment([int(t) for t in tokens], 0, 0, 0)]
        timing["fragment_split_seconds"] += perf_counter() - t0
        t0 = perf_counter()
        all_frags = [f for p in plans for f in p.fragments]
        texts = decode_batch(self.tokenizer, [f.token_ids for f in all_frags])
        at = 0
        for p in plans:
            n = len(p.fragments)
            kept = filter_decodable(p.fragments, texts[at : at + n], strip)
            at += n
            p.fragments = kept if kept else p.fragments[:1]
        timing["fragment_decode_seconds"] += perf_counter() - t0

    def _assemble_blocks(self, fragments: list[Fragment], query_len: int, sep_len: int) -> list[list[Fragment]]:
        """Greedy packing into blocks of at most ``max_length - 2`` tokens (standalone:2222-2259); a fragment
        that does not fit starts a new block and is truncated to the block capacity."""
        if not fragments:
            return []
        available = self.max_length - 2
        base = query_len + sep_len
        capacity = max(1, available - base)
        blocks: list[list[Fragment]] = []
        cur: list[Fragment] = []
        cur_len = base
        for frag in fragments:
            if cur_len + frag.token_length <= available:
                cur.append(frag)
                cur_len += frag.token_length
                continue
            if cur:
                blocks.append(cur)
            if frag.token_length > capacity:
                frag = Fragment(frag.token_ids[:capacity], frag.sentence_index, frag.fragment_index, frag.global_index)
            cur = [frag]
            cur_len = base + frag.token_length
        if cur:
            blocks.append(cur)
        return blocks

    def _block_ids(self, query_tokens: list[int], ctx_tokens: list[int]) -> tuple[list[int], int]:
        """``[CLS] q [SEP] ctx [SEP]`` and the index where the context starts (standalone:2104-2184)."""
        if self._manual_special_tokens_required:
            ids: list[int] = []
            if self._manual_cls_token_id is not None:
                ids.append(self._manual_cls_token_id)
            ids.extend(query_tokens)
            if self._manual_sep_token_id is not None:
                ids.append(self._manual_sep_token_id)
            ids.extend(ctx_tokens)
            if self._manual_sep_token_id is not None and ctx_tokens:
                ids.append(self._manual_sep_token_id)
        else:
            ids = [int(t) for t in self.tokenizer.build_inputs_with_special_tokens(query_tokens, ctx_tokens)]
            if not ids:
                ids = list(query_tokens) + list(ctx_tokens)
        if not ctx_tokens:
            return ids, len(ids)
        # the reference locates the context by its FIRST occurrence in the block (standalone:2159-2178)
        n, m = len(ids), len(ctx_tokens)
        first = ctx_tokens[0]
        start = -1
        for i in range(0, n - m + 1):
            if ids[i] == first and ids[i : i + m] == ctx_tokens:
                start = i
                break
        if start < 0:
            start = len(self.tokenizer.build_inputs_with_special_tokens(query_tokens, []))
        return ids, start

    def _build_table(self, plans: list[_ContextPlan], query_tokens: list[list[int]], sep_len: int) -> BlockTable:
        table = BlockTable()
        for p in plans:
            q = query_tokens[p.query_idx]
            p.blocks = self._assemble_blocks(p.fragments, len(q), sep_len)
            slots_of: dict[int, list[int]] = {}
            for block in p.blocks:
                ctx = [t for f in block for t in f.token_ids]
                ids, cursor = self._block_ids(q, ctx)
                b = table.n_blocks
                table.block_ids.append(np.asarray(ids, dtype=np.int32))
                p.block_slots.append(b)
                for f in (block if ctx else []):  # no context tokens -> the reference records no ranges (standalone:2183)
                    start, end = cursor, cursor + f.token_length
                    cursor = end
                    offset = sum(p.prefix_token_counts[: f.sentence_index])  # title quirk, standalone:3076-3080
                    start = max(0, start - offset)
                    end = max(start, end - offset)
                    end = min(end, len(ids))
                    start = min(start, len(ids))
                    slots_of.setdefault(f.global_index, []).append(len(table.frag_block))
                    table.frag_block.append(b)
                    table.frag_local.append((start, end))
            p.sentence_base = table.n_sentences
            per_sentence: dict[int, list[int]] = {}
            for f in p.fragments:  # standalone:3094-3099
                if f.global_index in slots_of:
                    per_sentence.setdefault(f.sentence_index, []).extend(slots_of[f.global_index])
            for s in range(len(p.sentences)):
                table.sent_frag_index.extend(per_sentence.get(s, []))
                table.sent_offsets.append(len(table.sent_frag_index))
            p.n_fragments, p.n_blocks = len(p.fragments), len(p.blocks)
        return table

    def _sep_token_length(self) -> int:
        """``len(tokenizer.encode(sep_token))`` (standalone:2232), cached per tokenizer / separator."""
        sep = getattr(self.tokenizer, "sep_token", None) or ""
        key = (id(self.tokenizer), sep)
        cached = getattr(self, "_sep_len_cache", None)
        if cached is None or cached[0] != key:
            cached = self._sep_len_cache = (key, len(self.tokenizer.encode(sep, add_special_tokens=False)))
        return cached[1]

    def _pack_template(self):
        """Special-token template for the native packer, derived once per tokenizer state."""
        key = (id(self.tokenizer), self._manual_special_tokens_required, self._manual_cls_token_id,
               self._manual_sep_token_id)
        cached = getattr(self, "_pack_template_cache", None)
        if cached is None or cached[0] != key:
            template = host_pack.special_token_template(
                self.tokenizer, self._manual_special_tokens_required, self._manual_cls_token_id,
                self._manual_sep_token_id)
            cached = self._pack_template_cache = (key, template)
        return cached[1]

    def _pack_native(self, plans: list[_ContextPlan], query_tokens: list[list[int]], sep_len: int,
                     max_fragment_tokens: int, strip: bool, respect: bool) -> BlockTable | None:
        """Fragment windows -> filter -> blocks -> table in one ``opv_pack_build`` call (csrc/host_pack.cu).

        Returns None for the inputs the native packer leaves to the Python restatement below: a tokenizer whose
        ``build_inputs_with_special_tokens`` is not ``head + q + mid + ctx + tail``, and contexts whose sentences
        all tokenise to nothing (the reference's whole-context fallback fragment, standalone:813-816)."""
        if getattr(self, "host_pack_mode", "native") != "native" or not plans:
            return None
        template = self._pack_template()
        if template is None or any(not any(p.token_lists) for p in plans):
            return None
        packed = host_pack.pack_blocks(
            self.tokenizer, [t for p in plans for t in p.token_lists], [len(p.token_lists) for p in plans],
            [p.query_idx for p in plans], [len(p.prefix_sentences) for p in plans], query_tokens,
            template=template, max_length=self.max_length, max_fragment_tokens=max_fragment_tokens,
            keep_sentence_boundaries=respect, sep_len=sep_len, strip_sentences=strip)
        for c, p in enumerate(plans):
            b0, b1 = int(packed.ctx_block_offsets[c]), int(packed.ctx_block_offsets[c + 1])
            p.block_slots = list(range(b0, b1))
            p.sentence_base = int(packed.ctx_sentence_base[c])
            p.n_blocks = b1 - b0
            p.n_fragments = max(1, p.n_blocks)
        return packed.table

    # ------------------------------------------------------------------ process()
    def process(
        self,
        question: str | Sequence[str],
        context: str | Sequence[str] | Sequence[Sequence[str]],
        title: None | str | Sequence[str] | Sequence[Sequence[str]] = "first_sentence",
        first_line_as_title: bool = False,
        *,
        batch_size: int = 32,
        threshold: float | None = None,
        always_select_title: bool = False,
        reorder: bool = False,
        top_k: int | None = None,
        sentence_splitter: SentenceSplitter | Mapping[str, SentenceSplitter] | None = None,
        language: str | None = None,
        use_best_reranker_score: bool = True,
        zero_score_when_empty: bool = True,
        show_progress: bool = True,
        debug_messages: bool | Callable[[str], None] = False,
        enable_warnings: bool = True,
        strip_sentences: bool = False,
        respect_sentence_boundaries: bool = False,
        return_sentence_metrics: bool = False,
        return_sentence_texts: bool = False,
        show_inference_progress: bool | None = None,
        preprocess_workers: int | None = None,
        preprocess_batch_size: int | None = None,
        torch_dataloader_kwargs: Mapping[str, Any] | None = None,
    ) -> dict[str, Any]:
        """Same arguments and result dictionary as the reference (standalone:3314-3808).

        ``batch_size`` bounds a launch to ``batch_size * max_length`` packed tokens (at least 131072);
        results do not depend on it.  ``preprocess_workers`` / ``preprocess_batch_size`` /
        ``torch_dataloader_kwargs`` / progress flags are accepted for signature compatibility: host
        preprocessing here is one batched tokenizer call, not a DataLoader of per-context jobs."""
        del show_progress, show_inference_progress, enable_warnings, preprocess_workers
        del torch_dataloader_kwargs
        if self._scorer is None:
            raise RuntimeError("this OpenProvenceModel has no engine (constructed without weights)")
        batch_size = max(1, int(batch_size))
        threshold = self._resolve_process_threshold(threshold)
        t_start = perf_counter()
        splitter = resolve_sentence_splitter(sentence_splitter, language, getattr(self, "default_splitter_language", None))
        if isinstance(debug_messages, bool):
            debug = (lambda m: print(m, flush=True)) if debug_messages else None
        elif callable(debug_messages):
            debug = debug_messages
        else:
            raise TypeError("debug_messages must be a bool or a callable that accepts a string")

        timing = {k: 0.0 for k in ("sentence_collect_seconds", "sentence_normalize_seconds", "tokenize_seconds",
                                   "fragment_split_seconds", "fragment_decode_seconds")}
        queries, contexts, structure = self._normalize_inputs(question, context)
        contexts, titles = self._resolve_titles(queries, contexts, title, first_line_as_title=first_line_as_title)
        max_fragment_tokens = max(16, self.max_length - 2) if respect_sentence_boundaries else max(16, self.max_length // 2)
        sep_len = self._sep_token_length()

        # Host preparation (sentences -> tokens -> fragments -> packed block table) of chunk k+1 runs in a worker
        # thread while the device scores chunk k: the tokenizer releases the GIL, and so does this thread while it
        # waits for the GPU.  (The reference streams its jobs through a DataLoader for the same reason,
        # standalone:3510-3605.)  One chunk when the input is small.
        all_pairs = [(qi, ci) for qi in range(len(queries)) for ci in range(len(contexts[qi]))]
        if preprocess_batch_size:
            chunk_size = max(1, int(preprocess_batch_size))
            chunks = [all_pairs[i : i + chunk_size] for i in range(0, len(all_pairs), chunk_size)] or [[]]
        else:
            # 16, 32, 64, 64, ... contexts: the device is idle while the first chunk is prepared, so that one is small
            chunks, at, size = [], 0, 16
            while at < len(all_pairs):
                chunks.append(all_pairs[at : at + size])
                at += size
                size = min(64, size * 2)
            chunks = chunks or [[]]
        query_tokens = tokenize_batch(self.tokenizer, queries)  # same ids as tokenizer.encode(q, add_special_tokens=False)
        stage = {"assembly": 0.0, "inference": 0.0}

        def prepare(pairs_k):
            t_local = {k: 0.0 for k in timing}
            plans_k, _ = self._plan_contexts(queries, contexts, titles, splitter, strip_sentences, t_local,
                                             pairs=pairs_k, query_tokens=query_tokens)
            t0 = perf_counter()
            table_k = self._pack_native(plans_k, query_tokens, sep_len, max_fragment_tokens, strip_sentences,
                                        respect_sentence_boundaries)
            if table_k is None:
                self._fragmentize(plans_k, max_fragment_tokens, strip_sentences, respect_sentence_boundaries, t_local)
                t0 = perf_counter()
                table_k = self._build_table(plans_k, query_tokens, sep_len)
            return plans_k, table_k, t_local, perf_counter() - t0

        if hasattr(self._scorer, "max_tokens"):
            self._scorer.max_tokens = max(131072, batch_size * max(self.max_length, 1))
        plans: list[_ContextPlan] = []
        parts: list[dict[str, np.ndarray]] = []
        tickets: list[Any] = []
        two_phase = hasattr(self._scorer, "submit") and hasattr(self._scorer, "collect")
        block_base = sentence_base = 0
        from concurrent.futures import ThreadPoolExecutor

        class _Done:  # a single chunk is prepared inline: starting a worker thread costs more than it hides
            def __init__(self, value):
                self._value = value

            def result(self):
                return self._value

        with ThreadPoolExecutor(max_workers=1) if len(chunks) > 1 else contextlib.nullcontext() as pool:
            pending = pool.submit(prepare, chunks[0]) if pool is not None else _Done(prepare(chunks[0]))
            for k in range(len(chunks)):
                plans_k, table_k, t_local, t_asm = pending.result()
                if k + 1 < len(chunks):
                    pending = pool.submit(prepare, chunks[k + 1])
                for key, value in t_local.items():
                    timing[key] += value
                stage["assembly"] += t_asm
                t0 = perf_counter()
                if table_k.n_blocks:
                    if two_phase:  # data-parallel scorer: local device work now, ONE collective after the last chunk
                        tickets.append(self._scorer.submit(table_k, threshold))
                    else:
                        parts.append(self._scorer.run(table_k, threshold))
                stage["inference"] += perf_counter() - t0
                for p in plans_k:  # chunk-local slots -> positions in the concatenated result arrays
                    p.block_slots = [b + block_base for b in p.block_slots]
                    p.sentence_base += sentence_base
                block_base += table_k.n_blocks
                sentence_base += table_k.n_sentences
                plans.extend(plans_k)
        if two_phase:
            t0 = perf_counter()
            parts = self._scorer.collect(tickets)
            stage["inference"] += perf_counter() - t0
        preprocess_time = sum(timing.values())
        assembly_time, inference_time = stage["assembly"], stage["inference"]
        if parts:
            scored = {key: np.concatenate([part[key] for part in parts]) for key in ("rank_score", "sent_prob", "keep")}
        else:
            scored = {"rank_score": np.zeros(0, np.float32), "sent_prob": np.zeros(0), "keep": np.zeros(0, bool)}

        t0 = perf_counter()
        per_query = self._postprocess(
            queries, contexts, plans, scored, threshold=threshold, always_select_title=always_select_title,
            use_best_reranker_score=use_best_reranker_score, first_line_as_title=first_line_as_title,
            zero_score_when_empty=zero_score_when_empty, want_probs=return_sentence_metrics,
            want_texts=return_sentence_texts)
        post_time = perf_counter() - t0
        total_time = perf_counter() - t_start
        trace = ProcessPerformanceTrace(
            preprocess_seconds=preprocess_time, assembly_seconds=assembly_time, inference_seconds=inference_time,
            postprocess_seconds=post_time, total_seconds=total_time, **timing)
        if debug is not None:
            debug("[OpenProvenceModel] Timing: " + " ".join(f"{k}={v:.3f}s" for k, v in trace.as_dict().items()))

        if reorder:
            per_query = self._apply_reordering(per_query, top_k)
        return self._shape_result(per_query, structure, trace, return_sentence_metrics, return_sentence_texts)

    # ------------------------------------------------------------------ process(): postprocess
    def _postprocess(self, queries, contexts, plans, scored, *, threshold, always_select_title,
                     use_best_reranker_score, first_line_as_title, zero_score_when_empty, want_probs, want_texts):
        """Keep flags -> strings (standalone:2962-3202).  Returns per query a dict of parallel lists."""
        by_key = {(p.query_idx, p.context_idx): p for p in plans}
        rank_score, sent_prob, keep = scored["rank_score"], scored["sent_prob"], scored["keep"]
        out = []
        for qi in range(len(queries)):
            q = {"pruned": [], "score": [], "compression": [], "kept": [], "removed": [], "title": [], "probs": []}
            for ci, entry in enumerate(contexts[qi]):
                p = by_key.get((qi, ci))
                prefix = list(p.prefix_sentences) if p else []
                fallback_title: Any = None
                if first_line_as_title and prefix:
                    fallback_title = prefix[0] if len(prefix) == 1 else list(prefix)
                if p is None or not p.n_fragments:
                    q["pruned"].append(entry)
                    q["score"].append(None)
                    q["compression"].append(0.0)
                    q["kept"].append([entry] if entry else [])
                    q["removed"].append([])
                    q["title"].append(fallback_title)
                    q["probs"].append([])
                    continue
                if not p.n_blocks:
                    q["pruned"].append(entry)
                    q["score"].append(None)
                    q["compression"].append(0.0)
                    q["kept"].append(p.sentences)
                    q["removed"].append([])
                    q["title"].append(fallback_title)
                    q["probs"].append([1.0] * len(p.sentences))
                    continue
                scores = [float(rank_score[b]) for b in p.block_slots]
                ranking: float | None = (max(scores) if use_best_reranker_score else scores[0]) if scores else None
                n = len(p.sentences)
                probs = [float(v) for v in sent_prob[p.sentence_base : p.sentence_base + n]]
                flags = [bool(v) for v in keep[p.sentence_base : p.sentence_base + n]]
                prefix_len = len(prefix)
                title_idx = None
                if always_select_title:  # standalone:3108-3112
                    if prefix_len > 0:
                        title_idx = 0
                    elif p.title_is_first_sentence and n > prefix_len:
                        title_idx = prefix_len
                if title_idx is not None and any(flags):  # forced only if something passes (standalone:3124-3132)
                    flags[title_idx] = True
                kept_s = [s for s, k in zip(p.sentences, flags) if k]
                removed_s = [s for s, k in zip(p.sentences, flags) if not k]
                pruned = "".join(s for i, (s, k) in enumerate(zip(p.sentences, flags)) if k and i >= prefix_len)
                original = p.context_text
                compression = (len(original) - len(pruned)) / max(len(original), 1) * 100.0
                if zero_score_when_empty and not pruned.strip():
                    ranking = 0.0
                q["pruned"].append(pruned)
                q["score"].append(ranking)
                q["compression"].append(compression)
                q["kept"].append(kept_s)
                q["removed"].append(removed_s)
                q["title"].append((prefix[0] if len(prefix) == 1 else list(prefix)) if prefix else None)
                q["probs"].append(probs)
            out.append(q)
        return out

    @staticmethod
    def _apply_reordering(per_query: list[dict[str, list]], top_k: int | None) -> list[dict[str, list]]:
        """Stable sort by score descending, ``None`` last, optional top-k (standalone:3204-3312)."""
        limit = None if top_k is None else max(0, int(top_k))
        out = []
        for q in per_query:
            scores = q["score"]
            if not scores:
                out.append(q)
                continue
            order = sorted(range(len(scores)), key=lambda i: float("-inf") if scores[i] is None else float(scores[i]),
                           reverse=True)
            if limit is not None:
                order = order[:limit]
            out.append({k: [v[i] for i in order] for k, v in q.items()})
        return out

    @staticmethod
    def _shape_result(per_query, structure: str, trace: ProcessPerformanceTrace, want_probs: bool, want_texts: bool):
        """Back to the caller's input structure (standalone:3740-3805)."""
        fields = {"pruned": "pruned_context", "score": "reranking_score", "compression": "compression_rate",
                  "title": "title"}
        if want_texts:
            fields.update(kept="kept_sentences", removed="removed_sentences")
        if want_probs:
            fields.update(probs="sentence_probabilities")
        empty = {"pruned": "", "score": None, "compression": 0.0, "title": None, "kept": [], "removed": [], "probs": []}
        payload: dict[str, Any] = {}
        for key, name in fields.items():
            nested = [q[key] for q in per_query]
            if structure == "str" and per_query:
                first = nested[0]
                if key == "probs":
                    value: Any = first[0] if (nested and first) else nested
                else:
                    value = first[0] if first else empty[key]
            elif structure == "list" and per_query:
                value = nested[0]
            elif structure == "aligned" and per_query:
                value = [v[0] if v else empty[key] for v in nested]
            else:
                value = nested
            payload[name] = value
        ordered = {k: payload[k] for k in ("pruned_context", "reranking_score", "compression_rate", "title")}
        ordered["timing"] = trace.as_dict()
        ordered["performance_trace"] = trace
        for name in ("kept_sentences", "removed_sentences", "sentence_probabilities"):
            if name in payload:
                ordered[name] = payload[name]
        return ordered


__all__ = [
    "OpenProvenceModel",
    "OpenProvenceConfig",
    "OpenProvenceRawPrediction",
    "OpenProvenceOutput",
    "ProcessPerformanceTrace",
]
