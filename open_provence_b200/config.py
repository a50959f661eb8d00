"""``OpenProvenceConfig`` -- the checkpoint's ``config.json`` (reference: standalone:1246-1302).

A plain Python object (the engine does not need ``transformers.PretrainedConfig``); it reads and
writes the same JSON keys, including the legacy mis-spelt ``default_threadshold``.
"""

from __future__ import annotations

import json
import warnings
from pathlib import Path
from typing import Any, Mapping

DEFAULT_PROCESS_THRESHOLD = 0.1  # standalone:57


class OpenProvenceConfig:
    model_type = "open_provence"

    def __init__(
        self,
        mode: str = "reranking_pruning",
        base_model_name_or_path: str | None = None,
        base_model_config: Mapping[str, Any] | None = None,
        tokenizer_name_or_path: str | None = None,
        pruning_config: Mapping[str, Any] | None = None,
        max_length: int = 512,
        num_labels: int | None = None,
        num_pruning_labels: int | None = None,
        encoder_architecture: str | None = None,
        **kwargs: Any,
    ) -> None:
        raw_threadshold = kwargs.pop("default_threadshold", None)
        alt_threshold = kwargs.pop("default_threshold", None)
        kwargs.pop("splitter_default_language", None)
        kwargs.pop("standalone_process_default_language", None)
        id2label = kwargs.get("id2label")
        self.mode = mode
        self.base_model_name_or_path = base_model_name_or_path
        if base_model_config is not None and hasattr(base_model_config, "to_dict"):
            base_model_config = base_model_config.to_dict()
        self.base_model_config = dict(base_model_config) if base_model_config is not None else None
        self.tokenizer_name_or_path = tokenizer_name_or_path
        self.pruning_config = dict(pruning_config or {})
        self.max_length = int(max_length)
        self.encoder_architecture = encoder_architecture
        if num_labels is None:
            num_labels = len(id2label) if id2label else 1  # HF serialises num_labels as id2label
        self.num_labels = int(num_labels)
        self.num_pruning_labels = 2 if num_pruning_labels is None else int(num_pruning_labels)
        self.default_threadshold: float | None = None
        if raw_threadshold is not None:
            try:
                self.default_threadshold = float(raw_threadshold)
            except (TypeError, ValueError) as exc:
                raise TypeError(
                    "Config value 'default_threadshold' must be a numeric type convertible to float."
                ) from exc
        elif alt_threshold is not None:
            warnings.warn(
                "Config key 'default_threshold' detected. Did you intend 'default_threadshold'? "
                "Using the provided value for backwards compatibility.",
                RuntimeWarning,
                stacklevel=2,
            )
            try:
                self.default_threadshold = float(alt_threshold)
            except (TypeError, ValueError) as exc:
                raise TypeError(
                    "Config value 'default_threshold' must be a numeric type convertible to float."
                ) from exc
        self.default_threshold = self.default_threadshold
        self._name_or_path = kwargs.pop("_name_or_path", "")
        self.extra = kwargs

    @classmethod
    def from_dict(cls, data: Mapping[str, Any]) -> "OpenProvenceConfig":
        payload = dict(data)
        payload.pop("model_type", None)
        if payload.get("default_threadshold") is not None:
            payload.pop("default_threshold", None)  # HF writes both spellings; the legacy one wins silently
        return cls(**payload)

    @classmethod
    def from_pretrained(cls, path: str | Path) -> "OpenProvenceConfig":
        cfg_path = Path(path) / "config.json"
        if not cfg_path.exists():
            raise FileNotFoundError(f"{cfg_path} not found")
        cfg = cls.from_dict(json.loads(cfg_path.read_text()))
        cfg._name_or_path = str(path)
        return cfg

    def to_dict(self) -> dict[str, Any]:
        """Known keys + every unknown key the loaded ``config.json`` carried (``self.extra``: vocab_size, hidden_size,
        transformers_version, ... -- the reference's config round-trips them, encoder.py:1050-1088)."""
        out = dict(self.extra)
        out.update({
            "model_type": self.model_type,
            "mode": self.mode,
            "base_model_name_or_path": self.base_model_name_or_path,
            "base_model_config": self.base_model_config,
            "tokenizer_name_or_path": self.tokenizer_name_or_path,
            "pruning_config": self.pruning_config,
            "max_length": self.max_length,
            "num_labels": self.num_labels,
            "num_pruning_labels": self.num_pruning_labels,
            "encoder_architecture": self.encoder_architecture,
            "default_threadshold": self.default_threadshold,
        })
        return out

    def resolve_default_threshold(self) -> float:
        """standalone:1422-1431."""
        return DEFAULT_PROCESS_THRESHOLD if self.default_threadshold is None else float(self.default_threadshold)
