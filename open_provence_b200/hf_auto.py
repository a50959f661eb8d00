"""Hugging Face ``Auto*`` integration of the drop-in (reference: standalone:3810-3906, encoder.py:1079-1085).

The reference ships its checkpoints with ``auto_map`` entries that make ``AutoModel.from_pretrained(...,
trust_remote_code=True)`` import ``modeling_open_provence_standalone.py`` from the checkpoint directory.  Here the same
four entry points resolve to the sm_100a engine instead:

  * :class:`OpenProvenceForSequenceClassification` / :class:`OpenProvenceForTokenClassification` -- the two wrapper
    classes the reference's ``auto_map`` names (plus the ``OpenProvenceEncoder*`` aliases);
  * :func:`register_auto_classes` -- registers ``model_type = "open_provence"`` with ``AutoConfig`` / ``AutoModel`` /
    ``AutoModelForSequenceClassification`` / ``AutoModelForTokenClassification`` so that, after one call, callers that
    load through the ``Auto*`` factories (``scripts/eval_datasets.py:300-312``) get this engine without remote code;
  * :data:`AUTO_MAP` / :func:`with_auto_map` -- what ``save_pretrained`` writes so a saved checkpoint stays loadable by
    the reference.
"""

from __future__ import annotations

from typing import Any

import torch

from .modeling import OpenProvenceModel, OpenProvenceOutput

AUTO_MAP = {
    "AutoConfig": "modeling_open_provence_standalone.OpenProvenceConfig",
    "AutoModel": "modeling_open_provence_standalone.OpenProvenceForSequenceClassification",
    "AutoModelForSequenceClassification": "modeling_open_provence_standalone.OpenProvenceForSequenceClassification",
    "AutoModelForTokenClassification": "modeling_open_provence_standalone.OpenProvenceForTokenClassification",
}
ARCHITECTURES = ["OpenProvenceForSequenceClassification"]


def with_auto_map(config_dict: dict[str, Any]) -> dict[str, Any]:
    """``config.json`` payload with the reference's ``architectures`` / ``auto_map`` (encoder.py:1078-1085)."""
    out = dict(config_dict)
    out["architectures"] = list(ARCHITECTURES)
    out["auto_map"] = dict(AUTO_MAP)
    return out


class OpenProvenceForSequenceClassification(OpenProvenceModel):
    """standalone:3813-3830: same forward; ``.logits`` are the ranking logits."""


class OpenProvenceForTokenClassification(OpenProvenceModel):
    """standalone:3833-3899: ``.logits`` are the pruning logits ``[B, S, 2]``; the ranking logits ride along."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.num_ranking_labels = self.num_labels
        self.num_labels = self.num_pruning_labels

    def forward(self, input_ids: torch.Tensor | None = None, attention_mask: torch.Tensor | None = None,
                labels: torch.Tensor | None = None, return_dict: bool | None = None, **kwargs: Any):
        base = OpenProvenceModel.forward(self, input_ids=input_ids, attention_mask=attention_mask, labels=None,
                                         return_dict=True, **kwargs)
        prune, rank = base["pruning_logits"], base["ranking_logits"]
        loss = None
        if labels is not None:  # standalone:3868-3878
            labels = labels.to(prune.device)
            if attention_mask is not None:
                active = attention_mask.to(prune.device).reshape(-1) == 1
                logits, target = prune.reshape(-1, self.num_labels)[active], labels.reshape(-1)[active]
                loss = (torch.nn.functional.cross_entropy(logits, target) if logits.numel() > 0
                        else torch.tensor(0.0, device=prune.device))
            else:
                loss = torch.nn.functional.cross_entropy(prune.reshape(-1, self.num_labels), labels.reshape(-1))
        if return_dict is not None and not return_dict:
            return ((loss,) if loss is not None else ()) + (prune,)
        return OpenProvenceOutput(loss=loss, logits=prune, ranking_logits=rank, pruning_logits=prune,
                                  hidden_states=None, attentions=None)

    __call__ = forward


OpenProvenceEncoderForSequenceClassification = OpenProvenceForSequenceClassification
OpenProvenceEncoderForTokenClassification = OpenProvenceForTokenClassification

_REGISTERED = False


def _dtype_name(value: Any) -> str | None:
    return None if value is None else str(value).replace("torch.", "")


def _stored_dtype(path: Any) -> Any:
    import json
    from pathlib import Path

    cfg = Path(str(path)) / "config.json"
    if not cfg.exists():
        return None
    data = json.loads(cfg.read_text())
    return data.get("dtype", data.get("torch_dtype"))


def register_auto_classes() -> None:
    """Make ``AutoConfig`` / ``AutoModel*`` resolve ``model_type == "open_provence"`` to this package.

    ``transformers`` only accepts ``PretrainedConfig`` subclasses in its registries, so a thin config carrying the
    checkpoint's keys is registered; the model classes ignore it and read ``config.json`` themselves
    (``OpenProvenceModel.from_pretrained``).  Idempotent."""
    global _REGISTERED
    if _REGISTERED:
        return
    from transformers import (AutoConfig, AutoModel, AutoModelForSequenceClassification,
                              AutoModelForTokenClassification, PretrainedConfig)

    class OpenProvenceAutoConfig(PretrainedConfig):
        model_type = "open_provence"

        def __init__(self, **kwargs: Any) -> None:
            for key in ("mode", "base_model_name_or_path", "base_model_config", "tokenizer_name_or_path",
                        "pruning_config", "max_length", "num_pruning_labels", "encoder_architecture",
                        "default_threadshold"):
                setattr(self, key, kwargs.pop(key, None))
            super().__init__(**kwargs)

    def adopt(cls: type) -> type:
        """Subclass whose ``from_pretrained`` accepts what the Auto factories pass (``config=...``, positional
        model args) and that names its config class, as ``_BaseAutoModelClass.register`` requires."""

        class _Auto(cls):  # type: ignore[misc, valid-type]
            config_class = OpenProvenceAutoConfig

            @classmethod
            def from_pretrained(cls_, pretrained_model_name_or_path, *model_args: Any, **kwargs: Any):
                config = kwargs.get("config")
                if config is not None and "dtype" not in kwargs and "torch_dtype" not in kwargs:
                    # AutoConfig swallows a caller's dtype= / torch_dtype= into config.dtype: hand it back unless it
                    # is just what config.json stores
                    requested = getattr(config, "dtype", None)
                    stored = _stored_dtype(pretrained_model_name_or_path)
                    if requested is not None and _dtype_name(requested) != _dtype_name(stored):
                        kwargs["dtype"] = _dtype_name(requested)
                for key in ("config", "_from_auto", "_commit_hash", "adapter_kwargs", "use_safetensors", "revision",
                            "cache_dir", "force_download", "local_files_only", "token", "code_revision", "subfolder",
                            "proxies", "resume_download", "use_auth_token"):
                    kwargs.pop(key, None)
                return super().from_pretrained(pretrained_model_name_or_path, **kwargs)

        _Auto.__name__ = cls.__name__
        _Auto.__qualname__ = cls.__qualname__
        return _Auto

    seq = adopt(OpenProvenceForSequenceClassification)
    tok = adopt(OpenProvenceForTokenClassification)
    AutoConfig.register("open_provence", OpenProvenceAutoConfig, exist_ok=True)
    AutoModel.register(OpenProvenceAutoConfig, seq, exist_ok=True)
    AutoModelForSequenceClassification.register(OpenProvenceAutoConfig, seq, exist_ok=True)
    AutoModelForTokenClassification.register(OpenProvenceAutoConfig, tok, exist_ok=True)
    _REGISTERED = True


__all__ = [
    "AUTO_MAP", "ARCHITECTURES", "with_auto_map", "register_auto_classes",
    "OpenProvenceForSequenceClassification", "OpenProvenceForTokenClassification",
    "OpenProvenceEncoderForSequenceClassification", "OpenProvenceEncoderForTokenClassification",
]
