#!/usr/bin/env python
"""Mean-pooling fixture: the reference forward (fp64) on the tiny golden checkpoint's weights with
``classifier_pooling = "mean"`` in the backbone config (HF ModernBertForSequenceClassification pools the masked
mean of the final hidden states instead of the CLS row; ModernBERT-base derived checkpoints ship with it).
Ragged right-padded batch, lengths on both sides of the 256-row chunk the device kernel reduces in.
Build container only.  -> ``forward_tiny_mean.npz``"""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden as mg  # noqa: E402


def main() -> None:
    ref = mg.load_reference_module()
    from safetensors.torch import load_file
    from transformers import AutoTokenizer

    fast = AutoTokenizer.from_pretrained(str(HERE / "tiny_ckpt"))
    ref.AutoTokenizer.from_pretrained = staticmethod(lambda *_a, **_k: mg.Tokenizer457Shim(fast))
    cfg = json.loads((HERE / "tiny_ckpt" / "config.json").read_text())
    backbone = dict(cfg["base_model_config"], classifier_pooling="mean")
    config = ref.OpenProvenceConfig(base_model_config=backbone, tokenizer_name_or_path="tiny_ckpt",
                                    pruning_config=cfg["pruning_config"], max_length=1024)
    model = ref.OpenProvenceModel(config)
    assert model.ranking_model.config.classifier_pooling == "mean"
    model.load_state_dict(load_file(str(HERE / "tiny_ckpt" / "model.safetensors")), strict=True)
    torch.nn.Module.to(model, dtype=torch.float64)
    model.eval()
    rng = np.random.default_rng(123)
    lengths = [1, 2, 31, 255, 256, 257, 513, 700, 1024]
    width = max(lengths)
    ids = np.zeros((len(lengths), width), dtype=np.int64)
    mask = np.zeros((len(lengths), width), dtype=np.int64)
    for b, n in enumerate(lengths):
        ids[b, :n] = rng.integers(5, backbone["vocab_size"], size=n)
        ids[b, 0] = 1
        mask[b, :n] = 1
    with torch.inference_mode():
        out = model.forward(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask), return_dict=True)
        cls_cfg = ref.OpenProvenceConfig(base_model_config=cfg["base_model_config"], tokenizer_name_or_path="tiny_ckpt",
                                         pruning_config=cfg["pruning_config"], max_length=1024)
        cls_model = ref.OpenProvenceModel(cls_cfg)
        cls_model.load_state_dict(load_file(str(HERE / "tiny_ckpt" / "model.safetensors")), strict=True)
        torch.nn.Module.to(cls_model, dtype=torch.float64)
        cls_out = cls_model.eval().forward(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask),
                                           return_dict=True)
    prune = out.pruning_logits.numpy()
    gap = float(np.abs(out.ranking_logits.numpy() - cls_out.ranking_logits.numpy()).max())
    assert gap > 1e-3, "mean and cls pooling must differ for the fixture to mean anything"
    np.savez_compressed(HERE / "forward_tiny_mean.npz", lengths=np.asarray(lengths),
                        input_ids=np.concatenate([ids[b, :n] for b, n in enumerate(lengths)]).astype(np.int32),
                        ranking_logits_f64=out.ranking_logits.numpy(),
                        pruning_logits_f64=np.concatenate([prune[b, :n] for b, n in enumerate(lengths)]),
                        versions=json.dumps({"torch": torch.__version__, "transformers": __import__("transformers").__version__}))
    print("wrote forward_tiny_mean.npz", out.ranking_logits.numpy().ravel()[:4], "cls/mean gap", gap)


if __name__ == "__main__":
    main()
