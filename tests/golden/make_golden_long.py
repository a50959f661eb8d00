#!/usr/bin/env python
"""Long-sequence fixture: the reference forward (fp64) on the tiny golden checkpoint at max_position_embeddings
(one 8192-token block and one 4097-token block).  Build container only.  -> ``forward_long_tiny.npz``"""

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
    config = ref.OpenProvenceConfig(base_model_config=cfg["base_model_config"], tokenizer_name_or_path="tiny_ckpt",
                                    pruning_config=cfg["pruning_config"], max_length=8192)
    model = ref.OpenProvenceModel(config)
    model.load_state_dict(load_file(str(HERE / "tiny_ckpt" / "model.safetensors")), strict=True)
    torch.nn.Module.to(model, dtype=torch.float64)
    model.eval()
    rng = np.random.default_rng(99)
    lengths = [8192, 4097]
    ids = np.zeros((2, 8192), dtype=np.int64)
    mask = np.zeros((2, 8192), dtype=np.int64)
    for b, n in enumerate(lengths):
        ids[b, :n] = rng.integers(5, cfg["base_model_config"]["vocab_size"], size=n)
        ids[b, 0] = 1
        mask[b, :n] = 1
    with torch.inference_mode():
        out = model.forward(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask), return_dict=True)
    prune = out.pruning_logits.numpy()
    np.savez_compressed(HERE / "forward_long_tiny.npz", lengths=np.asarray(lengths),
                        input_ids=np.concatenate([ids[b, :n] for b, n in enumerate(lengths)]).astype(np.int32),
                        ranking_logits_f64=out.ranking_logits.numpy(),
                        pruning_logits_f64=np.concatenate([prune[b, :n] for b, n in enumerate(lengths)]),
                        versions=json.dumps({"torch": torch.__version__, "transformers": __import__("transformers").__version__}))
    print("wrote forward_long_tiny.npz", prune.shape, float(np.abs(prune).max()))


if __name__ == "__main__":
    main()
