#!/usr/bin/env python
"""The reference's OWN bf16 forward on the forward fixture's inputs (build container only).

north_star states the bf16 tolerance against "the reference PyTorch/HF forward"; the reference runs bf16 on
CUDA by default (standalone:219-233) with bf16 weights, a bf16 residual stream and bf16-rounded RoPE tables
(HF:172).  This script loads the tiny golden checkpoint with the reference's ``from_pretrained`` in
``torch.bfloat16`` (on CPU: same arithmetic, no GPU here), runs its forward on ``forward_tiny.npz``'s batch and
stores the logits (the model is built by the reference's constructor and filled from the checkpoint, as
``make_golden.py`` does: the reference's ``from_pretrained`` trips over a transformers 5.x tied-weights API), so that the bf16 engine can be judged against the reference's bf16 error, not only
against an absolute number.  -> ``forward_tiny_bf16.npz``
"""

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
    from transformers import AutoTokenizer

    fast = AutoTokenizer.from_pretrained(str(HERE / "tiny_ckpt"))
    ref.AutoTokenizer.from_pretrained = staticmethod(lambda *_a, **_k: mg.Tokenizer457Shim(fast))
    golden = np.load(HERE / "forward_tiny.npz")
    ids, mask = torch.from_numpy(golden["input_ids"]), torch.from_numpy(golden["attention_mask"])
    out = {}
    for name, dtype in (("bf16", torch.bfloat16), ("f32check", torch.float32)):
        cfg = json.loads((HERE / "tiny_ckpt" / "config.json").read_text())
        config = ref.OpenProvenceConfig(base_model_config=cfg["base_model_config"], tokenizer_name_or_path="tiny_ckpt",
                                        pruning_config=cfg["pruning_config"], max_length=cfg["max_length"])
        model = ref.OpenProvenceModel(config)
        from safetensors.torch import load_file

        model.load_state_dict(load_file(str(HERE / "tiny_ckpt" / "model.safetensors")), strict=True)
        torch.nn.Module.to(model, dtype=dtype)  # what the reference does after loading (standalone:1649-1653)
        model.eval()
        with torch.inference_mode():
            fw = model.forward(input_ids=ids, attention_mask=mask, return_dict=True)
        out[name] = (fw.ranking_logits.float().numpy(), fw.pruning_logits.float().numpy())
    # the fp32 reload must reproduce the stored fp32 fixture bit for bit: proves the checkpoint round trip
    assert np.array_equal(out["f32check"][0], golden["ranking_logits_f32"])
    valid = golden["attention_mask"][..., None].astype(bool)
    assert np.array_equal(np.where(valid, out["f32check"][1], 0), np.where(valid, golden["pruning_logits_f32"], 0))
    rank, prune = out["bf16"]
    e_rank = np.abs(rank - golden["ranking_logits_f64"]).max()
    e_prune = np.abs((prune - golden["pruning_logits_f64"]) * golden["attention_mask"][..., None]).max()
    np.savez_compressed(HERE / "forward_tiny_bf16.npz", ranking_logits_bf16=rank, pruning_logits_bf16=prune,
                        versions=json.dumps({"torch": torch.__version__, "transformers": __import__("transformers").__version__}))
    print(f"reference bf16 forward vs its fp64 forward: rank {e_rank:.3e}  prune {e_prune:.3e}")


if __name__ == "__main__":
    main()
