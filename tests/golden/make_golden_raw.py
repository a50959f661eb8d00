#!/usr/bin/env python
"""Fixtures for ``get_raw_predictions[_batch]`` / ``predict_with_thresholds`` (standalone:1742-1881), produced by
running the reference's methods on the tiny golden checkpoint (build container only).  -> ``raw_tiny.json``"""

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
                                    pruning_config=cfg["pruning_config"], max_length=160)
    model = ref.OpenProvenceModel(config)
    model.load_state_dict(load_file(str(HERE / "tiny_ckpt" / "model.safetensors")), strict=True)
    model.eval()

    recorded = []
    orig_forward = ref.OpenProvenceModel.forward

    def recording_forward(self, input_ids=None, attention_mask=None, **kw):
        out = orig_forward(self, input_ids=input_ids, attention_mask=attention_mask, **kw)
        for b in range(input_ids.shape[0]):
            n = int(attention_mask[b].sum())
            recorded.append({"ids": input_ids[b, :n].tolist(),
                             "rank_logits": out.ranking_logits[b].double().tolist(),
                             "prune_logits": out.pruning_logits[b, :n].double().tolist()})
        return out

    ref.OpenProvenceModel.forward = recording_forward
    query = "What are bananas?"
    contexts = ["Bananas are botanically berries. ", "They grow in clusters! ", mg.EN_DOC]  # third one is truncated at 160
    batch = [contexts, [mg.JA_DOC], ["short."]]
    queries = [query, "東京タワーの高さは？", "q"]
    out = {"max_length": 160, "query": query, "contexts": contexts, "batch_queries": queries, "batch_contexts": batch}
    raws = model.get_raw_predictions_batch(queries, batch)
    out["raw_batch"] = [{"ranking_score": r.ranking_score, "pruning_probs": np.asarray(r.pruning_probs, dtype=np.float64).tolist(),
                         "context_ranges": [list(x) for x in r.context_ranges]} for r in raws]
    single = model.get_raw_predictions(query, contexts)
    out["raw_single"] = {"ranking_score": single.ranking_score, "context_ranges": [list(x) for x in single.context_ranges]}
    out["thresholds"] = []
    for majority in (False, True):
        res = model.predict_with_thresholds(query, contexts, [0.05, 0.1, 0.5], use_majority=majority)
        out["thresholds"].append({"use_majority": majority, "ranking_score": res["ranking_score"],
                                  "predictions": {str(k): v for k, v in res["predictions"].items()},
                                  "context_ranges": [list(x) for x in res["context_ranges"]]})
    out["blocks"] = recorded
    (HERE / "raw_tiny.json").write_text(json.dumps(out, ensure_ascii=False))
    print("wrote raw_tiny.json:", len(recorded), "forward rows;", out["thresholds"][0]["predictions"], out["raw_single"])


if __name__ == "__main__":
    main()
