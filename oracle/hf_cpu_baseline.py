"""CPU baseline: the reference's own arithmetic path, timed on the host cores.

TEST / BENCH INFRASTRUCTURE -- see ``oracle/__init__.py``.  Not imported by the product.

The reference's forward (standalone:1666-1739) is
``AutoModelForSequenceClassification.from_config(ModernBertConfig)`` (standalone:1341) + ``Linear(H, 2)``
(standalone:420), run in fp32 on CPU (standalone:238-244) with HF's default attention (sdpa), followed by
the CPU sigmoid / softmax (standalone:2893-2924) and the numpy per-sentence means (standalone:3065-3134).
``/root/reference`` cannot travel to the GPU box, but the arithmetic lives in the installed
``transformers`` package, so this module drives exactly that library path (kind = "port": the glue is
restated here, the kernels are the reference's).
"""

from __future__ import annotations

import time
from typing import Any, Mapping

import numpy as np
import torch

from . import postprocess_numpy as opp


def build_hf_model(backbone_cfg: Mapping[str, Any], state_dict: Mapping[str, torch.Tensor], num_labels: int = 1):
    """(ModernBertForSequenceClassification, pruning Linear) in fp32 eval mode with ``state_dict`` loaded."""
    from transformers import AutoConfig, AutoModelForSequenceClassification

    cfg = dict(backbone_cfg)
    model_type = cfg.pop("model_type", "modernbert")
    hf_cfg = AutoConfig.for_model(model_type, **cfg)  # standalone:1362
    hf_cfg.num_labels = num_labels
    model = AutoModelForSequenceClassification.from_config(hf_cfg)  # standalone:1341
    prefix = "ranking_model."
    backbone = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
    missing, unexpected = model.load_state_dict(backbone, strict=False)
    if unexpected or [m for m in missing if "inv_freq" not in m]:
        raise RuntimeError(f"state dict mismatch: missing={missing} unexpected={unexpected}")
    head = torch.nn.Linear(int(backbone_cfg["hidden_size"]), 2)
    head.weight.data.copy_(state_dict["pruning_head.classifier.weight"])
    head.bias.data.copy_(state_dict["pruning_head.classifier.bias"])
    return model.float().eval(), head.float().eval()


@torch.inference_mode()
def forward_padded(model, head, sequences: list[list[int]], pad_id: int = 0):
    """Right-padded batch through the HF model (standalone:2832-2903) -> (rank [B, labels], list of prune [n_i, 2])."""
    lengths = [len(s) for s in sequences]
    S = max(lengths)
    ids = torch.full((len(sequences), S), pad_id, dtype=torch.long)
    mask = torch.zeros((len(sequences), S), dtype=torch.long)
    for b, s in enumerate(sequences):
        ids[b, : len(s)] = torch.tensor(s, dtype=torch.long)
        mask[b, : len(s)] = 1
    out = model(input_ids=ids, attention_mask=mask, output_hidden_states=True, return_dict=True)
    prune = head(out.hidden_states[-1])
    return out.logits.float().numpy(), [prune[b, :n].float().numpy() for b, n in enumerate(lengths)]


def score_workload(model, head, workload: Mapping[str, Any], pair_slice: slice, threshold: float, batch_size: int = 32):
    """The hot path on CPU for a slice of a synthetic workload: forward, score conversion, sentence prune."""
    cu = workload["cu_seqlens"]
    blocks = list(range(*pair_slice.indices(len(cu) - 1)))
    results = {"rank_score": [], "sent_prob": [], "keep": [], "rank_logits": [], "prune_logits": [], "frag_index": []}
    for at in range(0, len(blocks), batch_size):
        chunk = blocks[at : at + batch_size]
        seqs = [workload["ids"][cu[b] : cu[b + 1]].tolist() for b in chunk]
        rank, prunes = forward_padded(model, head, seqs)
        for b, r, pr in zip(chunk, rank, prunes):
            results["rank_logits"].append(np.asarray(r, dtype=np.float32))
            results["prune_logits"].append(np.asarray(pr, dtype=np.float32))
            results["rank_score"].append(opp.ranking_score_from_logits(r))
            probs = opp.keep_probs_from_logits(pr)
            sel = np.nonzero(workload["frag_block"] == b)[0]
            for f in sel:
                s, e = workload["frag_ranges"][f] - cu[b]
                mean = 1.0 if e <= s else float(probs[s:e].mean())
                mean = max(0.0, min(float(np.mean([mean])), 1.0))
                results["sent_prob"].append(mean)
                results["frag_index"].append(int(f))
                results["keep"].append(mean > threshold)
    return results


def time_cpu_baseline(model, head, workload, n_pairs: int, threshold: float, warmup_pairs: int = 1) -> dict[str, Any]:
    """pairs/s of the CPU path on the first ``n_pairs`` blocks (after ``warmup_pairs`` untimed)."""
    if warmup_pairs:
        score_workload(model, head, workload, slice(0, warmup_pairs), threshold)
    t0 = time.perf_counter()
    results = score_workload(model, head, workload, slice(0, n_pairs), threshold)
    dt = time.perf_counter() - t0
    return {"pairs": n_pairs, "seconds": dt, "pairs_per_s": n_pairs / dt, "threads": torch.get_num_threads(),
            "results": results}
