#!/usr/bin/env python
"""CPU baseline cross-check (build container only: needs /root/reference): the UNMODIFIED reference's own inference
loop -- ``OpenProvenceModel.process()`` -> ``_run_inference_batches`` (standalone:2761-2939: right-padding, per-row
tensor fills, ``forward``, ``.detach().cpu()``, torch softmax / sigmoid per block) -- timed beside the "port" loop that
``bench.py --impl reference`` and the ``cpu_baseline`` leg run on the GPU box (oracle/hf_cpu_baseline.py), on the SAME
token blocks with the SAME weights (base-130M dims, random init, 2048-token blocks).

    python tools/reference_loop_compare.py [n_blocks] [seq_len]  ->  profiles/r2_reference_loop_vs_port.md
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

import make_golden as mg  # noqa: E402  (reference loader with the nltk stub, tokenizer shim)
from open_provence_b200 import synthetic as syn  # noqa: E402
from oracle import hf_cpu_baseline as hb  # noqa: E402

n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
torch.set_num_threads(max(1, len(__import__("os").sched_getaffinity(0))))

ref = mg.load_reference_module()
from transformers import AutoTokenizer  # noqa: E402

ckpt = ROOT / "tests" / "golden" / "tiny_ckpt"
fast = AutoTokenizer.from_pretrained(str(ckpt))
ref.AutoTokenizer.from_pretrained = staticmethod(lambda *_a, **_k: mg.Tokenizer457Shim(fast))
cfg = syn.backbone_config("base-130M")
sd = syn.random_state_dict(cfg, seed=0)
tiny_cfg = json.loads((ckpt / "config.json").read_text())
model = ref.OpenProvenceModel(ref.OpenProvenceConfig(
    base_model_config=cfg, tokenizer_name_or_path="tiny_ckpt", pruning_config={**tiny_cfg["pruning_config"], "hidden_size": cfg["hidden_size"]}, max_length=S,
    default_threadshold=0.1))
missing, unexpected = model.load_state_dict(sd, strict=False)
assert not unexpected and all("inv_freq" in m for m in missing), (missing, unexpected)
model.eval()
model.max_length = S

# one context per block: byte-level tokenizer, one token per character -> [CLS] q [SEP] ctx [SEP] = S tokens
rng = np.random.default_rng(5)
question = "what is the topic of this long passage?"
words = ["alpha", "beta", "gamma", "tower", "river", "banana", "pruning", "context"]
contexts = []
for _ in range(n_blocks):
    text = ""
    while len(text) < S - len(question) - 3 - 60:
        text += " ".join(rng.choice(words) for _ in range(int(rng.integers(4, 12)))) + ". "
    contexts.append(text[: S - len(question) - 3])

seen = []
orig_forward = ref.OpenProvenceModel.forward


def spy(self, *a, **kw):
    seen.append((kw["input_ids"].clone(), kw["attention_mask"].clone()))
    return orig_forward(self, *a, **kw)


ref.OpenProvenceModel.forward = spy
kw = dict(question=question, context=contexts, threshold=0.1, batch_size=32, sentence_splitter=ref.simple_sentence_splitter,
          show_progress=False, return_sentence_metrics=True)
model.process(**kw)  # warm-up (threads, allocator)
seen.clear()
t0 = time.perf_counter()
out = model.process(**kw)
t_process = time.perf_counter() - t0
timing = out["timing"]
blocks = [ids[b, : int(m[b].sum())].tolist() for ids, m in seen for b in range(ids.shape[0])]
lengths = [len(b) for b in blocks]

# the port loop on the very same blocks
hf_model, head = hb.build_hf_model(cfg, sd)
wl = {"ids": np.concatenate([np.asarray(b, dtype=np.int32) for b in blocks]),
      "cu_seqlens": np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32),
      "frag_block": np.arange(len(blocks), dtype=np.int32),
      "frag_ranges": np.stack([np.cumsum([0] + lengths[:-1]) + 50, np.cumsum(lengths) - 1], axis=1).astype(np.int32)}
hb.score_workload(hf_model, head, wl, slice(0, 1), 0.1)
t0 = time.perf_counter()
res = hb.score_workload(hf_model, head, wl, slice(0, len(blocks)), 0.1)
t_port = time.perf_counter() - t0

ref_scores = out["reranking_score"]
port_scores = res["rank_score"]
lines = [
    "# Reference CPU loop vs the timed 'port' loop (build container, CPU)",
    "",
    f"base-130M dims, random init, {len(blocks)} blocks of {lengths} tokens, {torch.get_num_threads()} threads, fp32.",
    "",
    "| arm | seconds | blocks/s |",
    "|---|---|---|",
    f"| reference `process()` total (tokenise + assemble + `_run_inference_batches` + postprocess) | {t_process:.3f} | {len(blocks) / t_process:.3f} |",
    f"| reference `inference_seconds` (the `forward` calls inside `_run_inference_batches`, standalone:2885-2891) | {timing['inference_seconds']:.3f} | {len(blocks) / timing['inference_seconds']:.3f} |",
    f"| port: `oracle/hf_cpu_baseline.score_workload` on the same blocks (what `bench.py --impl reference` runs) | {t_port:.3f} | {len(blocks) / t_port:.3f} |",
    "",
    f"port / reference-process time ratio: {t_port / t_process:.3f}; rerank scores agree to "
    f"{max(abs(float(a) - float(b)) for a, b in zip(ref_scores, port_scores)):.2e}.",
    "",
    "The GPU box has no /root/reference, so the timed CPU arm there is the port; this file pins how close the two are.",
]
(ROOT / "profiles" / "r2_reference_loop_vs_port.md").write_text("\n".join(lines) + "\n")
print("\n".join(lines))
