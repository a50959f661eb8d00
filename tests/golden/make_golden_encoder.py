#!/usr/bin/env python
"""Golden fixtures for the ``encoder.py`` secondary APIs, produced by running THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden_encoder.py

Imports the unmodified ``/root/reference/open_provence/encoder.py`` (through a stub ``open_provence`` package
object, so that the reference's ``__init__`` -- which pulls in the training stack -- is not executed, and with
the ``nltk`` stub of ``make_golden.py``), assembles an ``OpenProvenceEncoder`` around the tiny golden
checkpoint exactly as the reference's ``from_pretrained`` does after loading (encoder.py:1168-1201), and
records ``predict`` / ``predict_with_pruning`` / ``predict_context`` / ``prune`` / ``prune_texts`` outputs together
with the unpadded token ids and the reference forward's logits for every pair.  Nothing is copied from the
reference: it is executed, and its outputs are stored (-> ``encoder_tiny.json``).
"""

from __future__ import annotations

import importlib
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden as mg  # noqa: E402

CKPT = HERE / "tiny_ckpt"

PAIRS = [
    ("How tall is Tokyo Tower?", mg.EN_DOC),
    ("What are bananas?", mg.EN_DOC2),
    ("東京タワーの高さは？", mg.JA_DOC),
    ("Which topic?", mg.LONG_DOC),  # longer than max_length: truncated by the tokenizer
    ("q", "x"),
]


def build_reference_encoder():
    ref = mg.load_reference_module()  # installs the nltk stub, loads the standalone module
    pkg = types.ModuleType("open_provence")
    pkg.__path__ = ["/root/reference/open_provence"]
    sys.modules["open_provence"] = pkg
    enc_mod = importlib.import_module("open_provence.encoder")
    head_mod = importlib.import_module("open_provence.models.open_provence_head")
    from safetensors.torch import load_file
    from transformers import AutoModelForSequenceClassification, AutoTokenizer, ModernBertConfig

    cfg = json.loads((CKPT / "config.json").read_text())
    sd = load_file(str(CKPT / "model.safetensors"))
    backbone_cfg = ModernBertConfig(**{k: v for k, v in cfg["base_model_config"].items() if k != "model_type"}, num_labels=1)
    ranking = AutoModelForSequenceClassification.from_config(backbone_cfg)  # what standalone:1341 does
    ranking.load_state_dict({k[len("ranking_model."):]: v for k, v in sd.items() if k.startswith("ranking_model.")}, strict=True)
    # transformers 5.x shim: the reference head's constructor calls ``self.init_weights()`` (a 4.57 idiom that
    # needs ``post_init`` bookkeeping in 5.5), so the module is assembled with the same attributes its
    # ``__init__`` sets (open_provence_head.py:61-77) and its weights come from the checkpoint anyway.
    head_cfg = head_mod.OpenProvenceHeadConfig(**cfg["pruning_config"])
    head = head_mod.OpenProvenceHead.__new__(head_mod.OpenProvenceHead)
    from transformers.modeling_utils import PreTrainedModel
    PreTrainedModel.__init__(head, head_cfg)
    head.num_labels = head_cfg.num_labels
    head.sentence_pooling = head_cfg.sentence_pooling
    head.use_weighted_pooling = head_cfg.use_weighted_pooling
    head.dropout = torch.nn.Dropout(head_cfg.classifier_dropout)
    head.classifier = torch.nn.Linear(head_cfg.hidden_size, head_cfg.num_labels)
    head.load_state_dict({k[len("pruning_head."):]: v for k, v in sd.items() if k.startswith("pruning_head.")})
    enc = enc_mod.OpenProvenceEncoder.__new__(enc_mod.OpenProvenceEncoder)  # encoder.py:1168-1201
    torch.nn.Module.__init__(enc)
    enc.model_name_or_path = str(CKPT)
    enc.mode = "reranking_pruning"
    enc.num_labels = 1
    enc.max_length = 192
    enc.device = "cpu"
    enc.cache_dir = None
    enc.config = backbone_cfg
    enc.use_raw_logits = True
    enc.text_chunker = None
    enc._original_num_labels = 1
    enc._keys_to_ignore_on_save = []
    enc.ranking_model = ranking
    enc.pruning_head = head
    enc.tokenizer = AutoTokenizer.from_pretrained(str(CKPT))
    enc.to("cpu")
    enc.eval()
    return enc, ref


def simple_chunks(text: str):
    """Character spans of the sentences of ``text`` (split on . ! ? 。 ！ ？ and newlines)."""
    spans, start = [], 0
    for i, ch in enumerate(text):
        if ch in ".!?。！？\n":
            if i + 1 > start:
                spans.append((start, i + 1))
            start = i + 1
    if start < len(text):
        spans.append((start, len(text)))
    return spans


def main():
    torch.manual_seed(0)
    enc, _ = build_reference_encoder()
    tok = enc.tokenizer
    out = {"versions": {"torch": torch.__version__, "transformers": __import__("transformers").__version__},
           "max_length": enc.max_length, "pairs": [list(p) for p in PAIRS], "examples": []}
    # the reference tokenises a batch with padding; logits on valid tokens do not depend on the padding
    # (tests/test_oracle_golden.py), so the per-example records below hold the unpadded rows
    with torch.no_grad():
        for q, d in PAIRS:
            encd = tok([(q, d)], padding=True, truncation=True, max_length=enc.max_length, return_tensors="pt")
            fw = enc.forward(input_ids=encd["input_ids"], attention_mask=encd["attention_mask"])
            out["examples"].append({
                "ids": encd["input_ids"][0].tolist(),
                "rank_logits": fw["ranking_logits"][0].double().tolist(),
                "prune_logits": fw["pruning_logits"][0].double().tolist(),
            })
    scores = enc.predict([tuple(p) for p in PAIRS], batch_size=2)
    out["predict"] = np.asarray(scores, dtype=np.float64).tolist()
    out["predict_single"] = np.asarray(enc.predict(PAIRS[0]), dtype=np.float64).tolist()
    pw = []
    for thr in (0.5, 0.2):
        res = enc.predict_with_pruning([tuple(p) for p in PAIRS], batch_size=2, pruning_threshold=thr, return_documents=True)
        pw.append({"threshold": thr, "outputs": [
            {"ranking_scores": np.asarray(r.ranking_scores, dtype=np.float64).tolist(),
             "pruning_masks": np.asarray(r.pruning_masks).astype(int).tolist(),
             "tokens": r.sentences, "compression_ratio": float(r.compression_ratio),
             "num_pruned_sentences": int(r.num_pruned_sentences), "pruned_documents": r.pruned_documents}
            for r in res]})
    out["predict_with_pruning"] = pw
    single = enc.predict_with_pruning(PAIRS[1], pruning_threshold=0.5, return_documents=True)
    out["predict_with_pruning_single"] = {"pruned_documents": single.pruned_documents,
                                          "compression_ratio": float(single.compression_ratio)}
    chunks = [simple_chunks(d) for _, d in PAIRS]
    pc = enc.predict_context([tuple(p) for p in PAIRS], chunks, batch_size=3, token_threshold=0.5, chunk_threshold=0.5)
    out["predict_context"] = {"chunks": [[list(c) for c in ch] for ch in chunks], "outputs": [
        {"ranking_scores": float(r.ranking_scores), "chunk_predictions": np.asarray(r.chunk_predictions).astype(int).tolist(),
         "chunk_scores": np.asarray(r.chunk_scores, dtype=np.float64).tolist(),
         "token_scores": np.asarray(r.token_scores, dtype=np.float64).tolist(),
         "compression_ratio": float(r.compression_ratio)} for r in pc]}
    out["prune"] = {"plain": enc.prune(*PAIRS[0], threshold=0.5),
                    "detail": {k: (float(v) if isinstance(v, (float, np.floating)) else v)
                               for k, v in enc.prune(*PAIRS[0], threshold=0.5, return_sentences=True).items()}}
    out["prune_texts"] = [{"pruned_text": r["pruned_text"], "kept_ratio": float(r["kept_ratio"])}
                          for r in enc.prune_texts([p[0] for p in PAIRS], [p[1] for p in PAIRS], threshold=0.5, batch_size=2)]
    (HERE / "encoder_tiny.json").write_text(json.dumps(out, ensure_ascii=False))
    print(f"wrote encoder_tiny.json: {len(out['examples'])} examples; predict = {out['predict']}")
    print("keep fractions @0.5:", [1 - o["compression_ratio"] for o in pw[0]["outputs"]])


if __name__ == "__main__":
    main()
