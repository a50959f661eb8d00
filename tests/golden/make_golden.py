#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

It imports the unmodified /root/reference/open_provence/modeling_open_provence_standalone.py
(with a stub ``nltk`` module, which is a mandatory import there but is not installed) and the
installed ``transformers`` ModernBERT, builds a tiny ModernBERT OpenProvence checkpoint with the
reference's own constructor, and records

* ``tiny_ckpt/``            config.json + model.safetensors (reference ``save_pretrained``) + tokenizer files
* ``forward_tiny.npz``      ragged right-padded batch -> ranking/pruning logits of the reference forward
                            (fp32 as the reference runs on CPU, and fp64 via ``model.double()`` as "truth")
* ``process_tiny.json``     ``OpenProvenceModel.process`` on the input shapes the reference's tests and
                            ``scripts/hf_utils/hf_model_process_check.py:42-64`` exercise: results, the
                            unpadded block token ids the reference fed to ``forward`` and per-sentence
                            probabilities.

Nothing here is copied from the reference: it is executed, and its outputs are stored.
Library versions are written into every fixture.
"""

from __future__ import annotations

import importlib.util
import json
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REF_FILE = Path("/root/reference/open_provence/modeling_open_provence_standalone.py")

TINY_BACKBONE = {
    "model_type": "modernbert",
    "vocab_size": 264,
    "hidden_size": 128,
    "intermediate_size": 128,
    "num_hidden_layers": 4,
    "num_attention_heads": 2,
    "max_position_embeddings": 8192,
    "local_attention": 128,
    "global_attn_every_n_layers": 3,
    "norm_eps": 1e-5,
    "pad_token_id": 0,
    "bos_token_id": 1,
    "cls_token_id": 1,
    "eos_token_id": 2,
    "sep_token_id": 2,
    "classifier_pooling": "cls",
}


def load_reference_module():
    nltk = types.ModuleType("nltk")
    tok = types.ModuleType("nltk.tokenize")

    class PunktSentenceTokenizer:  # only referenced for isinstance checks in the English splitter
        pass

    tok.PunktSentenceTokenizer = PunktSentenceTokenizer
    nltk.tokenize = tok

    def _no_punkt(*_a, **_k):
        raise LookupError("punkt is not available in this container")

    nltk.data = types.SimpleNamespace(load=_no_punkt)
    sys.modules.setdefault("nltk", nltk)
    sys.modules.setdefault("nltk.tokenize", tok)
    spec = importlib.util.spec_from_file_location("ref_standalone", REF_FILE)
    module = importlib.util.module_from_spec(spec)
    sys.modules["ref_standalone"] = module
    spec.loader.exec_module(module)
    return module


def build_tiny_tokenizer(out_dir: Path):
    """Byte-level tokenizer with BERT-style specials; 5 specials + 256 byte symbols."""
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers, processors
    from transformers import PreTrainedTokenizerFast

    specials = ["[PAD]", "[CLS]", "[SEP]", "[UNK]", "[MASK]"]
    vocab = {s: i for i, s in enumerate(specials)}
    for ch in sorted(pre_tokenizers.ByteLevel.alphabet()):
        vocab[ch] = len(vocab)
    tok = Tokenizer(models.BPE(vocab=vocab, merges=[], unk_token="[UNK]"))
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)
    tok.decoder = decoders.ByteLevel()
    tok.post_processor = processors.TemplateProcessing(
        single="[CLS] $A [SEP]",
        pair="[CLS] $A [SEP] $B:1 [SEP]:1",
        special_tokens=[("[CLS]", 1), ("[SEP]", 2)],
    )
    fast = PreTrainedTokenizerFast(
        tokenizer_object=tok,
        cls_token="[CLS]",
        sep_token="[SEP]",
        pad_token="[PAD]",
        unk_token="[UNK]",
        mask_token="[MASK]",
    )
    fast.save_pretrained(str(out_dir))
    return fast


class Tokenizer457Shim:
    """Adds the two methods transformers 4.57.1 tokenizers had and 5.x dropped.

    The reference calls ``build_inputs_with_special_tokens`` / ``create_token_type_ids_from_sequences``
    (standalone:1514,2114,2149); with transformers 5.5.0 a fast tokenizer no longer has them, so the
    reference can only run here behind this shim (BERT template: [CLS] a [SEP] b [SEP]).
    """

    def __init__(self, fast):
        self._fast = fast

    def __getattr__(self, name):
        return getattr(self._fast, name)

    def __call__(self, *args, **kwargs):
        return self._fast(*args, **kwargs)

    @property
    def model_max_length(self):
        return self._fast.model_max_length

    @model_max_length.setter
    def model_max_length(self, value):
        self._fast.model_max_length = value

    def build_inputs_with_special_tokens(self, a, b=None):
        cls_id, sep_id = self._fast.cls_token_id, self._fast.sep_token_id
        if b:
            return [cls_id, *a, sep_id, *b, sep_id]
        return [cls_id, *a, sep_id]

    def create_token_type_ids_from_sequences(self, a, b=None):
        if b:
            return [0] * (len(a) + 2) + [1] * (len(b) + 1)
        return [0] * (len(a) + 2)


EN_DOC = (
    "Tokyo Tower\n"
    "Tokyo Tower is a communications and observation tower in Minato. "
    "It was completed in 1958! At 332.9 meters it is the second-tallest structure in Japan? "
    "The structure is an Eiffel Tower-inspired lattice tower.\n"
    "It is painted white and international orange to comply with air safety regulations. "
    "Over 150 million people have visited the tower."
)
EN_DOC2 = (
    "Bananas are botanically berries. They grow in clusters near the top of the plant! "
    "Almost all modern edible bananas come from two wild species? "
    "The plant is the largest herbaceous flowering plant.\n"
)
JA_DOC = "東京タワーは東京都港区にある電波塔である。高さは333メートル！1958年に完成した。観光名所として知られている？はい。"
LONG_DOC = " ".join(
    f"Sentence number {i} talks about topic {i % 7} and adds some filler words to be longer." for i in range(24)
)


def process_cases(ref):
    ss = ref.simple_sentence_splitter
    cases = [
        dict(name="str_str", max_length=512, kwargs=dict(question="How tall is Tokyo Tower?", context=EN_DOC, threshold=0.25)),
        dict(name="str_list", max_length=512, kwargs=dict(question="What are bananas?", context=[EN_DOC2, EN_DOC, JA_DOC], threshold=0.25)),
        dict(name="aligned", max_length=512, kwargs=dict(question=["How tall is Tokyo Tower?", "What are bananas?"], context=[EN_DOC, EN_DOC2], threshold=0.4)),
        dict(name="nested", max_length=512, kwargs=dict(question=["q one", "second question?"], context=[[EN_DOC, JA_DOC], [EN_DOC2]], threshold=0.1)),
        dict(
            name="presplit_sentences",
            max_length=512,
            kwargs=dict(
                question=["q one", "q two"],
                context=[[["First sentence. ", "Second one here! ", "  ", "Third?"]], [["Alpha beta. ", "Gamma delta."], EN_DOC2]],
                threshold=0.25,
            ),
        ),
        dict(name="explicit_titles", max_length=512, kwargs=dict(question="What are bananas?", context=[EN_DOC2, EN_DOC], title=["Bananas", "Tokyo Tower"], always_select_title=True, threshold=0.4)),
        dict(name="first_line_title", max_length=512, kwargs=dict(question="How tall is Tokyo Tower?", context=EN_DOC, first_line_as_title=True, title=None, always_select_title=True, threshold=0.4)),
        dict(name="title_none", max_length=512, kwargs=dict(question="How tall is Tokyo Tower?", context=[EN_DOC], title=None, threshold=0.25)),
        dict(name="multi_block", max_length=96, kwargs=dict(question="Which topic?", context=[LONG_DOC, EN_DOC], threshold=0.25)),
        dict(name="multi_block_respect", max_length=96, kwargs=dict(question="Which topic?", context=[LONG_DOC], respect_sentence_boundaries=True, threshold=0.25)),
        dict(name="overlong_sentence", max_length=48, kwargs=dict(question="Which topic?", context="word " * 120 + ". tail sentence.", threshold=0.25)),
        dict(name="strip_sentences", max_length=512, kwargs=dict(question="What are bananas?", context=[EN_DOC2 + "   \n  ", EN_DOC], strip_sentences=True, threshold=0.25)),
        dict(name="reorder_topk", max_length=512, kwargs=dict(question="What are bananas?", context=[EN_DOC2, EN_DOC, JA_DOC, LONG_DOC[:300]], reorder=True, top_k=2, threshold=0.25)),
        dict(name="no_best_score", max_length=96, kwargs=dict(question="Which topic?", context=[LONG_DOC], use_best_reranker_score=False, zero_score_when_empty=False, threshold=0.9)),
        dict(name="japanese", max_length=512, kwargs=dict(question="東京タワーの高さは？", context=JA_DOC, threshold=0.25)),
        dict(name="empty_context", max_length=512, kwargs=dict(question="q", context=["", EN_DOC2], threshold=0.25)),
        dict(name="batch_size_1", max_length=96, kwargs=dict(question=["Which topic?", "What are bananas?"], context=[LONG_DOC, EN_DOC2], batch_size=1, threshold=0.25)),
    ]
    for case in cases:
        case["kwargs"].setdefault("sentence_splitter", ss)
        case["kwargs"]["return_sentence_metrics"] = True
        case["kwargs"]["return_sentence_texts"] = True
        case["kwargs"]["show_progress"] = False
    return cases


def main() -> None:
    import safetensors.torch
    import transformers

    ref = load_reference_module()
    ckpt_dir = HERE / "tiny_ckpt"
    ckpt_dir.mkdir(parents=True, exist_ok=True)
    fast = build_tiny_tokenizer(ckpt_dir)
    ref.AutoTokenizer.from_pretrained = staticmethod(lambda *_a, **_k: Tokenizer457Shim(fast))

    torch.manual_seed(0)
    config = ref.OpenProvenceConfig(
        base_model_config=dict(TINY_BACKBONE),
        tokenizer_name_or_path="tiny_ckpt",
        pruning_config={"hidden_size": TINY_BACKBONE["hidden_size"], "num_labels": 2},
        max_length=512,
        default_threadshold=0.1,
    )
    model = ref.OpenProvenceModel(config)
    with torch.no_grad():
        # Spread keep-probabilities so thresholds actually cut (default init gives p ~ 0.5 everywhere).
        gen = torch.Generator().manual_seed(7)
        model.pruning_head.classifier.weight.copy_(torch.randn(2, TINY_BACKBONE["hidden_size"], generator=gen) * 0.3)
        model.pruning_head.classifier.bias.copy_(torch.tensor([0.0, -3.0]))
        model.ranking_model.classifier.bias.copy_(torch.tensor([0.25]))
        # Non-trivial LayerNorm gains (HF init leaves them at 1.0, which would hide gain bugs).
        for name, param in model.named_parameters():
            if name.endswith("norm.weight"):
                param.copy_(1.0 + 0.1 * torch.randn(param.shape, generator=gen))
    model.eval()

    state = {k: v.detach().clone().contiguous() for k, v in model.state_dict().items()}
    safetensors.torch.save_file(state, str(ckpt_dir / "model.safetensors"), metadata={"format": "pt"})
    cfg_dict = config.to_dict()
    cfg_dict["architectures"] = ["OpenProvenceForSequenceClassification"]
    (ckpt_dir / "config.json").write_text(json.dumps(cfg_dict, indent=2, sort_keys=True, default=str))

    versions = {
        "torch": torch.__version__,
        "transformers": transformers.__version__,
        "numpy": np.__version__,
        "reference_file": str(REF_FILE),
    }

    # ---- forward fixture: ragged right-padded batch through the reference forward ----
    rng = np.random.default_rng(1234)
    lengths = [1, 2, 7, 64, 65, 129, 130, 200, 257, 333]
    S = max(lengths)
    ids = np.zeros((len(lengths), S), dtype=np.int64)
    mask = np.zeros((len(lengths), S), dtype=np.int64)
    for b, n in enumerate(lengths):
        row = rng.integers(5, TINY_BACKBONE["vocab_size"], size=n)
        row[0] = 1
        if n > 3:
            row[min(n - 2, 5)] = 2
            row[-1] = 2
        ids[b, :n] = row
        mask[b, :n] = 1
    with torch.inference_mode():
        out32 = model.forward(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask), return_dict=True)
        rank32 = out32.ranking_logits.float().numpy()
        prune32 = out32.pruning_logits.float().numpy()
        model.double()
        out64 = model.forward(input_ids=torch.from_numpy(ids), attention_mask=torch.from_numpy(mask), return_dict=True)
        rank64 = out64.ranking_logits.numpy()
        prune64 = out64.pruning_logits.numpy()
        model.float()
    np.savez_compressed(
        HERE / "forward_tiny.npz",
        input_ids=ids,
        attention_mask=mask,
        lengths=np.asarray(lengths),
        ranking_logits_f32=rank32,
        pruning_logits_f32=prune32,
        ranking_logits_f64=rank64,
        pruning_logits_f64=prune64,
        versions=json.dumps(versions),
    )
    print("forward fixture: fp32-vs-fp64 max abs", np.abs(rank32 - rank64).max(), np.abs((prune32 - prune64) * mask[..., None]).max())

    # ---- process fixtures ----
    recorded: list[dict] = []
    orig_forward = ref.OpenProvenceModel.forward

    def recording_forward(self, input_ids=None, attention_mask=None, **kw):
        out = orig_forward(self, input_ids=input_ids, attention_mask=attention_mask, **kw)
        lens = attention_mask.sum(dim=1).tolist()
        for row, n, rl, pl in zip(input_ids.tolist(), lens, out.ranking_logits.float().tolist(), out.pruning_logits.float()):
            recorded.append({"ids": row[: int(n)], "rank_logits": rl, "prune_logits": pl[: int(n)].tolist()})
        return out

    ref.OpenProvenceModel.forward = recording_forward
    results = []
    for case in process_cases(ref):
        recorded.clear()
        model.max_length = case["max_length"]
        with torch.inference_mode():
            res = model.process(**case["kwargs"])
        kwargs = {k: v for k, v in case["kwargs"].items() if k != "sentence_splitter"}
        probs = res.get("sentence_probabilities")
        results.append(
            {
                "name": case["name"],
                "max_length": case["max_length"],
                "kwargs": kwargs,
                "result": {k: v for k, v in res.items() if k not in ("timing", "performance_trace")},
                "blocks": [dict(r) for r in recorded],
            }
        )
        print(case["name"], "blocks:", len(recorded), "score:", res["reranking_score"] if not isinstance(res["reranking_score"], list) else "...")
        del probs
    ref.OpenProvenceModel.forward = orig_forward
    (HERE / "process_tiny.json").write_text(json.dumps({"versions": versions, "cases": results}, ensure_ascii=False))
    total = sum(os.path.getsize(p) for p in HERE.rglob("*") if p.is_file())
    print("fixtures written, total bytes:", total)


if __name__ == "__main__":
    main()
