#!/usr/bin/env python
"""Host-side cost of process() (no GPU needed): the device stage is replaced by a scorer that returns
constant probabilities, so what is timed is input normalisation, sentence split, tokenisation, fragmentising,
block assembly (-> BlockTable) and the string post-processing.

    python tools/process_host_profile.py [n_contexts] [max_length] [--profile]
"""
import cProfile, pstats, sys, time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from open_provence_b200.config import OpenProvenceConfig  # noqa: E402
from open_provence_b200.host_text import simple_sentence_splitter  # noqa: E402
from open_provence_b200.modeling import OpenProvenceModel  # noqa: E402


class ConstantScorer:
    def run(self, table, threshold):
        n = table.n_sentences
        prob = np.linspace(0.0, 1.0, n) if n else np.zeros(0)
        return {"rank_score": np.full(table.n_blocks, 0.5, np.float32), "sent_prob": prob, "keep": prob > threshold}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_ctx = int(args[0]) if args else 256
    max_length = int(args[1]) if len(args) > 1 else 2048
    from transformers import AutoTokenizer

    ckpt = ROOT / "tests" / "golden" / "tiny_ckpt"
    tok = AutoTokenizer.from_pretrained(str(ckpt))
    model = OpenProvenceModel(OpenProvenceConfig.from_pretrained(ckpt), None, tok, scorer=ConstantScorer())
    model.max_length = max_length
    rng = np.random.default_rng(0)
    words = ["alpha", "beta", "gamma", "delta", "pruning", "context", "question", "answer", "tokyo", "river"]
    def sentence():
        return " ".join(rng.choice(words, size=int(rng.integers(6, 16)))) + ". "
    contexts = ["".join(sentence() for _ in range(int(rng.integers(30, 60)))) for _ in range(n_ctx)]
    questions = ["what is " + " ".join(rng.choice(words, size=4)) + "?" for _ in range(n_ctx)]
    kw = dict(question=questions, context=contexts, threshold=0.1, sentence_splitter=simple_sentence_splitter,
              show_progress=False)
    model.process(**kw)  # warm-up
    t0 = time.perf_counter()
    out = model.process(**kw)
    dt = time.perf_counter() - t0
    n_blocks = len(model._last_table.block_ids) if hasattr(model, "_last_table") else None
    print(f"{n_ctx} contexts, max_length {max_length}: {dt * 1e3:.1f} ms total -> {n_ctx / dt:.0f} contexts/s (host only)")
    print({k: round(v, 4) for k, v in out["timing"].items()})
    if "--profile" in sys.argv:
        pr = cProfile.Profile(); pr.enable(); model.process(**kw); pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(18)


if __name__ == "__main__":
    main()
