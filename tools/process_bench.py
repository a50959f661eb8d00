#!/usr/bin/env python
"""End-to-end ``OpenProvenceModel.process()`` on the GPU: text in, pruned text out (base-130M dims, random
weights, the tiny golden tokenizer).  Prints the stage timings the reference reports (result["timing"]).

    python tools/process_bench.py [n_contexts] [max_length] [model] [native|python]
"""
import sys, time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from open_provence_b200 import synthetic as syn  # noqa: E402
from open_provence_b200.config import OpenProvenceConfig  # noqa: E402
from open_provence_b200.host_text import simple_sentence_splitter  # noqa: E402
from open_provence_b200.modeling import OpenProvenceModel  # noqa: E402


def main():
    n_ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    max_length = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    name = sys.argv[3] if len(sys.argv) > 3 else "base-130M"
    from transformers import AutoTokenizer

    ckpt = ROOT / "tests" / "golden" / "tiny_ckpt"
    tok = AutoTokenizer.from_pretrained(str(ckpt))
    config = OpenProvenceConfig.from_pretrained(ckpt)
    config.base_model_config = syn.backbone_config(name)
    config.max_length = max_length
    model = OpenProvenceModel(config, syn.random_state_dict(config.base_model_config, seed=0), tok, device="cuda", dtype="bf16")
    model.max_length = max_length
    model.host_pack_mode = sys.argv[4] if len(sys.argv) > 4 else "native"  # block assembly: C++ (default) or Python
    rng = np.random.default_rng(0)
    words = ["alpha", "beta", "gamma", "delta", "pruning", "context", "question", "answer", "tokyo", "river"]
    def sentence():
        return " ".join(rng.choice(words, size=int(rng.integers(6, 16)))) + ". "
    contexts = ["".join(sentence() for _ in range(int(rng.integers(30, 60)))) for _ in range(n_ctx)]
    questions = ["what is " + " ".join(rng.choice(words, size=4)) + "?" for _ in range(n_ctx)]
    kw = dict(question=questions, context=contexts, threshold=0.1, sentence_splitter=simple_sentence_splitter,
              show_progress=False, batch_size=64)
    model.process(**kw)
    t0 = time.perf_counter()
    out = model.process(**kw)
    dt = time.perf_counter() - t0
    print(f"{name} [{model.host_pack_mode} pack]: {n_ctx} contexts, max_length {max_length}: {dt * 1e3:.1f} ms -> {n_ctx / dt:.0f} contexts/s end to end")
    print({k: round(v, 4) for k, v in out["timing"].items()})
    kept = sum(len(p) for p in out["pruned_context"])
    print(f"pruned characters kept: {kept} of {sum(len(c) for c in contexts)}")


if __name__ == "__main__":
    main()
