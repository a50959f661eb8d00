#!/usr/bin/env python
"""Single-call latency of ``OpenProvenceModel.process()`` (BASELINE.json config 0 shape: one question, one short
context) on the GPU, split into host and device time.

    python tools/latency_bench.py [model] [n_calls] [option=value]
"""
import statistics, sys, time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from open_provence_b200 import synthetic as syn  # noqa: E402
from open_provence_b200.config import OpenProvenceConfig  # noqa: E402
from open_provence_b200.host_text import simple_sentence_splitter  # noqa: E402
from open_provence_b200.modeling import OpenProvenceModel  # noqa: E402

QUESTION = "What's your favorite Japanese food?"
CONTEXT = ("Work deadlines piled up today, and I kept rambling about budget spreadsheets to my roommate. "
           "Next spring I'm planning a trip to Japan so I can wander Kyoto's markets and taste every regional dish I find. "
           "Sushi is honestly my favourite, I want to grab a counter seat and let the chef serve endless nigiri until I'm smiling through soy sauce. "
           "Later I remembered to water the plants and pay the electricity bill before finally getting some sleep.")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "xsmall-30M"
    n_calls = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    option = sys.argv[3] if len(sys.argv) > 3 else ""
    if option:
        from open_provence_b200 import _native

        key, value = option.split("=")
        _native.check(_native.load().opv_set_option(key.encode(), int(value)), "opv_set_option")
    from transformers import AutoTokenizer

    ckpt = ROOT / "tests" / "golden" / "tiny_ckpt"
    tok = AutoTokenizer.from_pretrained(str(ckpt))
    config = OpenProvenceConfig.from_pretrained(ckpt)
    config.base_model_config = syn.backbone_config(name)
    config.max_length = 512
    model = OpenProvenceModel(config, syn.random_state_dict(config.base_model_config, seed=0), tok, device="cuda", dtype="bf16")
    kw = dict(question=QUESTION, context=CONTEXT, threshold=0.1, sentence_splitter=simple_sentence_splitter,
              show_progress=False)
    for _ in range(10):
        out = model.process(**kw)
    wall, stages = [], []
    for _ in range(n_calls):
        t0 = time.perf_counter()
        out = model.process(**kw)
        wall.append((time.perf_counter() - t0) * 1e3)
        stages.append(out["timing"])
    # device time of the same block alone (forward + fragment means), CUDA events on the launch stream
    n_tok = len(tok.encode(QUESTION, add_special_tokens=False)) + len(tok.encode(CONTEXT, add_special_tokens=False)) + 3
    ids = torch.randint(5, 200, (n_tok,), dtype=torch.int32, device="cuda")
    cu = torch.tensor([0, n_tok], dtype=torch.int32, device="cuda")
    eng = model.engine
    for _ in range(10):
        eng.forward_packed(ids, cu, n_tok)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        eng.forward_packed(ids, cu, n_tok)
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / 50
    med = lambda key: statistics.median(s[key] for s in stages) * 1e3
    print(f"{name} {option}: one (question, context) pair, {n_tok} tokens, {n_calls} calls: "
          f"process() p50 {statistics.median(wall):.2f} ms  p90 {sorted(wall)[int(0.9 * n_calls)]:.2f} ms | "
          f"preprocess {med('preprocess_seconds'):.2f}  assembly {med('assembly_seconds'):.2f}  "
          f"inference {med('inference_seconds'):.2f}  postprocess {med('postprocess_seconds'):.2f} ms | "
          f"forward alone (back-to-back launches, device time) {dev_ms:.3f} ms")


if __name__ == "__main__":
    main()
