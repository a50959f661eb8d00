"""One rank of the 2-GPU NCCL test (tests/test_gpu_multi.py): ``process()`` through the product API with the blocks
sharded over the ranks must return, on every rank, exactly what a single-GPU model returns.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
        tests/multi/nccl_process_worker.py <out_dir>
"""

from __future__ import annotations

import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from open_provence_b200.host_text import simple_sentence_splitter  # noqa: E402
from open_provence_b200.modeling import OpenProvenceModel  # noqa: E402


def main() -> None:
    out_dir = Path(sys.argv[1])
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ckpt = ROOT / "tests" / "golden" / "tiny_ckpt"
    golden = json.loads((ROOT / "tests" / "golden" / "process_tiny.json").read_text())
    sharded = OpenProvenceModel.from_pretrained(ckpt, device=f"cuda:{local}", data_parallel=True)
    single = OpenProvenceModel.from_pretrained(ckpt, device=f"cuda:{local}")
    report = {"rank": rank, "world": world, "cases": {}}
    for name in ("str_list", "nested", "multi_block", "multi_block_respect", "reorder_topk", "empty_context"):
        case = next(c for c in golden["cases"] if c["name"] == name)
        kwargs = dict(case["kwargs"])
        kwargs["sentence_splitter"] = simple_sentence_splitter
        for chunk in (None, 1):  # one host-preparation chunk / one context per chunk (several submit() calls)
            kw = dict(kwargs)
            if chunk:
                kw["preprocess_batch_size"] = chunk
            sharded.max_length = single.max_length = case["max_length"]
            before = sharded._scorer.collectives
            a = sharded.process(**kw)
            collectives = sharded._scorer.collectives - before
            b = single.process(**kw)
            keys = ("pruned_context", "kept_sentences", "removed_sentences", "reranking_score", "sentence_probabilities",
                    "compression_rate", "title")
            report["cases"][f"{name}/chunk={chunk}"] = {
                "identical": all(a[k] == b[k] for k in keys),
                "collectives": collectives,
                "result": {k: a[k] for k in keys},
            }
    (out_dir / f"rank{rank}.json").write_text(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
