"""N > 1 path on CPU: world_size-2 ``gloo`` run of the shard + single all-gather.

The device stage is a recorded-logit scorer (see tests/test_process_host.py); what is under test is the
deterministic LPT assignment, the fixed-width all-gather records and that every rank ends with the same,
reference-identical ``process()`` result.
"""

from __future__ import annotations

import json
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent

from open_provence_b200.sharding import block_cost, lpt_assign  # noqa: E402


def test_lpt_is_deterministic_and_balanced():
    rng = np.random.default_rng(0)
    lengths = rng.integers(64, 4096, size=200)
    costs = [block_cost(int(n)) for n in lengths]
    a = lpt_assign(costs, 8)
    b = lpt_assign(costs, 8)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert sorted(int(i) for s in a for i in s) == list(range(200))
    loads = [sum(costs[i] for i in s) for s in a]
    assert max(loads) / (sum(loads) / 8) < 1.05
    assert [len(s) for s in lpt_assign([1.0, 2.0], 4)] == [1, 1, 0, 0]


class _CpuRecordedScorer:
    """score_blocks / prune with the reference's recorded logits, torch-CPU tensors (gloo-compatible)."""

    def __init__(self, recorded_blocks):
        from oracle import postprocess_numpy as opp

        self.opp = opp
        self.by_ids = {tuple(b["ids"]): b for b in recorded_blocks}
        self.max_tokens = 1 << 20
        self.scored: list[int] = []

    def score_blocks(self, table, blocks):
        rank_score = np.zeros(table.n_blocks, dtype=np.float32)
        frag_mean = torch.zeros(max(len(table.frag_block), 1), dtype=torch.float32)
        for b in blocks:
            b = int(b)
            self.scored.append(b)
            rec = self.by_ids[tuple(int(t) for t in table.block_ids[b])]
            rank_score[b] = self.opp.ranking_score_from_logits(np.asarray(rec["rank_logits"], dtype=np.float32))
            probs = self.opp.keep_probs_from_logits(np.asarray(rec["prune_logits"], dtype=np.float32))
            for slot, (blk, (s, e)) in enumerate(zip(table.frag_block, table.frag_local)):
                if blk == b:
                    frag_mean[slot] = 1.0 if e <= s else float(probs[s:e].mean())
        return rank_score, frag_mean, []

    def prune(self, table, rank_score, frag_mean, kept, threshold):
        fm = frag_mean.numpy()
        prob, keep = [], []
        for s in range(table.n_sentences):
            members = list(table.sent_frag_index[table.sent_offsets[s] : table.sent_offsets[s + 1]])
            p = max(0.0, min(float(np.mean([float(fm[k]) for k in members])) if members else 0.0, 1.0))
            prob.append(p)
            keep.append(p > threshold)
        return {"rank_score": rank_score, "sent_prob": np.asarray(prob), "keep": np.asarray(keep, dtype=bool)}


def _worker(rank: int, world: int, port: int, case_name: str, out_dir: str):
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist
    from transformers import AutoTokenizer

    from open_provence_b200.config import OpenProvenceConfig
    from open_provence_b200.host_text import simple_sentence_splitter
    from open_provence_b200.modeling import OpenProvenceModel
    from open_provence_b200.sharding import ShardedScorer

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    golden = json.loads((ROOT / "tests" / "golden" / "process_tiny.json").read_text())
    case = next(c for c in golden["cases"] if c["name"] == case_name)
    ckpt = ROOT / "tests" / "golden" / "tiny_ckpt"
    inner = _CpuRecordedScorer(case["blocks"])
    model = OpenProvenceModel(OpenProvenceConfig.from_pretrained(ckpt), None, AutoTokenizer.from_pretrained(str(ckpt)),
                              scorer=ShardedScorer(inner, hidden=128, inter=128))
    model.max_length = case["max_length"]
    kwargs = dict(case["kwargs"])
    kwargs["sentence_splitter"] = simple_sentence_splitter
    res = model.process(**kwargs)
    payload = {k: res[k] for k in ("pruned_context", "reranking_score", "kept_sentences", "sentence_probabilities")}
    payload["scored_blocks"] = sorted(inner.scored)
    Path(out_dir, f"rank{rank}.json").write_text(json.dumps(payload))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case_name", ["multi_block", "nested"])
def test_two_rank_gloo_process_matches_reference(case_name, tmp_path, process_golden):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")
    mp.spawn(_worker, args=(2, port, case_name, str(tmp_path)), nprocs=2, join=True)
    case = next(c for c in process_golden["cases"] if c["name"] == case_name)
    r0 = json.loads((tmp_path / "rank0.json").read_text())
    r1 = json.loads((tmp_path / "rank1.json").read_text())
    # each block scored on exactly one rank; both ranks hold the identical, reference-identical result
    assert sorted(r0["scored_blocks"] + r1["scored_blocks"]) == list(range(len(case["blocks"])))
    assert r0["scored_blocks"] and r1["scored_blocks"]
    for key in ("pruned_context", "kept_sentences"):
        assert r0[key] == r1[key] == case["result"][key]
    flat = lambda x: [v for y in x for v in (flat(y) if isinstance(y, list) else [y])]  # noqa: E731
    for a, b in zip(flat(r0["sentence_probabilities"]), flat(case["result"]["sentence_probabilities"])):
        assert abs(a - b) < 1e-6
    for a, b in zip(flat(r0["reranking_score"]), flat(case["result"]["reranking_score"])):
        assert abs(a - b) < 1e-6
