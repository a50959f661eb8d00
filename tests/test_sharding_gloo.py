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
    """score_blocks / prune with the reference's recorded logits, torch-CPU tensors (gloo-compatible).

    ``guard`` > 0 marks sentences within that distance of the threshold as ``near`` and leaves them to the
    exact-mean exchange, like DeviceScorer.prune(reevaluate=False); the "device" means are perturbed by
    ``noise`` so that only the exchanged exact values can reproduce the reference."""

    def __init__(self, recorded_blocks, guard: float = 0.0, noise: float = 0.0):
        from oracle import postprocess_numpy as opp

        self.opp = opp
        self.by_ids = {tuple(b["ids"]): b for b in recorded_blocks}
        self.max_tokens = 1 << 20
        self.scored: list[int] = []
        self.guard, self.noise = guard, noise

    def _probs(self, table, b):
        rec = self.by_ids[tuple(int(t) for t in table.block_ids[b])]
        return rec, self.opp.keep_probs_from_logits(np.asarray(rec["prune_logits"], dtype=np.float32))

    def score_blocks(self, table, blocks, host_scores=True):
        rank_score = torch.zeros(max(table.n_blocks, 1), dtype=torch.float32)
        frag_mean = torch.zeros(max(len(table.frag_block), 1), dtype=torch.float32)
        kept = []
        for b in blocks:
            b = int(b)
            self.scored.append(b)
            rec, probs = self._probs(table, b)
            rank_score[b] = self.opp.ranking_score_from_logits(np.asarray(rec["rank_logits"], dtype=np.float32))
            for slot, (blk, (s, e)) in enumerate(zip(table.frag_block, table.frag_local)):
                if blk == b:
                    exact = 1.0 if e <= s else float(probs[s:e].mean())
                    frag_mean[slot] = exact + self.noise  # what a device reduction in another order could return
                    kept.append((slot, exact))
        rank_score = rank_score[: table.n_blocks]
        return (rank_score.numpy() if host_scores else rank_score), frag_mean, kept

    def prune(self, table, rank_score, frag_mean, kept, threshold, reevaluate=True):
        fm = frag_mean.numpy().copy()
        prob, keep, near = [], [], []
        for s in range(table.n_sentences):
            members = list(table.sent_frag_index[table.sent_offsets[s] : table.sent_offsets[s + 1]])
            p = max(0.0, min(float(np.mean([float(fm[k]) for k in members])) if members else 0.0, 1.0))
            prob.append(p)
            keep.append(p > threshold)
            near.append(abs(p - threshold) <= self.guard)
        out = {"rank_score": rank_score, "frag_mean": fm, "sent_prob": np.asarray(prob), "keep": np.asarray(keep, dtype=bool),
               "near": np.asarray(near, dtype=bool)}
        if reevaluate and out["near"].any():
            self.apply_exact(out, table, dict(kept), threshold)
        return out

    def exact_slot_means(self, needed, kept):
        return {slot: exact for slot, exact in kept if slot in needed}

    def apply_exact(self, out, table, slot_mean, threshold):
        for s in np.nonzero(out["near"])[0]:
            members = list(table.sent_frag_index[table.sent_offsets[s] : table.sent_offsets[s + 1]])
            assert all(int(k) in slot_mean for k in members), "exact means of a guard-band sentence did not arrive"
            p = max(0.0, min(float(np.mean([slot_mean[int(k)] for k in members])), 1.0))
            out["sent_prob"][s] = p
            out["keep"][s] = p > threshold


def _worker(rank: int, world: int, port: int, case_name: str, out_dir: str, chunk: int = 0, guard: float = 0.0,
            noise: float = 0.0):
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
    inner = _CpuRecordedScorer(case["blocks"], guard=guard, noise=noise)
    model = OpenProvenceModel(OpenProvenceConfig.from_pretrained(ckpt), None, AutoTokenizer.from_pretrained(str(ckpt)),
                              scorer=ShardedScorer(inner, hidden=128, inter=128))
    model.max_length = case["max_length"]
    kwargs = dict(case["kwargs"])
    kwargs["sentence_splitter"] = simple_sentence_splitter
    if chunk:
        kwargs["preprocess_batch_size"] = chunk
    res = model.process(**kwargs)
    payload = {k: res[k] for k in ("pruned_context", "reranking_score", "kept_sentences", "sentence_probabilities")}
    payload["scored_blocks"] = sorted(inner.scored)
    payload["collectives"] = model._scorer.collectives
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


def _spawn(tmp_path, case_name, **kw):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")
    mp.spawn(_worker, args=(2, port, case_name, str(tmp_path), kw.get("chunk", 0), kw.get("guard", 0.0), kw.get("noise", 0.0)),
             nprocs=2, join=True)
    return [json.loads((tmp_path / f"rank{r}.json").read_text()) for r in range(2)]


def test_chunked_process_issues_one_collective(tmp_path, process_golden):
    """One context per host-preparation chunk: every chunk is scored locally (submit), and the whole call still ends
    with ONE all-gather (the gather used to run once per chunk)."""
    case = next(c for c in process_golden["cases"] if c["name"] == "str_list")
    r0, r1 = _spawn(tmp_path, "str_list", chunk=1)
    assert r0["collectives"] == r1["collectives"] == 1
    assert len(case["kwargs"]["context"]) > 1  # several chunks did go through submit()
    for key in ("pruned_context", "kept_sentences"):
        assert r0[key] == r1[key] == case["result"][key]


def test_guard_band_sentences_resolved_identically_on_all_ranks(tmp_path, process_golden):
    """Every sentence is declared 'near' (guard = 1) and the local 'device' means carry an error: only the exact means
    from the owning ranks, shared by the all-reduce, reproduce the reference -- on BOTH ranks."""
    case = next(c for c in process_golden["cases"] if c["name"] == "multi_block")
    r0, r1 = _spawn(tmp_path, "multi_block", guard=1.0, noise=3e-3)
    assert r0["collectives"] == r1["collectives"] == 2  # the gather + the exact-mean exchange
    assert r0["sentence_probabilities"] == r1["sentence_probabilities"]
    flat = lambda x: [v for y in x for v in (flat(y) if isinstance(y, list) else [y])]  # noqa: E731
    for a, b in zip(flat(r0["sentence_probabilities"]), flat(case["result"]["sentence_probabilities"])):
        assert abs(a - b) < 1e-6
    assert r0["kept_sentences"] == r1["kept_sentences"] == case["result"]["kept_sentences"]
