"""Multi-GPU data parallelism for ``process()``: blocks are independent, so they are dealt to the ranks
of one NVSwitch box and the per-block results come back with ONE all-gather per ``process()`` call
(SURVEY.md section 8e).

One process per GPU (``torchrun``), weights replicated, every rank calls ``process()`` with the SAME
arguments.  Every rank holds the whole (cheap, host-side) :class:`~open_provence_b200.scoring.BlockTable`
of each host-preparation chunk, derives the same deterministic assignment, scores its own blocks
(``submit``: device work only, nothing crosses PCIe or NVLink), and at the end of the call contributes one
fixed-width record per chunk ``[rank scores | fragment means]`` -- all chunks concatenated -- to a single
``all_gather_into_tensor`` (``collect``).  The per-sentence prune then runs on every rank from identical
inputs.  Sentences inside the 1e-5 guard band are re-evaluated with the reference's exact CPU arithmetic
by the rank that owns their logits and shared with one small all-reduce -- only in calls that have such a
sentence, and with the same outcome on every rank.
"""

from __future__ import annotations

from typing import Any, Sequence

import numpy as np
import torch

from .scoring import BlockTable


def block_cost(n_tokens: int, hidden: int = 512, inter: int = 2048, global_fraction: float = 1.0 / 3.0) -> float:
    """Relative FLOPs of one block of ``n_tokens``: linear GEMM term + quadratic
    global-attention term + banded local-attention term."""
    n = float(n_tokens)
    gemm = n * (8.0 * hidden * hidden + 6.0 * hidden * inter)
    attn_global = global_fraction * 4.0 * hidden * n * n
    attn_local = (1.0 - global_fraction) * 4.0 * hidden * n * min(n, 129.0)
    return gemm + attn_global + attn_local


def lpt_assign(costs: Sequence[float], world: int) -> list[np.ndarray]:
    """Longest-processing-time-first: heaviest block to the currently lightest rank.  Deterministic
    (ties broken by index), so every rank derives the same plan without communicating."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.lexsort((np.arange(costs.size), -costs))
    loads = np.zeros(world)
    bins: list[list[int]] = [[] for _ in range(world)]
    for b in order:
        r = int(np.argmin(loads))  # first minimum -> deterministic
        bins[r].append(int(b))
        loads[r] += costs[b]
    return [np.asarray(sorted(b), dtype=np.int64) for b in bins]


class _Ticket:
    """One submitted chunk: its table, the plan every rank derived, and this rank's record on the device."""

    __slots__ = ("table", "threshold", "shards", "slots_of", "max_blocks", "max_slots", "record", "kept", "frag_mean")

    def __init__(self, **kw: Any) -> None:
        for key, value in kw.items():
            setattr(self, key, value)


class ShardedScorer:
    """Wraps a scorer (``score_blocks`` / ``prune`` / ``engine.device``) with the shard + single all-gather.

    ``submit(table, threshold)`` scores this rank's blocks of one chunk; ``collect(tickets)`` performs the one
    collective of the call and returns the per-chunk results; ``run`` is both for a single table."""

    def __init__(self, scorer: Any, group: Any = None, hidden: int = 512, inter: int = 2048) -> None:
        self.scorer = scorer
        self.group = group
        self.hidden, self.inter = hidden, inter
        self.collectives = 0  # all-gather / all-reduce calls issued so far (tests, profiling)

    @property
    def max_tokens(self) -> int:
        return self.scorer.max_tokens

    @max_tokens.setter
    def max_tokens(self, value: int) -> None:
        self.scorer.max_tokens = value

    @property
    def engine(self) -> Any:
        return getattr(self.scorer, "engine", None)

    # ------------------------------------------------------------------ phase 1: local device work
    def submit(self, table: BlockTable, threshold: float) -> _Ticket:
        import torch.distributed as dist

        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        lengths = [int(b.shape[0]) for b in table.block_ids]
        shards = lpt_assign([block_cost(n, self.hidden, self.inter) for n in lengths], world)
        owner = np.zeros(table.n_blocks, dtype=np.int64)
        for r, shard in enumerate(shards):
            owner[shard] = r
        frag_owner = owner[np.asarray(table.frag_block, dtype=np.int64)] if len(table.frag_block) else np.zeros(0, np.int64)
        order = np.argsort(frag_owner, kind="stable")  # fragment slots grouped by owning rank, ascending inside
        first = np.searchsorted(frag_owner[order], np.arange(world + 1))
        slots_of = [order[first[r] : first[r + 1]] for r in range(world)]
        max_blocks = max((len(s) for s in shards), default=0)
        max_slots = max((len(s) for s in slots_of), default=0)

        rank_score, frag_mean, kept = self.scorer.score_blocks(table, shards[rank], host_scores=False)
        dev = frag_mean.device
        record = torch.zeros(max(max_blocks + max_slots, 1), dtype=torch.float32, device=dev)
        mine_b, mine_s = shards[rank], slots_of[rank]
        if len(mine_b):
            record[: len(mine_b)] = rank_score[torch.from_numpy(mine_b).to(dev)]
        if len(mine_s):
            record[max_blocks : max_blocks + len(mine_s)] = frag_mean[torch.from_numpy(mine_s).to(dev)]
        return _Ticket(table=table, threshold=threshold, shards=shards, slots_of=slots_of, max_blocks=max_blocks,
                       max_slots=max_slots, record=record, kept=kept, frag_mean=frag_mean)

    # ------------------------------------------------------------------ phase 2: the one collective + prune
    def collect(self, tickets: Sequence[_Ticket]) -> list[dict[str, np.ndarray]]:
        import torch.distributed as dist

        if not tickets:
            return []
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        mine = torch.cat([t.record for t in tickets]) if len(tickets) > 1 else tickets[0].record
        gathered = torch.empty(world * mine.numel(), dtype=torch.float32, device=mine.device)
        dist.all_gather_into_tensor(gathered, mine, group=self.group)  # the one collective of the call
        self.collectives += 1
        gathered = gathered.view(world, -1)
        host = gathered.cpu().numpy()  # one device -> host copy for the rank scores of every chunk
        dev = mine.device

        results, pending_near = [], []
        at = 0
        for t in tickets:
            width = t.record.numel()
            table = t.table
            full_rank = np.zeros(table.n_blocks, dtype=np.float32)
            full_frag = torch.zeros_like(t.frag_mean)
            for r in range(world):
                if len(t.shards[r]):
                    full_rank[t.shards[r]] = host[r, at : at + len(t.shards[r])]
                if len(t.slots_of[r]):
                    lo = at + t.max_blocks
                    full_frag[torch.from_numpy(t.slots_of[r]).to(dev)] = gathered[r, lo : lo + len(t.slots_of[r])]
            at += width
            out = self.scorer.prune(table, full_rank, full_frag, t.kept, t.threshold, reevaluate=False)
            results.append(out)
            near = out.get("near")
            if near is not None and np.any(near):
                pending_near.append((t, out))

        # Guard band (|p - threshold| <= 1e-5, rare): the rank that scored a fragment recomputes its mean from the
        # fp32 logits exactly as the reference does on the CPU; one all-reduce shares the values, so every rank takes
        # the same keep decisions.  `near` comes from identical gathered inputs, so all ranks agree on taking this branch.
        if pending_near:
            needed_all, spans = [], []
            for t, out in pending_near:
                sent_index = np.asarray(t.table.sent_frag_index, dtype=np.int64)
                sent_offsets = np.asarray(t.table.sent_offsets, dtype=np.int64)
                needed = sorted({int(k) for s in np.nonzero(out["near"])[0]
                                 for k in sent_index[sent_offsets[s] : sent_offsets[s + 1]]})
                spans.append((len(needed_all), needed))
                needed_all.extend(needed)
            buf = torch.zeros(2, max(len(needed_all), 1), dtype=torch.float64)
            for (t, out), (base, needed) in zip(pending_near, spans):
                local = self.scorer.exact_slot_means(set(needed), t.kept)
                for j, slot in enumerate(needed):
                    if slot in local:
                        buf[0, base + j] = local[slot]
                        buf[1, base + j] = 1.0
            buf = buf.to(dev)
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            self.collectives += 1
            vals = buf.cpu().numpy()
            for (t, out), (base, needed) in zip(pending_near, spans):
                slot_mean = {slot: float(vals[0, base + j]) for j, slot in enumerate(needed) if vals[1, base + j] > 0.5}
                self.scorer.apply_exact(out, t.table, slot_mean, t.threshold)
        del rank
        return results

    def run(self, table: BlockTable, threshold: float) -> dict[str, np.ndarray]:
        return self.collect([self.submit(table, threshold)])[0]
