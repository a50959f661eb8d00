"""Multi-GPU data parallelism for ``process()``: blocks are independent, so they are dealt to the ranks
of one NVSwitch box and the per-block results come back with ONE all-gather.

One process per GPU (``torchrun``), weights replicated.  Every rank holds the whole (cheap, host-side)
:class:`~open_provence_b200.scoring.BlockTable`, computes the same deterministic assignment, scores its
own blocks, and contributes a fixed-width record ``[rank scores | fragment means]`` (padded to the largest
shard) to ``all_gather_into_tensor``.  No other exchange happens on the data path.
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


class ShardedScorer:
    """Wraps a scorer (``score_blocks`` / ``prune`` / ``engine.device``) with the shard + all-gather."""

    def __init__(self, scorer: Any, group: Any = None, hidden: int = 512, inter: int = 2048) -> None:
        self.scorer = scorer
        self.group = group
        self.hidden, self.inter = hidden, inter

    @property
    def max_tokens(self) -> int:
        return self.scorer.max_tokens

    @max_tokens.setter
    def max_tokens(self, value: int) -> None:
        self.scorer.max_tokens = value

    def run(self, table: BlockTable, threshold: float) -> dict[str, np.ndarray]:
        import torch.distributed as dist

        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        lengths = [int(b.shape[0]) for b in table.block_ids]
        shards = lpt_assign([block_cost(n, self.hidden, self.inter) for n in lengths], world)
        frag_block = np.asarray(table.frag_block, dtype=np.int64)
        slots_of = [np.nonzero(np.isin(frag_block, s))[0] for s in shards]
        max_blocks = max((len(s) for s in shards), default=0)
        max_slots = max((len(s) for s in slots_of), default=0)

        rank_score, frag_mean, kept = self.scorer.score_blocks(table, shards[rank])
        dev = frag_mean.device
        width = max_blocks + max_slots
        record = torch.zeros(max(width, 1), dtype=torch.float32, device=dev)
        mine_b, mine_s = shards[rank], slots_of[rank]
        if len(mine_b):
            record[: len(mine_b)] = torch.from_numpy(rank_score[mine_b]).to(dev)
        if len(mine_s):
            record[max_blocks : max_blocks + len(mine_s)] = frag_mean[torch.from_numpy(mine_s).to(dev)]
        gathered = torch.empty(world * record.numel(), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(gathered, record, group=self.group)  # the one collective
        gathered = gathered.view(world, -1)
        full_rank = np.zeros(table.n_blocks, dtype=np.float32)
        full_frag = torch.zeros_like(frag_mean)
        host = gathered[:, :max_blocks].cpu().numpy() if max_blocks else np.zeros((world, 0), np.float32)
        for r in range(world):
            if len(shards[r]):
                full_rank[shards[r]] = host[r, : len(shards[r])]
            if len(slots_of[r]):
                full_frag[torch.from_numpy(slots_of[r]).to(dev)] = gathered[r, max_blocks : max_blocks + len(slots_of[r])]
        return self.scorer.prune(table, full_rank, full_frag, kept, threshold)
