"""Block scoring: the device side of ``process()``.

``process()`` hands a :class:`BlockTable` (token ids per block, fragment ranges, sentence -> fragment
CSR) to a *scorer* and gets back per-block rerank scores and per-sentence keep decisions.  The only
scorer shipped is :class:`DeviceScorer` (the sm_100a engine); tests inject recorded-logit scorers at the
same seam the reference's tests use (they monkeypatch ``OpenProvenceModel.forward``).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import numpy as np
import torch


@dataclass
class BlockTable:
    """Everything the hot path needs about a ``process()`` call, in flat arrays."""

    block_ids: list[np.ndarray] = field(default_factory=list)  # int32 token ids per block
    frag_block: list[int] = field(default_factory=list)  # block index of every fragment slot
    frag_local: list[tuple[int, int]] = field(default_factory=list)  # block-local [start, end)
    sent_offsets: list[int] = field(default_factory=lambda: [0])  # CSR over sentences
    sent_frag_index: list[int] = field(default_factory=list)  # fragment slots of each sentence
    # set by the native packer (host_pack.pack_blocks): all blocks in one array, block_ids are views into it
    packed_ids: np.ndarray | None = None
    block_offsets: np.ndarray | None = None

    @property
    def n_blocks(self) -> int:
        return len(self.block_ids)

    @property
    def n_sentences(self) -> int:
        return len(self.sent_offsets) - 1


def plan_launches(lengths: list[int], max_tokens: int, max_blocks: int = 65535) -> list[tuple[int, int]]:
    """Consecutive [begin, end) block ranges with at most ``max_tokens`` packed tokens each."""
    out: list[tuple[int, int]] = []
    begin, tokens = 0, 0
    for i, n in enumerate(lengths):
        if i > begin and (tokens + n > max_tokens or i - begin >= max_blocks):
            out.append((begin, i))
            begin, tokens = i, 0
        tokens += n
    if begin < len(lengths):
        out.append((begin, len(lengths)))
    return out


def exact_fragment_mean(logits: np.ndarray) -> float:
    """The reference's arithmetic for one fragment, bit for bit (standalone:2918-2920, 3081):
    torch CPU fp32 softmax over the two logits, column 1, numpy fp32 mean."""
    if logits.shape[0] == 0:
        return 1.0
    probs = torch.softmax(torch.from_numpy(np.ascontiguousarray(logits, dtype=np.float32)), dim=-1).numpy()[:, 1]
    return float(probs.mean())


def exact_sentence_probability(frag_means: list[float]) -> float:
    """standalone:3118-3119."""
    avg = float(np.mean(frag_means)) if frag_means else 0.0
    return max(0.0, min(avg, 1.0))


class DeviceScorer:
    """Runs a :class:`BlockTable` on an :class:`~open_provence_b200.engine.Engine`.

    Blocks are packed unpadded into launches of at most ``max_tokens`` tokens (host -> device once per
    launch), fragment means are reduced on the device, the per-sentence prune runs on the device for all
    sentences of the call, and only ``(rank_score, sentence_prob, keep)`` cross PCIe.  Sentences whose
    probability lands within ``guard`` of the threshold are re-evaluated on the host from their fp32
    logits with the reference's exact procedure, so keep decisions are a pure function of the logits.
    """

    def __init__(self, engine: Any, max_tokens: int = 131072, guard: float = 1e-5) -> None:
        self.engine = engine
        self.max_tokens = int(max_tokens)
        self.guard = float(guard)

    def run(self, table: BlockTable, threshold: float) -> dict[str, np.ndarray]:
        """Single-GPU path: score every block, then prune."""
        self.last_n_blocks = table.n_blocks
        if self.single_launch_fast_path:
            out = self._run_single_launch(table, threshold)
            if out is not None:
                return out
        rank_score, frag_mean_dev, kept = self.score_blocks(table, np.arange(table.n_blocks))
        return self.prune(table, rank_score, frag_mean_dev, kept, threshold)

    # ------------------------------------------------------------------ one launch, one copy each way
    single_launch_fast_path = True

    def _staging(self, n_in: int, n_out: int):
        """Pinned host + device staging buffers (int32 inputs, byte outputs), grown on demand and reused: the call
        synchronises before it returns, so nothing is in flight when the next call overwrites them."""
        dev = self.engine.device
        st = getattr(self, "_stage", None)
        if st is None or st["h_in"].numel() < n_in or st["h_out"].numel() < n_out:
            cap_in = max(n_in, 2 * (st["h_in"].numel() if st else 0), 1 << 16)
            cap_out = max(n_out, 2 * (st["h_out"].numel() if st else 0), 1 << 14)
            st = self._stage = {
                "h_in": torch.empty(cap_in, dtype=torch.int32).pin_memory(),
                "d_in": torch.empty(cap_in, dtype=torch.int32, device=dev),
                "h_out": torch.empty(cap_out, dtype=torch.uint8).pin_memory(),
                "d_out": torch.empty(cap_out, dtype=torch.uint8, device=dev),
            }
        return st

    def _run_single_launch(self, table: BlockTable, threshold: float) -> dict[str, np.ndarray] | None:
        """The whole table in ONE forward launch with ONE host->device and ONE device->host copy (the general
        path makes six small copies each way, which is a fifth of the latency of a single-pair call).  Same
        kernels and same arithmetic as score_blocks() + prune(); returns None when the table needs several
        launches or has no fragments / sentences."""
        from . import _native as N

        n_blocks, n_sent = table.n_blocks, table.n_sentences
        frag_block = np.asarray(table.frag_block, dtype=np.int64)
        n_frags = int(frag_block.shape[0])
        if n_blocks == 0 or n_blocks > 65535 or n_frags == 0 or n_sent == 0:
            return None
        if table.block_offsets is not None:
            cu64 = np.asarray(table.block_offsets, dtype=np.int64)
        else:
            cu64 = np.zeros(n_blocks + 1, dtype=np.int64)
            np.cumsum([int(b.shape[0]) for b in table.block_ids], out=cu64[1:])
        n_tokens = int(cu64[-1])
        if n_tokens == 0 or n_tokens > self.max_tokens:
            return None
        eng = self.engine
        dev = eng.device
        frag_local = np.asarray(table.frag_local, dtype=np.int64).reshape(-1, 2)
        ranges = (frag_local + cu64[:-1][frag_block][:, None]).astype(np.int32)
        sent_index = np.asarray(table.sent_frag_index, dtype=np.int32)
        sent_offsets = np.asarray(table.sent_offsets, dtype=np.int32)
        max_seqlen = int(np.diff(cu64).max())

        def up4(n: int) -> int:
            return (n + 3) & ~3

        o_ids = 0
        o_cu = o_ids + up4(n_tokens)
        o_rng = o_cu + up4(n_blocks + 1)
        o_off = o_rng + up4(2 * n_frags)
        o_idx = o_off + up4(n_sent + 1)
        n_in = o_idx + up4(max(int(sent_index.shape[0]), 1))
        b_prob = 0
        b_score = b_prob + 8 * n_sent
        b_mean = b_score + 4 * up4(n_blocks)
        b_keep = b_mean + 4 * up4(n_frags)
        b_near = b_keep + up4(n_sent)
        n_out = b_near + up4(n_sent)
        st = self._staging(n_in, n_out)
        h_in = st["h_in"].numpy()
        if table.packed_ids is not None:
            h_in[o_ids : o_ids + n_tokens] = table.packed_ids
        else:
            np.concatenate(table.block_ids, out=h_in[o_ids : o_ids + n_tokens])
        h_in[o_cu : o_cu + n_blocks + 1] = cu64
        h_in[o_rng : o_rng + 2 * n_frags] = ranges.reshape(-1)
        h_in[o_off : o_off + n_sent + 1] = sent_offsets
        h_in[o_idx : o_idx + sent_index.shape[0]] = sent_index
        d_in, d_out = st["d_in"], st["d_out"]
        d_in[:n_in].copy_(st["h_in"][:n_in], non_blocking=True)
        prune, rank = eng.forward_packed(d_in[o_ids : o_ids + n_tokens], d_in[o_cu : o_cu + n_blocks + 1], max_seqlen)
        p_in, p_out, stream = d_in.data_ptr(), d_out.data_ptr(), eng._stream()
        with torch.cuda.device(dev):
            N.check(eng.lib.opv_fragment_means(
                prune.data_ptr(), n_tokens, p_in + 4 * o_rng, n_frags, p_out + b_mean, rank.data_ptr(), n_blocks,
                int(rank.shape[1]), p_out + b_score, stream), "opv_fragment_means")
            N.check(eng.lib.opv_sentence_prune(
                p_out + b_mean, p_in + 4 * o_off, p_in + 4 * o_idx, n_sent, float(threshold), float(self.guard),
                p_out + b_prob, p_out + b_keep, p_out + b_near, stream), "opv_sentence_prune")
        st["h_out"][:n_out].copy_(d_out[:n_out], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        h_out = st["h_out"].numpy()
        prob_h = h_out[b_prob : b_prob + 8 * n_sent].view(np.float64).copy()
        rank_score = h_out[b_score : b_score + 4 * n_blocks].view(np.float32).copy()
        frag_mean_h = h_out[b_mean : b_mean + 4 * n_frags].view(np.float32).copy()
        keep_h = h_out[b_keep : b_keep + n_sent].astype(bool)
        near_h = h_out[b_near : b_near + n_sent].astype(bool)
        if near_h.any():  # rare: re-evaluate with the reference's exact CPU arithmetic (same as prune())
            self._reevaluate_near(near_h, sent_index, sent_offsets, [(ranges, np.arange(n_frags), prune)], frag_mean_h,
                                  prob_h, keep_h, threshold)
        return {"rank_score": rank_score, "frag_mean": frag_mean_h, "sent_prob": prob_h, "keep": keep_h, "near": near_h}

    @staticmethod
    def exact_slot_means(needed: set, kept_logits) -> dict[int, float]:
        """Fragment slots of ``needed`` whose logits are on THIS device -> mean keep-probability computed from the
        fp32 logits exactly as the reference does on the CPU (standalone:2918-2920, 3081)."""
        slot_mean: dict[int, float] = {}
        for ranges, slots, prune in kept_logits:
            for j, slot in enumerate(slots):
                if int(slot) in needed:
                    a, b = int(ranges[j, 0]), int(ranges[j, 1])
                    logits = prune[a:b].cpu().numpy() if b > a else np.zeros((0, 2), np.float32)
                    slot_mean[int(slot)] = exact_fragment_mean(logits)
        return slot_mean

    @staticmethod
    def _apply_exact_means(near_h, sent_index, sent_offsets, slot_mean, frag_mean_h, prob_h, keep_h, threshold) -> None:
        """standalone:3118-3119 for the guard-band sentences, from exact fragment means; in place."""
        for s in np.nonzero(near_h)[0]:
            members = [int(k) for k in sent_index[sent_offsets[s] : sent_offsets[s + 1]]]
            exact = exact_sentence_probability([slot_mean.get(k, float(frag_mean_h[k])) for k in members])
            prob_h[s] = exact
            keep_h[s] = exact > threshold

    def apply_exact(self, out: dict, table: BlockTable, slot_mean: dict[int, float], threshold: float) -> None:
        """Sharded path: finish the guard-band sentences of ``out`` (a prune(reevaluate=False) result) from exact
        fragment means gathered over all ranks."""
        self._apply_exact_means(out["near"], np.asarray(table.sent_frag_index, dtype=np.int64),
                                np.asarray(table.sent_offsets, dtype=np.int64), slot_mean, out["frag_mean"],
                                out["sent_prob"], out["keep"], threshold)

    @classmethod
    def _reevaluate_near(cls, near_h, sent_index, sent_offsets, kept_logits, frag_mean_h, prob_h, keep_h, threshold) -> None:
        """Sentences within the guard band of the threshold: recompute from the fp32 logits exactly as the
        reference does on the CPU (standalone:2918-2920, 3081, 3118-3119); updates prob_h / keep_h in place."""
        needed = set()
        for s in np.nonzero(near_h)[0]:
            needed.update(int(k) for k in sent_index[sent_offsets[s] : sent_offsets[s + 1]])
        slot_mean = cls.exact_slot_means(needed, kept_logits)
        cls._apply_exact_means(near_h, sent_index, sent_offsets, slot_mean, frag_mean_h, prob_h, keep_h, threshold)

    def score_blocks(self, table: BlockTable, blocks: np.ndarray, host_scores: bool = True):
        """Forward + score conversion + fragment means for ``blocks`` (indices into the table).

        Returns ``(rank_score [n_blocks], frag_mean cuda fp32 [F], kept_logits)``; entries of blocks that were
        not requested stay 0 (filled in by the all-gather in the sharded path).  ``rank_score`` is a numpy fp32
        array (one host sync) or, with ``host_scores=False``, a CUDA tensor (nothing synchronises: the sharded
        path copies scores to the host once per process() call, after its all-gather)."""
        eng = self.engine
        dev = eng.device
        n_blocks = table.n_blocks
        lengths = [int(b.shape[0]) for b in table.block_ids]
        frag_block = np.asarray(table.frag_block, dtype=np.int64)
        frag_local = np.asarray(table.frag_local, dtype=np.int64).reshape(-1, 2)
        n_frags = frag_block.shape[0]
        rank_score_dev = torch.zeros(max(n_blocks, 1), dtype=torch.float32, device=dev)
        frag_mean_dev = torch.zeros(max(n_frags, 1), dtype=torch.float32, device=dev)
        order = np.argsort(frag_block, kind="stable")  # fragment slots grouped by block
        first_slot = np.searchsorted(frag_block[order], np.arange(n_blocks + 1))
        kept_logits: list[tuple] = []

        blocks = np.asarray(blocks, dtype=np.int64)
        sub_lengths = [lengths[b] for b in blocks]
        for lo, hi in plan_launches(sub_lengths, self.max_tokens):
            chunk = blocks[lo:hi]
            cu = np.zeros(len(chunk) + 1, dtype=np.int32)
            np.cumsum([lengths[b] for b in chunk], out=cu[1:])
            if table.packed_ids is not None and int(chunk[-1]) - int(chunk[0]) + 1 == len(chunk) and np.all(np.diff(chunk) == 1):
                # natively packed table, consecutive blocks: the launch input is a slice of the packed array
                ids = table.packed_ids[int(table.block_offsets[chunk[0]]) : int(table.block_offsets[chunk[-1] + 1])]
            else:
                ids = np.concatenate([table.block_ids[b] for b in chunk]).astype(np.int32, copy=False)
            counts = [int(first_slot[b + 1] - first_slot[b]) for b in chunk]
            slots = np.concatenate([order[first_slot[b] : first_slot[b + 1]] for b in chunk]) if n_frags else np.zeros(0, np.int64)
            base = np.repeat(cu[:-1].astype(np.int64), counts)
            ranges = (frag_local[slots] + base[:, None]).astype(np.int32) if slots.size else np.zeros((0, 2), np.int32)
            d_ids = torch.from_numpy(ids).to(dev, non_blocking=True)
            d_cu = torch.from_numpy(cu).to(dev, non_blocking=True)
            d_ranges = torch.from_numpy(np.ascontiguousarray(ranges)).to(dev, non_blocking=True)
            prune, rank = eng.forward_packed(d_ids, d_cu, int(max(lengths[b] for b in chunk)))
            means, score = eng.fragment_means(prune, d_ranges, rank)
            if slots.size:
                frag_mean_dev[torch.from_numpy(slots).to(dev)] = means
            rank_score_dev[torch.from_numpy(chunk).to(dev)] = score
            kept_logits.append((ranges, slots, prune))
        if not host_scores:
            return rank_score_dev[:n_blocks], frag_mean_dev, kept_logits
        return rank_score_dev[:n_blocks].cpu().numpy(), frag_mean_dev, kept_logits  # one sync, after the last launch

    def prune(self, table: BlockTable, rank_score: np.ndarray, frag_mean_dev: torch.Tensor, kept_logits: list,
              threshold: float, reevaluate: bool = True) -> dict[str, np.ndarray]:
        """Per-sentence mean / threshold on the device for every sentence of the call.  ``reevaluate=False`` leaves
        the guard-band sentences (``near``) to the caller (sharded path: exact means come from the owning ranks)."""
        eng = self.engine
        dev = eng.device
        n_frags = len(table.frag_block)
        sent_offsets = np.asarray(table.sent_offsets, dtype=np.int32)
        sent_index = np.asarray(table.sent_frag_index, dtype=np.int32)
        n_sent = table.n_sentences
        if n_sent > 0:
            d_off = torch.from_numpy(sent_offsets).to(dev)
            d_idx = torch.from_numpy(sent_index if sent_index.size else np.zeros(1, np.int32)).to(dev)
            prob, keep, near = eng.sentence_prune(frag_mean_dev, d_off, d_idx, threshold, self.guard)
            prob_h = prob.cpu().numpy()
            keep_h = keep.cpu().numpy().astype(bool)
            near_h = near.cpu().numpy().astype(bool)
        else:
            prob_h, keep_h, near_h = np.zeros(0), np.zeros(0, bool), np.zeros(0, bool)
        frag_mean_h = frag_mean_dev[:n_frags].cpu().numpy() if n_frags else np.zeros(0, np.float32)

        if reevaluate and near_h.any():  # rare: re-evaluate with the reference's exact CPU arithmetic
            self._reevaluate_near(near_h, sent_index, sent_offsets, kept_logits, frag_mean_h, prob_h, keep_h, threshold)
        return {"rank_score": rank_score, "frag_mean": frag_mean_h, "sent_prob": prob_h, "keep": keep_h, "near": near_h}
