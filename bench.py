#!/usr/bin/env python
"""Benchmark of the OpenProvence scoring-and-pruning hot path on B200 (contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic pre-tokenised (question, context)
blocks: ModernBERT forward -> rerank score + token keep-probabilities -> per-sentence mean /
threshold / prune.  Metric: query-context pairs/sec at seq_len=2048, base-130M (BASELINE.json).
Weak scaling: every GPU gets its own batch of 64 blocks; for N > 1 each step ends with the single
NCCL all-gather of the per-block scores and per-sentence results.

Rank 0 prints ONE JSON line.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402



def metric_name(args) -> str:
    """BASELINE.json's metric for the default workload; other --model/--seq-len runs name themselves."""
    return f"query-context pairs/sec at seq_len={args.seq_len}, {args.model}"


UNIT = "pairs/s"
# /root/reference cannot travel to the GPU box, so the timed CPU arm is the port (oracle/hf_cpu_baseline.py: the reference's
# library forward + its numpy / torch post-processing).  Checked once in the build container against the UNMODIFIED
# reference's own loop (process() -> _run_inference_batches, standalone:2761-2939) on the same blocks and weights:
PORT_NOTE = ("port loop = 0.99 x the time of the unmodified reference's process() on the same 4 x 2048-token blocks, identical "
             "scores (profiles/r2_reference_loop_vs_port.md, tools/reference_loop_compare.py)")
THRESHOLD = 0.1


_REAL_STDOUT = None


def quiet_stdout() -> None:
    """From here on everything libraries print to stdout (NCCL's version banner, ...) goes to stderr; the ONE
    JSON line is written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="base-130M")
    ap.add_argument("--seq-len", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=64, help="blocks per GPU per step")
    ap.add_argument("--mode", default="dense", choices=["dense", "ragged"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-pairs", type=int, default=32, help="pairs in the bounded CPU-baseline sample")
    ap.add_argument("--ref-pairs-per-step", type=int, default=4)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch blocks PER GPU; strong: --batch blocks IN TOTAL, dealt to the ranks by "
                         "longest-processing-time-first on the FLOP cost model (sharding.lpt_assign)")
    ap.add_argument("--launch-tokens", type=int, default=262144, help="packed tokens per forward launch (strong scaling)")
    ap.add_argument("--set-option", action="append", default=[], metavar="NAME=VALUE",
                    help="library tuning switch for A/B runs (opv_set_option), e.g. zigzag=0")
    return ap.parse_args()


def measured_peaks() -> dict:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        data = json.loads(path.read_text())
        data["source"] = "measured"
        return data
    # B200_PROFILING.md fallback (an earlier measurement on this pool)
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    FIELDS = (
        "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
        "clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower() == "active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {
            "sm_mhz": statistics.median(sm),
            "sm_max_mhz": max(smax),
            "power_w_max": max(power),
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def run_reference(args, world: int, rank: int) -> None:
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    from open_provence_b200 import synthetic as syn
    from oracle import hf_cpu_baseline as hb

    cfg = syn.backbone_config(args.model)
    sd = syn.random_state_dict(cfg, seed=0)
    model, head = hb.build_hf_model(cfg, sd)
    per_step = max(1, args.ref_pairs_per_step)
    n_pairs = per_step * (args.steps + args.warmup)
    wl = syn.make_workload(cfg, n_pairs, args.seq_len, mode=args.mode, seed=1234)
    at = 0
    for _ in range(args.warmup):
        hb.score_workload(model, head, wl, slice(at, at + per_step), THRESHOLD)
        at += per_step
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hb.score_workload(model, head, wl, slice(at, at + per_step), THRESHOLD)
        at += per_step
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"{per_step} blocks of {args.seq_len} tokens per step, {args.steps} steps, HF ModernBERT fp32 sdpa on CPU"
    line = {
        "impl": "reference",
        "metric": metric_name(args),
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.model} seq_len={args.seq_len} {args.mode}, {per_step} blocks per step (bounded CPU sample)",
                   "transformers": __import__("transformers").__version__},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count(), "port_vs_reference_loop": PORT_NOTE},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def parity_against_cpu(cpu: dict, gpu: dict, wl: dict, n_pairs: int) -> dict:
    """The GPU results of the benchmarked step against the reference's fp32 CPU arithmetic (HF ModernBERT fp32 +
    numpy prune, ``cpu_baseline`` leg) on the same first ``n_pairs`` blocks: standalone:2893-2924, 3116-3134."""
    cu = wl["cu_seqlens"]
    t_end = int(cu[n_pairs])
    cpu_prune = np.concatenate(cpu["prune_logits"])
    cpu_rank = np.stack(cpu["rank_logits"]).reshape(n_pairs, -1)
    frag = np.asarray(cpu["frag_index"], dtype=np.int64)
    cpu_prob = np.asarray(cpu["sent_prob"], dtype=np.float64)
    cpu_keep = np.asarray(cpu["keep"], dtype=bool)
    g_prob, g_keep = gpu["sent_prob"][frag], gpu["keep"][frag]
    mism = g_keep != cpu_keep
    band = np.abs(cpu_prob - THRESHOLD) <= 1e-2
    scale = float(np.abs(cpu_prune).max())
    return {
        "against": "reference CPU arithmetic in fp32 (cpu_baseline leg), same blocks, same weights",
        "blocks": n_pairs,
        "sentences": int(frag.size),
        "prune_logit_max_abs": float(np.abs(gpu["prune_logits"][:t_end] - cpu_prune).max()),
        "prune_logit_mean_abs": float(np.abs(gpu["prune_logits"][:t_end] - cpu_prune).mean()),
        "prune_logit_scale": scale,
        "prune_logit_max_rel_to_scale": float(np.abs(gpu["prune_logits"][:t_end] - cpu_prune).max() / scale),
        "rank_logit_max_abs": float(np.abs(gpu["rank_logits"][:n_pairs].reshape(n_pairs, -1) - cpu_rank).max()),
        "rank_score_max_abs": float(np.abs(gpu["rank_score"][:n_pairs] - np.asarray(cpu["rank_score"])).max()),
        "keep_prob_max_abs": float(np.abs(g_prob - cpu_prob).max()),
        "keep_mismatches": int(mism.sum()),
        "mismatches_outside_1e-2_band": int((mism & ~band).sum()),
        "sentences_inside_1e-2_band": int(band.sum()),
        "kept_fraction_cpu": float(cpu_keep.mean()),
        "e2e_equals_device_path": bool(np.array_equal(gpu["e2e_keep"], gpu["keep"])
                                       and np.array_equal(gpu["e2e_sent_prob"], gpu["sent_prob"])),
    }


def main() -> None:
    args = parse_args()
    quiet_stdout()
    if args.impl == "reference":
        # CPU only: no process group; under torchrun rank 0 alone runs it, with every host thread it can use
        # (torchrun exports OMP_NUM_THREADS=1, which would otherwise pin the baseline to one core)
        rank = int(os.environ.get("RANK", "0"))
        if rank == 0:
            torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
            run_reference(args, int(os.environ.get("WORLD_SIZE", "1")), rank)
        return
    world, rank, local = dist_setup(args)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback)")
    from open_provence_b200 import synthetic as syn
    from open_provence_b200.engine import Engine

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if args.set_option:
        from open_provence_b200 import ops

        for item in args.set_option:
            name, _, value = item.partition("=")
            ops.set_option(name, int(value))
    cfg = syn.backbone_config(args.model)
    sd = syn.random_state_dict(cfg, seed=0)
    eng = Engine(cfg, sd, device=dev, dtype=args.dtype, num_labels=1)
    del sd
    strong = args.scaling == "strong"
    if strong:
        # ONE global block list, identical on every rank; blocks dealt by LPT on the cost model, ragged or dense
        from open_provence_b200.scoring import plan_launches
        from open_provence_b200.sharding import block_cost, lpt_assign

        wl_global = syn.make_workload(cfg, args.batch, args.seq_len, mode=args.mode, seed=1234)
        costs = [block_cost(int(n), cfg["hidden_size"], cfg["intermediate_size"]) for n in wl_global["lengths"]]
        shards = lpt_assign(costs, world)
        mine = shards[rank]
        wl = syn.slice_workload(wl_global, mine)
        spans = plan_launches([int(n) for n in wl["lengths"]], args.launch_tokens)
        parts = [syn.slice_workload(wl, np.arange(lo, hi)) for lo, hi in spans]
        blocks_total = args.batch
        shard_blocks = [int(len(sh)) for sh in shards]
    else:
        wl = syn.make_workload(cfg, args.batch, args.seq_len, mode=args.mode, seed=1234 + rank)
        parts = [wl]
        blocks_total = world * args.batch
        shard_blocks = [args.batch] * world
    T = int(wl["ids"].shape[0])
    n_frags = int(wl["frag_ranges"].shape[0])
    n_mine = int(wl["n_pairs"])

    # host (pinned) and device-resident copies of the step inputs, one set per forward launch
    KEYS = ("ids", "cu_seqlens", "frag_ranges", "sent_offsets", "sent_frag_index")
    hosts = [{k: torch.from_numpy(np.ascontiguousarray(pt[k])).pin_memory() for k in KEYS} for pt in parts]
    devs = [{k: v.to(dev) for k, v in h.items()} for h in hosts]
    host, d = hosts[0], devs[0]

    frag_max, blk_max = n_frags, n_mine
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([n_frags, n_mine], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        frag_max, blk_max = int(t[0].item()), int(t[1].item())
    rec_len = blk_max + 2 * frag_max
    record = torch.zeros(max(rec_len, 1), dtype=torch.float32, device=dev)
    gathered = torch.empty(world * record.numel(), dtype=torch.float32, device=dev) if world > 1 else None

    def step_device():
        b_at = f_at = 0
        for pt, dd in zip(parts, devs):
            prune, rank_logits = eng.forward_packed(dd["ids"], dd["cu_seqlens"], pt["max_seqlen"])
            frag_mean, score = eng.fragment_means(prune, dd["frag_ranges"], rank_logits)
            prob, keep, near = eng.sentence_prune(frag_mean, dd["sent_offsets"], dd["sent_frag_index"], THRESHOLD)
            if world > 1:
                nb, nf = int(pt["n_pairs"]), int(pt["frag_ranges"].shape[0])
                record[b_at : b_at + nb] = score
                record[blk_max + f_at : blk_max + f_at + nf] = prob.float()
                record[blk_max + frag_max + f_at : blk_max + frag_max + f_at + nf] = keep.float()
                b_at, f_at = b_at + nb, f_at + nf
        if world > 1:
            dist.all_gather_into_tensor(gathered, record)  # the one collective of the step
        return score, prob, keep

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(3, args.warmup)):
        step_device()
    sync_all()
    launches_before = eng.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    sync_all()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    # forward launches (counted in the engine) + 3 prune-path kernels per step
    launches = (eng.launch_count() - launches_before) // args.steps + 3 * len(parts)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = blocks_total / (ms_per_step * 1e-3)

    # ---- e2e: the host-buffer call, H2D of the step's inputs and D2H of its results inside the timed region
    h2d = sum(int(v.numel() * v.element_size()) for h in hosts for v in h.values())

    def step_host():
        outs = []
        for pt, h in zip(parts, hosts):
            outs.append(eng.score_packed_host(h["ids"], h["cu_seqlens"], pt["max_seqlen"], h["frag_ranges"],
                                              h["sent_offsets"], h["sent_frag_index"], THRESHOLD))
        if world > 1:
            b_at = 0
            for pt, o in zip(parts, outs):
                record[b_at : b_at + int(pt["n_pairs"])] = o["rank_score"].to(dev, non_blocking=True)
                b_at += int(pt["n_pairs"])
            dist.all_gather_into_tensor(gathered, record)
        return outs

    for _ in range(2):
        outs = step_host()
    out = outs[0]
    d2h = sum(int(v.numel() * v.element_size()) for o in outs for v in o.values())
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    sync_all()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = blocks_total * args.steps / e2e_s

    # ---- outputs of the benchmarked step itself (device-resident inputs, same kernels) for the in-run parity check
    gpu_check = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not strong:
        prune_l, rank_l = eng.forward_packed(d["ids"], d["cu_seqlens"], wl["max_seqlen"])
        frag_mean, score = eng.fragment_means(prune_l, d["frag_ranges"], rank_l)
        prob, keep, near = eng.sentence_prune(frag_mean, d["sent_offsets"], d["sent_frag_index"], THRESHOLD)
        torch.cuda.synchronize(dev)
        gpu_check = {
            "prune_logits": prune_l.float().cpu().numpy(), "rank_logits": rank_l.float().cpu().numpy(),
            "rank_score": score.float().cpu().numpy(), "sent_prob": prob.double().cpu().numpy(),
            "keep": keep.cpu().numpy().astype(bool), "e2e_keep": out["keep"].numpy().astype(bool),
            "e2e_sent_prob": out["sent_prob"].double().numpy(),
        }

    # ---- per-kernel-class device times of one step (CUDA events on the launch stream, inside the library)
    eng.profile(True)
    prof_steps = min(3, args.steps)
    for _ in range(prof_steps):
        for pt, dd in zip(parts, devs):
            eng.forward_packed(dd["ids"], dd["cu_seqlens"], pt["max_seqlen"])
    prof = eng.profile_collect()
    eng.profile(False)
    profile_ms = {k: round(v["ms"] / prof_steps, 4) for k, v in prof.items()}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    H, I, L = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"]
    Tl = T / max(1, len(parts))  # tokens per forward launch (average when a strong-scaling shard needs several)
    gemm_flops = {
        "gemm_qkv_rope": 2.0 * Tl * 3 * H * H,
        "gemm_wo_residual": 2.0 * Tl * H * H,
        "gemm_wi_geglu": 2.0 * Tl * 2 * I * H,
        "gemm_wo2_residual": 2.0 * Tl * H * I,
    }
    dom = "gemm_wi_geglu"  # largest share of the algorithmic FLOPs (6HI of 8H^2+6HI per token per layer)
    dom_ms = prof[dom]["ms"] / max(1, prof[dom]["launches"])
    achieved_tf = gemm_flops[dom] / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    flops_pair = syn.algorithmic_flops_per_pair(cfg, args.seq_len) if args.mode == "dense" else None
    step_tf = (flops_pair * n_mine / (ms_per_step * 1e-3) / 1e12) if flops_pair else None  # rank 0's share of the step
    traffic = None
    traffic_file = ROOT / "profiles" / "roofline_traffic.json"
    if traffic_file.exists() and args.model == "base-130M" and args.seq_len == 2048 and args.batch == 64 and args.mode == "dense":
        traffic = json.loads(traffic_file.read_text())  # one ncu --set full capture of this kernel at this shape
    roofline = {
        "bound": "tensor",
        "kernel": "gemm_bf16_tcgen05_pair_kernel<GEGLU> (mlp.Wi + GeGLU, 38 % of the step's algorithmic FLOPs)",
        "achieved": round(achieved_tf, 2),
        "peak": peak_tf,
        "unit": "TFLOP/s",
        "frac": round(achieved_tf / peak_tf, 4),
        "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
        "traffic_source": traffic["source"] if traffic else None,
        "algorithmic_bytes_per_launch": int(Tl * (2 * H + 2 * I) + 2 * 2 * I * H),
        "peak_kind": f"bf16 sustained, {peaks['source']} (MEASURED_PEAKS.json)",
        "flops_per_launch": gemm_flops[dom],
        "ms_per_launch": round(dom_ms, 4),
        "all_gemms_tflops": {k: round(v / (prof[k]["ms"] / max(1, prof[k]["launches"]) * 1e-3) / 1e12, 1) for k, v in gemm_flops.items() if prof[k]["ms"] > 0},
        "step_algorithmic_tflops": round(step_tf, 2) if step_tf else None,
        "step_frac_of_burst_peak": round(step_tf / float(peaks["bf16_tflops"]), 4) if step_tf else None,
    }

    cpu_baseline = None
    parity = None
    if world == 1 and not args.no_cpu_baseline and not strong:
        from oracle import hf_cpu_baseline as hb

        sd_cpu = syn.random_state_dict(cfg, seed=0)
        model, head = hb.build_hf_model(cfg, sd_cpu)
        res = hb.time_cpu_baseline(model, head, wl, args.cpu_pairs, THRESHOLD, warmup_pairs=1)
        parity = parity_against_cpu(res["results"], gpu_check, wl, args.cpu_pairs)
        cpu_baseline = {
            "value": round(res["pairs_per_s"], 4),
            "unit": UNIT,
            "cores": res["threads"],
            "kind": "port",
            "sample": f"first {args.cpu_pairs} blocks of the same workload ({res['seconds']:.1f} s), HF ModernBERT fp32 sdpa on CPU + numpy prune",
            "host_cpus": os.cpu_count(),
            "port_vs_reference_loop": PORT_NOTE,
        }

    line = {
        "metric": metric_name(args),
        "value": round(value, 2),
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(3, args.warmup),
        "ms_per_step": round(ms_per_step, 4),
        "higher_is_better": True,
        "scaling": args.scaling,
        "vs_baseline": None,
        "dtype": args.dtype,
        "data": "synthetic",
        "config": {
            "workload": (f"{args.model} seq_len={args.seq_len} batch={args.batch} blocks "
                         f"{'in total (LPT-sharded over the ranks)' if strong else 'per GPU'}, {args.mode}, pre-tokenised, "
                         "random-init weights"),
            "tokens_per_step_per_gpu": T,
            "sentences_per_step_per_gpu": n_frags,
            "blocks_per_rank": shard_blocks,
            "forward_launches_per_step_per_gpu": len(parts),
            "threshold": THRESHOLD,
            "parallelism": f"dp{world} (blocks sharded, weights replicated, one all-gather of scores per step)" if world > 1 else "single GPU",
            "l2": "per-step activations (>1 GB) exceed the 126 MB L2; no explicit flush",
        },
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches) * args.steps,
        "gpu_launches_per_step": int(launches),
        "roofline": roofline,
        "profile_ms_per_step": profile_ms,
        "cpu_baseline": cpu_baseline,
        "parity": parity,
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
