#!/usr/bin/env python
"""Timing (and a correctness spot check) of the four projection GEMMs at the bench shape, per kernel variant.

    python tools/gemm_check.py <gemm_pair 0|1> [M] [case] [iters]     (iters = 0: one launch per case, for ncu)
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_provence_b200 import _native as N  # noqa: E402
from open_provence_b200 import ops  # noqa: E402
from open_provence_b200.engine import interleave_wi, rope_table  # noqa: E402

pair = int(sys.argv[1])
M = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
ONLY = sys.argv[3] if len(sys.argv) > 3 else ""
ITERS = int(sys.argv[4]) if len(sys.argv) > 4 else 20
ops.set_option("gemm_pair", pair)
dev = "cuda"
H, I = 512, 2048
g = torch.Generator().manual_seed(0)


def rnd(shape, scale=1.0):
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(dev)


cos, sin = rope_table(8192, 64, 160000.0)
cos, sin = cos.to(dev), sin.to(dev)
pos = (torch.arange(M, dtype=torch.int32) % 2048).to(dev)
cases = {
    "qkv_rope": dict(a=rnd((M, H)), w=rnd((3 * H, H), 0.05), kw=dict(epilogue=N.EPI_ROPE, pos=pos, cos=cos, sin=sin, hidden_size=H), flops=2.0 * M * 3 * H * H),
    "wo_residual": dict(a=rnd((M, H)), w=rnd((H, H), 0.05), kw=dict(epilogue=N.EPI_RESIDUAL), flops=2.0 * M * H * H),
    "wi_geglu": dict(a=rnd((M, H)), w=interleave_wi(rnd((2 * I, H), 0.05)), kw=dict(epilogue=N.EPI_GEGLU), flops=2.0 * M * 2 * I * H),
    "wo2_residual": dict(a=rnd((M, I)), w=rnd((H, I), 0.05), kw=dict(epilogue=N.EPI_RESIDUAL), flops=2.0 * M * H * I),
    "store_4096": dict(a=rnd((M, H)), w=rnd((2 * I, H), 0.05), kw=dict(), flops=2.0 * M * 2 * I * H),
}
for name, c in cases.items():
    if ONLY and name != ONLY:
        continue
    kw = dict(c["kw"])
    if kw.get("epilogue") == N.EPI_RESIDUAL:
        kw["out"] = torch.zeros((M, c["w"].shape[0]), dtype=torch.float32, device=dev)
    out = ops.gemm(c["a"], c["w"], **kw)
    torch.cuda.synchronize()
    # spot check 256 rows against torch
    rows = torch.randint(0, M, (256,), generator=g).to(dev)
    ref = c["a"][rows].float() @ c["w"].float().T
    if name in ("wo_residual", "wo2_residual", "store_4096"):
        err = (out[rows].float() - ref).abs().max().item() / max(1.0, ref.abs().max().item())
    else:
        err = float("nan")  # rope / geglu are covered by tests/test_gpu_kernels.py
    if ITERS <= 0:
        continue
    for _ in range(3):
        ops.gemm(c["a"], c["w"], **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ITERS):
        ops.gemm(c["a"], c["w"], **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / ITERS
    print(f"pair={pair} M={M} {name:13s}: {ms * 1e3:8.1f} us  {c['flops'] / ms / 1e9:7.1f} TFLOP/s  rel err {err:.2e}", flush=True)
