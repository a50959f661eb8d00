#!/usr/bin/env python
"""Attention at the bench shape: this repo's tcgen05 kernels against flash-attn 2 (the kernel the reference uses on
CUDA, standalone:1597-1604; the wheel ships recompiled mma.sync SASS for sm_100).

    python tools/flashattn_compare.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_provence_b200 import ops  # noqa: E402

B, S, heads = 64, 2048, 8
dev = "cuda"
qkv = torch.randn((B * S, 3 * heads * 64), device=dev).to(torch.bfloat16)
cu = torch.arange(0, B * S + 1, S, dtype=torch.int32, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


try:
    from flash_attn import flash_attn_varlen_func
except Exception as exc:  # pragma: no cover
    print(f"flash_attn is not importable here: {exc}")
    flash_attn_varlen_func = None

q, k, v = (t.reshape(B * S, heads, 64) for t in qkv.split(heads * 64, dim=1))
print(f"{B} x {S} tokens, {heads} heads x 64, bf16, one B200")
print("| attention | this repo (ms, TFLOP/s) | flash-attn 2 (ms, TFLOP/s) | max abs diff |\n|---|---|---|---|")
for name, hw in (("global", -1), ("sliding window 64+1+64", 64)):
    ours = timeit(lambda: ops.attention(qkv, cu, S, heads, hw))
    if hw < 0:
        flops = 4.0 * heads * 64 * S * S * B
    else:
        flops = 4.0 * heads * 64 * sum(min(S - 1, i + hw) - max(0, i - hw) + 1 for i in range(S)) * B
    line = f"| {name} | {ours:.3f} ms, {flops / ours / 1e9:.0f} |"
    if flash_attn_varlen_func is not None:
        window = (-1, -1) if hw < 0 else (hw, hw)
        fa = lambda: flash_attn_varlen_func(q, k, v, cu, cu, S, S, softmax_scale=0.125, causal=False, window_size=window)  # noqa: E731
        t_fa = timeit(fa)
        diff = (fa().reshape(B * S, heads * 64).float() - ops.attention(qkv, cu, S, heads, hw).float()).abs().max().item()
        line += f" {t_fa:.3f} ms, {flops / t_fa / 1e9:.0f} | {diff:.1e} |"
    else:
        line += " n/a | n/a |"
    print(line)
