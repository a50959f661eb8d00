#!/usr/bin/env python
"""Correctness + timing of one bf16 attention kernel (run once per impl, each in its own process so that a
device trap in one variant cannot poison the others).

    python tools/attn_check.py <impl 1..5> [half_window] [trace]
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_provence_b200 import ops  # noqa: E402

impl = int(sys.argv[1])
hw = int(sys.argv[2]) if len(sys.argv) > 2 else -1
dev = torch.device("cuda:0")
ops.set_option("attention_impl", impl)
import os
dbg = int(os.environ.get("ATTN_DEBUG", "0"))
if os.environ.get("ATTN_Q4_POLY"):
    ops.set_option("attention_q4_poly", int(os.environ["ATTN_Q4_POLY"]))


def ref_attention(qkv, lengths, heads, half_window):
    outs, start = [], 0
    for n in lengths:
        blk = qkv[start:start + n].double().view(n, 3, heads, 64)
        q, k, v = blk[:, 0], blk[:, 1], blk[:, 2]
        s = torch.einsum("ihd,jhd->hij", q, k) * 0.125
        if half_window >= 0:
            idx = torch.arange(n, device=qkv.device)
            s = s.masked_fill(~((idx[:, None] - idx[None, :]).abs() <= half_window)[None], float("-inf"))
        outs.append(torch.einsum("hij,jhd->ihd", torch.softmax(s, -1), v).reshape(n, heads * 64))
        start += n
    return torch.cat(outs, 0)


for lengths, heads in (([128], 1), ([256], 1), ([100, 300, 129, 1, 640], 2)):
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn((sum(lengths), 3 * heads * 64), generator=g).to(torch.bfloat16).to(dev)
    cu = torch.tensor([0] + list(np.cumsum(lengths)), dtype=torch.int32, device=dev)
    out = ops.attention(qkv, cu, max(lengths), heads, hw)
    torch.cuda.synchronize()
    ref = ref_attention(qkv, lengths, heads, hw)
    err = (out.double() - ref).abs()
    print(f"impl={impl} hw={hw} lengths={lengths} heads={heads}: max err {err.max().item():.3e} mean {err.mean().item():.3e}"
          f" finite={bool(torch.isfinite(out.float()).all())}", flush=True)
    if err.max().item() > 2e-2:
        rows = (err.max(dim=1).values > 2e-2).nonzero().flatten()[:8].tolist()
        cols = (err.max(dim=0).values > 2e-2).nonzero().flatten()[:16].tolist()
        print(f"   bad rows {rows} ... bad cols {cols} ...; out[0,:8]={out[0,:8].float().tolist()} ref[0,:8]={ref[0,:8].tolist()}")

# timing at the bench shape: 64 x 2048, 8 heads
if dbg:
    ops.set_option("attention_debug", dbg)
    print(f"attention_debug={dbg}: results below are WRONG by construction, timing only")
heads, S, B = 8, 2048, 64
qkv = torch.randn((B * S, 3 * heads * 64), device=dev).to(torch.bfloat16)
cu = torch.arange(0, B * S + 1, S, dtype=torch.int32, device=dev)
for _ in range(3):
    ops.attention(qkv, cu, S, heads, hw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attention(qkv, cu, S, heads, hw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
if hw < 0:
    flops = 4.0 * heads * 64 * S * S * B
else:
    pairs = sum(min(S - 1, i + hw) - max(0, i - hw) + 1 for i in range(S))
    flops = 4.0 * heads * 64 * pairs * B
print(f"impl={impl} hw={hw} B={B} S={S} heads={heads}: {ms:.3f} ms/launch, {flops / ms / 1e9:.1f} TFLOP/s (algorithmic)", flush=True)

if len(sys.argv) > 3 and sys.argv[3] == "trace":
    nb = (S + 127) // 128 if hw < 0 else 3
    if impl == 5:
        buf = torch.zeros(32 * 64, dtype=torch.int64, device=dev)
        ops.set_option("attention_trace_ptr", buf.data_ptr())
        ops.attention(qkv, cu, S, heads, hw)
        torch.cuda.synchronize()
        ops.set_option("attention_trace_ptr", 0)
        t = buf.cpu().numpy().reshape(64, 32)[:nb]
        t0 = t[0, 0]
        print("two-Q-tile kernel, CTA 0, first super tile; cycles since tile A saw s_full(0)")
        print("blk | A: s_full loaded max_ok exp_done pv_done P_pub | B: s_full loaded max_ok exp_done pv_done P_pub | MMA: S_A S_B PV_A PV_B | TMA: K V")
        for i in range(nb):
            def f(v):
                return f"{int(v - t0):7d}" if v else "      -"
            print(f"{i:3d} | " + " ".join(f(v) for v in t[i][0:6]) + " | " + " ".join(f(v) for v in t[i][8:14]) + " | "
                  + " ".join(f(v) for v in t[i][16:20]) + " | " + " ".join(f(v) for v in t[i][24:26]))
    else:
        buf = torch.zeros(16 * 64, dtype=torch.int64, device=dev)
        ops.set_option("attention_trace_ptr", buf.data_ptr())
        ops.attention(qkv, cu, S, heads, hw)
        torch.cuda.synchronize()
        ops.set_option("attention_trace_ptr", 0)
        t = buf.cpu().numpy().reshape(64, 16)[:nb]
        t0 = t[0, 0]
        print("block: s_full  S_loaded  max_done  exp_done  pv_done  P_stored | S_issued(i) PV_issued(i) | per softmax warp 0..2: S_loaded, P_stored   (cycles since block 0 s_full)")
        for i in range(nb):
            print(f"{i:3d}: " + " ".join(f"{int(v - t0):8d}" if v else "       -" for v in t[i][:14]))
        print(f"kernel entry {int(t[0][14] - t0)}, exit {int(t[0][15] - t0)} (cycles relative to block 0 s_full)")
