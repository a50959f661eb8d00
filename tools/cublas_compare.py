#!/usr/bin/env python
"""The projection GEMMs at the bench shape: this repo's fused tcgen05 kernels against torch.matmul (cuBLAS) +
the separate elementwise kernels the reference would launch ("cuBLAS is the bar to beat").

    python tools/cublas_compare.py [M]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_provence_b200 import _native as N  # noqa: E402
from open_provence_b200 import ops  # noqa: E402
from open_provence_b200.engine import interleave_wi, rope_table  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
H, I = 512, 2048
dev = "cuda"
g = torch.Generator().manual_seed(0)
rnd = lambda shape, s=1.0: (torch.randn(shape, generator=g) * s).to(torch.bfloat16).to(dev)  # noqa: E731


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


cos, sin = rope_table(8192, 64, 160000.0)
cos, sin = cos.to(dev), sin.to(dev)
pos = (torch.arange(M, dtype=torch.int32) % 2048).to(dev)
x, attn, act = rnd((M, H)), rnd((M, H)), rnd((M, I))
wqkv, wo, wi, wo2 = rnd((3 * H, H), 0.05), rnd((H, H), 0.05), rnd((2 * I, H), 0.05), rnd((H, I), 0.05)
wi_il = interleave_wi(wi)
h = torch.zeros((M, H), dtype=torch.float32, device=dev)

rows = []
# Wqkv (+ RoPE)
ours = timeit(lambda: ops.gemm(x, wqkv, epilogue=N.EPI_ROPE, pos=pos, cos=cos, sin=sin, hidden_size=H))
lib = timeit(lambda: x @ wqkv.T)
rows.append(("Wqkv [+RoPE fused]", 2.0 * M * 3 * H * H, ours, lib, "matmul only (RoPE would be extra kernels)"))
# Wi (+ GeGLU)
ours = timeit(lambda: ops.gemm(x, wi_il, epilogue=N.EPI_GEGLU))
lib_mm = timeit(lambda: x @ wi.T)
u = x @ wi.T
lib_act = timeit(lambda: torch.nn.functional.gelu(u[:, :I]) * u[:, I:])
rows.append(("Wi [+GeGLU fused]", 2.0 * M * 2 * I * H, ours, lib_mm + lib_act, f"matmul {lib_mm:.0f} us + gelu*gate {lib_act:.0f} us"))
# Wo2 (+ residual)
ours = timeit(lambda: ops.gemm(act, wo2, epilogue=N.EPI_RESIDUAL, out=h))
lib_mm = timeit(lambda: act @ wo2.T)
y = act @ wo2.T
lib_add = timeit(lambda: h.add_(y))
rows.append(("Wo2 [+fp32 residual fused]", 2.0 * M * H * I, ours, lib_mm + lib_add, f"matmul {lib_mm:.0f} us + fp32 add {lib_add:.0f} us"))
# Wo (+ residual)
ours = timeit(lambda: ops.gemm(attn, wo, epilogue=N.EPI_RESIDUAL, out=h))
lib_mm = timeit(lambda: attn @ wo.T)
y = attn @ wo.T
lib_add = timeit(lambda: h.add_(y))
rows.append(("Wo [+fp32 residual fused]", 2.0 * M * H * H, ours, lib_mm + lib_add, f"matmul {lib_mm:.0f} us + fp32 add {lib_add:.0f} us"))
print(f"M = {M} tokens, bf16, one B200 ({torch.cuda.get_device_name(0)})")
print("| GEMM | this repo (us, TFLOP/s) | cuBLAS path (us) | note |\n|---|---|---|---|")
for name, flops, o, l, note in rows:
    print(f"| {name} | {o:.0f} us, {flops / o / 1e6:.0f} | {l:.0f} us | {note} |")
