"""Each CUDA kernel against a plain PyTorch fp32 reference of the same op (through the C ABI)."""

from __future__ import annotations

import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from open_provence_b200 import _native as N  # noqa: E402
from open_provence_b200 import ops  # noqa: E402
from open_provence_b200.engine import interleave_wi, rope_table  # noqa: E402

DEV = "cuda"


def _rand_bf16(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(DEV)


def _bf16_close(out: torch.Tensor, ref: torch.Tensor, what: str):
    """out was rounded to bf16 once from an fp32 accumulator: |err| <= 2^-8 |ref| + small."""
    err = (out.float() - ref).abs()
    bound = ref.abs() * 2.0**-8 + 2e-3 * ref.abs().max().clamp(min=1e-6) * 2.0**-4
    bad = (err > bound).sum().item()
    assert bad == 0, f"{what}: {bad} elements out of tolerance, max err {err.max().item():.4e}, ref max {ref.abs().max().item():.3e}"


@pytest.fixture(params=[1, 0], ids=["pair", "single"])
def gemm_kernel(request):
    """Run the bf16 GEMM tests on both kernels: CTA pair (cta_group::2, default) and single CTA."""
    ops.set_option("gemm_pair", request.param)
    yield request.param
    ops.set_option("gemm_pair", 1)


GEMM_SHAPES = [
    (128, 128, 64),
    (128, 256, 64),
    (300, 384, 128),
    (1000, 512, 512),
    (4113, 1536, 512),
    (257, 256, 2048),
    (20000, 768, 256),
    (19201, 512, 512),  # >= one 256-row block per CTA pair, K <= 512: the A-stationary pair kernel (ragged last block)
    (1, 128, 64),
]


@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
def test_gemm_bf16_store(m, n, k, gemm_kernel):
    a = _rand_bf16((m, k), 1)
    w = _rand_bf16((n, k), 2, 0.05)
    out = ops.gemm(a, w)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().T
    _bf16_close(out, ref, f"gemm store {m}x{n}x{k}")


@pytest.mark.parametrize("m,n,k", [(300, 128, 128), (1000, 512, 2048), (4113, 256, 512)])
def test_gemm_bf16_residual(m, n, k, gemm_kernel):
    a = _rand_bf16((m, k), 3)
    w = _rand_bf16((n, k), 4, 0.05)
    g = torch.Generator().manual_seed(5)
    r0 = torch.randn((m, n), generator=g).to(DEV)
    r = r0.clone()
    ops.gemm(a, w, epilogue=N.EPI_RESIDUAL, out=r)
    torch.cuda.synchronize()
    ref = r0 + a.float() @ w.float().T
    err = (r - ref).abs().max().item()
    assert err < 1e-4 * max(1.0, ref.abs().max().item()), f"residual epilogue max err {err:.3e}"


@pytest.mark.parametrize("m,n,k", [(300, 256, 256), (1000, 512, 512), (4113, 256, 1024), (40001, 512, 512),
                                   (76033, 256, 256), (1, 512, 64)])
def test_gemm_residual_layernorm_rows(m, n, k):
    """Full-row residual GEMM with the following LayerNorm from TMEM (``opv_op_gemm_residual_ln``, gemm_rowln.cuh):
    the new residual against fp32 torch, X against ``F.layer_norm`` of the residual the kernel itself wrote (so the
    bound is the bf16 rounding of X, 2^-8 relative), ragged last row block, more row blocks than CTA pairs."""
    a = _rand_bf16((m, k), 31)
    w = _rand_bf16((n, k), 32, 0.05)
    g = torch.Generator().manual_seed(33)
    r0 = (torch.randn((m, n), generator=g) * 3.0 + torch.randn((m, 1), generator=g)).to(DEV)  # rows with a mean
    gamma = (1.0 + 0.2 * torch.randn((n,), generator=g)).to(DEV)
    r = r0.clone()
    x = ops.gemm_residual_ln(a, w, r, gamma, 1e-5)
    torch.cuda.synchronize()
    ref_r = r0 + a.float() @ w.float().T
    err = (r - ref_r).abs().max().item()
    assert err < 1e-4 * max(1.0, ref_r.abs().max().item()), f"residual max err {err:.3e}"
    ref_x = torch.nn.functional.layer_norm(r, (n,), gamma, None, 1e-5)
    _bf16_close(x, ref_x, f"row LayerNorm {m}x{n}x{k}")
    # and the statistics themselves, against fp64
    r64 = r.double()
    ref64 = ((r64 - r64.mean(1, keepdim=True)) / torch.sqrt(r64.var(1, unbiased=False, keepdim=True) + 1e-5) * gamma.double())
    assert (x.double() - ref64).abs().max().item() < 2.0**-7 * ref64.abs().max().item()


def _rope_ref(qkv: torch.Tensor, pos: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, hidden: int) -> torch.Tensor:
    t = qkv.shape[0]
    heads = hidden // 64
    x = qkv.float().clone().view(t, 3, heads, 64)
    c = torch.cat([cos[pos.long()], cos[pos.long()]], dim=-1)[:, None, :]
    s = torch.cat([sin[pos.long()], sin[pos.long()]], dim=-1)[:, None, :]
    for part in (0, 1):
        v = x[:, part]
        rot = torch.cat([-v[..., 32:], v[..., :32]], dim=-1)
        x[:, part] = v * c + rot * s
    return x.view(t, 3 * hidden)


@pytest.mark.parametrize("m,hidden", [(333, 128), (1500, 256), (2000, 512), (76033, 256), (19300, 512)])  # last two: row-grouped tile order
def test_gemm_bf16_rope(m, hidden, gemm_kernel):
    a = _rand_bf16((m, hidden), 6)
    w = _rand_bf16((3 * hidden, hidden), 7, 0.05)
    cos, sin = rope_table(4096, 64, 160000.0)
    cos, sin = cos.to(DEV), sin.to(DEV)
    g = torch.Generator().manual_seed(8)
    pos = torch.randint(0, 4096, (m,), generator=g, dtype=torch.int32).to(DEV)
    out = ops.gemm(a, w, epilogue=N.EPI_ROPE, pos=pos, cos=cos, sin=sin, hidden_size=hidden)
    torch.cuda.synchronize()
    ref = _rope_ref(a.float() @ w.float().T, pos, cos, sin, hidden)
    _bf16_close(out, ref, f"gemm rope m={m} H={hidden}")


@pytest.mark.parametrize("m,k,inter", [(300, 128, 128), (1000, 512, 2048), (4113, 256, 1152), (19201, 512, 2048),
                                       (40000, 256, 1024)])  # the last two: A-stationary pair kernel
def test_gemm_bf16_geglu(m, k, inter, gemm_kernel):
    a = _rand_bf16((m, k), 9)
    wi = _rand_bf16((2 * inter, k), 10, 0.08)
    out = ops.gemm(a, interleave_wi(wi), epilogue=N.EPI_GEGLU)
    torch.cuda.synchronize()
    u = a.float() @ wi.float().T
    ref = torch.nn.functional.gelu(u[:, :inter]) * u[:, inter:]
    assert out.shape == (m, inter)
    _bf16_close(out, ref, f"gemm geglu m={m} k={k} I={inter}")


@pytest.mark.parametrize("m,n,k", [(70, 64, 16), (300, 384, 128), (1000, 512, 2048)])
def test_gemm_f32(m, n, k):
    g = torch.Generator().manual_seed(11)
    a = torch.randn((m, k), generator=g).to(DEV)
    w = (torch.randn((n, k), generator=g) * 0.05).to(DEV)
    out = ops.gemm(a, w)
    r0 = torch.randn((m, n), generator=g).to(DEV)
    r = r0.clone()
    ops.gemm(a, w, epilogue=N.EPI_RESIDUAL, out=r)
    torch.cuda.synchronize()
    ref = (a.double() @ w.double().T)
    assert (out.double() - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    assert (r.double() - (r0.double() + ref)).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("hidden", [128, 256, 512, 768, 1024])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_layernorm(hidden, out_dtype):
    g = torch.Generator().manual_seed(12)
    h = (torch.randn((777, hidden), generator=g) * 3 + 0.5).to(DEV)
    w = (1 + 0.1 * torch.randn(hidden, generator=g)).to(DEV)
    out = ops.layernorm(h, w, 1e-5, out_dtype)
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(h.double(), (hidden,), w.double(), None, 1e-5)
    tol = 2e-6 if out_dtype == torch.float32 else 2.0**-8
    err = ((out.double() - ref).abs() / (ref.abs() + 1.0)).max().item()
    assert err < tol * 4, f"layernorm H={hidden} {out_dtype}: rel err {err:.3e}"


@pytest.mark.parametrize("emb_dtype", [torch.float32, torch.bfloat16])
def test_embed_ln(emb_dtype):
    hidden, vocab = 256, 1000
    g = torch.Generator().manual_seed(13)
    emb = (torch.randn((vocab, hidden), generator=g) * 0.02).to(emb_dtype).to(DEV)
    w = (1 + 0.1 * torch.randn(hidden, generator=g)).to(DEV)
    ids = torch.randint(0, vocab, (999,), generator=g, dtype=torch.int32).to(DEV)
    h, x = ops.embed_ln(ids, emb, w, 1e-5)
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(emb[ids.long()].double(), (hidden,), w.double(), None, 1e-5)
    assert (h.double() - ref).abs().max().item() < 1e-5
    tol = 1e-5 if emb_dtype == torch.float32 else 4e-2
    assert (x.double() - ref).abs().max().item() < tol


def _attention_ref(qkv: torch.Tensor, lengths: list[int], heads: int, half_window: int) -> torch.Tensor:
    hidden = heads * 64
    outs = []
    start = 0
    for n in lengths:
        blk = qkv[start : start + n].double().view(n, 3, heads, 64)
        q, k, v = blk[:, 0], blk[:, 1], blk[:, 2]
        s = torch.einsum("ihd,jhd->hij", q, k) * 0.125
        if half_window >= 0:
            idx = torch.arange(n, device=qkv.device)
            band = (idx[:, None] - idx[None, :]).abs() <= half_window
            s = s.masked_fill(~band[None], float("-inf"))
        p = torch.softmax(s, dim=-1)
        outs.append(torch.einsum("hij,jhd->ihd", p, v).reshape(n, hidden))
        start += n
    return torch.cat(outs, dim=0)


ATT_LENGTHS = [1, 2, 63, 64, 65, 127, 128, 129, 130, 200, 257, 513]


@pytest.mark.parametrize("half_window", [-1, 64])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attention(dtype, half_window):
    heads = 2
    total = sum(ATT_LENGTHS)
    g = torch.Generator().manual_seed(14)
    qkv = torch.randn((total, 3 * heads * 64), generator=g).to(dtype).to(DEV)
    cu = torch.tensor([0] + list(np.cumsum(ATT_LENGTHS)), dtype=torch.int32, device=DEV)
    out = ops.attention(qkv, cu, max(ATT_LENGTHS), heads, half_window)
    torch.cuda.synchronize()
    ref = _attention_ref(qkv, ATT_LENGTHS, heads, half_window)
    err = (out.double() - ref).abs().max().item()
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    assert err < tol, f"attention {dtype} window={half_window}: max err {err:.3e}"
    assert torch.isfinite(out.float()).all()


@pytest.mark.parametrize("impl", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("half_window", [-1, 64, 8, 100, 192])
def test_attention_bf16_impls(impl, half_window):
    """All three bf16 kernels (mma.sync v1, tcgen05 with P in TMEM, tcgen05 with P in smem) on ragged lengths
    that cross the 128-row tile and 128-key block boundaries, plus a multi-block sequence."""
    heads = 3
    lengths = ATT_LENGTHS + [1100]
    total = sum(lengths)
    g = torch.Generator().manual_seed(15)
    qkv = (torch.randn((total, 3 * heads * 64), generator=g) * 1.5).to(torch.bfloat16).to(DEV)
    cu = torch.tensor([0] + list(np.cumsum(lengths)), dtype=torch.int32, device=DEV)
    ops.set_option("attention_impl", impl)
    try:
        out = ops.attention(qkv, cu, max(lengths), heads, half_window)
        torch.cuda.synchronize()
    finally:
        ops.set_option("attention_impl", 1)
    ref = _attention_ref(qkv, lengths, heads, half_window)
    err = (out.double() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all()
    # bf16 output rounding alone is 2^-9 relative; scale the bound with the magnitude of the outputs
    tol = 6e-3 * max(1.0, ref.abs().max().item())
    assert err < tol, f"attention impl={impl} window={half_window}: max err {err:.3e} (tol {tol:.3e})"


@pytest.mark.parametrize("lengths", [[1, 2, 63, 64, 65, 127, 128, 129, 130, 200, 257, 513], [1100, 2048, 511, 512, 640], [4097]])
def test_attention_global_four_q_tiles(lengths):
    """Four-Q-tile kernel (attention_impl 7): 512-row work units, 64-key blocks, P aliased onto S; ragged lengths that
    leave 1..4 tiles active and end inside a key block."""
    heads = 3
    g = torch.Generator().manual_seed(23)
    qkv = (torch.randn((sum(lengths), 3 * heads * 64), generator=g) * 1.5).to(torch.bfloat16).to(DEV)
    cu = torch.tensor([0] + list(np.cumsum(lengths)), dtype=torch.int32, device=DEV)
    ops.set_option("attention_impl", 7)
    try:
        out = ops.attention(qkv, cu, max(lengths), heads, -1)
        torch.cuda.synchronize()
        with pytest.raises(NotImplementedError, match="global attention only"):
            ops.attention(qkv, cu, max(lengths), heads, 64)
    finally:
        ops.set_option("attention_impl", 1)
    ref = _attention_ref(qkv, lengths, heads, -1)
    err = (out.double() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all()
    tol = 6e-3 * max(1.0, ref.abs().max().item())
    assert err < tol, f"four-Q-tile attention lengths={lengths}: max err {err:.3e} (tol {tol:.3e})"


@pytest.mark.parametrize("half_window", [64, 8, 1, 0, 33])
def test_attention_local_onepass(half_window):
    """One-pass sliding-window kernel (window <= 128): the whole band of a 128-query tile in one 256-key score tile."""
    heads = 3
    lengths = ATT_LENGTHS + [1100, 2048]
    total = sum(lengths)
    g = torch.Generator().manual_seed(21)
    qkv = (torch.randn((total, 3 * heads * 64), generator=g) * 1.5).to(torch.bfloat16).to(DEV)
    cu = torch.tensor([0] + list(np.cumsum(lengths)), dtype=torch.int32, device=DEV)
    ops.set_option("attention_impl", 6)
    try:
        out = ops.attention(qkv, cu, max(lengths), heads, half_window)
        torch.cuda.synchronize()
        with pytest.raises(NotImplementedError, match="half_window <= 64"):
            ops.attention(qkv, cu, max(lengths), heads, 65)
    finally:
        ops.set_option("attention_impl", 1)
    ref = _attention_ref(qkv, lengths, heads, half_window)
    err = (out.double() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all()
    tol = 6e-3 * max(1.0, ref.abs().max().item())
    assert err < tol, f"one-pass local attention window={half_window}: max err {err:.3e} (tol {tol:.3e})"


def test_rope_and_geglu_unfused():
    hidden, m, inter = 128, 500, 256
    cos, sin = rope_table(1024, 64, 10000.0)
    cos, sin = cos.to(DEV), sin.to(DEV)
    g = torch.Generator().manual_seed(15)
    pos = torch.randint(0, 1024, (m,), generator=g, dtype=torch.int32).to(DEV)
    qkv = torch.randn((m, 3 * hidden), generator=g).to(DEV)
    ref = _rope_ref(qkv, pos, cos, sin, hidden)
    out = ops.rope_(qkv.clone(), pos, cos, sin, hidden)
    u = torch.randn((m, 2 * inter), generator=g).to(DEV)
    act = ops.geglu(u)
    torch.cuda.synchronize()
    assert (out - ref).abs().max().item() < 1e-5
    ref_act = torch.nn.functional.gelu(u[:, :inter].double()) * u[:, inter:].double()
    assert (act.double() - ref_act).abs().max().item() < 1e-5


def test_positions():
    lengths = [3, 1, 700, 64]
    cu = torch.tensor([0] + list(np.cumsum(lengths)), dtype=torch.int32, device=DEV)
    pos = ops.positions(cu, sum(lengths))
    torch.cuda.synchronize()
    ref = torch.cat([torch.arange(n) for n in lengths]).to(torch.int32)
    assert torch.equal(pos.cpu(), ref)


def test_library_reports_errors():
    a = _rand_bf16((128, 96), 1)  # K not a multiple of 64
    w = _rand_bf16((128, 96), 2)
    with pytest.raises(NotImplementedError, match="multiple of 64"):
        ops.gemm(a, w)
    assert math.isfinite(1.0)
