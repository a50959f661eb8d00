"""Single-kernel entry points of libopv_sm100.so on torch CUDA tensors (unit tests, profiling).

These call the same kernels the engine's forward launches; they exist so each kernel can be checked
against a plain fp32 reference in isolation.  No CPU fallback.
"""

from __future__ import annotations

import torch

from . import _native as N


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _code(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return N.OPV_DTYPE_BF16
    if t.dtype == torch.float32:
        return N.OPV_DTYPE_F32
    raise TypeError(f"operand dtype must be bf16 or fp32, got {t.dtype}")


def _ptr(t: torch.Tensor | None):
    return None if t is None else t.data_ptr()


def gemm(
    a: torch.Tensor,
    w: torch.Tensor,
    *,
    epilogue: int = N.EPI_STORE,
    out: torch.Tensor | None = None,
    pos: torch.Tensor | None = None,
    cos: torch.Tensor | None = None,
    sin: torch.Tensor | None = None,
    hidden_size: int = 0,
) -> torch.Tensor:
    """epilogue(A[M,K] . W[N,K]^T).  RESIDUAL accumulates into ``out`` (fp32 [M,N])."""
    lib = N.load()
    assert a.is_cuda and w.is_cuda and a.is_contiguous() and w.is_contiguous()
    m, k = a.shape
    n = w.shape[0]
    if epilogue == N.EPI_RESIDUAL:
        assert out is not None and out.dtype == torch.float32 and out.shape == (m, n)
    elif epilogue == N.EPI_GEGLU:
        out = torch.empty((m, n // 2), dtype=a.dtype, device=a.device) if out is None else out
    else:
        out = torch.empty((m, n), dtype=a.dtype, device=a.device) if out is None else out
    with torch.cuda.device(a.device):
        rc = lib.opv_op_gemm(_code(a), epilogue, a.data_ptr(), w.data_ptr(), out.data_ptr(), m, n, k, _ptr(pos),
                             _ptr(cos), _ptr(sin), hidden_size, _stream(a))
    N.check(rc, "opv_op_gemm")
    return out


def gemm_residual_ln(a: torch.Tensor, w: torch.Tensor, r: torch.Tensor, ln_weight: torch.Tensor, eps: float) -> torch.Tensor:
    """``r += a @ w.T`` in place (fp32) and returns ``LayerNorm(r) * ln_weight`` as bf16, from one kernel
    (``opv_op_gemm_residual_ln``; hidden size 256 or 512)."""
    lib = N.load()
    assert a.is_cuda and a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.is_contiguous() and w.is_contiguous()
    m, k = a.shape
    n = w.shape[0]
    assert r.dtype == torch.float32 and r.shape == (m, n) and r.is_contiguous()
    x = torch.empty((m, n), dtype=torch.bfloat16, device=a.device)
    with torch.cuda.device(a.device):
        rc = lib.opv_op_gemm_residual_ln(a.data_ptr(), w.data_ptr(), r.data_ptr(), x.data_ptr(), ln_weight.data_ptr(),
                                         float(eps), m, n, k, _stream(a))
    N.check(rc, "opv_op_gemm_residual_ln")
    return x


def layernorm(h: torch.Tensor, weight: torch.Tensor, eps: float, out_dtype: torch.dtype) -> torch.Tensor:
    lib = N.load()
    m, hidden = h.shape
    out = torch.empty((m, hidden), dtype=out_dtype, device=h.device)
    with torch.cuda.device(h.device):
        rc = lib.opv_op_layernorm(_code(out), h.data_ptr(), weight.data_ptr(), out.data_ptr(), m, hidden, eps, _stream(h))
    N.check(rc, "opv_op_layernorm")
    return out


def embed_ln(ids: torch.Tensor, emb: torch.Tensor, weight: torch.Tensor, eps: float) -> tuple[torch.Tensor, torch.Tensor]:
    lib = N.load()
    m = ids.numel()
    vocab, hidden = emb.shape
    h = torch.empty((m, hidden), dtype=torch.float32, device=emb.device)
    x = torch.empty((m, hidden), dtype=emb.dtype, device=emb.device)
    with torch.cuda.device(emb.device):
        rc = lib.opv_op_embed_ln(_code(emb), ids.data_ptr(), emb.data_ptr(), weight.data_ptr(), h.data_ptr(),
                                 x.data_ptr(), m, hidden, vocab, eps, _stream(emb))
    N.check(rc, "opv_op_embed_ln")
    return h, x


def attention(qkv: torch.Tensor, cu_seqlens: torch.Tensor, max_seqlen: int, num_heads: int, half_window: int) -> torch.Tensor:
    """qkv [T, 3H] (RoPE applied) -> [T, H]; half_window < 0 = global attention."""
    lib = N.load()
    t = qkv.shape[0]
    out = torch.empty((t, num_heads * 64), dtype=qkv.dtype, device=qkv.device)
    with torch.cuda.device(qkv.device):
        rc = lib.opv_op_attention(_code(qkv), qkv.data_ptr(), out.data_ptr(), cu_seqlens.data_ptr(),
                                  cu_seqlens.numel() - 1, t, max_seqlen, num_heads, half_window, _stream(qkv))
    N.check(rc, "opv_op_attention")
    return out


def rope_(qkv: torch.Tensor, pos: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, hidden: int) -> torch.Tensor:
    lib = N.load()
    with torch.cuda.device(qkv.device):
        rc = lib.opv_op_rope(_code(qkv), qkv.data_ptr(), pos.data_ptr(), cos.data_ptr(), sin.data_ptr(), qkv.shape[0],
                             hidden, _stream(qkv))
    N.check(rc, "opv_op_rope")
    return qkv


def geglu(u: torch.Tensor) -> torch.Tensor:
    lib = N.load()
    m, two_i = u.shape
    act = torch.empty((m, two_i // 2), dtype=u.dtype, device=u.device)
    with torch.cuda.device(u.device):
        rc = lib.opv_op_geglu(_code(u), u.data_ptr(), act.data_ptr(), m, two_i // 2, _stream(u))
    N.check(rc, "opv_op_geglu")
    return act


def positions(cu_seqlens: torch.Tensor, n_tokens: int) -> torch.Tensor:
    lib = N.load()
    pos = torch.empty(n_tokens, dtype=torch.int32, device=cu_seqlens.device)
    with torch.cuda.device(cu_seqlens.device):
        rc = lib.opv_op_positions(cu_seqlens.data_ptr(), cu_seqlens.numel() - 1, pos.data_ptr(), _stream(cu_seqlens))
    N.check(rc, "opv_op_positions")
    return pos


def set_option(name: str, value: int) -> None:
    """Process-wide tuning switch of the library (see ``opv_set_option`` in include/opv.h)."""
    N.check(N.load().opv_set_option(name.encode(), int(value)), "opv_set_option")
