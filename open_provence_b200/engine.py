"""Host wrapper around one ``opv_engine`` handle: weight packing, workspace, launches.

PyTorch is used for device memory, streams and (elsewhere) ``torch.distributed`` only; all
arithmetic of the hot path runs in ``libopv_sm100.so``.
"""

from __future__ import annotations

import ctypes as C
from typing import Any, Mapping

import torch

from . import _native as N

_DTYPE_ALIASES = {
    "bf16": N.OPV_DTYPE_BF16,
    "bfloat16": N.OPV_DTYPE_BF16,
    torch.bfloat16: N.OPV_DTYPE_BF16,
    "fp32": N.OPV_DTYPE_F32,
    "float32": N.OPV_DTYPE_F32,
    "f32": N.OPV_DTYPE_F32,
    torch.float32: N.OPV_DTYPE_F32,
    # fp32 parity through the tensor-core pipeline: six bf16 tcgen05 passes per projection (include/opv.h)
    "fp32_tc": N.OPV_DTYPE_F32_TC,
    "f32_tc": N.OPV_DTYPE_F32_TC,
}


def split3_bf16(w: torch.Tensor) -> torch.Tensor:
    """fp32 [out, in] -> bf16 [3, out, in] with w = hi + mid + lo up to 2^-24 |w| (OPV_DTYPE_F32_TC weights)."""
    w = w.detach().to(torch.float32)
    hi = w.to(torch.bfloat16)
    r1 = w - hi.to(torch.float32)
    mid = r1.to(torch.bfloat16)
    lo = (r1 - mid.to(torch.float32)).to(torch.bfloat16)
    return torch.stack([hi, mid, lo]).contiguous()


def resolve_engine_dtype(dtype: Any) -> int:
    if dtype is None:
        return N.OPV_DTYPE_BF16
    key = dtype.lower() if isinstance(dtype, str) else dtype
    if key in (torch.float16, "fp16", "float16", "half"):
        raise NotImplementedError("the sm_100a engine computes in bf16 or fp32; fp16 is not implemented")
    if key not in _DTYPE_ALIASES:
        raise TypeError(f"Unsupported dtype for the sm_100a engine: {dtype!r}")
    return _DTYPE_ALIASES[key]


def layer_is_global(backbone_cfg: Mapping[str, Any], layer: int) -> bool:
    """configuration_modernbert.py:113-120 -- layer i is full attention iff i % every_n == 0."""
    layer_types = backbone_cfg.get("layer_types")
    if layer_types:
        return layer_types[layer] == "full_attention"
    return layer % int(backbone_cfg.get("global_attn_every_n_layers", 3)) == 0


def rope_thetas(backbone_cfg: Mapping[str, Any]) -> tuple[float, float]:
    """configuration_modernbert.py:77,141-148 (defaults 160000 global / 10000 local)."""
    rp = backbone_cfg.get("rope_parameters") or {}
    g = (rp.get("full_attention") or {}).get("rope_theta", backbone_cfg.get("global_rope_theta", 160000.0))
    l = (rp.get("sliding_attention") or {}).get("rope_theta", backbone_cfg.get("local_rope_theta", 10000.0))
    return float(g), float(l)


def rope_table(n_pos: int, head_dim: int, theta: float) -> tuple[torch.Tensor, torch.Tensor]:
    """cos/sin [n_pos, head_dim/2] fp32, computed the way HF does (HF:139-172): fp32 inv_freq, fp32 angles."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).to(dtype=torch.float) / head_dim))
    pos = torch.arange(n_pos, dtype=torch.float32)
    freqs = pos[:, None] * inv_freq[None, :]
    return freqs.cos().contiguous(), freqs.sin().contiguous()


def interleave_wi(wi: torch.Tensor, block: int = 128) -> torch.Tensor:
    """[2I, H] (input rows then gate rows, HF:90) -> blocks of `block` input rows followed by the matching
    gate rows, so one 256-wide GEMM tile holds both halves of 128 GeGLU features."""
    two_i, _ = wi.shape
    inter = two_i // 2
    idx = torch.arange(inter).view(inter // block, block)
    order = torch.cat([idx, idx + inter], dim=1).reshape(-1)
    return wi.index_select(0, order.to(wi.device)).contiguous()


class Engine:
    """One sm_100a engine bound to one CUDA device."""

    def __init__(
        self,
        backbone_cfg: Mapping[str, Any],
        state_dict: Mapping[str, torch.Tensor],
        *,
        device: torch.device | str = "cuda",
        dtype: Any = "bf16",
        num_labels: int = 1,
        fuse_epilogues: bool = True,
        max_positions: int | None = None,
    ) -> None:
        self.lib = N.load()
        if not torch.cuda.is_available():
            raise N.OpvError("the sm_100a engine needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError(f"the sm_100a engine runs on CUDA devices only, got {self.device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        model_type = backbone_cfg.get("model_type", "modernbert")
        if model_type != "modernbert":
            raise NotImplementedError(f"unsupported backbone architecture {model_type!r}: only ModernBERT is implemented")
        pooling = backbone_cfg.get("classifier_pooling", "cls")  # configuration_modernbert.py: "cls" | "mean"
        if pooling not in ("cls", "mean"):
            raise ValueError(f'classifier_pooling must be "cls" or "mean", got {pooling!r}')
        self.classifier_pooling = pooling
        for flag in ("attention_bias", "mlp_bias", "norm_bias", "classifier_bias"):
            if backbone_cfg.get(flag, False):
                raise NotImplementedError(f"ModernBERT option {flag}=True is not implemented")
        if backbone_cfg.get("hidden_activation", "gelu") != "gelu" or backbone_cfg.get("classifier_activation", "gelu") != "gelu":
            raise NotImplementedError("only exact-erf GELU activations are implemented")

        self.dtype_code = resolve_engine_dtype(dtype)
        self.op_dtype = torch.bfloat16 if self.dtype_code == N.OPV_DTYPE_BF16 else torch.float32
        self.split_gemm = self.dtype_code == N.OPV_DTYPE_F32_TC
        self.fused = bool(fuse_epilogues) and self.dtype_code == N.OPV_DTYPE_BF16
        self.hidden = int(backbone_cfg["hidden_size"])
        self.layers = int(backbone_cfg["num_hidden_layers"])
        self.heads = int(backbone_cfg["num_attention_heads"])
        self.inter = int(backbone_cfg["intermediate_size"])
        self.vocab = int(backbone_cfg["vocab_size"])
        self.num_labels = int(num_labels)
        self.local_window = int(backbone_cfg.get("local_attention", 128))
        self.eps = float(backbone_cfg.get("norm_eps", 1e-5))
        self.max_positions = int(max_positions or backbone_cfg.get("max_position_embeddings", 8192))
        self.global_flags = [layer_is_global(backbone_cfg, l) for l in range(self.layers)]
        if self.hidden != self.heads * 64:
            raise NotImplementedError(f"head_dim must be 64 (hidden_size={self.hidden}, heads={self.heads})")

        self._keep: list[torch.Tensor] = []  # device tensors the engine borrows pointers into
        self._handle = C.c_void_p()
        self._workspace: torch.Tensor | None = None
        self._pack(state_dict, backbone_cfg)

    # ------------------------------------------------------------------ weights
    def _dev(self, t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
        out = t.detach().to(device=self.device, dtype=dtype).contiguous()
        self._keep.append(out)
        return out

    def _pack(self, sd: Mapping[str, torch.Tensor], backbone_cfg: Mapping[str, Any]) -> None:
        """Map the reference's state-dict keys to the C weight structs."""
        p = "ranking_model."

        def get(name: str) -> torch.Tensor:
            if name not in sd:
                raise KeyError(f"missing weight {name!r} in checkpoint")
            return sd[name]

        f32, op = torch.float32, self.op_dtype
        w = N.OpvWeights()
        emb = get(p + "model.embeddings.tok_embeddings.weight")
        if tuple(emb.shape) != (self.vocab, self.hidden):
            raise ValueError(f"tok_embeddings has shape {tuple(emb.shape)}, config says {(self.vocab, self.hidden)}")
        w.d_tok_embeddings = self._dev(emb, op).data_ptr()
        w.d_emb_norm = self._dev(get(p + "model.embeddings.norm.weight"), f32).data_ptr()
        w.d_final_norm = self._dev(get(p + "model.final_norm.weight"), f32).data_ptr()
        w.d_head_dense = self._dev(get(p + "head.dense.weight"), f32).data_ptr()
        w.d_head_norm = self._dev(get(p + "head.norm.weight"), f32).data_ptr()
        cls_w = get(p + "classifier.weight")
        if cls_w.shape[0] != self.num_labels:
            raise ValueError(f"classifier has {cls_w.shape[0]} labels, config says {self.num_labels}")
        w.d_cls_weight = self._dev(cls_w, f32).data_ptr()
        w.d_cls_bias = self._dev(get(p + "classifier.bias"), f32).data_ptr()
        prune_w = get("pruning_head.classifier.weight")
        if tuple(prune_w.shape) != (2, self.hidden):
            raise NotImplementedError(f"pruning head must be Linear({self.hidden}, 2), got {tuple(prune_w.shape)}")
        w.d_prune_weight = self._dev(prune_w, f32).data_ptr()
        w.d_prune_bias = self._dev(get("pruning_head.classifier.bias"), f32).data_ptr()

        theta_g, theta_l = rope_thetas(backbone_cfg)
        cos_g, sin_g = rope_table(self.max_positions, 64, theta_g)
        cos_l, sin_l = rope_table(self.max_positions, 64, theta_l)
        w.d_rope_cos_global = self._dev(cos_g, f32).data_ptr()
        w.d_rope_sin_global = self._dev(sin_g, f32).data_ptr()
        w.d_rope_cos_local = self._dev(cos_l, f32).data_ptr()
        w.d_rope_sin_local = self._dev(sin_l, f32).data_ptr()

        layers = (N.OpvLayerWeights * self.layers)()
        for l in range(self.layers):
            lp = f"{p}model.layers.{l}."
            lw = layers[l]
            lw.d_attn_norm = self._dev(get(lp + "attn_norm.weight"), f32).data_ptr() if l > 0 else None
            def gw(t: torch.Tensor) -> int:  # GEMM weight: operand dtype, or hi | mid | lo bf16 planes
                if self.split_gemm:
                    return self._dev(split3_bf16(t), torch.bfloat16).data_ptr()
                return self._dev(t, op).data_ptr()

            lw.d_wqkv = gw(get(lp + "attn.Wqkv.weight"))
            lw.d_wo = gw(get(lp + "attn.Wo.weight"))
            lw.d_mlp_norm = self._dev(get(lp + "mlp_norm.weight"), f32).data_ptr()
            wi = get(lp + "mlp.Wi.weight")
            if tuple(wi.shape) != (2 * self.inter, self.hidden):
                raise ValueError(f"layer {l}: Wi has shape {tuple(wi.shape)}")
            if self.split_gemm:
                lw.d_wi = gw(wi)
            else:
                wi = wi.detach().to(dtype=op)
                lw.d_wi = self._dev(interleave_wi(wi) if self.fused else wi, op).data_ptr()
            lw.d_wo2 = gw(get(lp + "mlp.Wo.weight"))
        w.h_layers = C.cast(layers, C.POINTER(N.OpvLayerWeights))

        cfg = N.OpvConfig()
        cfg.abi_version = N.OPV_ABI_VERSION
        cfg.hidden_size = self.hidden
        cfg.num_layers = self.layers
        cfg.num_heads = self.heads
        cfg.intermediate_size = self.inter
        cfg.vocab_size = self.vocab
        cfg.num_labels = self.num_labels
        cfg.local_window = self.local_window
        cfg.max_positions = self.max_positions
        cfg.norm_eps = self.eps
        cfg.dtype = self.dtype_code
        cfg.fuse_epilogues = 1 if self.fused else 0
        cfg.classifier_pooling = 1 if self.classifier_pooling == "mean" else 0
        if self.layers > N.OPV_MAX_LAYERS:
            raise NotImplementedError(f"at most {N.OPV_MAX_LAYERS} layers are supported")
        for l, flag in enumerate(self.global_flags):
            cfg.layer_is_global[l] = 1 if flag else 0
        with torch.cuda.device(self.device):
            N.check(self.lib.opv_create(C.byref(cfg), C.byref(w), self.device.index, C.byref(self._handle)), "opv_create")

    def close(self) -> None:
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self.lib.opv_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self) -> None:  # pragma: no cover - interpreter shutdown ordering
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ profiling
    def profile(self, on: bool) -> None:
        """Bracket every forward launch with CUDA events on the launch stream (see ``profile_collect``)."""
        N.check(self.lib.opv_profile_enable(self._handle, 1 if on else 0), "opv_profile_enable")

    def profile_collect(self) -> dict[str, dict[str, float]]:
        """{class: {"ms": total device ms, "launches": n}} since the last collect (synchronises)."""
        n = len(N.PROF_CLASSES)
        ms = (C.c_float * n)()
        launches = (C.c_int32 * n)()
        N.check(self.lib.opv_profile_collect(self._handle, ms, launches, n), "opv_profile_collect")
        return {name: {"ms": float(ms[i]), "launches": int(launches[i])} for i, name in enumerate(N.PROF_CLASSES)}

    def launch_count(self) -> int:
        return int(self.lib.opv_launch_count(self._handle))

    # ------------------------------------------------------------------ launches
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _ensure_workspace(self, n_tokens: int, n_seqs: int) -> torch.Tensor:
        need = int(self.lib.opv_workspace_bytes(self._handle, n_tokens, n_seqs))
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = None
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace

    def forward_packed(
        self, ids: torch.Tensor, cu_seqlens: torch.Tensor, max_seqlen: int
    ) -> tuple[torch.Tensor, torch.Tensor]:
        """ids int32 [T], cu_seqlens int32 [n+1] (both on the engine's device) ->
        (pruning_logits fp32 [T, 2], ranking_logits fp32 [n, num_labels]).  Asynchronous."""
        if ids.dtype != torch.int32 or cu_seqlens.dtype != torch.int32:
            raise TypeError("ids and cu_seqlens must be int32")
        if ids.device != self.device or cu_seqlens.device != self.device:
            raise ValueError("ids and cu_seqlens must live on the engine's device")
        n_tokens = int(ids.numel())
        n_seqs = int(cu_seqlens.numel()) - 1
        prune = torch.empty((n_tokens, 2), dtype=torch.float32, device=self.device)
        rank = torch.empty((max(n_seqs, 0), self.num_labels), dtype=torch.float32, device=self.device)
        if n_tokens == 0 or n_seqs <= 0:
            return prune, rank
        ws = self._ensure_workspace(n_tokens, n_seqs)
        with torch.cuda.device(self.device):
            rc = self.lib.opv_forward_packed(
                self._handle, ids.data_ptr(), cu_seqlens.data_ptr(), n_seqs, n_tokens, int(max_seqlen),
                prune.data_ptr(), rank.data_ptr(), ws.data_ptr(), ws.numel(), self._stream(),
            )
        N.check(rc, "opv_forward_packed")
        self._last_forward = (ws, n_seqs, n_tokens)
        return prune, rank

    def forward_status(self) -> int:
        """Number of sequences of the last ``forward_packed`` call whose ``cu_seqlens`` boundaries broke the contract
        (0 = well formed).  The forward itself clamps them and stays memory-safe; this synchronises the stream."""
        last = getattr(self, "_last_forward", None)
        if last is None:
            return 0
        ws, n_seqs, n_tokens = last
        bad = N.C.c_int32(0)
        with torch.cuda.device(self.device):
            rc = self.lib.opv_forward_status(self._handle, ws.data_ptr(), n_seqs, n_tokens, N.C.byref(bad), self._stream())
        N.check(rc, "opv_forward_status")
        return int(bad.value)

    def fragment_means(
        self, prune_logits: torch.Tensor, frag_ranges: torch.Tensor, rank_logits: torch.Tensor
    ) -> tuple[torch.Tensor, torch.Tensor]:
        """Per-fragment mean keep-probability and per-block sigmoid rank score (device, async)."""
        n_frags = int(frag_ranges.shape[0])
        n_seqs = int(rank_logits.shape[0])
        frag_mean = torch.empty(n_frags, dtype=torch.float32, device=self.device)
        score = torch.empty(n_seqs, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.opv_fragment_means(
                prune_logits.data_ptr(), int(prune_logits.shape[0]), frag_ranges.data_ptr(), n_frags,
                frag_mean.data_ptr(), rank_logits.data_ptr(), n_seqs, int(rank_logits.shape[1]) if n_seqs else 1,
                score.data_ptr(), self._stream(),
            )
        N.check(rc, "opv_fragment_means")
        return frag_mean, score

    def token_keep_probs(self, prune_logits: torch.Tensor) -> torch.Tensor:
        """softmax(prune_logits)[:, 1] per packed token (device, async) -- encoder.py:429-430."""
        n = int(prune_logits.shape[0])
        prob = torch.empty(n, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.opv_token_keep_probs(prune_logits.data_ptr(), n, prob.data_ptr(), self._stream())
        N.check(rc, "opv_token_keep_probs")
        return prob

    def sentence_prune(
        self,
        frag_mean: torch.Tensor,
        sent_offsets: torch.Tensor,
        sent_frag_index: torch.Tensor,
        threshold: float,
        guard: float = 1e-5,
    ) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        n_sents = int(sent_offsets.numel()) - 1
        prob = torch.empty(max(n_sents, 0), dtype=torch.float64, device=self.device)
        keep = torch.empty(max(n_sents, 0), dtype=torch.uint8, device=self.device)
        near = torch.empty(max(n_sents, 0), dtype=torch.uint8, device=self.device)
        if n_sents > 0:
            with torch.cuda.device(self.device):
                rc = self.lib.opv_sentence_prune(
                    frag_mean.data_ptr(), sent_offsets.data_ptr(), sent_frag_index.data_ptr(), n_sents,
                    float(threshold), float(guard), prob.data_ptr(), keep.data_ptr(), near.data_ptr(), self._stream(),
                )
            N.check(rc, "opv_sentence_prune")
        return prob, keep, near

    # ------------------------------------------------------------------ host-buffer entry point
    def score_packed_host(
        self,
        ids: torch.Tensor,
        cu_seqlens: torch.Tensor,
        max_seqlen: int,
        frag_ranges: torch.Tensor,
        sent_offsets: torch.Tensor,
        sent_frag_index: torch.Tensor,
        threshold: float,
        guard: float = 1e-5,
    ) -> dict[str, torch.Tensor]:
        """The whole hot path on HOST buffers (CPU int32 tensors, ideally pinned): H2D copies, forward,
        score conversion, per-sentence prune, D2H of the results.  Synchronises before returning.

        Returns CPU tensors: ``rank_score`` fp32 [n_blocks], ``sent_prob`` fp64 [n_sents],
        ``keep`` uint8 [n_sents], ``near`` uint8 [n_sents] (|prob - threshold| <= guard).
        """
        dev = self.device
        d_ids = ids.to(dev, non_blocking=True)
        d_cu = cu_seqlens.to(dev, non_blocking=True)
        d_ranges = frag_ranges.to(dev, non_blocking=True)
        d_off = sent_offsets.to(dev, non_blocking=True)
        d_idx = sent_frag_index.to(dev, non_blocking=True)
        prune, rank = self.forward_packed(d_ids, d_cu, max_seqlen)
        frag_mean, score = self.fragment_means(prune, d_ranges, rank)
        prob, keep, near = self.sentence_prune(frag_mean, d_off, d_idx, threshold, guard)
        out = {
            "rank_score": score.to("cpu", non_blocking=True),
            "sent_prob": prob.to("cpu", non_blocking=True),
            "keep": keep.to("cpu", non_blocking=True),
            "near": near.to("cpu", non_blocking=True),
        }
        torch.cuda.current_stream(dev).synchronize()
        return out
