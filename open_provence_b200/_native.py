"""ctypes binding of ``libopv_sm100.so`` (the C ABI declared in ``include/opv.h``).

There is NO fallback: if the shared library is missing or a call fails, this module raises.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libopv_sm100.so"

OPV_ABI_VERSION = 2
OPV_MAX_LAYERS = 64
OPV_DTYPE_BF16 = 0
OPV_DTYPE_F32 = 1
OPV_DTYPE_F32_TC = 2
EPI_STORE, EPI_ROPE, EPI_RESIDUAL, EPI_GEGLU = 0, 1, 2, 3

PROF_CLASSES = (
    "misc", "embed_ln", "layernorm", "gemm_qkv_rope", "attention_global", "attention_local", "gemm_wo_residual",
    "gemm_wi_geglu", "gemm_wo2_residual", "heads",
)

OPV_ERR_INVALID_ARGUMENT = -1
OPV_ERR_UNSUPPORTED = -2
OPV_ERR_CUDA = -3
OPV_ERR_WORKSPACE = -4


class OpvConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("hidden_size", C.c_int32),
        ("num_layers", C.c_int32),
        ("num_heads", C.c_int32),
        ("intermediate_size", C.c_int32),
        ("vocab_size", C.c_int32),
        ("num_labels", C.c_int32),
        ("local_window", C.c_int32),
        ("max_positions", C.c_int32),
        ("norm_eps", C.c_float),
        ("dtype", C.c_int32),
        ("fuse_epilogues", C.c_int32),
        ("classifier_pooling", C.c_int32),
        ("layer_is_global", C.c_uint8 * OPV_MAX_LAYERS),
    ]


class OpvLayerWeights(C.Structure):
    _fields_ = [
        ("d_attn_norm", C.c_void_p),
        ("d_wqkv", C.c_void_p),
        ("d_wo", C.c_void_p),
        ("d_mlp_norm", C.c_void_p),
        ("d_wi", C.c_void_p),
        ("d_wo2", C.c_void_p),
    ]


class OpvWeights(C.Structure):
    _fields_ = [
        ("d_tok_embeddings", C.c_void_p),
        ("d_emb_norm", C.c_void_p),
        ("d_final_norm", C.c_void_p),
        ("d_head_dense", C.c_void_p),
        ("d_head_norm", C.c_void_p),
        ("d_cls_weight", C.c_void_p),
        ("d_cls_bias", C.c_void_p),
        ("d_prune_weight", C.c_void_p),
        ("d_prune_bias", C.c_void_p),
        ("d_rope_cos_global", C.c_void_p),
        ("d_rope_sin_global", C.c_void_p),
        ("d_rope_cos_local", C.c_void_p),
        ("d_rope_sin_local", C.c_void_p),
        ("h_layers", C.POINTER(OpvLayerWeights)),
    ]


class OpvPackInput(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("max_length", C.c_int32),
        ("max_fragment_tokens", C.c_int32),
        ("keep_sentence_boundaries", C.c_int32),
        ("sep_len", C.c_int32),
        ("n_contexts", C.c_int32),
        ("n_queries", C.c_int32),
        ("vocab_size", C.c_int32),
        ("n_head", C.c_int32),
        ("n_mid", C.c_int32),
        ("n_tail", C.c_int32),
        ("h_head", C.c_void_p),
        ("h_mid", C.c_void_p),
        ("h_tail", C.c_void_p),
        ("h_tokens", C.c_void_p),
        ("h_sent_offsets", C.c_void_p),
        ("h_ctx_sent_offsets", C.c_void_p),
        ("h_ctx_query", C.c_void_p),
        ("h_ctx_prefix", C.c_void_p),
        ("h_query_tokens", C.c_void_p),
        ("h_query_offsets", C.c_void_p),
        ("h_token_visible", C.c_void_p),
        ("h_frag_drop", C.c_void_p),
    ]


class OpvPackView(C.Structure):
    _fields_ = [
        ("needs_decode", C.c_int32),
        ("n_raw_fragments", C.c_int64),
        ("n_uncertain", C.c_int64),
        ("h_raw_uncertain", C.c_void_p),
        ("h_raw_start", C.c_void_p),
        ("h_raw_len", C.c_void_p),
        ("n_blocks", C.c_int64),
        ("n_tokens", C.c_int64),
        ("n_slots", C.c_int64),
        ("n_sentences", C.c_int64),
        ("n_contexts", C.c_int64),
        ("h_ids", C.c_void_p),
        ("h_block_offsets", C.c_void_p),
        ("h_block_context", C.c_void_p),
        ("h_frag_block", C.c_void_p),
        ("h_frag_local", C.c_void_p),
        ("h_sent_slot_offsets", C.c_void_p),
        ("h_sent_slot_index", C.c_void_p),
        ("h_ctx_block_offsets", C.c_void_p),
    ]


# name -> (restype, argtypes); every symbol include/opv.h declares
SIGNATURES = {
    "opv_last_error": (C.c_char_p, []),
    "opv_abi_version": (C.c_int, []),
    "opv_create": (C.c_int, [C.POINTER(OpvConfig), C.POINTER(OpvWeights), C.c_int, C.POINTER(C.c_void_p)]),
    "opv_destroy": (C.c_int, [C.c_void_p]),
    "opv_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int32]),
    "opv_forward_packed": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_size_t, C.c_void_p],
    ),
    "opv_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "opv_profile_collect": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int32]),
    "opv_launch_count": (C.c_int64, [C.c_void_p]),
    "opv_fragment_means": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
         C.c_void_p],
    ),
    "opv_token_keep_probs": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "opv_sentence_prune": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p],
    ),
    "opv_pack_build": (C.c_int, [C.POINTER(OpvPackInput), C.POINTER(C.c_void_p)]),
    "opv_pack_view_get": (C.c_int, [C.c_void_p, C.POINTER(OpvPackView)]),
    "opv_pack_destroy": (C.c_int, [C.c_void_p]),
    "opv_forward_status": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "opv_op_gemm_residual_ln": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_int32, C.c_int32,
         C.c_void_p],
    ),
    "opv_op_gemm": (
        C.c_int,
        [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p],
    ),
    "opv_op_layernorm": (
        C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p]),
    "opv_op_embed_ln": (
        C.c_int,
        [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
         C.c_float, C.c_void_p],
    ),
    "opv_op_attention": (
        C.c_int,
        [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
         C.c_void_p]),
    "opv_set_option": (C.c_int, [C.c_char_p, C.c_int64]),
    "opv_engine_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "opv_op_rope": (
        C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "opv_op_geglu": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "opv_op_positions": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
}

_lib = None


class OpvError(RuntimeError):
    """A call into libopv_sm100.so failed (CUDA error, unsupported shape, ...)."""


def load() -> C.CDLL:
    """Load the shared library once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise OpvError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.opv_abi_version() != OPV_ABI_VERSION:
        raise OpvError(f"libopv_sm100.so ABI {lib.opv_abi_version()} != binding ABI {OPV_ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    """Translate a negative status into the exception type the reference would raise."""
    if rc == 0:
        return
    message = load().opv_last_error().decode("utf-8", "replace")
    text = f"{what}: {message}" if what else message
    if rc == OPV_ERR_INVALID_ARGUMENT:
        raise ValueError(message if not what else text)
    if rc == OPV_ERR_UNSUPPORTED:
        raise NotImplementedError(text)
    raise OpvError(text)
