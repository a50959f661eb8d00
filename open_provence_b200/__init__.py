"""open_provence_b200 -- B200 (sm_100a) engine for OpenProvence's scoring-and-pruning hot path.

Only what the hot path needs lives here:
  csrc/        hand-written CUDA for sm_100a + the C ABI (include/opv.h)
  _native.py   ctypes binding of libopv_sm100.so (no fallback)
  engine.py    weight packing + launches
  modeling.py  drop-in ``OpenProvenceModel`` (from_pretrained / forward / process)
  encoder.py   drop-in ``OpenProvenceEncoder`` inference APIs (token-level pruning, chunk votes)
  host_pack.py native block assembly between tokenizer and device (opv_pack_*)
  hf_auto.py   ``AutoModel`` / ``auto_map`` entry points of the reference, resolved to this engine
"""

__version__ = "0.1.0"

__all__ = ["__version__", "OpenProvenceModel", "OpenProvenceEncoder", "OpenProvenceConfig",
           "OpenProvenceForSequenceClassification", "OpenProvenceForTokenClassification", "register_auto_classes"]


def __getattr__(name):  # lazy: importing the package must not pull torch / the shared library in
    if name == "OpenProvenceModel":
        from .modeling import OpenProvenceModel

        return OpenProvenceModel
    if name == "OpenProvenceEncoder":
        from .encoder import OpenProvenceEncoder

        return OpenProvenceEncoder
    if name == "OpenProvenceConfig":
        from .config import OpenProvenceConfig

        return OpenProvenceConfig
    if name in ("OpenProvenceForSequenceClassification", "OpenProvenceForTokenClassification",
                "OpenProvenceEncoderForSequenceClassification", "OpenProvenceEncoderForTokenClassification",
                "register_auto_classes"):
        from . import hf_auto

        return getattr(hf_auto, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
