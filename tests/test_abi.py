"""The C-ABI library loads on a CPU-only box and exports every symbol include/opv.h declares."""

from __future__ import annotations

import ctypes
import re
import shutil
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib_path():
    from open_provence_b200._build import LIB_PATH, build_native

    if not LIB_PATH.exists():
        if shutil.which("nvcc") is None and not Path("/usr/local/cuda/bin/nvcc").exists():
            pytest.skip("libopv_sm100.so not built and nvcc not available")
        build_native()
    return LIB_PATH


def _declared_functions() -> list[str]:
    text = (ROOT / "include" / "opv.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(opv_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_the_documented_entry_points():
    names = _declared_functions()
    for required in ("opv_create", "opv_destroy", "opv_workspace_bytes", "opv_forward_packed", "opv_fragment_means",
                     "opv_sentence_prune", "opv_last_error"):
        assert required in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(str(lib_path))
    missing = [n for n in _declared_functions() if not hasattr(lib, n)]
    assert not missing, f"symbols declared in include/opv.h but not exported: {missing}"
    lib.opv_abi_version.restype = ctypes.c_int
    assert lib.opv_abi_version() == 2


def test_binding_covers_every_declared_symbol(lib_path):
    from open_provence_b200 import _native

    assert sorted(_native.SIGNATURES) == _declared_functions()
    _native.load()


def test_sass_uses_blackwell_tensor_and_tma_paths(lib_path):
    """tcgen05.mma -> UTCHMMA, TMA -> UTMALDG, tcgen05.ld -> LDTM (B200_PROFILING.md)."""
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    import subprocess

    sass = subprocess.run([cuobjdump, "-sass", str(lib_path)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} not found in SASS"


def test_product_has_no_oracle_import():
    """The product path must never route through the CPU oracle."""
    for path in (ROOT / "open_provence_b200").rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{path} imports the oracle"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from open_provence_b200 import _native

    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", tmp_path / "libopv_sm100.so")
    with pytest.raises(_native.OpvError, match="no CPU fallback"):
        _native.load()
