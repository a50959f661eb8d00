"""Build recipe for libopv_sm100.so (nvcc, sm_100a only).  Used by ``__graft_entry__.build()``.

The library is built IN-TREE (``open_provence_b200/lib/libopv_sm100.so``) so that it travels with the
repository snapshot to the GPU box; it is git-ignored.
"""

from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libopv_sm100.so"

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-std=c++17",
    "--shared",
    "-Xcompiler",
    "-fPIC",
    "-Xptxas",
    "-v",
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


STAMP_PATH = LIB_DIR / "libopv_sm100.stamp"


def source_digest() -> str:
    """sha256 over the flags and every file the library is compiled from (content, not mtime: a snapshot copy or a
    checkout changes mtimes without changing sources, and an edited header must always trigger a rebuild)."""
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    deps = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "opv.h"]
    for p in deps:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def _stale() -> bool:
    if not LIB_PATH.exists() or not STAMP_PATH.exists():
        return True
    return STAMP_PATH.read_text().strip() != source_digest()


def build_native(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu into lib/libopv_sm100.so with nvcc for sm_100a."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libopv_sm100.so")
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB_PATH), *[str(s) for s in sources()]]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    (LIB_DIR / "build.log").write_text(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{log[-4000:]}")
    if "bytes spill stores" in log:
        spills = [ln for ln in log.splitlines() if "spill stores" in ln and " 0 bytes spill stores" not in ln]
        if spills and verbose:
            print("register spills:\n" + "\n".join(spills))
    STAMP_PATH.write_text(source_digest() + "\n")
    if verbose:
        print(f"built {LIB_PATH} ({LIB_PATH.stat().st_size} bytes)")
    return LIB_PATH


if __name__ == "__main__":
    build_native(force=True, verbose=True)
