"""Memory safety of the native block assembly: csrc/host_pack.cu is plain C++, so it is rebuilt here with
``g++ -fsanitize=address,undefined`` together with the fuzz driver tests/native/pack_fuzz.cpp and run over a few
thousand random inputs (empty contexts / sentences, token ids outside the visibility table, queries longer than the
block, every special-token template shape, both passes of the decode protocol)."""

from __future__ import annotations

import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_pack_fuzz_under_asan_ubsan(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = tmp_path / "pack_fuzz"
    build = subprocess.run(
        [gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-x", "c++",
         str(ROOT / "open_provence_b200" / "csrc" / "host_pack.cu"), str(ROOT / "tests" / "native" / "pack_fuzz.cpp"),
         "-o", str(exe)],
        capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr.lower() and "cannot find" in build.stderr.lower():
        pytest.skip("sanitizer runtimes not installed")
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([str(exe), "3000"], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, (run.stdout + run.stderr)[-3000:]
    assert "3000 cases ok" in run.stdout
