"""bench.py prints exactly ONE JSON line with the keys the driver's contract names (reference arm on CPU here, the
engine arm on the GPU with a tiny workload)."""

from __future__ import annotations

import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(*args: str) -> dict:
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, f"stdout must carry exactly one line, got {len(lines)}: {proc.stdout[:500]}"
    return json.loads(lines[0])


def test_reference_arm_line_on_cpu():
    line = _run("--impl", "reference", "--model", "tiny", "--seq-len", "64", "--steps", "1", "--warmup", "0",
                "--ref-pairs-per-step", "1")
    assert BASE_KEYS <= set(line)
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["higher_is_better"] is True
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert "workload" in line["config"] and line["gpu_launches"] == 0


@pytest.mark.gpu
def test_engine_arm_line_on_gpu():
    line = _run("--model", "tiny", "--seq-len", "256", "--batch", "8", "--steps", "2", "--warmup", "3", "--cpu-pairs", "2")
    assert BASE_KEYS | {"roofline", "clocks", "profile_ms_per_step"} <= set(line)
    assert line["n_gpus"] == 1 and line["dtype"] == "bf16" and line["data"] == "synthetic" and line["scaling"] == "weak"
    assert line["value"] > 0 and line["e2e"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["gpu_launches"] > 0 and line["vs_baseline"] is None
    roof = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(roof) and roof["bound"] in ("hbm", "tensor")
    assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"])
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    # in-run correctness: the step's GPU outputs against the CPU results the cpu_baseline leg computed
    par = line["parity"]
    assert par["blocks"] == 2 and par["sentences"] > 0 and par["e2e_equals_device_path"] is True
    assert par["mismatches_outside_1e-2_band"] == 0
    assert par["prune_logit_max_rel_to_scale"] < 1e-2 and par["rank_score_max_abs"] < 1e-2
