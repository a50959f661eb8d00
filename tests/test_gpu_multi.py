"""Multi-GPU paths on real devices (skipped on a 1-GPU box): two engines on two devices in one process, and
``process()`` sharded over two NCCL ranks through the product API."""

from __future__ import annotations

import json
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent

needs_two = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")


@needs_two
def test_two_engines_on_two_devices_in_one_process():
    """Function attributes (230 KB dynamic shared memory) are per device and options per engine: a second engine on
    another GPU of the same process must launch and give the same bits (process-global state)."""
    from open_provence_b200 import synthetic as syn
    from open_provence_b200.engine import Engine

    cfg = syn.backbone_config("base-130M")
    cfg["num_hidden_layers"], cfg["vocab_size"] = 3, 2048
    sd = syn.random_state_dict(cfg, seed=5)
    lengths = [300, 1, 2048, 129]
    rng = np.random.default_rng(3)
    ids = torch.from_numpy(rng.integers(3, 2048, size=sum(lengths)).astype(np.int32))
    cu = torch.from_numpy(np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32))
    outs = []
    engines = [Engine(cfg, sd, device=f"cuda:{d}", dtype="bf16", num_labels=1) for d in (0, 1)]
    for d, eng in enumerate(engines):  # interleaved use, current device left at 0 on purpose
        prune, rank = eng.forward_packed(ids.to(f"cuda:{d}"), cu.to(f"cuda:{d}"), max(lengths))
        torch.cuda.synchronize(d)
        outs.append((prune.cpu(), rank.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.isfinite(outs[0][0]).all()


@needs_two
def test_two_rank_nccl_process_equals_single_gpu(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "multi" / "nccl_process_worker.py"), str(tmp_path)]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-3000:]
    r0 = json.loads((tmp_path / "rank0.json").read_text())
    r1 = json.loads((tmp_path / "rank1.json").read_text())
    assert r0["world"] == r1["world"] == 2
    for name, entry in r0["cases"].items():
        assert entry["identical"], f"{name}: sharded result differs from the single-GPU result on rank 0"
        assert r1["cases"][name]["identical"], f"{name}: sharded result differs from the single-GPU result on rank 1"
        assert entry["result"] == r1["cases"][name]["result"], f"{name}: ranks disagree"
        assert entry["collectives"] <= 2 and r1["cases"][name]["collectives"] == entry["collectives"]
    assert any(e["collectives"] == 1 for e in r0["cases"].values())
