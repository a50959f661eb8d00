"""Shared pytest configuration: the ``gpu`` marker and fixture loaders."""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tiny_ckpt_dir() -> Path:
    return GOLDEN / "tiny_ckpt"


@pytest.fixture(scope="session")
def tiny_config() -> dict:
    return json.loads((GOLDEN / "tiny_ckpt" / "config.json").read_text())


@pytest.fixture(scope="session")
def tiny_weights() -> dict:
    """fp32 numpy state dict of the tiny golden checkpoint (reference ``state_dict`` keys)."""
    from safetensors.numpy import load_file

    return load_file(str(GOLDEN / "tiny_ckpt" / "model.safetensors"))


@pytest.fixture(scope="session")
def forward_golden() -> dict:
    data = np.load(GOLDEN / "forward_tiny.npz")
    return {k: data[k] for k in data.files}


@pytest.fixture(scope="session")
def process_golden() -> dict:
    return json.loads((GOLDEN / "process_tiny.json").read_text())
