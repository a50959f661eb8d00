"""On-disk formats of a reference checkpoint: legacy un-prefixed keys
(standalone:1452-1464), ``pytorch_model.bin``, sharded safetensors with an index, and the explicit rejection of
backbones that are not ModernBERT (the reference's ``scripts/eval_mldr.py:68-125`` also loads DeBERTa baselines).

The CPU tests pin what the loader hands to the engine; the ``gpu`` tests load every format through
``from_pretrained`` and require the same logits as the canonical ``model.safetensors``.
"""

from __future__ import annotations

import json
import shutil
from pathlib import Path

import pytest
import torch
from safetensors.torch import load_file, save_file

from open_provence_b200.modeling import OpenProvenceModel, _load_state_dict, convert_legacy_state_dict


def _copy_ckpt(src: Path, dst: Path) -> Path:
    dst.mkdir(parents=True, exist_ok=True)
    for f in src.iterdir():
        if f.name != "model.safetensors":
            shutil.copy(f, dst / f.name)
    return dst


def _legacy_names(sd):
    """What a pre-``ranking_model.`` checkpoint looks like: backbone keys bare, pruning head keys unchanged."""
    return {(k[len("ranking_model."):] if k.startswith("ranking_model.") else k): v for k, v in sd.items()}


@pytest.fixture(scope="module")
def canonical(tiny_ckpt_dir):
    return load_file(str(tiny_ckpt_dir / "model.safetensors"))


def _variants(tiny_ckpt_dir, canonical, root: Path) -> dict[str, Path]:
    out = {}
    d = _copy_ckpt(tiny_ckpt_dir, root / "legacy_keys")
    save_file(_legacy_names(canonical), str(d / "model.safetensors"))
    out["legacy_keys"] = d
    d = _copy_ckpt(tiny_ckpt_dir, root / "torch_bin")
    torch.save(dict(canonical), str(d / "pytorch_model.bin"))
    out["torch_bin"] = d
    d = _copy_ckpt(tiny_ckpt_dir, root / "sharded")
    keys = sorted(canonical)
    halves = (keys[: len(keys) // 2], keys[len(keys) // 2 :])
    weight_map = {}
    for i, part in enumerate(halves, 1):
        name = f"model-{i:05d}-of-00002.safetensors"
        save_file({k: canonical[k] for k in part}, str(d / name))
        weight_map.update({k: name for k in part})
    (d / "model.safetensors.index.json").write_text(json.dumps({"metadata": {}, "weight_map": weight_map}))
    out["sharded"] = d
    return out


def test_legacy_keys_get_the_ranking_model_prefix(canonical):
    legacy = _legacy_names(canonical)
    assert not any(k.startswith("ranking_model.") for k in legacy)
    converted = convert_legacy_state_dict(legacy)
    assert set(converted) == set(canonical)
    assert all(torch.equal(converted[k], canonical[k]) for k in canonical)
    assert convert_legacy_state_dict(canonical) is canonical  # already prefixed: returned untouched (standalone:1455)


def test_every_format_resolves_to_the_same_state_dict(tiny_ckpt_dir, canonical, tmp_path):
    for name, path in _variants(tiny_ckpt_dir, canonical, tmp_path).items():
        sd = convert_legacy_state_dict(_load_state_dict(path))
        assert set(sd) == set(canonical), name
        assert all(torch.equal(sd[k], canonical[k]) for k in canonical), name
    empty = _copy_ckpt(tiny_ckpt_dir, tmp_path / "empty")
    with pytest.raises(FileNotFoundError, match="no model.safetensors"):
        _load_state_dict(empty)


@pytest.mark.gpu
def test_from_pretrained_loads_every_format_identically(tiny_ckpt_dir, canonical, tmp_path, forward_golden):
    ids = torch.from_numpy(forward_golden["input_ids"])
    mask = torch.from_numpy(forward_golden["attention_mask"])
    ref = OpenProvenceModel.from_pretrained(tiny_ckpt_dir, device="cuda", dtype="fp32").forward(
        input_ids=ids, attention_mask=mask, return_dict=True)
    for name, path in _variants(tiny_ckpt_dir, canonical, tmp_path).items():
        out = OpenProvenceModel.from_pretrained(path, device="cuda", dtype="fp32").forward(
            input_ids=ids, attention_mask=mask, return_dict=True)
        assert torch.equal(out.ranking_logits, ref.ranking_logits), name
        assert torch.equal(out.pruning_logits, ref.pruning_logits), name


@pytest.mark.gpu
@pytest.mark.parametrize("model_type", ["deberta-v2", "bert", "xlm-roberta"])
def test_non_modernbert_backbone_is_rejected_with_a_clear_message(tiny_ckpt_dir, tmp_path, model_type):
    d = tmp_path / model_type
    shutil.copytree(tiny_ckpt_dir, d)
    cfg = json.loads((d / "config.json").read_text())
    cfg["base_model_config"]["model_type"] = model_type
    (d / "config.json").write_text(json.dumps(cfg))
    with pytest.raises(NotImplementedError, match="only ModernBERT is implemented"):
        OpenProvenceModel.from_pretrained(d, device="cuda")
