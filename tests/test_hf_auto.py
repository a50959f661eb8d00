"""``AutoConfig`` / ``AutoModel*`` resolve ``model_type == "open_provence"`` to this package after
``register_auto_classes()`` (reference: standalone:3810-3906 wrapper classes, encoder.py:1078-1085 ``auto_map``)."""

from __future__ import annotations

import json

import numpy as np
import pytest

from open_provence_b200 import hf_auto
from open_provence_b200.modeling import OpenProvenceModel


def test_auto_factories_dispatch_to_the_wrappers(tiny_ckpt_dir, monkeypatch):
    from transformers import AutoConfig, AutoModel, AutoModelForSequenceClassification, AutoModelForTokenClassification

    hf_auto.register_auto_classes()
    hf_auto.register_auto_classes()  # idempotent
    cfg = AutoConfig.from_pretrained(str(tiny_ckpt_dir))
    assert cfg.model_type == "open_provence" and cfg.base_model_config["model_type"] == "modernbert"

    calls = []

    def fake(cls, path, **kwargs):
        calls.append((cls, str(path), kwargs))
        return "engine-backed model"

    monkeypatch.setattr(OpenProvenceModel, "from_pretrained", classmethod(fake))
    assert AutoModel.from_pretrained(str(tiny_ckpt_dir), device="cuda:0", dtype="float32") == "engine-backed model"
    cls, path, kwargs = calls[-1]
    assert issubclass(cls, hf_auto.OpenProvenceForSequenceClassification) and path == str(tiny_ckpt_dir)
    assert kwargs["device"] == "cuda:0" and kwargs["dtype"] == "float32" and "config" not in kwargs
    AutoModelForSequenceClassification.from_pretrained(str(tiny_ckpt_dir), trust_remote_code=True)
    assert issubclass(calls[-1][0], hf_auto.OpenProvenceForSequenceClassification)
    assert "dtype" not in calls[-1][2]  # nothing requested: the engine default (bf16) applies
    AutoModelForTokenClassification.from_pretrained(str(tiny_ckpt_dir))
    assert issubclass(calls[-1][0], hf_auto.OpenProvenceForTokenClassification)


def test_auto_map_matches_the_reference_layout():
    cfg = hf_auto.with_auto_map({"model_type": "open_provence"})
    assert cfg["architectures"] == ["OpenProvenceForSequenceClassification"]
    assert set(cfg["auto_map"]) == {"AutoConfig", "AutoModel", "AutoModelForSequenceClassification",
                                    "AutoModelForTokenClassification"}
    assert all(v.startswith("modeling_open_provence_standalone.") for v in cfg["auto_map"].values())
    assert hf_auto.OpenProvenceEncoderForTokenClassification is hf_auto.OpenProvenceForTokenClassification


@pytest.mark.gpu
def test_auto_model_loads_the_engine_and_token_wrapper_exposes_pruning_logits(tiny_ckpt_dir, forward_golden, tmp_path):
    import torch
    from transformers import AutoModel, AutoModelForTokenClassification

    from open_provence_b200.encoder import OpenProvenceEncoder

    hf_auto.register_auto_classes()
    seq = AutoModel.from_pretrained(str(tiny_ckpt_dir), device="cuda", dtype="float32")
    tok = AutoModelForTokenClassification.from_pretrained(str(tiny_ckpt_dir), device="cuda", dtype="float32")
    assert isinstance(seq, OpenProvenceModel) and seq.engine is not None
    ids = torch.from_numpy(forward_golden["input_ids"]).cuda()
    mask = torch.from_numpy(forward_golden["attention_mask"]).cuda()
    a = seq(input_ids=ids, attention_mask=mask, return_dict=True)
    b = tok(input_ids=ids, attention_mask=mask, return_dict=True)
    valid = forward_golden["attention_mask"].astype(bool)
    np.testing.assert_allclose(a.logits.cpu().numpy(), forward_golden["ranking_logits_f64"], atol=1e-5)
    np.testing.assert_allclose(b.logits.cpu().numpy()[valid], forward_golden["pruning_logits_f64"][valid], atol=1e-5)
    assert torch.equal(b.ranking_logits, a.ranking_logits) and tok.num_labels == 2
    labels = torch.zeros(ids.shape, dtype=torch.long, device="cuda")
    out = tok(input_ids=ids, attention_mask=mask, labels=labels)
    want = torch.nn.functional.cross_entropy(b.logits[mask.bool()], labels[mask.bool()])
    assert abs(float(out.loss) - float(want)) < 1e-6
    assert tok(input_ids=ids, attention_mask=mask, return_dict=False)[0].shape == b.logits.shape

    # a checkpoint written by this package loads back through AutoModel.  ``auto_map`` is only written when the remote
    # code module it names travels with the checkpoint (stock transformers fails on a dangling auto_map)
    enc = OpenProvenceEncoder.from_pretrained(tiny_ckpt_dir, device="cuda", dtype="fp32")
    enc.save_pretrained(tmp_path / "ckpt")
    saved = json.loads((tmp_path / "ckpt" / "config.json").read_text())
    assert "auto_map" not in saved and saved["architectures"] == hf_auto.ARCHITECTURES
    assert saved["transformers_version"] and saved["id2label"]  # unknown config keys round-trip (config.extra)
    import shutil

    src = tmp_path / "src_with_remote_code"
    shutil.copytree(tiny_ckpt_dir, src)
    (src / "modeling_open_provence_standalone.py").write_text("# the reference's remote-code module travels with the checkpoint\n")
    OpenProvenceEncoder.from_pretrained(src, device="cuda", dtype="fp32").save_pretrained(tmp_path / "ckpt2")
    saved2 = json.loads((tmp_path / "ckpt2" / "config.json").read_text())
    assert saved2["auto_map"] == hf_auto.AUTO_MAP and (tmp_path / "ckpt2" / "modeling_open_provence_standalone.py").is_file()
    again = AutoModel.from_pretrained(str(tmp_path / "ckpt"), device="cuda", dtype="float32")
    c = again(input_ids=ids, attention_mask=mask, return_dict=True)
    assert torch.equal(c.ranking_logits, a.ranking_logits)
