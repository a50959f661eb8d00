"""Engine forward at the dimensions of the four published model families (SURVEY.md section 8 dims
table), against the numpy oracle on the same seeded weights and ragged packed batches.

The golden fixtures produced by the reference cover a tiny 128-wide model; these cases exercise the
tile shapes the real checkpoints hit (H = 256 / 512 / 768, I = 1024 / 2048 / 3072 / 1152, 4-12 heads)
with a reduced layer count (global, local, local, global) so the fp64 oracle finishes in seconds.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from open_provence_b200 import synthetic as syn  # noqa: E402
from open_provence_b200.engine import Engine  # noqa: E402
from oracle import modernbert_numpy as onp  # noqa: E402

FAMILIES = ["xsmall-30M", "base-130M", "large-310M", "en-gte-149M"]
LENGTHS = [1, 7, 64, 129, 200, 257, 385, 130]  # below / at / across the 128-token tile and 64-token band edges


def _case(name: str, layers: int = 4, vocab: int = 1024):
    cfg = syn.backbone_config(name)
    cfg["num_hidden_layers"] = layers
    cfg["vocab_size"] = vocab
    sd = syn.random_state_dict(cfg, seed=7)
    rng = np.random.default_rng(11)
    seqs = [rng.integers(3, vocab, size=n).tolist() for n in LENGTHS]
    return cfg, sd, seqs


def _run(cfg, sd, seqs, dtype):
    eng = Engine(cfg, sd, device="cuda", dtype=dtype, num_labels=1)
    ids = torch.tensor([t for s in seqs for t in s], dtype=torch.int32, device="cuda")
    cu = torch.tensor(np.concatenate([[0], np.cumsum([len(s) for s in seqs])]), dtype=torch.int32, device="cuda")
    prune, rank = eng.forward_packed(ids, cu, max(len(s) for s in seqs))
    torch.cuda.synchronize()
    return prune.cpu().double().numpy(), rank.cpu().double().numpy()


@pytest.fixture(scope="module", params=FAMILIES)
def family(request):
    cfg, sd, seqs = _case(request.param)
    w64 = {k: v.double().numpy() for k, v in sd.items()}
    ref_rank, ref_prune = onp.forward_batch(seqs, w64, cfg)
    return request.param, cfg, sd, seqs, ref_rank, np.concatenate(ref_prune)


def test_family_fp32_within_1e5(family):
    name, cfg, sd, seqs, ref_rank, ref_prune = family
    prune, rank = _run(cfg, sd, seqs, "fp32")
    e_rank, e_prune = np.abs(rank - ref_rank).max(), np.abs(prune - ref_prune).max()
    print(f"{name} fp32 engine vs fp64 oracle: rank {e_rank:.2e} prune {e_prune:.2e} (|prune| max {np.abs(ref_prune).max():.1f})")
    assert e_rank < 1e-5 and e_prune < 2e-5 * max(1.0, np.abs(ref_prune).max())


def test_family_bf16_close(family):
    name, cfg, sd, seqs, ref_rank, ref_prune = family
    prune, rank = _run(cfg, sd, seqs, "bf16")
    scale = max(1.0, np.abs(ref_prune).max())
    e_rank, e_prune = np.abs(rank - ref_rank).max(), np.abs(prune - ref_prune).max()
    print(f"{name} bf16 engine vs fp64 oracle: rank {e_rank:.2e} prune {e_prune:.2e} (|prune| max {scale:.1f})")
    assert np.isfinite(prune).all() and np.isfinite(rank).all()
    # bf16 operands (2^-9 relative rounding) with fp32 accumulation / residual / LN / softmax
    assert e_rank < 2e-2 and e_prune < 1e-2 * scale


def test_maximum_sequence_length_8192(tiny_ckpt_dir, tiny_config):
    """One block at max_position_embeddings (8192 tokens: 64 key blocks per query tile, RoPE positions up to 8191
    with theta 160000 / 10000) next to a 4097-token one, against the reference's fp64 forward on the tiny golden
    checkpoint (tests/golden/make_golden_long.py)."""
    from safetensors.torch import load_file

    golden = np.load(tiny_ckpt_dir.parent / "forward_long_tiny.npz")
    sd = load_file(str(tiny_ckpt_dir / "model.safetensors"))
    lengths = golden["lengths"].tolist()
    ids = torch.from_numpy(golden["input_ids"]).to("cuda")
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lengths)]), dtype=torch.int32, device="cuda")
    ref_rank, ref_prune = golden["ranking_logits_f64"], golden["pruning_logits_f64"]
    scale = max(1.0, np.abs(ref_prune).max())
    for dtype, tol_rank, tol_prune in (("fp32", 1e-5, 2e-5 * scale), ("bf16", 2e-2, 1e-2 * scale)):
        eng = Engine(tiny_config["base_model_config"], sd, device="cuda", dtype=dtype, num_labels=1)
        prune, rank = eng.forward_packed(ids, cu, max(lengths))
        torch.cuda.synchronize()
        e_rank = np.abs(rank.cpu().double().numpy() - ref_rank).max()
        e_prune = np.abs(prune.cpu().double().numpy() - ref_prune).max()
        print(f"S=8192 {dtype} engine vs fp64 reference: rank {e_rank:.2e} prune {e_prune:.2e} (|prune| max {scale:.1f})")
        assert e_rank < tol_rank and e_prune < tol_prune


def test_empty_and_single_token_batches():
    cfg = syn.backbone_config("tiny")
    sd = syn.random_state_dict(cfg, seed=2)
    eng = Engine(cfg, sd, device="cuda", dtype="bf16", num_labels=1)
    empty_ids = torch.zeros(0, dtype=torch.int32, device="cuda")
    prune, rank = eng.forward_packed(empty_ids, torch.zeros(1, dtype=torch.int32, device="cuda"), 1)
    assert prune.shape == (0, 2) and rank.shape == (0, 1)
    # 300 one-token blocks: every sequence is its own CLS row
    ids = torch.arange(3, 303, dtype=torch.int32, device="cuda") % cfg["vocab_size"]
    cu = torch.arange(0, 301, dtype=torch.int32, device="cuda")
    prune, rank = eng.forward_packed(ids.contiguous(), cu, 1)
    torch.cuda.synchronize()
    w64 = {k: v.double().numpy() for k, v in sd.items()}
    ref_rank, ref_prune = onp.forward_batch([[int(t)] for t in ids.cpu().tolist()[:5]], w64, cfg)
    assert np.abs(rank[:5].cpu().double().numpy() - ref_rank).max() < 2e-2
    assert torch.isfinite(prune).all() and torch.isfinite(rank).all()


def test_ragged_batch_equals_single_sequences_at_base_dims():
    """Unpadded packing at bench-like sizes: every sequence of a ragged batch (1 .. 2048 tokens, base-130M widths,
    global + sliding-window layers) must come out bit-identical to running it alone -- tile decode of the
    persistent attention kernels, RoPE positions and the row-grouped GEMM tile order do not leak across blocks."""
    cfg = syn.backbone_config("base-130M")
    cfg["num_hidden_layers"], cfg["vocab_size"] = 3, 2048
    sd = syn.random_state_dict(cfg, seed=13)
    eng = Engine(cfg, sd, device="cuda", dtype="bf16", num_labels=1)
    lengths = [2048, 1, 777, 1500, 129, 64, 2047]
    rng = np.random.default_rng(17)
    ids = torch.from_numpy(rng.integers(3, 2048, size=sum(lengths)).astype(np.int32)).to("cuda")
    cu_host = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32)
    prune_all, rank_all = eng.forward_packed(ids, torch.from_numpy(cu_host).to("cuda"), max(lengths))
    for b, n in enumerate(lengths):
        lo, hi = int(cu_host[b]), int(cu_host[b + 1])
        one_cu = torch.tensor([0, n], dtype=torch.int32, device="cuda")
        prune_one, rank_one = eng.forward_packed(ids[lo:hi].contiguous(), one_cu, n)
        torch.cuda.synchronize()
        assert torch.equal(prune_all[lo:hi], prune_one), f"sequence {b} (len {n}) differs inside the batch"
        assert torch.equal(rank_all[b : b + 1], rank_one)


@pytest.mark.parametrize("name", ["xsmall-30M", "base-130M"])
def test_programmatic_dependent_launch_is_bit_identical(name):
    """The PDL attribute only moves kernel prologues ahead of the predecessor's tail (``griddepcontrol.wait`` comes
    before the first global access): outputs with it, without it, and past the token threshold are the same bits."""
    from open_provence_b200 import ops

    cfg, sd, seqs = _case(name)
    outs = []
    # early release (default for small forwards) | no PDL | late release (what large forwards get) | attribute off above
    # the token threshold
    for options in ((("pdl", 1),), (("pdl", 0),), (("pdl_max_tokens", 0), ("pdl_late", 1)),
                    (("pdl_max_tokens", 0), ("pdl_late", 0))):
        for option, value in options:
            ops.set_option(option, value)
        try:
            outs.append(_run(cfg, sd, seqs, "bf16"))
        finally:
            ops.set_option("pdl", 1)
            ops.set_option("pdl_max_tokens", 32768)
            ops.set_option("pdl_late", 1)
    for prune, rank in outs[1:]:
        assert np.array_equal(prune, outs[0][0]) and np.array_equal(rank, outs[0][1])


@pytest.mark.parametrize("name", FAMILIES)
def test_layernorm_in_the_residual_gemm_epilogue_is_bit_identical(name):
    """``ln_fuse=1`` (not the default: measured slower): forwards with at least one 256-row block per
    CTA pair run mlp_norm / the next layer's attn_norm inside the Wo / Wo2 GEMM (RESIDUAL_LN, H = 256 / 512 / 768 /
    1024) with the statistics code of ``layernorm_kernel``: the same bits as the standalone launches, including the
    ragged last row block."""
    from open_provence_b200 import ops

    cfg, sd, _ = _case(name, layers=3)
    rng = np.random.default_rng(23)
    lengths = [2048] * 9 + [777, 1, 300, 129]  # 19 639 tokens: 77 row pairs, the last one 183 rows
    seqs = [rng.integers(3, cfg["vocab_size"], size=n).tolist() for n in lengths]
    outs = []
    for fuse in (1, 0):
        ops.set_option("ln_fuse", fuse)
        try:
            outs.append(_run(cfg, sd, seqs, "bf16"))
        finally:
            ops.set_option("ln_fuse", 2)
    assert np.isfinite(outs[0][0]).all() and np.isfinite(outs[0][1]).all()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("name", ["xsmall-30M", "base-130M"])
def test_layernorm_from_tmem_in_the_residual_gemm(name):
    """``ln_fuse=2`` (default): attn.Wo (and mlp.Wo at H = 256) run on the full-row kernel that computes the following LayerNorm
    from TMEM (gemm_rowln.cuh).  Same fp32 residual arithmetic, two-pass statistics with a different summation order:
    the forward stays within bf16 rounding noise of the standalone-LayerNorm forward and of the fp64 oracle bounds."""
    from open_provence_b200 import ops

    cfg, sd, _ = _case(name, layers=3)
    rng = np.random.default_rng(23)
    lengths = [2048] * 9 + [777, 1, 300, 129]
    seqs = [rng.integers(3, cfg["vocab_size"], size=n).tolist() for n in lengths]
    outs = []
    for fuse in (2, 0):
        ops.set_option("ln_fuse", fuse)
        try:
            outs.append(_run(cfg, sd, seqs, "bf16"))
        finally:
            ops.set_option("ln_fuse", 2)
    (p2, r2), (p0, r0) = outs
    assert np.isfinite(p2).all() and np.isfinite(r2).all()
    scale = max(1.0, np.abs(p0).max())
    d_prune, d_rank = np.abs(p2 - p0).max(), np.abs(r2 - r0).max()
    print(f"{name}: ln_fuse=2 vs 0: prune {d_prune:.2e} (scale {scale:.1f}) rank {d_rank:.2e}")
    assert d_prune < 4e-3 * scale and d_rank < 4e-3


@pytest.mark.parametrize("name", ["xsmall-30M", "en-gte-149M"])
def test_family_mean_pooling_fp32(name):
    """classifier_pooling = "mean" at H = 256 / 768 against the fp64 oracle."""
    cfg, sd, seqs = _case(name)
    cfg["classifier_pooling"] = "mean"
    w64 = {k: v.double().numpy() for k, v in sd.items()}
    ref_rank, _ = onp.forward_batch(seqs, w64, cfg)
    _, rank = _run(cfg, sd, seqs, "fp32")
    assert np.abs(rank - ref_rank).max() < 1e-5
