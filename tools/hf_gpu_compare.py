#!/usr/bin/env python
"""The reference's GPU path on the same B200: HF ModernBERT in bf16 (sdpa, and flash_attention_2 when transformers
accepts it) + the pruning Linear, right-padded batches of 32 as the reference feeds them (standalone:2832-2903), on
the bench workload.  Test / bench infrastructure: uses oracle.hf_cpu_baseline to build the library model.

    python tools/hf_gpu_compare.py [n_blocks] [seq_len]
"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_provence_b200 import synthetic as syn  # noqa: E402
from oracle import hf_cpu_baseline as hb  # noqa: E402

n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
cfg = syn.backbone_config("base-130M")
sd = syn.random_state_dict(cfg, seed=0)
wl = syn.make_workload(cfg, n_blocks, S, mode="dense", seed=1234)
ids = torch.from_numpy(wl["ids"]).long().view(n_blocks, S).cuda()
mask = torch.ones_like(ids)

for impl in ("sdpa", "flash_attention_2"):
    try:
        cfg_impl = dict(cfg)
        model, head = hb.build_hf_model(cfg_impl, sd)
        model.config._attn_implementation = impl
        if hasattr(model, "set_attn_implementation"):
            model.set_attn_implementation(impl)
        model, head = model.to("cuda", torch.bfloat16), head.to("cuda", torch.bfloat16)

        @torch.inference_mode()
        def step():
            for at in range(0, n_blocks, 32):  # reference default batch_size = 32 (standalone:3321)
                out = model(input_ids=ids[at : at + 32], attention_mask=mask[at : at + 32], output_hidden_states=True,
                            return_dict=True)
                prune = head(out.hidden_states[-1])
                _ = out.logits.float().cpu(), prune.float().cpu()  # the reference's .detach().cpu() (standalone:2893-2903)

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"HF ModernBERT bf16 attn={impl}: {n_blocks} blocks x {S} tokens in {dt * 1e3:.1f} ms -> {n_blocks / dt:.1f} pairs/s")
    except Exception as exc:  # noqa: BLE001
        print(f"HF ModernBERT bf16 attn={impl}: not available here ({type(exc).__name__}: {str(exc)[:120]})")
