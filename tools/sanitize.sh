#!/usr/bin/env bash
# compute-sanitizer memcheck over the smoke forward (single-CTA GEMMs, both attention kernels) and one model-family
# parity case (CTA-pair GEMMs).  Usage on the GPU box: bash tools/sanitize.sh
set -u
compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_families.py -m gpu -x -q -k "xsmall and bf16" 2>&1 | tail -4
compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_forward.py -m gpu -x -q -k "mean" 2>&1 | tail -4
# round 2: one-pass sliding-window kernel, two-Q-tile kernel, fp32 through the tcgen05 GEMM pipeline (6 passes + split)
compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "local_onepass or four_q_tiles or (attention_bf16_impls and 5-)" 2>&1 | tail -4
compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_forward.py -m gpu -x -q -k "tensor_core_pipeline" 2>&1 | tail -4
compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "local_onepass and 64" 2>&1 | tail -6
# round 2 (later): four-Q-tile kernel covered above; full-row residual GEMM + LayerNorm from TMEM, RESIDUAL_LN epilogue
compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "residual_layernorm_rows and not 76033 and not 40001" 2>&1 | tail -4
compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "residual_layernorm_rows and 1000-512" 2>&1 | tail -6
compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "four_q_tiles" 2>&1 | tail -6
# malformed cu_seqlens (boundaries past the buffer, negative, decreasing): clamped on the device, no out-of-bounds access
compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_forward.py -m gpu -x -q -k "malformed" 2>&1 | tail -4
