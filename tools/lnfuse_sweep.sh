#!/usr/bin/env bash
# Full-row residual GEMM + LayerNorm (ln_fuse = 2) against standalone LayerNorm launches over the tokens per forward.
for M in base-130M xsmall-30M; do
for B in 1 2 4 8 16 32; do
  for OPT in "ln_fuse=0" "ln_fuse_min_blocks=0"; do
    timeout 200 python bench.py --model $M --steps 20 --warmup 5 --no-cpu-baseline --batch $B --set-option $OPT 2>/dev/null \
      | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$M batch $B tokens', d['config']['tokens_per_step_per_gpu'], '$OPT', d['ms_per_step'], 'ms/step')"
  done
done
done
python tools/latency_bench.py 2>&1 | tail -4
