#!/usr/bin/env bash
# Where does programmatic dependent launch stop paying?  base-130M, seq_len 2048, tokens per forward = 2048 * batch;
# pdl_max_tokens = 0 (never) against 1 << 40 (always).  Usage (GPU box): bash tools/pdl_sweep.sh <tag>
set -u
TAG=${1:-pdl}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
for B in 1 4 8 16 32; do
  for MAXTOK in 0 1099511627776; do
    timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch $B --set-option pdl_max_tokens=$MAXTOK 2>/dev/null \
      | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('batch $B tokens', d['config']['tokens_per_step_per_gpu'], 'pdl', 'on ' if $MAXTOK else 'off', d['ms_per_step'], 'ms/step', d['value'], 'pairs/s  e2e', d['e2e']['value'])" | tee -a "$OUT/pdl_sweep.log"
  done
done
