#!/usr/bin/env bash
# every attention variant in its own process (a trap in one cannot poison the next)
mkdir -p gpurun_out
for hw in -1 64; do for impl in ${IMPLS:-5 4 3}; do
  timeout 180 python tools/attn_check.py $impl $hw 2>&1 | grep -v Warning | tail -12
done; done | tee gpurun_out/attn_check.log
