#!/usr/bin/env bash
# A/B of the LayerNorm placement (ln_fuse 0 / 2) on the bench workload and on config 1 (xsmall 256 x 512).
for v in 2 0 2 0; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --set-option "ln_fuse=$v" > /tmp/ab_line.json
  python - "$v" <<'PY'
import json, sys
d = json.loads(open("/tmp/ab_line.json").read())
print(f"base ln_fuse={sys.argv[1]}: {d['value']} pairs/s {d['ms_per_step']} ms sm {d['clocks']['sm_mhz']} MHz  {d['profile_ms_per_step']}")
PY
done
for v in 2 0 2 0; do
  timeout 300 python bench.py --model xsmall-30M --seq-len 512 --batch 256 --steps 10 --warmup 3 --no-cpu-baseline --set-option "ln_fuse=$v" > /tmp/ab_line.json
  python - "$v" <<'PY'
import json, sys
d = json.loads(open("/tmp/ab_line.json").read())
print(f"xsmall ln_fuse={sys.argv[1]}: {d['value']} pairs/s {d['ms_per_step']} ms sm {d['clocks']['sm_mhz']} MHz  {d['profile_ms_per_step']}")
PY
done
