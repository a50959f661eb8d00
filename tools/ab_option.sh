#!/usr/bin/env bash
# A/B a library tuning switch with bench.py:  bash tools/ab_option.sh <option> <value>...   (alternating runs)
OPT=$1; shift
for v in "$@"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --set-option "$OPT=$v" > /tmp/ab_line.json
  python - "$OPT" "$v" <<'PY'
import json, sys
d = json.loads(open("/tmp/ab_line.json").read())
print(f"{sys.argv[1]}={sys.argv[2]}: {d['value']} pairs/s  sm {d['clocks']['sm_mhz']} MHz  {d['profile_ms_per_step']}")
PY
done
