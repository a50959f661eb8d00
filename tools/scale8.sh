#!/usr/bin/env bash
# One 8-GPU gpurun call: BASELINE.json configs 3 and 4 at 8 GPUs (+ the headline config, weak).
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/scale8.sh r2e'
set -u
TAG=${1:-scale8}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
N=${NGPUS:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # name, port, args...
  local name=$1 port=$2; shift 2
  timeout 240 $TR --master-port $port bench.py --gpus $N --no-cpu-baseline "$@" > "$OUT/$name.json" 2> "$OUT/$name.err" || echo "$name failed: $(tail -2 $OUT/$name.err)"
  python - "$OUT/$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    r = d["roofline"]
    print(sys.argv[1], d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "step_frac_burst", r.get("step_frac_of_burst_peak"), d["config"]["blocks_per_rank"][:2], d["clocks"]["sm_mhz"])
except Exception as exc:
    print(sys.argv[1], "unreadable:", exc)
PY
}
# config 3: large-310M, 512 blocks of 4096 tokens IN TOTAL, sharded over the ranks (strong scaling)
run c3_large_strong 29601 --scaling strong --model large-310M --seq-len 4096 --batch 512 --steps 4 --warmup 3
# config 4: en-gte-149M sequence-length sweep, 131072 tokens per GPU per step (weak)
port=29610
for sl in 512 1024 2048 4096 8192; do
  run c4_engte_$sl $port --model en-gte-149M --seq-len $sl --batch $((131072 / sl)) --steps 4 --warmup 3
  port=$((port + 1))
done
# headline config, weak
run c2_base_weak 29620 --steps 8 --warmup 3
ls -la "$OUT"
