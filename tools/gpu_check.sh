#!/usr/bin/env bash
# One gpurun call: GPU parity tests, a bench line, the ncu launch list and full captures of the top kernels.
# Usage (from the repo root on the GPU box): bash tools/gpu_check.sh [tag]
set -u
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/smi.csv" 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee "$OUT/pytest_gpu.log"
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 2> "$OUT/bench.err" | tee "$OUT/bench.json"
tail -5 "$OUT/bench.err"
if [ "${NCU:-1}" = "1" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 300 --csv \
      --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_launch.log" 2>&1
  echo "== ncu full: GEMMs"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 12 -c 4 \
      -f -o "$OUT/prof_gemm" python bench.py --steps 1 --warmup 3 --no-cpu-baseline > "$OUT/ncu_gemm.log" 2>&1
  echo "== ncu full: attention"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention -s 3 -c 4 \
      -f -o "$OUT/prof_attn" python bench.py --steps 1 --warmup 3 --no-cpu-baseline > "$OUT/ncu_attn.log" 2>&1
fi
ls -la "$OUT"
