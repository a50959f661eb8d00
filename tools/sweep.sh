#!/usr/bin/env bash
# BASELINE.json configs 1-4 on one GPU (tokens per step per GPU = 131072 unless stated); one JSON line each.
# Usage (GPU box): bash tools/sweep.sh <tag>
set -u
TAG=${1:-sweep}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
run() { echo "== $*" >&2; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>>"$OUT/sweep.err" | tee -a "$OUT/sweep.jsonl"; }
run --model xsmall-30M --seq-len 512 --batch 256
run --model base-130M --seq-len 2048 --batch 64 --mode ragged
run --model large-310M --seq-len 4096 --batch 64
for S in 512 1024 2048 4096 8192; do run --model en-gte-149M --seq-len $S --batch $((131072 / S)); done
run --model base-130M --seq-len 512 --batch 256
run --model base-130M --seq-len 8192 --batch 16
