#!/usr/bin/env bash
mkdir -p gpurun_out
for pair in 0 1; do timeout 300 python tools/gemm_check.py $pair 2>&1 | grep -v Warning | tail -8; done | tee gpurun_out/gemm_check.log
