#!/bin/bash
# Run under gpurun (1 GPU).  Produces the ncu launch list of one bench command and full captures of the top kernels.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 420 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --steps 1 --warmup 3 --no-render --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 120 -c 3 -f -o gpurun_out/prof_gemm \
    python bench.py --steps 1 --warmup 3 --no-render --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_eval -s 2 -c 1 -f -o gpurun_out/prof_fused \
    python scripts/run_fused_once.py > gpurun_out/fused_once.log 2>&1
ls -la gpurun_out
