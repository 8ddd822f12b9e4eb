#!/bin/bash
# Run under gpurun (1 GPU).  Produces the ncu launch list of one bench command and full captures of the top kernels.
# Numbers printed by commands under ncu are never bench values.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 1) launch list of the bench command (eager launches: the graph replays the same kernel sequence)
$NCU --metrics gpu__time_duration.sum -s 2600 -c 1300 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --steps 1 --warmup 3 --no-render --no-cpu --no-extras --no-graph --no-trunk > gpurun_out/bench_under_ncu.log 2>&1
# 2) full captures: the four GEMM variants of the training path on the trunk shape
$NCU --set full --import-source on -k regex:gemm2_bf16 -s 4 -c 4 -f -o gpurun_out/prof_gemm2 \
    python scripts/run_gemm_once.py 2 > gpurun_out/gemm_once.log 2>&1
# 3) fused render kernel
$NCU --set full --import-source on -k regex:fused_eval -s 2 -c 1 -f -o gpurun_out/prof_fused \
    python scripts/run_fused_once.py > gpurun_out/fused_once.log 2>&1
# 4) compositing kernels (HBM roofline, dram traffic)
$NCU --set full -k regex:composite_ -s 5 -c 3 -f -o gpurun_out/prof_composite \
    python scripts/bench_extras.py composite > gpurun_out/composite_under_ncu.log 2>&1
ls -la gpurun_out
