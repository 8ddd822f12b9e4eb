#!/bin/bash
# r02 session 8: step table of the current training step, ncu launch list of the bench command, full captures of the
# consumer-side-activation GEMM (gemm_tc3.cu) next to the four gemm2 variants
mkdir -p gpurun_out
timeout 300 python scripts/profile_step.py 4096 > gpurun_out/r02_step_table_v1.txt 2>&1
head -30 gpurun_out/r02_step_table_v1.txt
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -s 2000 -c 1500 --csv --log-file gpurun_out/r02_launches_train.csv \
    python bench.py --steps 1 --warmup 3 --no-render --no-cpu --no-extras --no-graph --no-trunk --no-configs3 > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/r02_launches_train.csv
timeout 600 $NCU --set full --import-source on -k regex:gemm -s 6 -c 6 -f -o gpurun_out/prof_gemm_r02 \
    python scripts/run_gemm_once.py 2 > gpurun_out/gemm_once.log 2>&1
tail -2 gpurun_out/gemm_once.log
ls -la gpurun_out | tail -5
