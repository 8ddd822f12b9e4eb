#!/bin/bash
# Run under gpurun (1 GPU).  Produces the ncu launch list of one bench command and full captures of the top kernels.
# Numbers printed by commands under ncu are never bench values.  Summaries: scripts/ncu_summary.py, scripts/ncu_traffic.py
# (writes profiles/r02_ncu_traffic.json, the file bench.py reads its `traffic` figures from), scripts/ncu_hot_k.py.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 1) launch list of the bench command (eager launches: the graph replays the same kernel sequence)
$NCU --metrics gpu__time_duration.sum -s 1000 -c 900 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --steps 1 --warmup 3 --no-render --no-cpu --no-extras --no-graph --no-trunk --no-configs3 > gpurun_out/bench_under_ncu.log 2>&1
# 2) full captures: the training GEMM variants on the trunk shape (CTA-pair kernel x4, resident-A kernel x2)
$NCU --set full --import-source on -k regex:gemm -s 6 -c 6 -f -o gpurun_out/prof_gemm \
    python scripts/run_gemm_once.py 2 > gpurun_out/gemm_once.log 2>&1
# 3) fused render kernel
$NCU --set full --import-source on -k regex:fused_eval -s 2 -c 1 -f -o gpurun_out/prof_fused \
    python scripts/run_fused_once.py > gpurun_out/fused_once.log 2>&1
# 4) compositing kernels (HBM roofline, dram traffic)
$NCU --set full -k regex:composite_ -s 5 -c 3 -f -o gpurun_out/prof_composite \
    python scripts/bench_extras.py composite > gpurun_out/composite_under_ncu.log 2>&1
# 5) in-kernel timelines (clock64 stamps per barrier wait of CTA 0)
for w in fwd_stats fwd_sin dgrad; do python scripts/tc2_timeline.py $w > gpurun_out/tc2_timeline_$w.txt 2>&1; done
python scripts/tc3_timeline.py 0 > gpurun_out/tc3_timeline_nostore.txt 2>&1
python scripts/tc3_timeline.py 1 > gpurun_out/tc3_timeline_store.txt 2>&1
ls -la gpurun_out
