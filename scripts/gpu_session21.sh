#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02_pytest9.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r02_pytest9.log | cut -c1-400
for FT in 1 0; do
  SNB_FUSED_TAIL=$FT timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-render --no-extras --no-configs3 --no-trunk > gpurun_out/r02_bench_v13_$FT.json 2>gpurun_out/r02_bench_v13_$FT.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v13_$FT.json')); print('fused tail $FT:', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
  tail -1 gpurun_out/r02_bench_v13_$FT.err | cut -c1-300
done
