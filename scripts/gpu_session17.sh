#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_pytest7.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r02_pytest7.log | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-render --no-extras --no-configs3 > gpurun_out/r02_bench_v9.json 2>gpurun_out/r02_bench_v9.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v9.json')); print('v9', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"
timeout 300 python scripts/profile_step.py 4096 > gpurun_out/r02_step_table_v3.txt 2>&1
sed -n 4,6p gpurun_out/r02_step_table_v3.txt
