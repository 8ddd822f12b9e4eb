#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_pytest4.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r02_pytest4.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-render --no-extras --no-configs3 --no-trunk > gpurun_out/r02_bench_v3.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v3.json')); print('fast', d['ms_per_step'], d['gpu_launches'])"
timeout 600 python scripts/glue_sources.py > gpurun_out/r02_glue_sources.txt 2>&1
tail -80 gpurun_out/r02_glue_sources.txt | cut -c1-200
