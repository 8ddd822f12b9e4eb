#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02_pytest3.log
tail -12 gpurun_out/r02_pytest3.log
for tj in 4 6 8; do SNB_SWEEP_TJ=$tj timeout 300 python scripts/bench_extras.py year=1024,365 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read())['year']; print('TJ=$tj sweep_kernel_ms', d['sweep_kernel_ms'])"; done > gpurun_out/r02_year_tj.txt 2>&1
cat gpurun_out/r02_year_tj.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_bench_v2.json 2> gpurun_out/r02_bench_v2.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v2.json')); print(json.dumps(d['summary']))"
SNB_FAST_LOSS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-render --no-extras --no-configs3 --no-trunk > gpurun_out/r02_bench_v2_nofast.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v2_nofast.json')); print('nofast', d['ms_per_step'], d['gpu_launches'])"
