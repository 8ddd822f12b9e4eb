#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_v12_full.json 2> gpurun_out/r02_bench_v12_full.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_v12_full.json').read().strip().splitlines()[-1]); r=d['roofline']; print('default bench', d['ms_per_step'], d['value'], d['e2e']['value'], 'frac', r['frac'], 'launches', r['launches_per_step'], 'kernel ms', r['kernel_ms_per_step']); print(d['summary'])"
tail -2 gpurun_out/r02_bench_v12_full.err
