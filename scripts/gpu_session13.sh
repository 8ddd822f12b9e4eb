#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "xf or consumer_side" 2>&1 | grep -E "passed|failed|FAILED|^E  |Error" | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-render --no-extras --no-configs3 > gpurun_out/r02_bench_v6.json 2>gpurun_out/r02_bench_v6.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v6.json')); print('v6', d['ms_per_step'], d['gpu_launches']); print({k:(round(v['ms_per_launch']*1e3,1), round(v['achieved'])) for k,v in d['roofline']['trunk_launches'].items()})"
tail -3 gpurun_out/r02_bench_v6.err
