#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_graph_gpu.py tests/test_net_tool_gpu.py -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|^E  " | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-render --no-extras --no-configs3 --no-trunk > gpurun_out/r02_bench_v14.json 2>gpurun_out/r02_bench_v14.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v14.json')); print('v14:', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
tail -1 gpurun_out/r02_bench_v14.err | cut -c1-300
timeout 300 python scripts/e2e_profile.py 2>&1 | grep -E "e2e ms|_step_graphed|replay|_draw_inputs|_fill_static" | cut -c1-150
