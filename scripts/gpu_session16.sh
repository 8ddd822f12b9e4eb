#!/bin/bash
mkdir -p gpurun_out
for CFG in "1 1" "0 1" "1 0" "0 0"; do set -- $CFG
  SNB_LOSS_OVERLAP=$1 SNB_FUSED_ADAM=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-render --no-extras --no-configs3 --no-trunk > gpurun_out/r02_bench_v8_$1$2.json 2>gpurun_out/r02_bench_v8_$1$2.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_v8_$1$2.json')); print('overlap $1 fusedadam $2:', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
  tail -2 gpurun_out/r02_bench_v8_$1$2.err | cut -c1-300
done
timeout 900 python -m pytest tests/test_train_graph_gpu.py tests/test_net_tool_gpu.py tests/test_baseline_size_gpu.py tests/test_network_gpu.py -m gpu -q -x 2>&1 | grep -E "passed|failed|FAILED|^E  |Error" | cut -c1-300
