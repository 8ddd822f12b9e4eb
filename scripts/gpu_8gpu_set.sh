#!/bin/bash
# r02: 8-GPU bench (train weak scaling, configs[3], ray-sharded render gathered to rank 0, sharded year sweep) + multi-GPU tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|^E  |skipped" | cut -c1-300
SNB_SHARD_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_bench_8gpu_v1.json 2> gpurun_out/r02_bench_8gpu_v1.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_8gpu_v1.json').read().strip().splitlines()[-1]); print('8gpu', d['ms_per_step'], d['value']); print('render_sharded', d.get('render_sharded')); print('configs3', d.get('configs3')); print('year', d.get('year_sweep'))"
grep -i "shard" gpurun_out/r02_bench_8gpu_v1.err | tail -20
tail -3 gpurun_out/r02_bench_8gpu_v1.err
