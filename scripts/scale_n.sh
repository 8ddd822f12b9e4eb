N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_bench_${N}gpu_v2.json 2> gpurun_out/r02_bench_${N}gpu_v2.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_${N}gpu_v2.json').read().strip().splitlines()[-1]); print('${N}gpu', d['ms_per_step'], d['value'], 'render', d['render_sharded']['ms'], 'year', d['render_sharded']['year_sweep']['ms'], 'c3', d['configs3']['ms_per_step'], d['configs3']['value'])"
tail -1 gpurun_out/r02_bench_${N}gpu_v2.err | cut -c1-200
