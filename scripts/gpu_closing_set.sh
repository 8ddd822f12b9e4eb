#!/bin/bash
# Closing set of a round on one GPU (run under gpurun): GPU suite, smoke, default bench + reference arm, step table of the graph replay, ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_pytest10.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r02_pytest10.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_bench_v15_full.json 2> gpurun_out/r02_bench_v15_full.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_v15_full.json').read().strip().splitlines()[-1]); r=d['roofline']; print('default bench', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', r['frac'], r['launches_per_step'], d['gpu_launches']); print(d['summary'])"
tail -1 gpurun_out/r02_bench_v15_full.err | cut -c1-200
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_ref_v4.json 2> gpurun_out/r02_ref_v4.err; tail -c 300 gpurun_out/r02_ref_v4.json
SNB_PROFILE_GRAPH=1 timeout 300 python scripts/profile_step.py 4096 > gpurun_out/r02_step_table_v7_graph.txt 2>&1
sed -n 3,5p gpurun_out/r02_step_table_v7_graph.txt
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -s 1000 -c 900 --csv --log-file gpurun_out/r02_launches_train_v4.csv \
    python bench.py --steps 1 --warmup 3 --no-render --no-cpu --no-extras --no-graph --no-trunk --no-configs3 > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/r02_launches_train_v4.csv
