#!/bin/bash
# r02 final 1-GPU set: GPU suite, smoke, default bench + reference arm, step tables, ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_pytest8.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r02_pytest8.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_v10_full.json 2> gpurun_out/r02_bench_v10_full.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_v10_full.json').read().strip().splitlines()[-1]); print('default bench', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks']); print(d['summary'])"
tail -2 gpurun_out/r02_bench_v10_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_ref_v3.json 2> gpurun_out/r02_ref_v3.err; tail -c 400 gpurun_out/r02_ref_v3.json
SNB_PROFILE_GRAPH=1 timeout 300 python scripts/profile_step.py 4096 > gpurun_out/r02_step_table_v5_graph.txt 2>&1
sed -n 3,5p gpurun_out/r02_step_table_v5_graph.txt
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -s 2000 -c 1500 --csv --log-file gpurun_out/r02_launches_train_v3.csv \
    python bench.py --steps 1 --warmup 3 --no-render --no-cpu --no-extras --no-graph --no-trunk --no-configs3 > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/r02_launches_train_v3.csv
