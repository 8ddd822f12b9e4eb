#!/bin/bash
# r02 session 14: full GPU suite, default bench line + reference arm, step table, ncu launch list, ncu --set full of the GEMM variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_pytest6.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r02_pytest6.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r02_bench_v7_full.json 2> gpurun_out/r02_bench_v7_full.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_v7_full.json').read().strip().splitlines()[-1]); print('default bench', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('cpu_baseline')); print({k:(round(v['ms_per_launch']*1e3,1), round(v['achieved'])) for k,v in d['roofline']['trunk_launches'].items()})"
tail -2 gpurun_out/r02_bench_v7_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_ref_v2.json 2> gpurun_out/r02_ref_v2.err; tail -c 600 gpurun_out/r02_ref_v2.json
timeout 300 python scripts/profile_step.py 4096 > gpurun_out/r02_step_table_v2.txt 2>&1
sed -n 4,12p gpurun_out/r02_step_table_v2.txt
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -s 2000 -c 1500 --csv --log-file gpurun_out/r02_launches_train_v2.csv \
    python bench.py --steps 1 --warmup 3 --no-render --no-cpu --no-extras --no-graph --no-trunk --no-configs3 > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:gemm -s 6 -c 6 -f -o gpurun_out/prof_gemm_r02b \
    python scripts/run_gemm_once.py 2 > gpurun_out/gemm_once.log 2>&1
ls -la gpurun_out | tail -4
