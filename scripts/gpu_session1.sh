#!/bin/bash
# round 2, GPU session 1: box facts, the GPU test suite, the bench (both arms), ncu of the year-sweep kernel
mkdir -p gpurun_out
{ nproc; free -g; nvidia-smi -L; } > gpurun_out/r02_box.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -5 gpurun_out/r02_pytest1.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_v1.json 2> gpurun_out/r02_bench_v1.err
echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_ref_v1.json 2> gpurun_out/r02_ref_v1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:year_sweep -c 1 -f -o gpurun_out/prof_year \
    python scripts/bench_extras.py year=1024,365 > gpurun_out/year_under_ncu.log 2>&1
ls -la gpurun_out | tail -20
