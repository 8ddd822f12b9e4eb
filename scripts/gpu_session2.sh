#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_net_tool_gpu.py tests/test_kernels_gpu.py tests/test_network_gpu.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r02_pytest2.log
tail -15 gpurun_out/r02_pytest2.log
timeout 600 python scripts/bench_extras.py year=1024,365 > gpurun_out/r02_year_v2.json 2> gpurun_out/r02_year_v2.err
cat gpurun_out/r02_year_v2.json | cut -c1-900
timeout 600 ncu --set full --clock-control none --import-source on -k regex:year_sweep -c 1 -f -o gpurun_out/prof_year_v2 \
    python scripts/bench_extras.py year=1024,365 > gpurun_out/year_under_ncu.log 2>&1
