#!/bin/bash
mkdir -p gpurun_out
for V in 4 7; do for PF in 0 1; do
  echo "== variant $V prefetch $PF"
  SNB_TC3_PREFETCH=$PF SNB_TC3_VARIANT=$V timeout 120 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "xf or consumer_side" 2>&1 | grep -E "passed|failed|FAILED|^E  |Error|gemm3:" | cut -c1-300
  SNB_TC3_PREFETCH=$PF SNB_TC3_VARIANT=$V timeout 60 python scripts/run_gemm_once.py 4 2>&1 | tail -2 | sed 's/.*wgrad/wgrad/'
done; done
SNB_TC3_VARIANT=4 timeout 60 python scripts/tc3_timeline.py 0 > gpurun_out/tc3_timeline_v4pf_nostore.txt 2>&1
SNB_TC3_VARIANT=7 timeout 60 python scripts/tc3_timeline.py 0 > gpurun_out/tc3_timeline_v7pf_nostore.txt 2>&1
