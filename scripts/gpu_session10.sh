#!/bin/bash
# r02 session 10: 4-CTA-cluster B multicast in the resident-A kernel (variants 7 / 8) vs the pair kernel (4)
mkdir -p gpurun_out
for V in 7 8 4; do
  echo "== variant $V"
  SNB_TC3_VERBOSE=1 SNB_TC3_VARIANT=$V timeout 120 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "xf or consumer_side" 2>&1 | grep -E "passed|failed|FAILED|^E  |Error|gemm3:" | cut -c1-300
  SNB_TC3_VARIANT=$V timeout 60 python scripts/run_gemm_once.py 4 2>&1 | tail -2 | sed 's/.*wgrad/wgrad/'
done
SNB_TC3_VARIANT=7 timeout 60 python scripts/tc3_timeline.py 0 > gpurun_out/tc3_timeline_v7_nostore.txt 2>&1
