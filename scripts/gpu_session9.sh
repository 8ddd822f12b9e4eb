#!/bin/bash
# r02 session 9: A/B of the resident-A kernel variants (epilogue warps / B ring depth / L2 prefetch of the next tile)
mkdir -p gpurun_out
for V in 0 4 5 6; do
  echo "== variant $V"
  SNB_TC3_VARIANT=$V timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "xf or consumer_side" 2>&1 | grep -E "passed|failed|FAILED|^E  |Error" | cut -c1-300
  SNB_TC3_VARIANT=$V timeout 120 python scripts/run_gemm_once.py 4 2>&1 | tail -2 | sed 's/.*wgrad/wgrad/'
done
