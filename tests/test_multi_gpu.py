"""Two-GPU parity of the data-parallel training step (runs only on a box with >= 2 GPUs: `gpurun --gpus 2`)."""
import json
import os
import subprocess
import sys

import pytest
import torch as t

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", [("fp32",), ("bf16",), ("bf16", "graph")])
def test_sync_bn_data_parallel_step_equals_single_device_step(mode):
    if t.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "tests", "multi_gpu", "syncbn_parity.py")] + list(mode)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, (r.stdout[-2000:], r.stderr[-2000:])
    res = json.loads(lines[-1])
    assert res["ok"], res
