#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_pytest5.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r02_pytest5.log | cut -c1-300
cat gpurun_out/parity_observed.json | python -c "
import json,sys; d=json.load(sys.stdin)
for k,v in d.items(): print(k, {a:(round(b,6) if isinstance(b,float) else b) for a,b in v.items() if a in ('loss_max_rel','grad_norm_max_rel','grad_elem_max_rel_l2','bn_running_max_rel','grad_elem_worst')})"
