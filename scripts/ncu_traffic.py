#!/usr/bin/env python
"""Write / update profiles/r02_ncu_traffic.json - the NAMED file bench.py reads its `roofline.traffic` figures from - out of
`ncu --set full` captures: dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by kernel.
usage: python scripts/ncu_traffic.py gpurun_out/prof_gemm2.ncu-rep [...]        (run in the build container: ncu -i needs no GPU)
Keys: gemm2_<variant> for the four trunk-shaped training GEMMs (template arguments -> variant), fused_eval2_full,
composite_fwd / composite_bwd, year_sweep; gemm2_trunk_mean = launch-count-weighted mean of the variants over one step."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
VARIANT = {"<0, 0, 3, 0>": "fwd_bn_stats", "<0, 0, 1, 0>": "fwd_sin", "<0, 1, 2, 0>": "dgrad_cos_bnsums", "<1, 1, 0, 0>": "wgrad_splitk",
           "<0, 0, 3, 1>": "fwd_xf_streamed_bn_stats"}
# GEMM launches of one training step by variant (profiles/r02_step_table_v1.txt); fwd_xf_* = resident-A kernel gemm_tc3.cu
STEP_COUNTS = {"fwd_bn_stats": 4, "fwd_sin": 15, "fwd_xf_bn_stats": 6, "fwd_xf_storeY": 6, "dgrad_cos_bnsums": 17, "wgrad_splitk": 15}
_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def key_of(name):
    if "gemm2_bf16_kernel" in name:
        m = re.search(r"<[^>]*>", name)
        return "gemm2_" + VARIANT.get(m.group(0), m.group(0)) if m else None
    if "gemm3_xf_kernel" in name:
        return "gemm3"
    if "fused_eval2" in name:
        return "fused_eval2_full"
    if "composite_fwd" in name:
        return "composite_fwd"
    if "composite_bwd" in name:
        return "composite_bwd"
    if "year_sweep" in name:
        return "year_sweep"
    if "heads_composite" in name:
        return "heads_composite"
    return None


def main(paths):
    d = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            continue
        head, units = rows[0], rows[1]
        col = {c: i for i, c in enumerate(head)}
        for r in rows[2:]:
            k = key_of(r[col["Kernel Name"]])
            if k is None:
                continue
            rd = float(r[col["dram__bytes_read.sum"]]) * _UNIT[units[col["dram__bytes_read.sum"]]]
            wr = float(r[col["dram__bytes_write.sum"]]) * _UNIT[units[col["dram__bytes_write.sum"]]]
            if k == "gemm3":      # the same kernel with / without the write-back of the activated operand (trunk shape: 403 MB)
                k = "gemm2_fwd_xf_storeY" if wr > 1.5 * rd else "gemm2_fwd_xf_bn_stats"
            d[k] = {"dram_bytes": rd + wr, "read": rd, "write": wr, "duration_us": float(r[col["gpu__time_duration.sum"]]) *
                    {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0),
                    "kernel": r[col["Kernel Name"]][:90], "source": os.path.basename(path)}
    have = [k for k in STEP_COUNTS if "gemm2_" + k in d]
    if have:
        tot = sum(STEP_COUNTS[k] for k in have)
        d["gemm2_trunk_mean"] = {"dram_bytes": sum(STEP_COUNTS[k] * d["gemm2_" + k]["dram_bytes"] for k in have) / tot,
                                 "source": "launch-count-weighted mean of " + ", ".join(have)}
    json.dump(d, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps({k: v["dram_bytes"] for k, v in d.items()}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
