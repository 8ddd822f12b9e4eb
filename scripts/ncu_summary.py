#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of numbers the roofline discussion needs.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [more.ncu-rep ...] > profiles/rNN_xxx.txt"""
import csv
import io
import subprocess
import sys

EXACT = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.max",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def summarize(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print("%s: no data" % path)
        return
    head, units = rows[0], rows[1]
    col = {c: i for i, c in enumerate(head)}
    for r in rows[2:]:
        print("== %s :: %s" % (path.split("/")[-1], r[col["Kernel Name"]][:100]))
        for k in EXACT:
            if k in col:
                print("  %-78s %s %s" % (k, r[col[k]], units[col[k]]))
        stalls = []
        for c, i in col.items():
            if c.startswith(STALL_PREFIX) and c.endswith("_per_issue_active.ratio") and "not_issued" not in c:
                try:
                    stalls.append((float(r[i]), c[len(STALL_PREFIX):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  top warp stall reasons (warps stalled per issue-active cycle): " +
              ", ".join("%s %.2f" % (n, v) for v, n in stalls[:6]))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        summarize(p)
