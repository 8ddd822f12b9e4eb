#!/usr/bin/env python
"""Top sampled SASS instructions (warp-stall samples) of the first kernel in an .ncu-rep (needs --import-source on)."""
import csv, io, subprocess, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
lines = out.splitlines()
# first kernel only
start = 1
end = len(lines)
for i in range(2, len(lines)):
    if lines[i].startswith('"Kernel Name"'):
        end = i
        break
rows = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
h = rows[0]
ci = {c: i for i, c in enumerate(h)}
tot = sum(int(r[ci["# Samples"]] or 0) for r in rows[1:])
stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
print(lines[0][:160]); print("total samples", tot)
idx = sorted(range(1, len(rows)), key=lambda i: -int(rows[i][ci["# Samples"]] or 0))[:top]
for i in sorted(idx):
    r = rows[i]
    st = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print("%5d %5.1f%%  line %4d  %-70s %s" % (int(r[ci["# Samples"]]), 100.0 * int(r[ci["# Samples"]]) / max(tot, 1), i, r[ci["Source"]].strip()[:70],
                                    " ".join("%s:%d" % (n, v) for v, n in st if v)))
