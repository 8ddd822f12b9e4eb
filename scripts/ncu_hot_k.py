#!/usr/bin/env python
"""Top sampled SASS instructions (warp-stall samples) and shared-memory bank-conflict sources of the k-th kernel in an
.ncu-rep (needs --import-source on).   usage: python scripts/ncu_hot_k.py rep k [top]"""
import csv, io, subprocess, sys
path, k = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
lines = out.splitlines()
starts = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
starts.append(len(lines))
s, e = starts[k], starts[k + 1]
print(lines[s][:150])
rows = list(csv.reader(io.StringIO("\n".join(lines[s + 1:e]))))
h = rows[0]
ci = {c: i for i, c in enumerate(h)}
I = lambda r, c: int(r[ci[c]] or 0)
tot = sum(I(r, "# Samples") for r in rows[1:])
stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
print("total samples", tot)
idx = sorted(range(1, len(rows)), key=lambda i: -I(rows[i], "# Samples"))[:top]
for i in sorted(idx):
    r = rows[i]
    st = sorted(((I(r, c), c[6:]) for c in stall_cols), reverse=True)[:3]
    print("%5d %5.1f%%  line %4d  %-80s %s" % (I(r, "# Samples"), 100.0 * I(r, "# Samples") / max(tot, 1), i, r[ci["Source"]].strip()[:80],
                                    " ".join("%s:%d" % (n, v) for v, n in st if v)))
print("shared-memory wavefronts: excessive / total per instruction (top 12)")
ex = sorted(range(1, len(rows)), key=lambda i: -I(rows[i], "L1 Wavefronts Shared Excessive"))[:12]
for i in sorted(ex):
    r = rows[i]
    if I(r, "L1 Wavefronts Shared Excessive"):
        print("   line %4d  %-80s excessive %d of %d (ideal %d)" % (i, r[ci["Source"]].strip()[:80], I(r, "L1 Wavefronts Shared Excessive"),
              I(r, "L1 Wavefronts Shared"), I(r, "L1 Wavefronts Shared Ideal")))
