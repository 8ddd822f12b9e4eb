"""Per-section warp-stall summary of an ncu report captured with --import-source on (source page, SASS view):
   python scripts/ncu_stalls.py gpurun_out/prof_fused.ncu-rep
Prints the share of samples spent in mbarrier wait loops, the stall mix of the sine-epilogue section of the fused render
kernel and its hottest instructions."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S = lambda r: int(r[ix["# Samples"]] or 0)
tot = sum(S(r) for r in data)
print(rows[0][1][:70], "total samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = [i for i, r in enumerate(data) if "LDTM.x32" in r[ix["Source"]]]
if idx:
    a = idx[0] - 12
    b = [i for i, r in enumerate(data) if "LDTM.x16" in r[ix["Source"]]][0] - 8
    sec = data[a:b]
    s = sum(S(r) for r in sec)
    print("sine section %.1f%% of all samples, %d instructions" % (100 * s / tot, len(sec)))
    agg = collections.Counter()
    for r in sec:
        for h in stalls:
            agg[h] += int(r[ix[h]] or 0)
    print("  " + ", ".join("%s %.1f%%" % (h[6:], 100 * v / s) for h, v in agg.most_common(7)))
    for r in sorted(sec, key=lambda r: -S(r))[:8]:
        print("    ", r[ix["Address"]][-5:], "%.1f%%" % (100 * S(r) / tot), r[ix["Source"]][:60])
for i, r in enumerate(data):
    src = r[ix["Source"]]
    if "TRYWAIT" in src:
        ss = sum(S(data[j]) for j in range(i, min(i + 4, len(data))))
        if ss > tot * 0.003:
            print("  wait loop @%s %.1f%%  %s" % (r[ix["Address"]][-5:], 100 * ss / tot, src[:70]))
