#!/usr/bin/env python
"""Opcode histogram per kernel of the shipped library (cuobjdump -sass): the tcgen05 / TMEM / TMA mnemonics that prove which
hardware paths the kernels use (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAREDG = TMA tensor copies,
UBLKCP = bulk copies, SYNCS = mbarrier, MUFU = special-function unit).
usage: python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.txt          (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "season_nerf_b200", "lib", "libseason_nerf_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP", "SYNCS", "MUFU", "HMMA", "FFMA", "DFMA",
       "LDG", "STG", "LDS", "STS", "RED", "ATOM", "SHFL", "BAR", "USETMAXREG", "UTCBAR", "UTCATOMSWS")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    kern, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", kern)
            hist[kern] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Za-z0-9_]+)*)", line)
        if m and kern:
            hist[kern][m.group(1)] += 1
    print("# %s\n# instructions per kernel; full mnemonics for the tensor-core / TMEM / TMA / mbarrier groups, totals for the rest" % os.path.relpath(LIB, ROOT))
    for k, h in hist.items():
        tot = sum(h.values())
        groups = collections.OrderedDict()
        for key in KEY:
            sel = {op: n for op, n in h.items() if op.split(".")[0].startswith(key)}
            if sel:
                groups[key] = sel
        print("\n%s  (%d instructions)" % (k, tot))
        for key, sel in groups.items():
            if key in ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP", "MUFU", "HMMA", "USETMAXREG"):
                print("   " + "  ".join("%s x%d" % (op, n) for op, n in sorted(sel.items())))
            else:
                print("   %s* x%d" % (key, sum(sel.values())))


if __name__ == "__main__":
    main()
