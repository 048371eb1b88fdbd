"""Cluster the SASS instructions of one kernel in an ncu report by execution count (loop nesting levels) and print
the hot regions.  python tools/ncu_sass_levels.py report.ncu-rep <kernel-substring> [--list]"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if "Instructions Executed" in r)
ix, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
ist = hdr.index("Warp Stall Sampling (All Samples)")
data = []
for r in rows[rows.index(hdr) + 1:]:
    try:
        data.append((int(r[ix]), r[isrc].strip(), int(r[ist] or 0)))
    except (ValueError, IndexError):
        continue
tot = sum(d[0] for d in data)
tst = sum(d[2] for d in data)
print(f"total warp instructions {tot}, {len(data)} SASS instructions, {tst} stall samples")
lv = collections.OrderedDict()
for n, s_, st in data:
    e = lv.setdefault(n, [0, 0])
    e[0] += 1
    e[1] += st
for n, (c, st) in sorted(lv.items(), key=lambda x: -x[0] * x[1][0])[:18]:
    print(f"count {n:>12} x{c:>4} instrs = {n*c/tot*100:5.2f}% of instructions, {st/max(tst,1)*100:5.2f}% of samples")
if "--list" in sys.argv:
    for n, s_, st in data:
        print(f"{n:>12} {st:>6}  {s_}")
