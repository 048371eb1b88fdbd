"""Hot regions of one kernel from an ncu report (source page, SASS view).
   python tools/ncu_hot.py report.ncu-rep <kernel regex> [min_pct]
Prints consecutive-instruction regions with similar execution counts: share of executed warp
instructions, share of stall samples, and the two dominant stall reasons."""
import csv, subprocess, sys, io

rep, rx = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr) and r[0].startswith("0x")]
ia, isrc = hdr.index("Address"), hdr.index("Source")
iss, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iss]) for r in data)
toti = sum(int(r[ie]) for r in data)
print(f"{rows[hi-1][1][:60] if hi else ''}  samples {tot}  warp-instr {toti}  sass lines {len(data)}")
base = int(data[0][ia], 16)
regs, cur = [], None
for r in data:
    s, e, off = int(r[iss]), int(r[ie]), int(r[ia], 16) - base
    if cur is None or abs(e - cur["e"]) > 0.2 * max(e, cur["e"], 1):
        cur = {"start": off, "e": e, "n": 0, "s": 0, "ex": 0, "st": [0] * len(stall_cols), "src": r[isrc].strip()[:44]}
        regs.append(cur)
    cur["n"] += 1; cur["s"] += s; cur["ex"] += e; cur["end"] = off
    for k, i in enumerate(stall_cols):
        cur["st"][k] += int(r[i])
for g in regs:
    if 100 * g["s"] / tot >= minpct or 100 * g["ex"] / toti >= minpct:
        top = sorted(zip(g["st"], [hdr[i][6:] for i in stall_cols]), reverse=True)[:2]
        print(f"{g['start']:05x}-{g['end']:05x} n={g['n']:3d} exec={g['e']/1e6:8.3f}M instr={100*g['ex']/toti:5.1f}% "
              f"samples={100*g['s']/tot:5.1f}%  {top[0][1]}:{100*top[0][0]/max(1,g['s']):.0f}% {top[1][1]}:{100*top[1][0]/max(1,g['s']):.0f}%  | {g['src']}")
