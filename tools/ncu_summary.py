"""Turn ncu outputs (launch-list CSV, raw CSV of a --set full capture) into the markdown tables kept
under profiles/.   python tools/ncu_summary.py launches.csv raw.csv > profiles/rNN_summary.md"""
import collections
import csv
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, mi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        per.setdefault(r[ii], {"k": r[ki].split("(")[0]})[r[mi]] = float(r[vi].replace(",", ""))
    agg = collections.OrderedDict()
    for p in per.values():
        a = agg.setdefault(p["k"], collections.Counter())
        a["n"] += 1
        for m in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum"):
            a[m] += p.get(m, 0.0)
    tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
    print("| kernel | launches | total us | us/launch | share | DRAM rd MB/launch | DRAM wr MB/launch | Mwarp-inst/launch |")
    print("|---|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1]["gpu__time_duration.sum"]):
        n = a["n"]
        print(f"| {k} | {n} | {a['gpu__time_duration.sum']/1e3:.1f} | {a['gpu__time_duration.sum']/1e3/n:.1f} | "
              f"{100*a['gpu__time_duration.sum']/tot:.1f}% | {a['dram__bytes_read.sum']/1e6/n:.1f} | "
              f"{a['dram__bytes_write.sum']/1e6/n:.1f} | {a['smsp__inst_executed.sum']/1e6/n:.2f} |")


WANT = [
    ("gpu__time_duration.sum", "duration (ncu unit)"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (SFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor inst %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per instruction"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/shared data-pipe wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "shared-load wavefronts"),
    ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "global-load wavefronts"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def full(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        print(f"\n### {name}  (grid {r[hdr.index('Grid Size')]}, block {r[hdr.index('Block Size')]})\n")
        print("| metric | value |")
        print("|---|---|")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print(f"| {label} | {r[i]} {units[i]} |")
        stalls = []
        for i, h in enumerate(hdr):
            if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h[34:-23]))
                except ValueError:
                    pass
        top = ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:5])
        print(f"| top stall reasons (warps per issue) | {top} |")


if __name__ == "__main__":
    print("## Launch list (ncu --metrics gpu__time_duration.sum ..., --clock-control none; cold-cache, serialised: compare shares)\n")
    launches(sys.argv[1])
    if len(sys.argv) > 2:
        print("\n## Full captures (ncu --set full --clock-control none --import-source on)")
        full(sys.argv[2])
