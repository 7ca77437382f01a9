#!/usr/bin/env python
"""One line per profiled launch of an ncu report: duration, DRAM bytes, L2 hit rate, warp instructions, active lanes per
instruction, issue-slot utilisation, occupancy.   python tools/ncu_table.py report.ncu-rep [kernel-regex]"""
import csv
import re
import subprocess
import sys

COLS = [("us", "gpu__time_duration.sum"), ("dramR_MB", "dram__bytes_read.sum"), ("dramW_MB", "dram__bytes_write.sum"),
        ("L2hit%", "lts__t_sector_hit_rate.pct"), ("L1hit%", "l1tex__t_sector_hit_rate.pct"), ("Minst", "smsp__inst_executed.sum"),
        ("lanes", "smsp__thread_inst_executed_per_inst_executed.ratio"), ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {}
    for name, key in COLS:
        for i, h in enumerate(hdr):
            if h == key:
                idx[name] = i
    print("%-22s %-14s " % ("kernel", "grid x block") + " ".join("%9s" % n for n, _ in COLS))
    for r in rows[2:]:
        kn = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("aicp::", "")
        if pat and not pat.search(kn):
            continue
        vals = []
        for name, _ in COLS:
            if name not in idx or r[idx[name]] == "":
                vals.append("%9s" % "-")
                continue
            v = float(r[idx[name]].replace(",", ""))
            u = units[idx[name]]
            if name in ("us", "dramR_MB", "dramW_MB"):
                v *= SCALE.get(u, 1.0)
            if name == "Minst":
                v *= 1e-6
            vals.append("%9.2f" % v)
        g = r[hdr.index("Grid Size")].replace(" ", "").strip("()").split(",")[0] + "x" + r[hdr.index("Block Size")].replace(" ", "").strip("()").split(",")[0]
        print("%-22s %-14s " % (kn[:22], g) + " ".join(vals))


if __name__ == "__main__":
    main()
