#!/usr/bin/env python
"""Aggregate an `ncu --metrics smsp__inst_executed.sum --csv` launch list by kernel: launches, warp instructions, share.
python tools/inst_shares.py gpurun_out/x_inst.csv"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    with open(sys.argv[1]) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    agg = OrderedDict()
    n = 0
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "smsp__inst_executed.sum":
            continue
        name = re.sub(r"<.*", "", re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += float(r["Metric Value"].replace(",", "")); n += 1
    tot = sum(a[1] for a in agg.values()) or 1.0
    print("launches %d  warp instructions %.1f M" % (n, tot / 1e6))
    print("%-42s %6s %12s %12s %6s" % ("kernel", "n", "Minst", "Minst/launch", "share"))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-42s %6d %12.2f %12.2f %5.1f%%" % (name[:42], a[0], a[1] / 1e6, a[1] / a[0] / 1e6, 100 * a[1] / tot))


if __name__ == "__main__":
    main()
