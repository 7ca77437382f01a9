#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total and mean device time,
share of the whole.  python tools/launch_shares.py gpurun_out/x_launches.csv [first_launch_id [last_launch_id]]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        i = int(r["ID"])
        if lo <= i <= hi:
            name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
            name = re.sub(r"<.*", "", name)
            rows.append((i, name, float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for i, name, ns, g, b in rows:
        a = agg.setdefault(name, [0, 0.0, g, b, ns, 0.0])
        a[0] += 1; a[1] += ns; a[4] = min(a[4], ns); a[5] = max(a[5], ns)
    tot = sum(a[1] for a in agg.values()) or 1.0
    print("launches %d  (ids %d..%d)  total device time %.1f us (cold-cache, serialised: compare shares)" %
          (len(rows), rows[0][0], rows[-1][0], tot / 1e3))
    print("%-42s %6s %10s %9s %9s %9s %6s  %s" % ("kernel", "n", "total_us", "mean_us", "min_us", "max_us", "share", "grid x block (first)"))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-42s %6d %10.1f %9.2f %9.2f %9.2f %5.1f%%  %s x %s" % (name[:42], a[0], a[1] / 1e3, a[1] / a[0] / 1e3, a[4] / 1e3, a[5] / 1e3,
                                                                 100 * a[1] / tot, a[2], a[3]))


if __name__ == "__main__":
    main()
