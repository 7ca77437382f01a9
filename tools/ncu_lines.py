#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel in an ncu report (needs -lineinfo and
--import-source on):  python tools/ncu_lines.py report.ncu-rep kernel_name [top]"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, lines = "", None, []
    seen_first = False
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) >= 2 and r[0] == "Function Name":
            if seen_first and False:
                break
            continue
        if len(r) > 5 and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
            d = dict(zip(hdr[4:], r[4:]))
            try:
                lines.append((cur_file, int(r[0]), r[1].strip(), float(d["Instructions Executed"]), float(d["# Samples"])))
            except ValueError:
                pass
    # the report repeats per launch instance; keep the first occurrence of each (file,line)
    first = {}
    for f, ln, src, ins, smp in lines:
        first.setdefault((f, ln), (src, ins, smp))
    tot_i = sum(v[1] for v in first.values()) or 1
    tot_s = sum(v[2] for v in first.values()) or 1
    print("total warp instructions %.0f, samples %.0f" % (tot_i, tot_s))
    for (f, ln), (src, ins, smp) in sorted(first.items(), key=lambda x: -x[1][1])[:top]:
        print("%5.1f%% inst %5.1f%% stall  %-12s:%-4d %s" % (100 * ins / tot_i, 100 * smp / tot_s, f, ln, src[:95]))


if __name__ == "__main__":
    main()
