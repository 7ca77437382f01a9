#!/usr/bin/env python
"""A/B throughput of experiment builds of the library (aicp_mapping_b200.build.build(variant=..., defines=...)):
    python tools/ab_bench.py base leaf4 ...      -> one line per variant ("base" = the product library)
Each variant runs in its own process (one library per process)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
from aicp_mapping_b200 import capi
v = sys.argv[1]
if v != "base":
    capi.LIB_PATH = capi.LIB_PATH.replace("libaicp_b200.so", "libaicp_b200_%%s.so" %% v)
sys.argv = ["bench.py", "--steps", "5", "--warmup", "3", "--streams", sys.argv[2], "--no-cpu", "--pairs", "32"] + sys.argv[3:]
import runpy
runpy.run_path(%r, run_name="__main__")
''' % (ROOT, os.path.join(ROOT, "bench.py"))

extra = [a for a in sys.argv[1:] if a.startswith("--")]
for v in [a for a in sys.argv[1:] if not a.startswith("--")]:
    for S in ("1", "8"):
        r = subprocess.run([sys.executable, "-c", CHILD, v, S] + extra, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            st = d["stage_ms_per_registration"]
            print("%-12s S=%s value %8.1f  lat %.3f ms  k_match %.4f ms  normals %.3f index %.3f" % (
                v, S, d["value"], d["latency_single_stream"]["ms_per_registration"], d["roofline"]["avg_launch_ms"], st["normals"], st["index"]), flush=True)
        except Exception as e:
            print(v, S, "failed", e, r.stderr[-800:], flush=True)
