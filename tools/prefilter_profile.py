#!/usr/bin/env python
"""One steady-state pre-filter call bracketed by cudaProfilerStart/Stop (the command profiled under ncu), plus a plain
timing line.   python tools/prefilter_profile.py [2|3] [reps]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import aicp_mapping_b200 as ab  # noqa: E402
from aicp_mapping_b200 import capi, synth  # noqa: E402


def main():
    cfgid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    raw = synth.raw_sweep(cfgid, 0)
    d = torch.from_numpy(capi.to_xyzw(raw["cloud"])).cuda()
    pf = ab.B200Prefilter(device=0)
    if len(sys.argv) > 3:
        pf._lib.aicp_b200_set_knn_schedule(pf._h, int(sys.argv[3]))      # 1 warp per query, 2 tile per warp
    ms = []
    for _ in range(3):
        pf.filter(d, keep_on_device=True)
        ms.append(round(pf.info.ms_total, 3))
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    pf.filter(d, keep_on_device=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    t0 = time.perf_counter()
    dev = 0.0
    for _ in range(reps):
        pf.filter(d, keep_on_device=True)
        dev += pf.info.ms_total
    wall = (time.perf_counter() - t0) / reps * 1e3
    print(json.dumps({"cloud": raw["name"], "n": int(d.shape[0]), "first_calls_ms": ms, "device_ms": dev / reps, "wall_ms": wall,
                      "n_sampled": int(pf.info.n_sampled), "n_out": int(pf.info.n_out), "passes": int(pf.info.passes),
                      "launches": int(pf.info.gpu_launches), "knn_schedule": int(sys.argv[3]) if len(sys.argv) > 3 else 0}))
    pf.close()


if __name__ == "__main__":
    main()
