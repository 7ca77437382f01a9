"""Single-registration timing of the ICP loop under both loop schedules (persistent kernel / multi-launch) and both match
schedules, C3 pair k: prints the stats the library reports.  python tools/loop_probe.py [pair] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_pairs  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
pairs = load_pairs(k + 1)
import torch  # noqa: E402
import aicp_mapping_b200 as ab  # noqa: E402
from aicp_mapping_b200 import capi  # noqa: E402

p = pairs[k]
ovl = ab.B200Overlap()
ovl.computeOverlap(p["ref"], p["read"], p["ref_origin"], p["read_origin"])
ratio = ab.autotune_ratio(float(ovl.getOverlap()))
ref = torch.from_numpy(capi.to_xyzw(p["ref"])).cuda()
read = torch.from_numpy(capi.to_xyzw(p["read"])).cuda()
reg = ab.B200Registration()
reg.setConfig(ratio=ratio)
ref_T = None
for ls in (1, 0):
    for ms in (1, 2):
        reg.setLoopSchedule(ls); reg.setMatchSchedule(ms); reg.setProfiling(2 if ls == 1 else 0)
        rows = []
        for r in range(reps + 1):
            torch.cuda.synchronize()
            T = reg.registerClouds(ref, read)
            s = reg.stats
            if r:
                rows.append([s.ms_total, s.ms_setup, s.ms_iterations, s.ms_match, s.ms_select, s.ms_accumulate, s.ms_tail_pick,
                             s.ms_tail_select, s.ms_tail_solve])
        if ref_T is None:
            ref_T = T.copy()
        m = np.median(np.array(rows), axis=0)
        print("loop %d match %d: total %.3f setup %.3f loop %.3f | search %.3f quantile %.3f normal-eq %.3f | tails pick %.3f sel %.3f solve %.3f"
              " | iters %d launches %d same %s" % ((ls, ms) + tuple(m) + (s.iterations, s.gpu_launches, bool(np.array_equal(T, ref_T)))))
