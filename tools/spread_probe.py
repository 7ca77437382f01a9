"""Search-phase latency of the persistent loop kernel against the number of reading points and the spread (lanes per query):
the C3 pair with the reading thinned to n points.  python tools/spread_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_pairs  # noqa: E402

pairs = load_pairs(1)
import torch  # noqa: E402
import aicp_mapping_b200 as ab  # noqa: E402
from aicp_mapping_b200 import capi  # noqa: E402

p = pairs[0]
ref = torch.from_numpy(capi.to_xyzw(p["ref"])).cuda()
for n in (131072, 65536, 32768, 16384):
    read = torch.from_numpy(capi.to_xyzw(p["read"][:: 131072 // n])).cuda()
    for spread in ("auto", "1", "2", "4", "8"):
        if spread == "auto":
            os.environ.pop("AICP_B200_SPREAD", None)
        else:
            os.environ["AICP_B200_SPREAD"] = spread
        reg = ab.B200Registration()
        reg.setConfig(ratio=0.6)
        reg.setLoopSchedule(2); reg.setMatchSchedule(1)
        rows = []
        for r in range(6):
            torch.cuda.synchronize()
            T = reg.registerClouds(ref, read)
            s = reg.stats
            if r:
                rows.append([s.ms_total, s.ms_setup, s.ms_iterations, s.ms_match, s.ms_select, s.ms_accumulate])
        m = np.median(np.array(rows), axis=0)
        it = s.iterations
        print("n %6d spread %4s: total %.3f setup %.3f loop %.3f | per iteration: search %.1f us quantile %.1f us normal-eq %.1f us | iters %d"
              % (n, spread, m[0], m[1], m[2], 1e3 * m[3] / it, 1e3 * m[4] / it, 1e3 * m[5] / it, it))
        reg.close()
