"""Reference-side k-NN schedules of one C3 registration (identical outputs): total / setup time per k-NN schedule.
python tools/knn_probe.py [pair] [reps]; AICP_B200_KNN_LISTS=heap|reg selects the list layout of the tile kernel."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_pairs  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
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
for ks in (1, 2):
    reg.setKnnSchedule(ks)
    rows = []
    for r in range(reps + 1):
        torch.cuda.synchronize()
        T = reg.registerClouds(ref, read)
        s = reg.stats
        if r:
            rows.append([s.ms_total, s.ms_setup, s.ms_iterations])
    if ref_T is None:
        ref_T = T.copy()
    m = np.median(np.array(rows), axis=0)
    print("lists %s knn schedule %d: total %.3f setup %.3f loop %.3f | launches %d same %s" % (
        os.environ.get("AICP_B200_KNN_LISTS", "auto"), ks, m[0], m[1], m[2], s.gpu_launches, bool(np.array_equal(T, ref_T))))
