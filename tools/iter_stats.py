#!/usr/bin/env python
"""Per-iteration statistics of one registration on the GPU (trimmed threshold, inliers, size of the pose increment)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import aicp_mapping_b200 as ab
from aicp_mapping_b200 import synth
cfg, trial = int(sys.argv[1]) if len(sys.argv) > 1 else 3, int(sys.argv[2]) if len(sys.argv) > 2 else 0
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_pair
p = load_pair(trial) if cfg == 3 else synth.make_pair(cfg, trial)
ov = ab.B200Overlap(); ov.computeOverlap(p["ref"], p["read"], p["ref_origin"], p["read_origin"])
ratio = ab.autotune_ratio(float(ov.getOverlap()))
reg = ab.B200Registration(); reg.setConfig(ratio=ratio); reg.setProfiling(2)
for rep in range(2):
    T = reg.registerClouds(p["ref"], p["read"])
print("ratio %.6f iterations %d ms_total %.3f match %.3f" % (ratio, reg.stats.iterations, reg.stats.ms_total, reg.stats.ms_match))
prev = np.eye(4)
for it, t in enumerate(reg.trace()):
    Ti = np.asarray(t["T_iter"], dtype=np.float64)
    step = np.linalg.norm((Ti @ np.linalg.inv(prev))[:3, 3]); prev = Ti
    print("it %2d  limit_d %.4f m  n_used %d / %d  step %.4f m" % (it, float(t["limit_d2"]) ** 0.5, t["n_used"], reg.stats.n_read, step))
