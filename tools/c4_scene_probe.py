"""How well does the C4 scene constrain the pose?  Registers the 122 880-point reading against the 10 485 760-point map for a
few trials, with and without the ground hits in the reading, and prints the pose error against the injected prior error."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aicp_mapping_b200 as ab
from aicp_mapping_b200 import synth

n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 10485760
reg = ab.B200Registration()
for trial in (0, 1):
    for rg in (False, True):
        case = synth.make_map_case(n_map=n_map, n_read=122880, trial=trial, n_poses=2, remove_ground=rg)
        reg.setReference(case["map"])
        for ratio in (0.5, 0.7):
            reg.setConfig(ratio=ratio, max_iterations=20)
            for k, rd in enumerate(case["readings"]):
                T = reg.registerToReference(rd["read"])
                d = T.astype(np.float64) @ np.linalg.inv(rd["T_true"])
                e0 = np.linalg.inv(rd["T_true"])
                ang = np.degrees(np.arccos(min(1.0, (np.trace(d[:3, :3]) - 1) / 2)))
                print("trial %d pose %d ground_removed %s ratio %.1f: iterations %2d  error %.4f m %.4f deg  (prior error %.3f m)  used %.3f" %
                      (trial, k, rg, ratio, reg.stats.iterations, np.linalg.norm(d[:3, 3]), ang, np.linalg.norm(e0[:3, 3]), reg.getWeightedPointUsedRatio()), flush=True)
reg.close()
