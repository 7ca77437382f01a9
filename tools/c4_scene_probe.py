"""How well does the C4 scene constrain the pose?  Registers 122 880-point readings against the 10 485 760-point map (campus +
street clutter) for several trials and prints the pose error left against the injected prior error."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aicp_mapping_b200 as ab
from aicp_mapping_b200 import synth

n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 10485760
reg = ab.B200Registration()
reg.setConfig(ratio=0.5, max_iterations=20)
for trial in range(5):
    for nc in (1500, 4000):
        case = synth.make_map_case(n_map=n_map, n_read=122880, trial=trial, n_poses=3, n_clutter=nc)
        reg.setReference(case["map"])
        for k, rd in enumerate(case["readings"]):
            T = reg.registerToReference(rd["read"])
            d = T.astype(np.float64) @ np.linalg.inv(rd["T_true"])
            e0 = np.linalg.inv(rd["T_true"])
            ang = np.degrees(np.arccos(min(1.0, (np.trace(d[:3, :3]) - 1) / 2)))
            print("trial %d clutter %d pose %d: iterations %2d  error %.4f m %.4f deg  (prior error %.3f m)" %
                  (trial, nc, k, reg.stats.iterations, np.linalg.norm(d[:3, 3]), ang, np.linalg.norm(e0[:3, 3])), flush=True)
reg.close()
