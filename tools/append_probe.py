"""Cost of appending an aligned 100k-point cloud to the 10 485 760-point reference: incremental update (csrc/append.cu) against
the full rebuild it replaces.  One JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import aicp_mapping_b200 as ab
from aicp_mapping_b200 import capi, synth

n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 10485760
case = synth.make_map_case(n_map=n_map, n_read=122880, trial=0, n_poses=3, n_clutter=1500)
mp = torch.from_numpy(capi.to_xyzw(case["map"])).cuda()
reg = ab.B200Registration()
reg.setConfig(ratio=0.5)
reg.setProfiling(2)
reg.setReference(mp)
rd = case["readings"]
T = reg.registerToReference(rd[0]["read"])          # first build: allocates
reg.setReference(mp)
T = reg.registerToReference(rd[0]["read"])          # the full rebuild an append replaces, buffers in place
full_build_ms = reg.stats.ms_index + reg.stats.ms_normals
out = []
for k in range(3):
    # the aligned reading (what App merges into the map), thinned to 100 000 points
    T = reg.registerToReference(rd[k]["read"])
    aligned = reg.getOutputReading()[:100000]
    inside = np.all((aligned[:, :3] >= case["map"].min(0)) & (aligned[:, :3] <= case["map"].max(0)), axis=1)
    aligned = np.ascontiguousarray(aligned[inside])
    dev = torch.from_numpy(aligned).cuda()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info = reg.appendToReference(dev)
    wall = (time.perf_counter() - t0) * 1e3
    out.append({"appended": int(len(aligned)), "incremental": int(info.incremental), "recomputed": int(info.n_recomputed),
                "ms_device": round(float(info.ms), 3), "ms_wall": round(wall, 3), "n_total": int(info.n_total)})
T = reg.registerToReference(rd[0]["read"])
print(json.dumps({"map_points": n_map, "full_rebuild_ms (index + normals)": round(float(full_build_ms), 3), "appends": out,
                  "registration_after_ms": round(float(reg.stats.ms_total), 3)}))
reg.close()
