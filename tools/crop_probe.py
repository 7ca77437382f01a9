#!/usr/bin/env python
"""Crop-box kernel probe: a random 10 485 760-point map on the device, crops of +-15 m; CUDA-event time per crop call and
achieved bytes/s against the measured HBM peak (one JSON line).  Also the command profiled under ncu for profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import aicp_mapping_b200 as ab

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10485760
rng = np.random.default_rng(1)
pts = np.ones((n, 4), dtype=np.float32)
pts[:, :3] = rng.uniform(-100, 100, (n, 3)).astype(np.float32)
pts[:, 2] *= 0.06
m = ab.B200Map()
m.updateCloud(torch.from_numpy(pts).cuda())
origin = np.eye(4, dtype=np.float32); origin[:3, 3] = [3.0, -2.0, 0.5]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    v = m.cropAround(15.0, origin)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms, wall = [], []
for _ in range(10):
    flush.zero_(); torch.cuda.synchronize()
    t0 = time.perf_counter(); ev0.record()
    v = m.cropAround(15.0, origin)
    ev1.record(); torch.cuda.synchronize()
    wall.append((time.perf_counter() - t0) * 1e3); ms.append(ev0.elapsed_time(ev1))
kept = v.shape[0]
b = 16.0 * n + 16.0 * kept
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
print(json.dumps({"kernel": "k_crop_box (+ reset, count read-back)", "points": n, "kept": kept, "algorithmic_bytes": b,
                  "ms_call_events_median": float(np.median(ms)), "ms_call_wall_median": float(np.median(wall)),
                  "achieved_GBps_events": b / (np.median(ms) * 1e-3) / 1e9, "peak_GBps": peak,
                  "frac": b / (np.median(ms) * 1e-3) / 1e9 / peak, "l2": "flushed before every call (256 MiB write)"}))
m.close()
