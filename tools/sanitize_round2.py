#!/usr/bin/env python
"""Small invocations of what round 2 added, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_round2.py
the persistent loop kernel (per-thread and tile search, spread 1 .. 8 through the cloud sizes), the fused quantile + normal-equation
kernel of the multi-launch loop, the batch over a device list, the incremental reference append, the TMA crop."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aicp_mapping_b200 as ab  # noqa: E402
from aicp_mapping_b200 import synth  # noqa: E402


def main():
    p = synth.make_pair(2, 0, 8192)
    reg = ab.B200Registration(device=0)
    reg.setConfig(ratio=0.6, max_iterations=6)
    for loop in (2, 1):
        for ms in (1, 2):
            reg.setLoopSchedule(loop); reg.setMatchSchedule(ms)
            for n in (8192, 1000, 300, 33):
                T = reg.registerClouds(p["ref"], p["read"][:: 8192 // n])
    print("loop kernels ok", reg.stats.iterations)
    reg.setLoopSchedule(0); reg.setMatchSchedule(0)
    pairs = [(p["ref"], p["read"])] * 5
    reg.registerBatch(pairs, ratios=[0.6] * 5, streams=2)
    reg.registerBatch(pairs, ratios=[0.6] * 5, streams=2, devices=[0])
    print("batches ok")
    base, extra = p["ref"][:6000], p["ref"][6000:]
    lo, hi = p["ref"].min(0), p["ref"].max(0)
    base = np.concatenate([base, lo[None], hi[None]], 0)
    reg.setReference(base)
    reg.registerToReference(p["read"])
    for part in np.array_split(extra, 3):
        info = reg.appendToReference(part)
    reg.registerToReference(p["read"])
    print("append ok", info.incremental, info.n_recomputed, info.n_total)
    crop = ab.B200CropBox(device=0)
    from aicp_mapping_b200 import filtering
    big = np.tile(p["ref"], (6, 1)).astype(np.float32)
    for n in (len(big), 4097, 2048, 5):
        g = crop.filter(big[:n], -3.0, 3.0, np.float32([0.0, 0.0, 0.3]), np.float32([0.5, 0.2, 0.6]))
    print("crop ok", g.shape)
    reg.close(); crop.close()


if __name__ == "__main__":
    main()
