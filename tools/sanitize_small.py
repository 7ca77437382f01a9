#!/usr/bin/env python
"""Small invocations of every entry point added for the widened rows, for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
(small clouds: the sanitizer slows kernels down by one to two orders of magnitude)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aicp_mapping_b200 as ab  # noqa: E402
from aicp_mapping_b200 import synth  # noqa: E402

MODEL = os.path.join(ROOT, "tests", "golden", "svm_models", "svm_1000training_thresh50_cross_validation_opencv3.xml")


def main():
    rng = np.random.default_rng(0)
    raw = synth.raw_sweep(2, 0, n_sweeps=1)["cloud"]
    raw[::91] = np.nan
    pf = ab.B200Prefilter(device=0)
    for cloud in (rng.uniform(0, 4, (1, 3)), rng.uniform(0, 4, (33, 3)), rng.uniform(0, 4, (700, 3)), raw[:5000], raw):
        out = pf.filter(cloud.astype(np.float32), view_point=[0.0, 0.0, 0.6])
        pf.segments()
        pf.voxelGrid(cloud.astype(np.float32))
    print("prefilter ok", out.shape, pf.info.n_clusters)
    p = synth.make_pair(2, 0, 8192)
    PA, PB = synth.rigid(*p["ref_origin"]), synth.rigid(*p["read_origin"])
    al = ab.B200Alignability(device=0, svm_model=MODEL)
    print("fov", al.overlapFilter(p["ref"], p["read"], PA, PB, 30.0, 270.0)[0])
    print("alignability", al.alignabilityFilter(p["ref"], p["read"], PA, PB))
    print("risk", al.computeAlignmentRisk(p["ref"], p["read"], PA, PB, 30.0, 270.0, 60.0))
    reg = ab.B200Registration(device=0)
    res = reg.pipelineBatch([(p["ref"], p["read"])] * 3, [(PA, PB)] * 3, MODEL, 30.0, 270.0, risk_threshold=1.0, streams=2)
    print("pipeline", res[1], res[2], res[3], res[5])
    acc = ab.B200VelodyneAccumulator(batch_size=3, device=0)
    for s in range(3):
        acc.processLidar(raw[s::3], synth.rigid(0.3 * s, 0, 0.6, 0, 0, 0.1 * s))
    print("accumulate", acc.download().shape)
    m = ab.B200Map(device=0)
    m.updateCloud(raw)
    print("map prefilter", m.prefilter().n_out, m.size())
    for x in (pf, al, reg, acc, m):
        x.close()


if __name__ == "__main__":
    main()
