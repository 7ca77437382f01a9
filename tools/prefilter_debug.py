#!/usr/bin/env python
"""Stage-by-stage comparison of the CUDA pre-filter with the CPU oracle (debug aid; run on the GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aicp_mapping_b200 as ab  # noqa: E402
from aicp_mapping_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def diff(name, a, b):
    if a.shape != b.shape:
        print("  %-8s SHAPE %s vs %s" % (name, a.shape, b.shape))
        return False
    if a.size == 0:
        print("  %-8s empty, equal" % name)
        return True
    av, bv = (a.view(np.uint32), b.view(np.uint32)) if a.dtype == np.float32 else (a, b)
    bad = np.flatnonzero((av != bv).reshape(a.shape[0], -1).any(axis=1))
    print("  %-8s rows differing: %d of %d" % (name, bad.size, a.shape[0]))
    for i in bad[:4]:
        print("     row %d: gpu %s  oracle %s" % (i, a[i], b[i]))
    return bad.size == 0


def main():
    pf = ab.B200Prefilter(device=0)
    rng = np.random.default_rng(0)
    clouds = [("uniform 500", rng.uniform(0, 4, (500, 3)).astype(np.float32)),
              ("vlp16 1 sweep", synth.raw_sweep(2, 0, n_sweeps=1)["cloud"]),
              ("vlp16 7 sweeps", synth.raw_sweep(2, 0)["cloud"]),
              ("hdl64", synth.raw_sweep(3, 0)["cloud"])]
    for name, cloud in clouds:
        print(name, cloud.shape)
        try:
            vg = pf.voxelGrid(cloud)
            diff("voxelgrid", vg, orc.voxel_grid(cloud))
            out = pf.filter(cloud)
            sampled, normals, labels, clusters = pf.segments()
            o = orc.prefilter(cloud, threads=os.cpu_count() or 1)
            print("  info: sampled %d clusters %d out %d passes %d launches %d ms %.3f | oracle: %d %d %d" %
                  (pf.info.n_sampled, pf.info.n_clusters, pf.info.n_out, pf.info.passes, pf.info.gpu_launches, pf.info.ms_total,
                   o.sampled.shape[0], o.n_clusters, o.cloud.shape[0]))
            diff("sampled", sampled, o.sampled)
            diff("normals", normals, o.normals)
            diff("labels", labels, o.labels)
            diff("output", out, o.cloud)
        except Exception as e:  # noqa: BLE001
            print("  FAILED:", repr(e))
    pf.close()


if __name__ == "__main__":
    main()
