"""Regenerates the small fixtures under tests/golden/ from the read-only reference checkout.

    python tests/golden/make_fixtures.py [/root/reference]

/root/reference does not exist on the GPU box, so the fixtures are committed:
  c1_scans.npz                  the three planar sample scans aicp_core/data/scan_0{0,1,2}.csv (2162 x 2 each) as float32 --
                                BASELINE.json config 1 ("sample point clouds from aicp_core/data"); the 3-D samples
                                cloud_0{0,1,2}.vtk are missing blobs (aicp_core/data/.MISSING_LARGE_BLOBS)
  icp_autotuned.yaml            aicp_core/config/icp/icp_autotuned.yaml          (the per-call rewritten chain file)
  icp_autotuned_default.yaml    aicp_core/config/icp/icp_autotuned_default.yaml  (the template)
  icp_3D_cfg_trimmed.yaml       aicp_core/config/icp/icp_3D_cfg_trimmed.yaml     (a chain the B200 path must REJECT: it holds
                                MaxDensity / RandomSampling filters)
These are configuration/data files the parser and the C1 parity case must accept verbatim, not source code.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main(ref="/root/reference"):
    scans = {}
    for i in range(3):
        a = np.loadtxt(os.path.join(ref, "aicp_core/data/scan_0%d.csv" % i), delimiter=",").astype(np.float32)
        scans["scan_0%d" % i] = a
    np.savez_compressed(os.path.join(HERE, "c1_scans.npz"), **scans)
    for f in ("icp_autotuned.yaml", "icp_autotuned_default.yaml", "icp_3D_cfg_trimmed.yaml"):
        shutil.copyfile(os.path.join(ref, "aicp_core/config/icp", f), os.path.join(HERE, f))
    print({k: v.shape for k, v in scans.items()})


if __name__ == "__main__":
    main(*sys.argv[1:])
