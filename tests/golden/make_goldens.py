"""Freezes oracle outputs for seeded inputs as golden vectors: tests/golden/icp_goldens.json.

    python tests/golden/make_goldens.py

The reference's own implementation of the path cannot be built in this image (libpointmatcher / libnabo / octomap / Eigen /
PCL are absent) and its single test points to data outside the repository (aicp_core/test/aicp_test.cpp:50-57), so there
are no reference-produced vectors to commit.  These vectors are produced by oracle/ (the CPU restatement) and pin BOTH
sides: `-m "not gpu"` tests check that the oracle still reproduces them (no silent drift of the checker), `-m gpu` tests
check the CUDA path against them without running the oracle at all.  Inputs are regenerated from seeds
(aicp_mapping_b200.synth); a CRC of every input cloud is stored so that a change of the generators is detected as such.

Per case: final transform (float32 bit patterns), iteration count, stop reason, per-iteration trimmed threshold (bits) and
inlier count, CRC32 of the correspondence indices of every iteration, CRC32 of the reference normals and of the output
cloud, the octree-overlap voxel counts and the auto-tuned ratio derived from them."""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = [  # name, (config, trial, n_points) or "c1:<reading>", ratio (None: auto-tuned from the overlap), extra config
    ("cube_t0_4000", (5, 0, 4000), 0.7, {}),
    ("cube_t3_full_autotuned", (5, 3, None), None, {}),
    ("cube_t1_ratio025", (5, 1, 6000), 0.25, {}),
    ("vlp16_t0_8192", (2, 0, 8192), None, {}),
    ("hdl64_t0_20000", (3, 0, 20000), None, {}),
    ("hdl64_t2_8192_knn10", (3, 2, 8192), 0.6, {"knn_normals": 10}),
    ("c1_scan01", "c1:1", 0.7, {}),
    ("cube_t4_counter3", (5, 4, 5000), 0.6, {"max_iterations": 3}),
]


def crc(a):
    return int(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def inputs(spec):
    from aicp_mapping_b200 import synth
    if isinstance(spec, str):
        return synth.c1_pair(int(spec.split(":")[1]))
    return synth.make_pair(*spec)


def bits(a):
    return [int(x) for x in np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).ravel()]


def main():
    from oracle import oracle as orc
    out = {}
    for name, spec, ratio, extra in CASES:
        p = inputs(spec)
        ov, counts = orc.overlap(p["ref"], p["ref_origin"], p["read"], p["read_origin"])
        auto, _ = orc.autotune_ratio(float(ov))
        r = float(auto) if ratio is None else ratio
        o = orc.icp(p["ref"], p["read"], orc.default_config(ratio=r, threads=os.cpu_count() or 1, **extra),
                    want_trace_idx=True, want_normals=True)
        assert o.rc == 0, (name, o.error)
        out[name] = dict(spec=spec if isinstance(spec, str) else list(spec), ratio=r, config=extra,
                         input_crc=[crc(p["ref"]), crc(p["read"])],
                         overlap_bits=bits([ov])[0], overlap_counts=list(counts), autotuned_ratio_bits=bits([auto])[0],
                         iterations=int(o.iterations), stop_reason=int(o.stop_reason), T_bits=bits(o.T),
                         limit_bits=[bits([t["limit_d2"]])[0] for t in o.trace], n_used=[int(t["n_used"]) for t in o.trace],
                         match_crc=[crc(row) for row in o.trace_idx], normals_crc=crc(o.normals), reading_crc=crc(o.reading),
                         weighted_point_used_ratio_bits=bits([o.weighted_point_used_ratio])[0])
        print(name, "iterations", o.iterations, "ratio %.6f" % r, "overlap %.4f" % float(ov))
    with open(os.path.join(HERE, "icp_goldens.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
