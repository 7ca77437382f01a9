"""Freezes oracle outputs of the widened rows (pre-filter, FOV overlap, alignability, sweep accumulation) for seeded inputs:
tests/golden/filter_goldens.json.

    python tests/golden/make_filter_goldens.py

Same role as make_goldens.py: PCL / Eigen are not installed, so these are oracle-produced (parity unpinned); `-m "not gpu"`
checks that the oracle still reproduces them, `-m gpu` checks the CUDA path against them WITHOUT the oracle running."""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def crc(a):
    return int(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def fbits(x):
    return int(np.float32(x).view(np.uint32))


def pose_from(origin, yaw=0.0):
    from aicp_mapping_b200 import synth
    return synth.rigid(origin[0], origin[1], origin[2], 0.0, 0.0, yaw)


def prefilter_inputs(name):
    from aicp_mapping_b200 import synth
    if name == "vlp16_raw_3sweeps":
        r = synth.raw_sweep(2, 11, n_sweeps=3)
        return r["cloud"], r["origin"].astype(np.float32)
    if name == "hdl64_raw_third":
        r = synth.raw_sweep(3, 12)
        return r["cloud"][::3].copy(), None
    if name == "cube":
        return synth.cube_cloud(), None
    raise KeyError(name)


def pair_inputs(name):
    from aicp_mapping_b200 import synth
    if name == "vlp16_pair_16384":
        p = synth.make_pair(2, 4, 16384)
        return p["ref"], p["read"], pose_from(p["ref_origin"]), pose_from(p["read_origin"], 0.04), 30.0, 270.0
    if name == "hdl64_pair_40000":
        p = synth.make_pair(3, 5, 40000)
        return p["ref"], p["read"], pose_from(p["ref_origin"]), pose_from(p["read_origin"]), 100.0, 360.0
    raise KeyError(name)


def sweep_inputs():
    from aicp_mapping_b200 import synth
    rng = np.random.default_rng(4242)
    boxes = synth.room_scene(rng)
    sweeps, poses = [], []
    for s in range(4):
        pose = synth.rigid(-2.0 + 0.35 * s, rng.uniform(-0.05, 0.05), 0.6, rng.uniform(-0.02, 0.02), rng.uniform(-0.02, 0.02), rng.uniform(-3, 3))
        world = synth.lidar_scan(pose, boxes, synth.VLP16_ELEV, 600, rng, max_range=100.0)
        local = ((world - pose[:3, 3]) @ pose[:3, :3]).astype(np.float32)
        local[::37] *= np.float32(15.0)
        sweeps.append(local); poses.append(pose)
    return sweeps, poses


PREFILTER_CASES = ["vlp16_raw_3sweeps", "hdl64_raw_third", "cube"]
PAIR_CASES = ["vlp16_pair_16384", "hdl64_pair_40000"]


def main():
    from oracle import oracle as orc
    out = {"prefilter": {}, "pairs": {}, "accumulate": {}}
    for name in PREFILTER_CASES:
        cloud, vp = prefilter_inputs(name)
        o = orc.prefilter(cloud, viewpoint=vp, threads=os.cpu_count() or 1)
        assert o.rc == 0
        out["prefilter"][name] = dict(input_crc=crc(cloud), n_sampled=int(o.sampled.shape[0]), n_clusters=o.n_clusters, n_out=int(o.cloud.shape[0]),
                                      sampled_crc=crc(o.sampled), normals_crc=crc(o.normals), labels_crc=crc(o.labels), cloud_crc=crc(o.cloud))
    for name in PAIR_CASES:
        a, b, PA, PB, rng_m, view = pair_inputs(name)
        ov, fa, fb = orc.fov_overlap(a, b, PA, PB, rng_m, view)
        al, matching, info = orc.alignability(fa, fb, PA, PB, threads=os.cpu_count() or 1)
        out["pairs"][name] = dict(input_crc=[crc(a), crc(b)], fov_overlap_bits=fbits(ov), accepted=[int(fa.shape[0]), int(fb.shape[0])],
                                  accepted_crc=[crc(fa), crc(fb)], alignability_bits=fbits(al), matching=[int(m) for m in matching], info=list(info))
    sweeps, poses = sweep_inputs()
    acc = orc.accumulate_sweeps(sweeps, poses)
    out["accumulate"]["vlp16_4sweeps"] = dict(input_crc=[crc(s) for s in sweeps], n=int(acc.shape[0]), cloud_crc=crc(acc))
    path = os.path.join(HERE, "filter_goldens.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
