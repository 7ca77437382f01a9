#!/usr/bin/env python
"""One whole AICP frame as App::processCloud runs it (aicp_core/src/registration/app.cpp:283-420), stage by stage, on the GPU
and -- beside it -- through the CPU oracle: accumulate 7 VLP-16 sweeps (velodyne_accumulator.cpp:31-73) -> pre-filter
(app.cpp:102-110) -> octree overlap (:112-141) -> alignment risk (:143-185) -> auto-tuned registration (:187-216).
Every intermediate stays on the device; the per-stage wall times include the host synchronisation of each C-ABI call.

    python tools/bench_pipeline.py [--frames 20] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sweeps_for(rng, boxes, x0, synth):
    out = []
    for s in range(7):
        pose = synth.rigid(x0 + 0.35 * s, rng.uniform(-0.05, 0.05), 0.6, 0, 0, rng.uniform(-0.05, 0.05))
        world = synth.lidar_scan(pose, boxes, synth.VLP16_ELEV, 1800, rng, max_range=100.0, az_offset=rng.uniform(0, 0.01))
        local = ((world - pose[:3, 3]) @ pose[:3, :3]).astype(np.float32)
        out.append((local, pose))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import torch
    import aicp_mapping_b200 as ab
    from aicp_mapping_b200 import capi, synth
    model = os.path.join(ROOT, "tests", "golden", "svm_models", "svm_1000training_thresh50_cross_validation_opencv3.xml")
    rng = np.random.default_rng(77)
    boxes = synth.room_scene(rng)
    ref_batch = sweeps_for(rng, boxes, -3.0, synth)
    read_batch = sweeps_for(rng, boxes, -3.0 + 0.35 * 7, synth)
    E = synth.rigid(0.12, -0.08, 0.02, 0.005, -0.004, 0.03)           # odometry drift of the reading's poses
    read_batch = [(sw, E @ P) for sw, P in read_batch]
    ref_pose, read_pose = ref_batch[3][1], read_batch[3][1]

    acc_a, acc_b = ab.B200VelodyneAccumulator(device=0), ab.B200VelodyneAccumulator(device=0)
    pf_a, pf_b = ab.B200Prefilter(device=0), ab.B200Prefilter(device=0)
    ovl = ab.B200Overlap(device=0)
    al = ab.B200Alignability(device=0, svm_model=model)
    reg = ab.B200Registration(device=0)
    reg.setConfig(max_iterations=20)
    pinned = [[torch.from_numpy(capi.to_xyzw(sw)).pin_memory() for sw, _ in batch] for batch in (ref_batch, read_batch)]
    n_raw = sum(int(t.shape[0]) for t in pinned[1])

    def frame(timing=None):
        t = [time.perf_counter()]
        for acc, batch, pins in ((acc_a, ref_batch, pinned[0]), (acc_b, read_batch, pinned[1])):
            acc.clearCloud()
            for (sw, P), pin in zip(batch, pins):
                acc.processLidar(pin.numpy(), P)
        t.append(time.perf_counter())
        ref_f = pf_a.filter(acc_a.getCloud(), keep_on_device=True)
        read_f = pf_b.filter(acc_b.getCloud(), keep_on_device=True)
        t.append(time.perf_counter())
        ovl.computeOverlap(ref_f, read_f, ref_pose[:3, 3], read_pose[:3, 3])
        ov = float(ovl.getOverlap())
        t.append(time.perf_counter())
        fov, ali, risk = al.computeAlignmentRisk(ref_f, read_f, ref_pose, read_pose, 30.0, 270.0, ov)
        t.append(time.perf_counter())
        reg.setConfig(ratio=ab.autotune_ratio(ov))
        T = reg.registerClouds(ref_f, read_f)
        t.append(time.perf_counter())
        if timing is not None:
            timing.append(np.diff(t))
        return dict(n_ref=ref_f.shape[0], n_read=read_f.shape[0], overlap=ov, fov=float(fov), alignability=float(ali), risk=risk, T=T,
                    iterations=reg.stats.iterations)

    for _ in range(3):
        res = frame()
    torch.cuda.synchronize()
    timing = []
    t0 = time.perf_counter()
    for _ in range(args.frames):
        res = frame(timing)
    wall = (time.perf_counter() - t0) / args.frames * 1e3
    st = np.mean(timing, axis=0) * 1e3
    d = res["T"].astype(np.float64) @ E
    line = {"metric": "ms per AICP frame (accumulate 2 x 7 sweeps -> pre-filter x 2 -> overlap -> alignment risk -> registration)",
            "value": wall, "unit": "ms", "frames_per_s": 1e3 / wall,
            "stage_ms": {"accumulate_2x7_sweeps_from_pinned_host": st[0], "prefilter_x2": st[1], "octree_overlap": st[2],
                         "alignment_risk": st[3], "registration": st[4]},
            "raw_points_per_batch": n_raw, "prefiltered_points": [res["n_ref"], res["n_read"]], "overlap_pct": res["overlap"],
            "fov_overlap_pct": res["fov"], "alignability_pct": res["alignability"], "risk": res["risk"], "icp_iterations": res["iterations"],
            "residual_translation_m": float(np.linalg.norm(d[:3, 3])), "h2d_bytes_per_frame": 2 * n_raw * 16, "data": "synthetic VLP-16 room scene"}
    if not args.no_cpu:
        from oracle import oracle as orc                      # CPU baseline leg only
        from oracle import aicp_oracle_svm as svm_orc
        ncpu = os.cpu_count() or 1
        t0 = time.perf_counter()
        a = orc.accumulate_sweeps([s for s, _ in ref_batch], [p for _, p in ref_batch])
        b = orc.accumulate_sweeps([s for s, _ in read_batch], [p for _, p in read_batch])
        t1 = time.perf_counter()
        fa, fb = orc.prefilter(a, threads=ncpu).cloud, orc.prefilter(b, threads=ncpu).cloud
        t2 = time.perf_counter()
        o_ov, _ = orc.overlap(fa, ref_pose[:3, 3], fb, read_pose[:3, 3])
        t3 = time.perf_counter()
        o_fov, ka, kb = orc.fov_overlap(fa, fb, ref_pose, read_pose, 30.0, 270.0)
        o_al, _, _ = orc.alignability(ka, kb, ref_pose, read_pose, threads=ncpu)
        o_risk = svm_orc.test(svm_orc.load_model(model), np.array([[float(o_ov), float(o_al)]]))[0]
        t4 = time.perf_counter()
        o = orc.icp(fa, fb, orc.default_config(ratio=ab.autotune_ratio(float(o_ov)), threads=ncpu))
        t5 = time.perf_counter()
        line["cpu_oracle"] = {"ms": (t5 - t0) * 1e3, "cores": ncpu,
                              "stage_ms": {"accumulate": (t1 - t0) * 1e3, "prefilter_x2": (t2 - t1) * 1e3, "octree_overlap": (t3 - t2) * 1e3,
                                           "alignment_risk": (t4 - t3) * 1e3, "registration": (t5 - t4) * 1e3},
                              "identical": bool(res["overlap"] == float(o_ov) and res["alignability"] == float(o_al)
                                                and abs(res["risk"] - o_risk) < 1e-6 and np.array_equal(res["T"], o.T))}
    print(json.dumps(line), flush=True)
    for x in (acc_a, acc_b, pf_a, pf_b, ovl, al, reg):
        x.close()


if __name__ == "__main__":
    main()
