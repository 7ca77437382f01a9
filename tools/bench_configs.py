#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configurations that are not the bench.py headline (C3): one JSON line each.

    python tools/bench_configs.py [--configs 2,4,5,overlap] [--streams 8]

  C2  VLP-16 ANYmal-shaped 32 768 x 32 768 pairs, batched (registrations/s)
  C4  122 880-point readings against a fixed 10 485 760-point map on ONE GPU: map index + normals built once
      (reported), then registrations/s and ms per registration with the map resident
  C5  validation sweep: 4096 registrations of 38 400-point cube pairs (16 distinct perturbations cycled), batched
  overlap  the octree-overlap parameter of the C3 pair (ms per call, voxel counts)
  voxelmap  the periodic re-filter of the merged map (app.cpp:486-493) at C4 size: VoxelGrid alone and the whole pre-filter on the
      10 485 760-point map, device-resident, against the HBM roofline
  risk  App::computeAlignmentRisk (FOV overlap -> alignability -> SVM) for the C2 and C3 pairs, ms per call, CPU oracle beside it
  prefilter  regionGrowingUniformPlaneSegmentationFilter (VoxelGrid 0.08 + k-30 normals + region growing) on the raw clouds
      App feeds it: 7 accumulated VLP-16 sweeps (~200 k points) and one HDL-64 sweep (~250 k points); ms per call, the
      voxel-grid stage alone against the HBM roofline, and the CPU oracle on the host cores beside it
All inputs are device-resident when timing starts unless the line says e2e.  Single GPU; see bench.py for multi-GPU.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="2,4,4crop,5,5step,5risk,frames,overlap,prefilter,risk,voxelmap")
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--map-points", type=int, default=10485760)
    args = ap.parse_args()
    import torch
    import aicp_mapping_b200 as ab
    from aicp_mapping_b200 import capi, synth
    want = args.configs.split(",")
    reg = ab.B200Registration(device=0)
    ovl = ab.B200Overlap(device=0)

    def dev(a):
        return torch.from_numpy(capi.to_xyzw(a)).cuda()

    def batch_line(name, pairs, n_total, streams):
        ratios = []
        for p in pairs:
            ovl.computeOverlap(p["ref"], p["read"], p["ref_origin"], p["read_origin"])
            ratios.append(ab.autotune_ratio(float(ovl.getOverlap())))
        dpairs = [(dev(p["ref"]), dev(p["read"])) for p in pairs]
        order = [i % len(pairs) for i in range(n_total)]
        batch = [dpairs[i] for i in order]
        rat = [ratios[i] for i in order]
        reg.setConfig(max_iterations=20)
        reg.setProfiling(0)
        reg.registerBatch(batch[:4 * streams], ratios=rat[:4 * streams], streams=streams)       # warm-up (allocations)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        T, stats, status, ms = reg.registerBatch(batch, ratios=rat, streams=streams)
        wall = time.perf_counter() - t0
        iters = float(np.mean([s.iterations for s in stats]))
        err_t = max(float(np.linalg.norm((T[i].astype(np.float64) @ np.linalg.inv(pairs[order[i]]["T_true"]))[:3, 3])) for i in range(min(n_total, 64)))
        print(json.dumps({"config": name, "metric": "ICP registrations/sec", "value": n_total / (ms * 1e-3), "unit": "registrations/s",
                          "registrations": n_total, "distinct_pairs": len(pairs), "points_per_cloud": int(pairs[0]["ref"].shape[0]),
                          "streams": streams, "device_ms": ms, "wall_s": wall, "iterations_mean": iters,
                          "failed": int(np.count_nonzero(status)), "max_translation_error_m_first64": err_t,
                          "inputs": "device-resident"}), flush=True)

    if "2" in want:
        batch_line("C2 VLP-16 32768x32768", [synth.make_pair(2, t) for t in range(4)], 256, args.streams)
    if "5" in want:
        batch_line("C5 validation sweep, cube pairs 38400 pts", [synth.make_pair(5, t) for t in range(16)], 4096, args.streams)
    if "5step" in want:
        pairs5 = [synth.make_pair(5, t) for t in range(16)]
        d5 = [(dev(p["ref"]), dev(p["read"])) for p in pairs5]
        order = [i % 16 for i in range(4096)]
        reg.setConfig(max_iterations=20); reg.setProfiling(0)
        reg.aicpBatch([d5[i] for i in order[:64]], [(pairs5[i]["ref_origin"], pairs5[i]["read_origin"]) for i in order[:64]], streams=args.streams)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        T, ovp, stats, status, ms = reg.aicpBatch([d5[i] for i in order], [(pairs5[i]["ref_origin"], pairs5[i]["read_origin"]) for i in order], streams=args.streams)
        wall = time.perf_counter() - t0
        print(json.dumps({"config": "C5 validation sweep with the overlap parameter: 4096 AICP steps (overlap -> auto-tuned ratio -> registration) of cube pairs",
                          "metric": "AICP steps/sec", "value": 4096 / (ms * 1e-3), "unit": "steps/s", "device_ms": ms, "wall_s": wall,
                          "overlap_pct_range": [float(ovp.min()), float(ovp.max())], "failed": int(np.count_nonzero(status)),
                          "streams": args.streams, "inputs": "device-resident"}), flush=True)
    if "5risk" in want:
        # BASELINE config 5 as written: "overlap + alignment-risk" per pair, then the registration (App::runAicpPipeline with
        # failure_prediction_mode); threshold 1.0 so that every pair is also registered (the full cost per pair)
        model = os.path.join(ROOT, "tests", "golden", "svm_models", "svm_1000training_thresh50_cross_validation_opencv3.xml")
        pairs5 = [synth.make_pair(5, t) for t in range(16)]
        d5 = [(dev(p["ref"]), dev(p["read"])) for p in pairs5]
        poses5 = [(synth.rigid(*p["ref_origin"]), synth.rigid(*p["read_origin"])) for p in pairs5]
        n5 = 1024
        order = [i % 16 for i in range(n5)]
        reg.setConfig(max_iterations=20); reg.setProfiling(0)
        reg.pipelineBatch([d5[i] for i in order[:32]], [poses5[i] for i in order[:32]], model, 100.0, 360.0, 1.0, streams=args.streams)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        T, ovp, alp, risk, stats, status, ms = reg.pipelineBatch([d5[i] for i in order], [poses5[i] for i in order], model, 100.0, 360.0, 1.0,
                                                                 streams=args.streams)
        wall = time.perf_counter() - t0
        print(json.dumps({"config": "C5 validation sweep with overlap + alignment risk: %d pipeline steps (overlap -> FOV overlap -> alignability -> SVM -> registration) of cube pairs" % n5,
                          "metric": "AICP pipeline steps/sec", "value": n5 / (ms * 1e-3), "unit": "steps/s", "device_ms": ms, "wall_s": wall,
                          "overlap_pct_range": [float(ovp.min()), float(ovp.max())], "alignability_pct_range": [float(alp.min()), float(alp.max())],
                          "risk_range": [float(risk.min()), float(risk.max())], "pairs_with_risk_above_0.5": int((risk > 0.5).sum()),
                          "failed": int(np.count_nonzero(status)), "streams": args.streams, "inputs": "device-resident"}), flush=True)
    if "frames" in want:
        # whole frames, batched: raw 7-sweep VLP-16 accumulations (201 600 points) in, per pair pre-filter x 2 -> overlap ->
        # alignment risk -> registration, the filtered clouds never leaving the device (aicp_b200_pipeline_batch, prefilter_first)
        model = os.path.join(ROOT, "tests", "golden", "svm_models", "svm_1000training_thresh50_cross_validation_opencv3.xml")
        raws = [synth.raw_sweep(2, t) for t in range(4)]
        E = synth.rigid(0.08, -0.05, 0.01, 0.0, 0.0, 0.02)
        fr = [(dev(raws[k]["cloud"]), dev(synth.apply_T(E, raws[(k + 1) % 4]["cloud"]))) for k in range(4)]
        fp = [(synth.rigid(*raws[k]["origin"]), synth.rigid(*raws[(k + 1) % 4]["origin"])) for k in range(4)]
        nf = 128
        order = [i % 4 for i in range(nf)]
        reg.setConfig(max_iterations=20); reg.setProfiling(0)
        reg.pipelineBatch([fr[i] for i in order[:16]], [fp[i] for i in order[:16]], model, 30.0, 270.0, 1.0, streams=args.streams, prefilter_first=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        T, ovp, alp, risk, stats, status, ms = reg.pipelineBatch([fr[i] for i in order], [fp[i] for i in order], model, 30.0, 270.0, 1.0,
                                                                 streams=args.streams, prefilter_first=True)
        wall = time.perf_counter() - t0
        print(json.dumps({"config": "whole frames, batched: %d pairs of raw VLP-16 accumulations (201600 points each): pre-filter x 2 -> overlap -> alignment risk -> registration" % nf,
                          "metric": "AICP frames/sec", "value": nf / (ms * 1e-3), "unit": "frames/s", "device_ms": ms, "wall_s": wall,
                          "ms_per_frame": ms / nf, "filtered_points_first_pair": [int(x) for x in reg.n_filtered[0]],
                          "overlap_pct_range": [float(ovp.min()), float(ovp.max())], "alignability_pct_range": [float(alp.min()), float(alp.max())],
                          "iterations_mean": float(np.mean([s.iterations for s in stats])), "failed": int(np.count_nonzero(status)),
                          "streams": args.streams, "inputs": "raw clouds device-resident"}), flush=True)
    if "overlap" in want:
        p = synth.make_pair(3, 0)
        r, q = dev(p["ref"]), dev(p["read"])
        counts = ovl.computeOverlap(r, q, p["ref_origin"], p["read_origin"])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            ovl.computeOverlap(r, q, p["ref_origin"], p["read_origin"])
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 20 * 1e3
        print(json.dumps({"config": "overlap of the C3 pair (131072 + 131072 rays, 0.2 m voxels)", "metric": "ms per computeOverlap",
                          "value": ms, "unit": "ms", "overlap_pct": float(ovl.getOverlap()), "voxel_counts_AandB_A_B": [int(c) for c in counts],
                          "inputs": "device-resident, host synchronised per call"}), flush=True)
    if "prefilter" in want:
        from oracle import oracle as orc          # CPU baseline leg only
        ncpu = os.cpu_count() or 1
        pf = ab.B200Prefilter(device=0)
        for cfgid in (2, 3):
            raw = synth.raw_sweep(cfgid, 0)
            cloud = capi.to_xyzw(raw["cloud"])
            d = torch.from_numpy(cloud).cuda()
            for _ in range(3):
                pf.filter(d, keep_on_device=True)
            torch.cuda.synchronize()
            reps = 20
            dev_ms, t0 = 0.0, time.perf_counter()
            for _ in range(reps):
                pf.filter(d, keep_on_device=True)
                dev_ms += pf.info.ms_total
            wall_dev = (time.perf_counter() - t0) / reps * 1e3
            info = pf.info
            t0 = time.perf_counter()
            for _ in range(reps):
                out = pf.filter(cloud)
            wall_host = (time.perf_counter() - t0) / reps * 1e3
            t0 = time.perf_counter()
            for _ in range(reps):
                vg = pf.voxelGrid(d)
            vg_ms = (time.perf_counter() - t0) / reps * 1e3
            t0 = time.perf_counter()
            o = orc.prefilter(cloud, threads=ncpu)
            cpu_ms = (time.perf_counter() - t0) * 1e3
            same = bool(np.array_equal(out.view(np.uint32), o.cloud.view(np.uint32)))
            n = cloud.shape[0]
            vg_bytes = 16.0 * n + 8.0 * n + 16.0 * vg.shape[0]
            print(json.dumps({"config": "pre-filter (VoxelGrid 0.08 + k-30 normals + region growing) of a %s, %d points" % (raw["name"], n),
                              "metric": "ms per regionGrowingUniformPlaneSegmentationFilter", "value": dev_ms / reps, "unit": "ms",
                              "wall_ms_device_input": wall_dev, "wall_ms_host_input_and_output": wall_host,
                              "n_sampled": int(info.n_sampled), "n_clusters": int(info.n_clusters), "n_out": int(info.n_out),
                              "region_growing_passes": int(info.passes), "gpu_launches": int(info.gpu_launches),
                              "voxel_grid_alone": {"wall_ms_incl_download": vg_ms, "algorithmic_bytes": vg_bytes,
                                                   "note": "16 B/point read + 8 B/point key and index + 16 B/voxel written"},
                              "cpu_oracle_ms": cpu_ms, "cpu_cores": ncpu, "bit_identical_to_oracle": same,
                              "clouds_per_s": 1e3 / (dev_ms / reps)}), flush=True)
        pf.close()
    if "voxelmap" in want:
        case = synth.make_map_case(n_map=args.map_points, n_read=1024, trial=0, n_poses=1)
        mp = dev(case["map"])
        n = int(mp.shape[0])
        pf = ab.B200Prefilter(device=0)
        out = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        n_out = C.c_int64()
        L = capi.lib()

        def vg():
            rc = L.aicp_b200_voxel_grid(pf._h, C.c_void_p(mp.data_ptr()), n, C.c_float(0.08), C.c_void_p(out.data_ptr()), C.byref(n_out))
            assert rc == 0, L.aicp_b200_last_error(pf._h)
        for _ in range(3):
            vg()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = 10
        t0 = time.perf_counter()
        for _ in range(reps):
            vg()
        torch.cuda.synchronize()
        vg_ms = (time.perf_counter() - t0) / reps * 1e3
        m = int(n_out.value)
        # algorithmic bytes: read 16 B/point, write + read the 8 B (key, index) pair once, 4 radix passes of 16 B/pair each way are
        # the sort's own traffic (counted separately), gather 16 B/point for the centroids, 16 B/voxel written
        alg = 16.0 * n + 8.0 * n + 16.0 * n + 16.0 * m
        sort_bytes = 4 * 2 * 8.0 * n
        for _ in range(2):
            pf.filter(mp, keep_on_device=True)
        t0 = time.perf_counter()
        for _ in range(3):
            pf.filter(mp, keep_on_device=True)
        pf_ms = (time.perf_counter() - t0) / 3 * 1e3
        info = pf.info
        peak = 6549.4
        print(json.dumps({"config": "map re-filter at C4 size: %d-point map, 0.08 m voxels" % n, "metric": "ms per VoxelGrid of the map",
                          "value": vg_ms, "unit": "ms", "voxels": m,
                          "roofline": {"bound": "hbm", "algorithmic_bytes": alg, "achieved_GBps_wall": alg / (vg_ms * 1e-3) / 1e9, "peak_GBps": peak,
                                       "frac_wall": alg / (vg_ms * 1e-3) / 1e9 / peak, "radix_sort_bytes_not_counted": sort_bytes,
                                       "note": "wall clock of the synchronous C-ABI call, result left on the device"},
                          "whole_prefilter_ms_wall": pf_ms, "whole_prefilter_device_ms": float(info.ms_total), "n_clusters": int(info.n_clusters),
                          "n_out": int(info.n_out), "passes": int(info.passes), "inputs": "map device-resident"}), flush=True)
        pf.close()
        del mp, out
    if "risk" in want:
        from oracle import oracle as orc          # CPU baseline leg only
        from oracle import aicp_oracle_svm as svm_orc
        ncpu = os.cpu_count() or 1
        model = os.path.join(ROOT, "tests", "golden", "svm_models", "svm_1000training_thresh50_cross_validation_opencv3.xml")
        al = ab.B200Alignability(device=0, svm_model=model)
        for cfgid, rng_m, view in ((2, 30.0, 270.0), (3, 100.0, 360.0)):
            p = synth.make_pair(cfgid, 0)
            PA, PB = synth.rigid(*p["ref_origin"]), synth.rigid(*p["read_origin"])
            r, q = dev(p["ref"]), dev(p["read"])
            ov_pct = float(ovl.computeOverlap(r, q, p["ref_origin"], p["read_origin"]) and ovl.getOverlap())
            for _ in range(3):
                res = al.computeAlignmentRisk(r, q, PA, PB, rng_m, view, ov_pct)
            torch.cuda.synchronize()
            reps = 20
            t0 = time.perf_counter()
            for _ in range(reps):
                res = al.computeAlignmentRisk(r, q, PA, PB, rng_m, view, ov_pct)
            ms = (time.perf_counter() - t0) / reps * 1e3
            t0 = time.perf_counter()
            o_fov, o_a, o_b = orc.fov_overlap(p["ref"], p["read"], PA, PB, rng_m, view)
            o_al, o_match, o_info = orc.alignability(o_a, o_b, PA, PB, threads=ncpu)
            o_risk = svm_orc.test(svm_orc.load_model(model), np.array([[ov_pct, float(o_al)]]))[0]
            cpu_ms = (time.perf_counter() - t0) * 1e3
            print(json.dumps({"config": "computeAlignmentRisk (FOV overlap -> alignability -> SVM) of the %s" % p["name"],
                              "metric": "ms per computeAlignmentRisk", "value": ms, "unit": "ms", "fov_overlap_pct": float(res[0]),
                              "alignability_pct": float(res[1]), "risk": float(res[2]), "octree_overlap_pct": ov_pct,
                              "clusters_A_B_matched": list(o_info), "cpu_oracle_ms": cpu_ms, "cpu_cores": ncpu,
                              "identical_to_oracle": bool(res[0] == o_fov and res[1] == o_al and abs(res[2] - o_risk) < 1e-6),
                              "inputs": "device-resident, host synchronised per call (wall clock)"}), flush=True)
        al.close()
    if "4" in want:
        t0 = time.perf_counter()
        case = synth.make_map_case(n_map=args.map_points, n_read=122880, trial=0, n_poses=8, n_clutter=1500)
        gen_s = time.perf_counter() - t0
        mp = dev(case["map"])
        reads = [dev(r["read"]) for r in case["readings"]]
        reg.setConfig(ratio=0.5, max_iterations=20)       # app.cpp:123-127
        reg.setProfiling(2)
        reg.setReference(mp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reg.registerToReference(reads[0])                 # builds the map index + normals
        first_ms = (time.perf_counter() - t0) * 1e3
        build = dict(index_ms=reg.stats.ms_index, normals_ms=reg.stats.ms_normals)
        reg.setProfiling(0)
        for r in reads[:2]:
            reg.registerToReference(r)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n, iters, dev_ms, errs = 0, 0, 0.0, []
        for rep in range(4):
            for i, r in enumerate(reads):
                T = reg.registerToReference(r)
                n += 1; iters += reg.stats.iterations; dev_ms += reg.stats.ms_total
                d = T.astype(np.float64) @ np.linalg.inv(case["readings"][i]["T_true"])
                errs.append(float(np.linalg.norm(d[:3, 3])))
        wall = time.perf_counter() - t0
        print(json.dumps({"config": "C4 localisation: 122880-pt readings vs %d-pt fixed map, one GPU" % args.map_points,
                          "metric": "ICP registrations/sec", "value": n / wall, "unit": "registrations/s", "registrations": n,
                          "ms_per_registration_device": dev_ms / n, "ms_per_registration_wall": wall / n * 1e3,
                          "iterations_mean": iters / n, "map_build_once": build, "first_call_ms": first_ms,
                          "max_translation_error_m": max(errs), "synthetic_generation_s": gen_s,
                          "inputs": "device-resident, one registration at a time (latency mode)"}), flush=True)
    if "4crop" in want:
        # the way App does it (app.cpp:41-69): crop the whole map to +-15 m around the prior pose, register against the crop
        from aicp_mapping_b200 import filtering
        case = synth.make_map_case(n_map=args.map_points, n_read=122880, trial=0, n_poses=4, n_clutter=1500)
        mp = dev(case["map"])
        crop = ab.B200CropBox(device=0)
        reg.setConfig(ratio=0.5, max_iterations=20)
        reg.setProfiling(0)

        class DevView:
            def __init__(self, a, n): self._a, self.shape, self.dtype = a, (n, 4), "torch.float32"
            def data_ptr(self): return self._a
            def dim(self): return 2
            def is_contiguous(self): return True
        priors = []
        for r in case["readings"]:
            P = np.eye(4, dtype=np.float32); P[:3, 3] = np.asarray(r["read_origin"], dtype=np.float32)
            priors.append((filtering.euler_angles_xyz(P[:3, :3]), P[:3, 3].copy()))
        reads = [dev(r["read"]) for r in case["readings"]]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        crop_ms, reg_ms, kept, n = 0.0, 0.0, 0, 0
        for rep in range(5):
            for (rpy, t), rd in zip(priors, reads):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                addr, nk = crop.filter(mp, -15.0, 15.0, rpy, t, keep_on_device=True)
                t1 = time.perf_counter()
                reg.registerClouds(DevView(addr, nk), rd)
                t2 = time.perf_counter()
                if rep > 0:
                    crop_ms += (t1 - t0) * 1e3; reg_ms += (t2 - t1) * 1e3; kept += nk; n += 1
        peak = 6549.4
        bytes_per_crop = 16.0 * args.map_points + 16.0 * kept / n
        print(json.dumps({"config": "C4 as App runs it: crop the %d-pt map to +-15 m around the prior pose, register against the crop" % args.map_points,
                          "metric": "ms per (crop + registration)", "value": (crop_ms + reg_ms) / n, "unit": "ms",
                          "crop_ms_wall": crop_ms / n, "registration_ms_wall": reg_ms / n, "points_kept_mean": kept / n,
                          "crop_roofline": {"bound": "hbm", "algorithmic_bytes": bytes_per_crop, "achieved_GBps_wall": bytes_per_crop / (crop_ms / n * 1e-3) / 1e9,
                                            "peak_GBps": peak, "frac_wall": bytes_per_crop / (crop_ms / n * 1e-3) / 1e9 / peak,
                                            "note": "wall clock of the synchronous C-ABI call (launch + D2H of the count + sync included); kernel-only time is in profiles/"},
                          "inputs": "map device-resident"}), flush=True)
        crop.close()
    reg.close(); ovl.close()


if __name__ == "__main__":
    main()
