#!/usr/bin/env python
"""bench.py -- ICP registrations/second of the AICP hot path on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--pairs P]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload: BASELINE.json configs[2], "Velodyne HDL-64 KITTI-shaped synthetic clouds (~128k pts, ground removed)
point-to-plane ICP" -- the configuration the metric ("128k pts") is quoted on; 131 072 x 131 072 points per pair, chain of
aicp_core/config/icp/icp_autotuned.yaml with epsilon 0 and the ratio auto-tuned from the octree overlap.
A step = one pass of the hot path (registerClouds: index + normals + ICP loop + output cloud) over P cloud pairs per GPU,
registered concurrently on S CUDA streams through aicp_b200_register_batch (a single 128k-point registration does not
fill a B200; independent pairs are the unit the reference's validation sweeps and sequence runs iterate over).

  value     registrations/s with both clouds already resident in HBM (device pointers through the C ABI); device time of
            each step from CUDA events spanning all the library's streams; L2 flushed before every step.
  e2e       the same through the plugin call with pinned HOST buffers (H2D of both clouds and D2H of the result inside).
  roofline  dominant kernel k_match (exact NN + transform + histogram): algorithmic bytes 24 B/reading point per launch
            over its CUDA-event duration, against the measured HBM copy peak.
  cpu_baseline  the CPU oracle (a restatement of the reference's libpointmatcher path; the real one cannot be built here)
            on the same pair set on this box's host cores.
Every step registers the same D = 16 DISTINCT pairs (4 ray-cast scenes x 4 erroneous prior poses, separate buffers), P / D times
each; the reference arm (--impl reference) cycles through the same D pairs with the same ratios, so both arms average over the
same ICP trajectories.
Multi-GPU: independent pairs are sharded over ranks with no data-path collective (weak scaling).  With more than one rank
the line also carries "sharded": ONE registration whose reading is sharded over all ranks (BASELINE.json configs[3]) -- the C3
pair and a 122 880-point reading against a 10 485 760-point map -- timed against the same registration on one GPU and
checked bit for bit against it.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS = 131072
CACHE = os.path.join(tempfile.gettempdir(), "aicp_b200_bench_cache")


N_VARIANTS = 4          # erroneous prior poses per ray-cast scene
D_MAX = 16              # distinct pairs per step


def _scene_path(scene, n_points):
    return os.path.join(CACHE, "c3_s%d_n%d_v%d.npz" % (scene, n_points, N_VARIANTS))


def _make_scene(job):
    scene, n_points = job
    from aicp_mapping_b200 import synth
    path = _scene_path(scene, n_points)
    if os.path.exists(path):
        return path
    vs = synth.make_pair(3, scene, n_points, variants=N_VARIANTS)
    tmp = path + ".%d.tmp.npz" % os.getpid()
    np.savez(tmp, ref=vs[0]["ref"], ref_origin=vs[0]["ref_origin"], read=np.stack([v["read"] for v in vs]),
             read_origin=np.stack([v["read_origin"] for v in vs]))
    os.replace(tmp, path)
    return path


def load_pairs(n_distinct, n_points=N_POINTS, generate=True):
    """The D distinct C3 pairs of a step: pair k = scene k // 4 (one ray cast of synth.make_pair(3, scene)), prior-pose variant
    k % 4.  Scenes are cached under the temp directory; missing ones are ray-cast in parallel by forked workers (call this
    before CUDA is initialised).  generate=False (ranks other than local rank 0): wait for the files instead."""
    import multiprocessing as mp
    os.makedirs(CACHE, exist_ok=True)
    n_scenes = (n_distinct + N_VARIANTS - 1) // N_VARIANTS
    missing = [(sc, n_points) for sc in range(n_scenes) if not os.path.exists(_scene_path(sc, n_points))]
    if missing and generate:
        if len(missing) == 1:
            _make_scene(missing[0])
        else:
            with mp.get_context("fork").Pool(min(len(missing), host_cores())) as pool:
                pool.map(_make_scene, missing)
    t0 = time.time()
    while any(not os.path.exists(_scene_path(sc, n_points)) for sc in range(n_scenes)):
        if time.time() - t0 > 900:
            raise RuntimeError("bench inputs were not generated within 900 s")
        time.sleep(0.5)
    pairs = []
    for k in range(n_distinct):
        z = np.load(_scene_path(k // N_VARIANTS, n_points))
        v = k % N_VARIANTS
        pairs.append(dict(ref=z["ref"], read=np.ascontiguousarray(z["read"][v]), ref_origin=z["ref_origin"],
                          read_origin=z["read_origin"][v]))
    return pairs


def load_pair(trial, n_points=N_POINTS):
    """Pair `trial` of the distinct set (tools/)."""
    return load_pairs(trial + 1, n_points)[trial]


def host_cores():
    """Host threads this process may use (the affinity mask, not the machine's core count)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def profiled_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            d = json.load(f)["k_match_tile"]
            return float(d["dram_bytes_per_launch"]), d["source"]
    except Exception:
        return None, None


def profiled_instructions(iterations_mean):
    """Warp instructions of one registration with `iterations_mean` iterations, from the committed ncu launch list of the batch
    schedule (profiles/roofline_traffic.json), or (None, None)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            d = json.load(f)["warp_instructions"]
        return d["setup"] + d["cold_iteration"] + max(0.0, iterations_mean - 1.0) * d["warm_iteration"], d["source"]
    except Exception:
        return None, None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def issue_slots(iterations_mean, regs_per_s_per_gpu, clocks):
    """The path is bound by instruction issue and latency, not by HBM: warp instructions per second of the batched leg against
    the GPU's issue peak (148 SMs x 4 schedulers x SM clock).  Instruction counts are the profiled ones (ncu), not live."""
    inst, src = profiled_instructions(iterations_mean)
    mhz = (clocks or {}).get("sm_mhz") or 0.0
    if inst is None or mhz <= 0:
        return None
    peak = 148 * 4 * mhz * 1e6
    achieved = inst * regs_per_s_per_gpu
    return {"warp_instructions_per_registration": inst, "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "G warp-instructions/s",
            "frac": achieved / peak, "source": src}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken in [t0, t1] (perf_counter times; the sampler is started well before the timed
        region because nvidia-smi needs a few hundred ms to deliver its first line)."""
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons = [], [], set()
        for ts, ln in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.05):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smmax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smmax)), "reasons": sorted(reasons), "samples": len(sm)}


def oracle_register(pair, ratio, threads, reading_normals):
    """One registration by the CPU oracle; returns (seconds, ICP iterations)."""
    from oracle import oracle as orc
    cfg = orc.default_config(ratio=ratio, threads=threads, reading_normals=reading_normals, use_kdtree=1)
    t0 = time.perf_counter()
    out = orc.icp(pair["ref"], pair["read"], cfg, want_reading=True)
    sec = time.perf_counter() - t0
    if out.rc != 0:
        raise RuntimeError("oracle failed: " + out.error)
    return sec, out.iterations


def oracle_ratios(pairs):
    """The auto-tuned ratio of every pair (octree overlap -> clamp -> 6-digit text round trip) from the CPU oracle; equal to
    the GPU arm's by the overlap parity tests."""
    from oracle import oracle as orc
    out = []
    for p in pairs:
        ov, _ = orc.overlap(p["ref"], p["ref_origin"], p["read"], p["read_origin"])
        out.append(float(orc.autotune_ratio(float(ov))[0]))
    return out


REF_REGS_PER_STEP = 4      # the reference arm's step: a bounded sample of the GPU arm's step (4 of its P registrations)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  libpointmatcher/libnabo/octomap cannot be built in
    this image (no Eigen/PCL/yaml-cpp either), so this times the oracle port with every host thread it can use, on the
    GPU arm's pair set: step s registers pairs 4s .. 4s+3 (mod D) of the same D distinct pairs with the same ratios, and
    --steps / --warmup are honoured.  The reading-side SurfaceNormal filter of the reference's chain, which PointToPlane never
    reads and the GPU arm skips, is NOT in the timed value (it is reported beside it)."""
    if rank != 0:
        return
    D = min(args.pairs, D_MAX)
    pairs = load_pairs(D)
    ratios = oracle_ratios(pairs)
    cores = host_cores()
    R = REF_REGS_PER_STEP
    k = 0
    for _ in range(args.warmup):
        for _ in range(R):
            oracle_register(pairs[k % D], ratios[k % D], cores, 0)
            k += 1
    t_wall0 = time.perf_counter()
    sec, iters = 0.0, 0
    for _ in range(args.steps):
        for _ in range(R):
            s_, it_ = oracle_register(pairs[k % D], ratios[k % D], cores, 0)
            sec += s_; iters += it_
            k += 1
    n_reg = args.steps * R
    value = n_reg / sec
    sec_full, _ = oracle_register(pairs[0], ratios[0], cores, 1)
    sec_lean, _ = oracle_register(pairs[0], ratios[0], cores, 0)
    sample = ("%d registrations per step, cycling through the GPU arm's %d distinct C3 pairs (131072 x 131072 pts, same auto-tuned "
              "ratios, mean %.2f ICP iterations), kd-tree oracle WITHOUT the dead reading-side SurfaceNormal filter, OpenMP over "
              "queries on %d threads" % (R, D, iters / n_reg, cores))
    cfg = workload_config(R, "n/a (CPU)", D, iters / n_reg)
    line = {"impl": "reference", "metric": "ICP registrations/sec (128k pts)", "value": value, "unit": "registrations/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": "registrations/s", "cores": cores, "kind": "port", "sample": sample,
                             "pair0_with_reading_normals": 1.0 / sec_full, "pair0_without_reading_normals": 1.0 / sec_lean},
            "e2e": {"value": value, "unit": "registrations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_wall0}
    print(json.dumps(line), flush=True)


def workload_config(pairs, l2, distinct, iterations_mean):
    return {"workload": "C3: Velodyne HDL-64 KITTI-shaped synthetic clouds, ground removed, 131072 reading x 131072 reference points, "
                        "point-to-plane ICP chain of icp_autotuned.yaml (knn 20 normals, exact 1-NN epsilon 0, trimmed ratio auto-tuned "
                        "from the octree overlap, <=20 iterations, differential stop 0.001 rad / 0.01 m / 4)",
            "pairs_per_gpu_per_step": pairs, "pairs_distinct": distinct, "iterations_mean": iterations_mean,
            "points_per_cloud": N_POINTS, "l2": l2, "parallelism": "independent pairs sharded over GPUs"}


def run_b200(args, rank, world, local_rank):
    P = args.pairs
    D = min(P, D_MAX)
    # weak scaling: every rank registers the SAME P pairs per step -- D distinct pairs (4 scenes x 4 prior poses, 5 to 9 ICP
    # iterations), each P / D times -- so that per-GPU work is identical for every N
    pairs = load_pairs(D, generate=(local_rank == 0))      # before CUDA is initialised: the ray casts run in forked workers
    import torch
    import aicp_mapping_b200 as ab
    from aicp_mapping_b200 import capi

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    reg = ab.B200Registration(device=local_rank)
    ovl = ab.B200Overlap(device=local_rank)
    reg.setMatchSchedule(args.match_schedule)
    reg.setKnnSchedule(args.knn_schedule)
    reg.setLoopSchedule(args.loop_schedule)
    reg.setProfiling(0 if args.no_profile else 1)     # CUDA events around k_match only inside the timed region
    dev, host, ratios = [], [], []
    for p in pairs:
        ovl.computeOverlap(p["ref"], p["read"], p["ref_origin"], p["read_origin"])
        ratios.append(ab.autotune_ratio(float(ovl.getOverlap())))
        r4, q4 = capi.to_xyzw(p["ref"]), capi.to_xyzw(p["read"])
        dev.append((torch.from_numpy(r4).cuda(), torch.from_numpy(q4).cuda()))      # one buffer per pair and cloud, also for shared scans
        hr, hq = torch.from_numpy(r4).pin_memory(), torch.from_numpy(q4).pin_memory()
        host.append((hr, hq, hr.numpy(), hq.numpy()))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    S = args.streams
    order = [j % len(pairs) for j in range(P)]
    batch_ratios = [ratios[k] for k in order]
    dev_batch = [(dev[k][0], dev[k][1]) for k in order]
    host_batch = [(host[k][2], host[k][3]) for k in order]

    def one_step(use_host):
        """P registrations, concurrently on S streams; returns (device ms of the batch, per-stage sums)."""
        flush.zero_()
        torch.cuda.synchronize()
        T, stats, status, ms = reg.registerBatch(host_batch if use_host else dev_batch, ratios=batch_ratios, streams=S)
        agg = dict(match=0.0, select=0.0, accumulate=0.0, index=0.0, normals=0.0, iters=0, launches=0, reg_ms=0.0,
                   tail_pick=0.0, tail_select=0.0, tail_solve=0.0, setup=0.0, loop=0.0)
        for s in stats:
            agg["match"] += s.ms_match; agg["select"] += s.ms_select; agg["accumulate"] += s.ms_accumulate
            agg["index"] += s.ms_index; agg["normals"] += s.ms_normals
            agg["iters"] += s.iterations; agg["launches"] += s.gpu_launches; agg["reg_ms"] += s.ms_total
            agg["tail_pick"] += s.ms_tail_pick; agg["tail_select"] += s.ms_tail_select; agg["tail_solve"] += s.ms_tail_solve
            agg["setup"] += s.ms_setup; agg["loop"] += s.ms_iterations
        return ms, agg

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        one_step(False)
    barrier()
    t_wall0 = time.perf_counter()
    dev_ms, agg = 0.0, None
    if args.profile_run:
        torch.cuda.profiler.start()          # ncu --profile-from-start off: capture the timed step only
    for _ in range(args.steps):
        ms, a = one_step(False)
        dev_ms += ms
        agg = a if agg is None else {k: agg[k] + a[k] for k in agg}
    barrier()
    if args.profile_run:
        torch.cuda.profiler.stop()
    t_wall1 = time.perf_counter()
    wall_s = t_wall1 - t_wall0
    clocks = sampler.stop(t_wall0, t_wall1)

    # e2e: pinned host buffers through the same plugin call, wall clock around the synchronous call (copies inside)
    for _ in range(0 if args.profile_run else min(args.warmup, 2)):
        one_step(True)
    barrier()
    e2e_s = 0.0
    for _ in range(0 if args.profile_run else args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reg.registerBatch(host_batch, ratios=batch_ratios, streams=S)
        e2e_s += time.perf_counter() - t0
    barrier()

    # per-stage breakdown: a separate, untimed pass with CUDA events around every stage (costs ~5 % throughput)
    stage = None
    if not args.profile_run and not args.no_profile:
        reg.setProfiling(2)
        for _ in range(2):
            _, a = one_step(False)
            stage = a if stage is None else {k: stage[k] + a[k] for k in stage}
        reg.setProfiling(1)
        barrier()

    # the whole AICP step (octree overlap -> auto-tuned ratio -> registration, app.cpp:218-247) per pair, batched the same way
    aicp_step = None
    if not args.profile_run:
        origins = [(pairs[k]["ref_origin"], pairs[k]["read_origin"]) for k in order]
        reg.setProfiling(0)
        reg.aicpBatch(dev_batch, origins, streams=S)
        step_ms = 0.0
        for _ in range(3):
            flush.zero_(); torch.cuda.synchronize()
            step_ms += reg.aicpBatch(dev_batch, origins, streams=S)[4]
        aicp_step = {"value": 3 * P / (step_ms * 1e-3), "unit": "AICP steps/s (overlap + auto-tune + registration) on this GPU",
                     "ms_per_step_amortised": step_ms / (3 * P)}
        reg.setProfiling(0 if args.no_profile else 1)
        barrier()

    # single-stream latency of one registration (no concurrency), for the roofline of the dominant kernel in isolation
    lat = dict(ms=0.0, match=0.0, iters=0, n=0)
    for k in range(0 if args.profile_run else min(4, len(pairs))):
        for j in range(4):
            reg.setConfig(ratio=ratios[k])
            flush.zero_()
            torch.cuda.synchronize()
            reg.registerClouds(dev[k][0], dev[k][1])
            if j == 0:
                continue                      # first call on this handle allocates its device buffers
            lat["ms"] += reg.stats.ms_total; lat["match"] += reg.stats.ms_match; lat["iters"] += reg.stats.iterations; lat["n"] += 1
    barrier()

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = float(t[0]), float(t[1])
    n_reg_total = world * P * args.steps
    value = n_reg_total / (dev_ms_max * 1e-3)
    e2e_value = n_reg_total / (e2e_ms_max * 1e-3) if e2e_ms_max > 0 else None

    if rank == 0:
        peak, peak_src = measured_peaks()
        iters_total = agg["iters"]
        n_launch_match = iters_total                                   # k_match launches that did work
        match_ms = agg["match"] / max(1, n_launch_match)
        alg_match = 24.0 * N_POINTS                                    # read point 16 + write pos 4 + d2 4
        achieved = alg_match / (match_ms * 1e-3) / 1e9 if match_ms > 0 else 0.0
        I = iters_total / (P * args.steps)
        b_reg = 72.0 * N_POINTS + I * 84.0 * N_POINTS + 32.0 * N_POINTS
        reg_ms = dev_ms / (P * args.steps)          # amortised device time per registration with S streams busy
        line = {"metric": "ICP registrations/sec (128k pts)", "value": value, "unit": "registrations/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload_config(P, "flushed before every step (256 MiB write)", D, I), streams_per_gpu=S),
                "roofline": {"bound": "hbm", "kernel": "k_match_tile (k_match for the cold first iteration)" if S > 1 else "k_match",
                             "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": profiled_traffic()[0], "traffic_source": profiled_traffic()[1],
                             "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_match, "avg_launch_ms": match_ms,
                             "launches_timed": n_launch_match,
                             "note": "timed live in the step with %d registrations in flight per GPU; k_match alone "
                                     "(one stream): %.4f ms per launch" % (S, lat["match"] / max(1, lat["iters"]))},
                "latency_single_stream": {"ms_per_registration": lat["ms"] / max(1, lat["n"]),
                                          "registrations_per_s": 1e3 * lat["n"] / max(1e-9, lat["ms"])},
                "roofline_registration": {"algorithmic_bytes": b_reg, "iterations_mean": I, "ms": reg_ms,
                                          "achieved": b_reg / (reg_ms * 1e-3) / 1e9, "frac": b_reg / (reg_ms * 1e-3) / 1e9 / peak,
                                          "unit": "GB/s"},
                "issue_slots": issue_slots(I, value / world, clocks),
                "aicp_step": aicp_step,
                "stage_ms_per_registration": None if stage is None else dict(
                    {k: stage[k] / (P * 2) for k in ("index", "normals", "match", "select", "accumulate", "tail_pick", "tail_select",
                                                     "tail_solve", "setup", "loop", "reg_ms")},
                    note="separate untimed pass with CUDA events around every stage; per-stream times while %d registrations share the GPU" % S),
                "e2e": {"value": e2e_value, "unit": "registrations/s", "h2d_bytes_per_step": int(P * 2 * N_POINTS * 16),
                        "d2h_bytes_per_step": int(P * 64)},
                "gpu_launches": int(agg["launches"]), "clocks": clocks, "wall_s": wall_s, "ratios": ratios}
        if world == 1 and not args.no_cpu:
            cores = host_cores()
            sec_all, it_all = 0.0, 0
            for k in range(D):                  # the step's D distinct pairs once each: ~5 s on 16 cores
                s_, i_ = oracle_register(pairs[k], ratios[k], cores, 0)
                sec_all += s_; it_all += i_
            sec_one, _ = oracle_register(pairs[0], ratios[0], 1, 0)
            sec_full, _ = oracle_register(pairs[0], ratios[0], cores, 1)
            line["cpu_baseline"] = {"value": D / sec_all, "unit": "registrations/s", "cores": cores, "kind": "port",
                                    "sample": "one registration of each of the step's %d distinct C3 pairs (131072 x 131072 pts, mean "
                                              "%.2f iterations) with the kd-tree oracle, OpenMP over queries on all cores, without the "
                                              "dead reading-side SurfaceNormal filter (skipped by the GPU arm too)" % (D, it_all / D),
                                    "value_1thread_pair0": 1.0 / sec_one, "value_pair0_with_reading_normals": 1.0 / sec_full}
    else:
        line = None
    if world > 1 and not args.no_sharded and not args.profile_run:
        # a hang of the extra leg (a rank lost, a collective mismatch) must not cost the round its headline line: after
        # 300 s every rank gives up, rank 0 printing the line it already has
        def give_up():
            if line is not None:
                line["sharded"] = {"error": "the sharded leg did not finish within 300 s"}
                print(json.dumps(line), flush=True)
            os._exit(0)
        dog = threading.Timer(300.0, give_up)
        dog.daemon = True
        dog.start()
        res = sharded_leg(args, rank, world, local_rank, dist, dev, pairs, ratios)
        dog.cancel()
        if line is not None:
            line["sharded"] = res
    if line is not None:
        print(json.dumps(line), flush=True)
    reg.close(); ovl.close()
    if dist is not None:
        dist.destroy_process_group()


def sharded_leg(args, rank, world, local_rank, dist, dev, pairs, ratios):
    """ONE registration with its reading sharded over all ranks (BASELINE.json configs[3]; reference shape app.cpp:41-69,
    123-127): the reference index is replicated, every rank matches its slice of the reading, and per iteration the ranks
    exchange the trimmed-quantile digits and the 27 normal-equation sums.  Two cases -- the C3 pair, and a 122 880-point
    reading against a --map-points map with the index built once -- each timed against the same registration on one GPU
    (device time of the call, max over ranks, median of the repetitions) and compared with it bit for bit."""
    import torch
    import aicp_mapping_b200 as ab
    from aicp_mapping_b200 import capi, synth
    from aicp_mapping_b200.registration import comm_unique_id
    out = {"world": world}
    u32 = lambda a: np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)

    def vmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def all_true(ok):
        t = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t[0]))

    single = sh = None
    try:
        uid = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        single = ab.B200Registration(device=local_rank)
        sh = ab.B200Registration(device=local_rank)
        sh.commInit(uid[0], rank, world)
        out["exchange"] = sh.commInfo()

        def case(ref_dev, read_dev, ratio, fixed_reference, reps=7):
            shard = read_dev[rank::world].contiguous()
            torch.cuda.synchronize()
            res = {}
            for name, r, q in (("single", single, read_dev), ("sharded", sh, shard)):
                r.setConfig(ratio=ratio)
                if fixed_reference:
                    r.setReference(ref_dev)
                ms, wall, T = [], [], None
                for k in range(reps + 1):
                    torch.cuda.synchronize()
                    dist.barrier()
                    t0 = time.perf_counter()
                    T = r.registerToReference(q) if fixed_reference else r.registerClouds(ref_dev, q)
                    w = time.perf_counter() - t0
                    if k > 0:                   # the first call allocates (and, with a fixed reference, builds the map index)
                        ms.append(vmax(r.stats.ms_total)); wall.append(vmax(w * 1e3))
                st = r.stats
                res[name] = (float(np.median(ms)), float(np.median(wall)), T, int(st.iterations),
                             {k: round(float(getattr(st, "ms_" + k)), 4) for k in ("setup", "iterations", "match", "select", "accumulate",
                                                                                   "exchange", "tail_pick", "tail_select", "tail_solve")})
            same = all_true(np.array_equal(u32(res["single"][2]), u32(res["sharded"][2])) and res["single"][3] == res["sharded"][3])
            return {"ms_per_registration": res["sharded"][0], "ms_single_gpu": res["single"][0],
                    "speedup": res["single"][0] / res["sharded"][0], "wall_ms_per_registration": res["sharded"][1],
                    "wall_ms_single_gpu": res["single"][1], "iterations": res["sharded"][3], "bit_identical_to_single_gpu": same,
                    "stage_ms_rank0_last_rep": {"sharded": res["sharded"][4], "single_gpu": res["single"][4]},
                    "n_reference": int(ref_dev.shape[0]), "n_reading": int(read_dev.shape[0])}

        out["c3"] = case(dev[0][0], dev[0][1], ratios[0], False)
        mp_ = synth.make_map_case(n_map=args.map_points, n_read=122880, trial=1, n_poses=1, n_clutter=1500)
        map_dev = torch.from_numpy(capi.to_xyzw(mp_["map"])).cuda()
        read_dev = torch.from_numpy(capi.to_xyzw(mp_["readings"][0]["read"])).cuda()
        out["c4"] = case(map_dev, read_dev, ab.autotune_ratio(50.0), True)      # app.cpp:123-127: overlap forced to 50 % against a prior map
        del map_dev
    except Exception as e:      # reported, not fatal: the headline line does not depend on this leg
        out["error"] = "%s: %s" % (type(e).__name__, e)
    for r in (sh, single):
        try:
            if r is not None:
                r.close()
        except Exception:
            pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=64, help="cloud pairs registered per GPU per step")
    ap.add_argument("--streams", type=int, default=8, help="concurrent registrations per GPU (CUDA streams)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded single-registration leg (world > 1)")
    ap.add_argument("--map-points", type=int, default=10485760, help="map size of the sharded C4 case")
    ap.add_argument("--match-schedule", type=int, default=0, help="0 auto, 1 per-thread search, 2 tile search (experiments)")
    ap.add_argument("--knn-schedule", type=int, default=0, help="0 auto, 1 warp-per-query k-NN, 2 tile k-NN (experiments)")
    ap.add_argument("--loop-schedule", type=int, default=0, help="0 persistent loop kernel, 1 three launches per iteration (experiments)")
    ap.add_argument("--no-profile", action="store_true", help="no per-stage CUDA events inside the registrations")
    ap.add_argument("--profile-run", action="store_true", help="device-resident leg only (the command profiled under ncu)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
