#!/usr/bin/env python
"""Single-process multi-GPU batch (aicp_b200_register_batch_devices): P pairs per GPU per step over every GPU of the box from
ONE process, against the torchrun figure of bench.py (one process per GPU).  One JSON line.
    python tools/bench_devices.py [--pairs 64] [--steps 5]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import load_pairs

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=64)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--streams", type=int, default=8)
args = ap.parse_args()
pairs = load_pairs(16)
import torch
import aicp_mapping_b200 as ab
from aicp_mapping_b200 import capi
G = torch.cuda.device_count()
ovl = ab.B200Overlap(device=0)
ratios, host = [], []
for p in pairs:
    ovl.computeOverlap(p["ref"], p["read"], p["ref_origin"], p["read_origin"])
    ratios.append(ab.autotune_ratio(float(ovl.getOverlap())))
    host.append((torch.from_numpy(capi.to_xyzw(p["ref"])).pin_memory().numpy(), torch.from_numpy(capi.to_xyzw(p["read"])).pin_memory().numpy()))
ovl.close()
reg = ab.B200Registration(device=0)
n = args.pairs * G
batch = [host[(i // G) % 16] for i in range(n)]          # device d gets pairs d, d + G, ...: every device the same 16 distinct pairs
rat = [ratios[(i // G) % 16] for i in range(n)]
devs = list(range(G))
for _ in range(2):
    reg.registerBatch(batch, ratios=rat, streams=args.streams, devices=devs)
wall, dev_ms = 0.0, 0.0
for _ in range(args.steps):
    t0 = time.perf_counter()
    T, stats, status, ms = reg.registerBatch(batch, ratios=rat, streams=args.streams, devices=devs)
    wall += time.perf_counter() - t0
    dev_ms += ms
    assert all(s == 0 for s in status)
print(json.dumps({"what": "aicp_b200_register_batch_devices, one process, host (pinned) inputs", "gpus": G, "pairs_per_gpu_per_step": args.pairs,
                  "registrations_per_s_wall (e2e)": n * args.steps / wall, "registrations_per_s_device_time": n * args.steps / (dev_ms * 1e-3),
                  "iterations_mean": float(np.mean([s.iterations for s in stats]))}))
reg.close()
