"""Distribution of per-warp durations of the search phase (iteration 3) of the persistent loop kernel, experiment build
libaicp_b200_wt3.so (-DAICP_DEBUG_WARP_TIMES=3).  python tools/warp_times_probe.py [pair]"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["AICP_B200_LIB"] = os.path.join(ROOT, "aicp_mapping_b200", "lib", "libaicp_b200_wt3.so")
from bench import load_pairs
k = int(sys.argv[1]) if len(sys.argv) > 1 else 0
pairs = load_pairs(k + 1)
import torch
import aicp_mapping_b200 as ab
from aicp_mapping_b200 import capi
p = pairs[k]
ovl = ab.B200Overlap(); ovl.computeOverlap(p["ref"], p["read"], p["ref_origin"], p["read_origin"])
ratio = ab.autotune_ratio(float(ovl.getOverlap())); ovl.close()
ref = torch.from_numpy(capi.to_xyzw(p["ref"])).cuda(); read = torch.from_numpy(capi.to_xyzw(p["read"])).cuda()
reg = ab.B200Registration(); reg.setConfig(ratio=ratio); reg.setLoopSchedule(2); reg.setMatchSchedule(1)
for _ in range(3):
    reg.registerClouds(ref, read)
buf = np.zeros(4 * 8192, dtype=np.uint32)
rc = capi.lib().aicp_b200_debug_warp_times(buf.ctypes.data_as(C.c_void_p), buf.size)
w = buf.reshape(-1, 4)[:4096].astype(np.float64)
dur, start, hard, sm = w[:, 0] / 1e3, w[:, 1] / 1e3, w[:, 2], w[:, 3]
print("rc", rc, "iterations", reg.stats.iterations, "search phase per iteration %.1f us" % (1e3 * reg.stats.ms_match / reg.stats.iterations))
print("warp duration us: mean %.1f median %.1f p90 %.1f p99 %.1f max %.1f; end (start + duration) max %.1f us; start max %.1f" %
      (dur.mean(), np.median(dur), np.percentile(dur, 90), np.percentile(dur, 99), dur.max(), (start + dur).max(), start.max()))
for lo, hi in ((0, 0), (1, 7), (8, 23), (24, 32)):
    m = (hard >= lo) & (hard <= hi)
    if m.any():
        print("warps with %2d..%2d outlier lanes: %4d  mean %.1f us  p90 %.1f  max %.1f" % (lo, hi, m.sum(), dur[m].mean(), np.percentile(dur[m], 90), dur[m].max()))
persm = np.array([dur[sm == s].sum() for s in np.unique(sm)])
print("per-SM sum of warp durations: mean %.0f max %.0f us; warps per SM max %d" % (persm.mean(), persm.max(), max((sm == s).sum() for s in np.unique(sm))))
order = np.argsort(-dur)[:8]
print("slowest warps (tile, dur us, outlier lanes):", [(int(i), round(dur[i], 1), int(hard[i])) for i in order])
reg.close()

# per-query work counters of the same iteration
reg = ab.B200Registration(); reg.setConfig(ratio=ratio); reg.setLoopSchedule(2); reg.setMatchSchedule(1)
for _ in range(2):
    reg.registerClouds(ref, read)
qb = np.zeros(4 * 131072, dtype=np.uint32)
capi.lib().aicp_b200_debug_query_stats(qb.ctypes.data_as(C.c_void_p), qb.size)
q = qb.reshape(-1, 4).astype(np.int64)
levels, desc, nodes, points, ns = q[:, 0] & 0xFFFF, q[:, 0] >> 16, q[:, 1], q[:, 2], q[:, 3] / 1e3
print("per query: climb levels mean %.1f p99 %d max %d | sibling descents mean %.2f max %d | nodes mean %.1f p99 %d max %d | points mean %.1f p99 %d max %d" %
      (levels.mean(), np.percentile(levels, 99), levels.max(), desc.mean(), desc.max(), nodes.mean(), np.percentile(nodes, 99), nodes.max(),
       points.mean(), np.percentile(points, 99), points.max()))
print("per query time us (includes waiting for the other lanes of the warp): mean %.1f p99 %.1f max %.1f" % (ns.mean(), np.percentile(ns, 99), ns.max()))
work = levels * 6 + nodes * 6 + points          # rough dependent-step proxy
wq = work.reshape(-1, 32)
print("per warp: sum of lane work mean %.0f max %.0f | max lane work mean %.0f max %.0f" % (wq.sum(1).mean(), wq.sum(1).max(), wq.max(1).mean(), wq.max(1).max()))
dw = q[:, 3].reshape(-1, 32).max(1) / 1e3
for name, v in (("sum of lane work", wq.sum(1)), ("max lane work", wq.max(1)), ("max levels", levels.reshape(-1, 32).max(1)), ("sum nodes", nodes.reshape(-1, 32).sum(1)),
                ("sum points", points.reshape(-1, 32).sum(1)), ("distinct lane paths (levels)", np.array([len(set(r)) for r in levels.reshape(-1, 32)]))):
    print("correlation of warp time with %s: %.2f" % (name, np.corrcoef(dw, v)[0, 1]))
slow = np.argsort(-dw)[:6]
for wi in slow:
    print("slow warp %d: %.1f us, levels %s nodes %s points %s" % (wi, dw[wi], levels.reshape(-1, 32)[wi].tolist(), nodes.reshape(-1, 32)[wi].tolist(), points.reshape(-1, 32)[wi].tolist()))
reg.close()
