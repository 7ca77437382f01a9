"""Stage-by-stage comparison of the CUDA path and the oracle on one pair: python tools/debug_parity.py cfg trial n ratio"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aicp_mapping_b200 as ab
from aicp_mapping_b200 import synth
from oracle import oracle as orc

cfg_id, trial, n, ratio = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
pair = synth.make_pair(cfg_id, trial, n)
u32 = lambda a: np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
reg = ab.B200Registration()
gn, gk = reg.surfaceNormals(pair["ref"], 20)
on, ok = orc.surface_normals(pair["ref"], 20, use_kdtree=False)
print("knn equal", np.array_equal(gk, ok), "normals equal", np.array_equal(u32(gn), u32(on)))
if not np.array_equal(gk, ok):
    bad = np.flatnonzero((gk != ok).any(1))
    print(" rows differing", len(bad), bad[:5]); i = bad[0]; print(gk[i]); print(ok[i])
    d = ((pair["ref"][ok[i]] - pair["ref"][i]) ** 2).sum(1); print(d)
    d = ((pair["ref"][gk[i]] - pair["ref"][i]) ** 2).sum(1); print(d)
elif not np.array_equal(u32(gn), u32(on)):
    bad = np.flatnonzero((u32(gn) != u32(on)).any(1)); print(" normal rows differing", len(bad), bad[:5]); i = bad[0]; print(gn[i], on[i])
reg.setConfig(ratio=ratio)
reg.enableMatchTrace(True)
T = reg.registerClouds(pair["ref"], pair["read"])
o = orc.icp(pair["ref"], pair["read"], orc.default_config(ratio=ratio, threads=8), want_trace_idx=True, want_normals=True)
print("iters", reg.stats.iterations, o.iterations, "stop", reg.stats.stop_reason, o.stop_reason)
print("ref normals equal", np.array_equal(u32(reg.getReferenceNormals()), u32(o.normals)))
gm = reg.getTraceMatches()
for it, (g, c) in enumerate(zip(reg.trace(), o.trace)):
    print(it, "match eq", np.array_equal(gm[it], o.trace_idx[it]), "limit", g["limit_d2"], c["limit_d2"], "nvalid", g["n_valid"], c["n_valid"],
          "nused", g["n_used"], c["n_used"], "T eq", np.array_equal(u32(g["T_iter"]), u32(c["T_iter"])), "err", g["rot_err"], c["rot_err"], g["trans_err"], c["trans_err"])
