"""What a hashed uniform grid would have to scan for the WARM 1-NN searches of one C3 registration (CPU simulation, numpy).

For every ICP iteration t >= 1 of the oracle's run: the query positions T_(t-1) * reading, the previous match as the seed,
the ball (query, |query - seed|), and -- as a multi-level Morton-cell grid would do it -- the finest level at which the ball's
box spans at most 2 cells per axis (<= 8 cells), counting the points in those cells.  That count is the number of
distance evaluations of an exact grid probe; compare with the radix-tree search's measured 23.6 points + 3.4 nodes per
query (profiles/round2_h_warp_times_probe.txt).  TEST/ANALYSIS TOOL: imports the oracle, never used by the product.
python tools/grid_probe_sim.py > profiles/round2_i_grid_probe_sim.txt"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aicp_mapping_b200 import synth
from oracle import oracle as orc
pair = synth.make_pair(3, 0)
ref, read = pair["ref"], pair["read"]
cfg = orc.default_config(ratio=0.7, threads=8, reading_normals=0, use_kdtree=1)
t0 = time.time(); out = orc.icp(ref, read, cfg, want_trace_idx=True, want_reading=False); print("oracle", time.time() - t0, out.iterations)
mu = out.mean_ref
refc = (ref[:, :3] - mu).astype(np.float32); read0 = (read[:, :3] - mu).astype(np.float32)
lo = refc.min(0); ext = (refc.max(0) - lo).max(); u = ext / 1023.0
print("ext", ext, "unit", u)
q = np.clip(((refc - lo) / u).astype(np.int64), 0, 1023)
def cell_id(c, sh):   # linear id at level (10 - sh)
    L = 10 - sh; c = c >> sh
    return (c[..., 0] << (2 * L)) | (c[..., 1] << L) | c[..., 2]
ids = {sh: np.sort(cell_id(q, sh)) for sh in range(0, 5)}
T = np.eye(4, dtype=np.float32)
for t in range(1, out.iterations):
    T = out.trace[t - 1]["T_iter"]
    p = read0 @ T[:3, :3].T + T[:3, 3]
    seed = refc[out.trace_idx[t - 1]]
    r = np.sqrt(((p - seed) ** 2).sum(1))
    clo = np.clip(((p - r[:, None] - lo) / u - 0.01).astype(np.int64), 0, 1023)
    chi = np.clip(((p + r[:, None] - lo) / u + 0.01).astype(np.int64), 0, 1023)
    sh = np.zeros(len(p), dtype=np.int64)
    for s in range(0, 10):
        bad = (((chi >> s) - (clo >> s)) > 1).any(1) & (sh == s)
        sh[bad] = s + 1
    cnt = np.zeros(len(p), dtype=np.int64); ncell = np.zeros(len(p), dtype=np.int64)
    for s in range(0, 5):
        m = sh == s
        if not m.any(): continue
        a, b = clo[m] >> s, chi[m] >> s
        L = 10 - s
        tot = np.zeros(m.sum(), dtype=np.int64); nc = np.zeros(m.sum(), dtype=np.int64)
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    c = np.stack([a[:, 0] + dx, a[:, 1] + dy, a[:, 2] + dz], 1)
                    ok = (c <= b).all(1)
                    cid = (c[:, 0] << (2 * L)) | (c[:, 1] << L) | c[:, 2]
                    k = np.searchsorted(ids[s], cid, "right") - np.searchsorted(ids[s], cid, "left")
                    tot += np.where(ok, k, 0); nc += ok
        cnt[m] = tot; ncell[m] = nc
    ok = sh <= 3
    print("iter %d: r median %.3f p90 %.3f | sh hist %s | fallback(sh>3) %.1f%% | cells mean %.2f | cand mean %.1f p50 %d p90 %d p99 %d max %d | >96: %.1f%%" % (
        t, np.median(r), np.percentile(r, 90), np.bincount(np.minimum(sh, 5), minlength=6).tolist(), 100 * (~ok).mean(), ncell[ok].mean(),
        cnt[ok].mean(), np.median(cnt[ok]), np.percentile(cnt[ok], 90), np.percentile(cnt[ok], 99), cnt[ok].max(), 100 * (cnt[ok] > 96).mean()))
