"""CPU tests that pin the oracle (oracle/) against analytic known-answer cases and the reference's in-repo text
KATs.  The reference ships no runnable golden vectors for this path (aicp_core/test/aicp_test.cpp:50-57 points
outside the repository), so these are the strongest pins available: PARITY UNPINNED."""
import math

import numpy as np
import pytest

from aicp_mapping_b200 import synth
from conftest import rot_angle


def test_sincos_atan_match_libm(orc):
    rng = np.random.default_rng(1)
    for x in list(rng.uniform(-7, 7, 200)) + [0.0, 1e-9, -1e-9, 1e-3, math.pi / 4, math.pi / 2, 3.0, 50.0]:
        s, c = orc.sincos(float(x))
        assert abs(s - math.sin(x)) < 4e-16 and abs(c - math.cos(x)) < 4e-16
    for y, x in rng.uniform(0, 2, (200, 2)):
        assert abs(orc.atan2_pos(float(y), float(x)) - math.atan2(y, x)) < 1e-15
    assert orc.atan2_pos(0.0, 0.0) == 0.0
    assert abs(orc.atan2_pos(1e-9, 1.0) - 1e-9) < 1e-24
    assert abs(orc.atan2_pos(1.0, 0.0) - math.pi / 2) < 1e-15


def test_match_kdtree_equals_bruteforce_with_ties(orc):
    rng = np.random.default_rng(2)
    ref = rng.uniform(-5, 5, (3000, 3)).astype(np.float32)
    ref[100:200] = ref[0:100]                      # exact duplicates -> ties must go to the lowest index
    ref = np.round(ref * 4) / 4                    # lattice -> many equidistant candidates
    qry = np.round(rng.uniform(-6, 6, (2000, 3)).astype(np.float32) * 8) / 8
    i_b, d_b = orc.match(ref, qry, use_kdtree=False)
    i_k, d_k = orc.match(ref, qry, use_kdtree=True)
    i_t, d_t = orc.match(ref, qry, use_kdtree=True, threads=4)
    assert np.array_equal(i_b, i_k) and np.array_equal(d_b.view(np.uint32), d_k.view(np.uint32))
    assert np.array_equal(i_b, i_t) and np.array_equal(d_b.view(np.uint32), d_t.view(np.uint32))
    # independent numpy check of the definition: argmin of (d2, index) with the float32 operation order
    d = (qry[:, None, :] - ref[None, :, :]).astype(np.float32)
    d2 = ((d[..., 0] * d[..., 0]) + (d[..., 1] * d[..., 1])) + (d[..., 2] * d[..., 2])
    assert np.array_equal(i_b, d2.argmin(1).astype(np.int32))       # argmin returns the first minimum
    assert np.array_equal(d_b, d2.min(1))


def test_normals_plane_and_kdtree_equals_bruteforce(orc):
    rng = np.random.default_rng(3)
    xy = rng.uniform(-2, 2, (1500, 2))
    n_true = np.array([0.3, -0.2, 0.933]); n_true /= np.linalg.norm(n_true)
    z = -(xy @ n_true[:2]) / n_true[2]
    pts = np.c_[xy, z].astype(np.float32)
    nb, kb = orc.surface_normals(pts, 20, use_kdtree=False)
    nk, kk = orc.surface_normals(pts, 20, use_kdtree=True)
    assert np.array_equal(kb, kk) and np.array_equal(nb.view(np.uint32), nk.view(np.uint32))
    assert np.all(kb[:, 0] == np.arange(1500))                      # self is the nearest (no duplicates here)
    assert np.abs(np.abs(nb[:, :3] @ n_true) - 1).max() < 1e-5
    assert np.all(nb[:, 2] > 0)                                     # canonical sign: largest component positive
    assert np.abs(np.linalg.norm(nb[:, :3].astype(np.float64), axis=1) - 1).max() < 1e-6
    # density = k / (4/3 pi r^3), r = farthest neighbour from the neighbourhood mean
    i = 7
    nn = pts[kb[i]].astype(np.float64)
    r = np.linalg.norm(nn - nn.mean(0), axis=1).max()
    assert abs(nb[i, 3] - 20 / (4 / 3 * math.pi * r ** 3)) / nb[i, 3] < 1e-6


def test_normals_degenerate_neighbourhood_is_unit_y(orc):
    # A.2: rank <= 1 neighbourhood -> eigenvalues (1,0,0), eigenvectors I -> normal (0,1,0)
    t = np.linspace(0, 1, 64, dtype=np.float32)
    line = np.c_[t, 2 * t, -t].astype(np.float32)
    n, _ = orc.surface_normals(line, 10)
    assert np.array_equal(n[:, :3], np.tile(np.array([0, 1, 0], dtype=np.float32), (64, 1)))


def test_normals_knn_must_be_smaller_than_n(orc):
    with pytest.raises(RuntimeError, match="KNN_TOO_LARGE"):
        orc.surface_normals(np.zeros((20, 3), dtype=np.float32), 20)


def test_trim_threshold_definition(orc):
    rng = np.random.default_rng(4)
    d2 = rng.exponential(1.0, 10001).astype(np.float32)
    d2[:10] = 0.0                 # zero distances are excluded from the quantile
    d2[10:15] = np.inf            # invalid matches too
    for ratio in (0.25, 0.358818, 0.7, 0.999, 1.0):
        limit, nv = orc.trim_threshold(d2, ratio)
        vals = np.sort(d2[(d2 > 0) & np.isfinite(d2)])
        assert nv == vals.size
        idx = vals.size - 1 if ratio == 1.0 else min(int(np.float32(vals.size) * np.float32(ratio)), vals.size - 1)
        assert limit == vals[idx]
    with pytest.raises(RuntimeError, match="NO_VALID_MATCH"):
        orc.trim_threshold(np.zeros(5, dtype=np.float32), 0.5)


def test_normal_equations_exact_and_order_independent(orc):
    rng = np.random.default_rng(5)
    n = 5000
    ref = rng.uniform(-50, 50, (n, 3)).astype(np.float32)
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = np.c_[nrm, np.ones(n)].astype(np.float32)
    p = (ref + rng.normal(0, 0.1, (n, 3))).astype(np.float32)
    idx = rng.permutation(n).astype(np.int32)
    d2 = rng.exponential(1.0, n).astype(np.float32)
    limit = np.float32(1.0)
    hi, lo, used = orc.normal_equations(p, ref, nrm, idx, d2, limit)
    perm = rng.permutation(n)
    hi2, lo2, used2 = orc.normal_equations(p[perm], ref, nrm, idx[perm], d2[perm], limit)
    assert used == used2 == int((d2 <= limit).sum())
    assert np.array_equal(hi, hi2) and np.array_equal(lo, lo2)
    # float64 check of the values
    w = d2 <= limit
    P, Q, N = p[w].astype(np.float64), ref[idx[w]].astype(np.float64), nrm[idx[w], :3].astype(np.float64)
    F = np.c_[np.cross(P, N), N]
    A = F.T @ F
    g = F.T @ np.einsum("ij,ij->i", P - Q, N)
    vals = np.array([float((int(h) << 64) + int(l)) / 2.0 ** 30 for h, l in zip(hi, lo)])
    iu = np.triu_indices(6)
    assert np.allclose(vals[:21], A[iu], rtol=1e-5, atol=1e-3)
    assert np.allclose(vals[21:], g, rtol=1e-4, atol=1e-2)
    x, path = orc.solve6(hi, lo)
    assert path == 1
    Aex = np.zeros((6, 6)); Aex[iu] = vals[:21]; Aex = Aex + Aex.T - np.diag(np.diag(Aex))
    assert np.allclose(x, np.linalg.solve(Aex, -vals[21:]), rtol=1e-9, atol=1e-12)


def test_solve6_rank_deficient_takes_minimum_norm_path(orc):
    # all points on one plane z=0 with normal z: only (rot_x, rot_y, trans_z) are observable
    rng = np.random.default_rng(6)
    n = 2000
    ref = np.c_[rng.uniform(-5, 5, (n, 2)), np.zeros(n)].astype(np.float32)
    nrm = np.tile(np.array([0, 0, 1, 1], dtype=np.float32), (n, 1))
    p = ref.copy(); p[:, 2] += 0.01
    hi, lo, _ = orc.normal_equations(p, ref, nrm, np.arange(n, dtype=np.int32), np.ones(n, dtype=np.float32), 2.0)
    x, path = orc.solve6(hi, lo)
    assert path == 2
    assert np.all(np.isfinite(x))
    assert abs(x[5] + 0.01) < 1e-6 and np.abs(x[[2, 3, 4]]).max() < 1e-9


def test_pose_increment_is_rodrigues(orc):
    x = np.array([0.01, -0.02, 0.03, 0.5, -0.25, 0.125])
    dT = orc.pose_increment(x).astype(np.float64)
    th = np.linalg.norm(x[:3]); u = x[:3] / th
    K = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
    R = np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * K @ K
    assert np.abs(dT[:3, :3] - R).max() < 1e-7 and np.array_equal(dT[:3, 3], x[3:].astype(np.float32))
    assert np.array_equal(orc.pose_increment(np.zeros(6)), np.eye(4, dtype=np.float32))   # zero rotation -> identity


def test_icp_recovers_cube_perturbation(orc):
    pair = synth.make_pair(5, trial=3)
    out = orc.icp(pair["ref"], pair["read"], orc.default_config(ratio=0.7))
    assert out.rc == 0 and 4 <= out.iterations <= 20
    err = out.T.astype(np.float64) @ np.linalg.inv(pair["T_true"])
    assert np.linalg.norm(err[:3, 3]) < 2e-3 and rot_angle(err[:3, :3]) < 1e-3
    # output reading = T * reading (pointmatcher_registration.cpp:128-131)
    assert np.array_equal(out.reading, orc.transform_points(out.T, pair["read"]))
    # brute force and kd-tree give the same trajectory bit for bit
    small = synth.make_pair(5, trial=3, n_points=3000)
    a = orc.icp(small["ref"], small["read"], orc.default_config(ratio=0.6, use_kdtree=0), want_trace_idx=True)
    b = orc.icp(small["ref"], small["read"], orc.default_config(ratio=0.6, use_kdtree=1, threads=3), want_trace_idx=True)
    assert a.iterations == b.iterations and np.array_equal(a.T, b.T) and np.array_equal(a.trace_idx, b.trace_idx)


def test_icp_checkers(orc):
    pair = synth.make_pair(5, trial=1, n_points=4000)
    out = orc.icp(pair["ref"], pair["read"], orc.default_config(ratio=0.7, max_iterations=3))
    assert out.iterations == 3 and out.stop_reason == orc.STOP_COUNTER
    out = orc.icp(pair["ref"], pair["read"], orc.default_config(ratio=0.7))
    assert out.stop_reason == orc.STOP_DIFFERENTIAL and out.iterations >= 4      # smoothLength 4 -> earliest stop
    assert all(math.isnan(t["rot_err"]) for t in out.trace[:3]) and not math.isnan(out.trace[3]["rot_err"])
    last = out.trace[-1]
    assert last["rot_err"] < 0.001 and last["trans_err"] < 0.01
    # identical clouds: every distance is zero -> no value enters the quantile -> ConvergenceError
    same = orc.icp(pair["ref"], pair["ref"], orc.default_config())
    assert same.error == "NO_VALID_MATCH"


def test_icp_init_transform_is_composed(orc):
    pair = synth.make_pair(5, trial=2, n_points=6000)
    init = synth.rigid(0.02, -0.01, 0.0, 0, 0, 0.01)
    read_pre = synth.apply_T(np.linalg.inv(init), pair["read"])        # so that init * read_pre == read
    a = orc.icp(pair["ref"], pair["read"], orc.default_config())
    b = orc.icp(pair["ref"], read_pre, orc.default_config(), init_T=init)
    assert np.abs(a.T.astype(np.float64) - b.T.astype(np.float64) @ np.linalg.inv(init)).max() < 1e-3


def test_autotune_ratio_text_roundtrip(orc):
    # the leftover of such a rewrite is still in aicp_core/config/icp/icp_autotuned.yaml:35
    r, text = orc.autotune_ratio(35.8818)
    assert text == "0.358818" and r == np.float32(0.358818)
    assert orc.autotune_ratio(10.0) == (np.float32(0.25), "0.25")     # app.cpp:199-200
    assert orc.autotune_ratio(93.0) == (np.float32(0.7), "0.7")       # app.cpp:201-202
    assert orc.autotune_ratio(50.0) == (np.float32(0.5), "0.5")       # prior-map mode, app.cpp:123-127


def test_overlap_single_ray_keys(orc):
    res = float(np.float32(0.2))
    keys = orc.ray_keys(np.array([[1.1, 0.1, 0.1]], dtype=np.float32), [0.1, 0.1, 0.1], res)
    xs = (keys >> np.uint64(32)).astype(np.int64) - 32768
    assert list(xs) == [0, 1, 2, 3, 4, 5]                              # origin voxel .. end voxel, all on one row
    assert np.all(((keys >> np.uint64(16)) & np.uint64(0xFFFF)) == 32768) and np.all((keys & np.uint64(0xFFFF)) == 32768)
    # end point in the origin voxel: only the occupied end voxel
    assert orc.ray_keys(np.array([[0.15, 0.1, 0.1]], dtype=np.float32), [0.1, 0.1, 0.1], res).size == 1
    # diagonal ray: a 3-D DDA visits |dx|+|dy|+|dz| + 1 voxels
    keys = orc.ray_keys(np.array([[1.05, 0.63, 0.47]], dtype=np.float32), [0.1, 0.1, 0.1], res)
    assert keys.size == 5 + 3 + 2 + 1


def test_overlap_identical_disjoint_and_min_rule(orc):
    rng = np.random.default_rng(7)
    a = rng.uniform(-3, 3, (500, 3)).astype(np.float32)
    ov, (ni, na, nb) = orc.overlap(a, [0, 0, 0], a, [0, 0, 0])
    assert ov == np.float32(100.0) and ni == na == nb
    b = a + np.float32(100.0)
    ov, (ni, na, nb) = orc.overlap(a, [0, 0, 0], b, [100, 100, 100])
    assert ov == 0.0 and ni == 0 and na == nb
    # reading covers half of the reference's rays: overlap = min(|A^B|/|A|, |A^B|/|B|)
    ov, (ni, na, nb) = orc.overlap(a, [0, 0, 0], a[:250], [0, 0, 0])
    assert ni == nb and nb < na
    assert ov == np.float32(float(np.float32(ni) / np.float32(na)) * 100.0)


def test_overlap_box_scene_is_plausible(orc, pair_cache):
    pair = pair_cache(2, 0, 8192)
    ov, (ni, na, nb) = orc.overlap(pair["ref"], pair["ref_origin"], pair["read"], pair["read_origin"])
    assert 40.0 < ov <= 100.0 and ni <= min(na, nb)
