"""An independent re-derivation of the path in numpy/scipy float64 (no shared code with oracle/ or the CUDA library),
written from SURVEY.md Appendix A: k-d tree normals (scipy cKDTree + numpy eigh), exact NN, trimmed quantile,
point-to-plane normal equations solved with numpy, angle-axis update, counter + differential checkers.  It anchors the
oracle's ALGORITHM (the strongest pin available: the reference's libpointmatcher cannot be built here and ships no vectors
for this path).  The oracle works in float32 with a fixed operation order (the reference is PointMatcher<float>); this file
is run twice: in float64 throughout (the two then differ by accumulated float32 rounding: transforms compared at 5e-5 m /
5e-5 rad, measured <= 1.5e-5), and with the GEOMETRY in float32 in the operation order DESIGN.md section 4 states
(transform row times column with k ascending, d2 = ((dx dx) + (dy dy)) + (dz dz), F = [p x n; n], centring on the rounded
fixed-point mean), float64 only where the oracle uses it (covariances, normal equations, solve, checkers) -- that run must
meet BASELINE.json's bar of 1e-5 m / 1e-5 rad.  The discrete outputs that are insensitive to rounding (iteration count,
inlier count of every iteration) are compared exactly in both."""
TOL = 5e-5          # float64 re-derivation
TOL_F32 = 1e-5      # float32-geometry re-derivation: the bar of BASELINE.json
import numpy as np
import pytest
from scipy.spatial import cKDTree

from aicp_mapping_b200 import synth
from conftest import rot_angle


def normals_knn(ref, k):
    tree = cKDTree(ref)
    _, idx = tree.query(ref, k=k)
    nb = ref[idx]                                   # n x k x 3
    d = nb - nb.mean(1, keepdims=True)
    C = np.einsum("nki,nkj->nij", d, d)
    w, v = np.linalg.eigh(C)
    return v[:, :, 0]                               # eigenvector of the smallest eigenvalue (sign irrelevant below)


def angle_axis(rv):
    th = np.linalg.norm(rv)
    if th == 0:
        return np.eye(3)
    u = rv / th
    K = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def quat_angle(Ra, Rb):
    R = Ra.T @ Rb
    return abs(np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1)))


def independent_icp(ref, read, ratio, knn=20, max_iter=20, min_rot=0.001, min_trans=0.01, smooth=4):
    ref = ref.astype(np.float64); read = read.astype(np.float64)
    nrm = normals_knn(ref, knn)
    mu = ref.mean(0)
    refc = ref - mu
    tree = cKDTree(refc)
    read0 = read - mu                               # T_init = identity
    T = np.eye(4)
    hist_R, hist_t = [np.eye(3)], [np.zeros(3)]
    n_used = []
    for it in range(max_iter):
        p = read0 @ T[:3, :3].T + T[:3, 3]
        d, idx = tree.query(p, k=1)
        d2 = (d * d).astype(np.float32)
        valid = d2[(d2 > 0) & np.isfinite(d2)]
        kth = int(np.float32(len(valid)) * np.float32(ratio))
        limit = np.partition(valid, kth)[kth]
        w = d2 <= limit
        n_used.append(int(w.sum()))
        pp, q, n = p[w], refc[idx[w]], nrm[idx[w]]
        F = np.hstack([np.cross(pp, n), n])
        A = F.T @ F
        b = -F.T @ np.einsum("ij,ij->i", pp - q, n)
        x = np.linalg.solve(A, b)
        dT = np.eye(4); dT[:3, :3] = angle_axis(x[:3]); dT[:3, 3] = x[3:]
        T = dT @ T
        hist_R.append(T[:3, :3].copy()); hist_t.append(T[:3, 3].copy())
        stop = it + 1 >= max_iter
        if len(hist_R) > smooth:
            re = np.mean([quat_angle(hist_R[-1 - j], hist_R[-2 - j]) for j in range(smooth)])
            te = np.mean([np.linalg.norm(hist_t[-1 - j] - hist_t[-2 - j]) for j in range(smooth)])
            if re < min_rot and te < min_trans:
                stop = True
        if stop:
            break
    Tmu = np.eye(4); Tmu[:3, 3] = mu
    Tmu_inv = np.eye(4); Tmu_inv[:3, 3] = -mu
    return Tmu @ T @ Tmu_inv, it + 1, n_used


def f32(a):
    return np.asarray(a, dtype=np.float32)


def xform32(T, pts):
    """T (4x4 float32) applied to n x 3 float32 points: ((T0 x + T1 y) + T2 z) + T3 per row, every operation rounded to
    float32 (numpy's elementwise arithmetic never fuses a multiply with an add)."""
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    return np.stack([((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3] for r in range(3)], 1)


def matmul32(A, B):
    """Rigid 4x4 product in float32, k ascending, the translation column adding A's translation last."""
    C = np.zeros((4, 4), dtype=np.float32); C[3, 3] = 1
    for c in range(4):
        for r in range(3):
            acc = A[r, 0] * B[0, c]
            acc = acc + A[r, 1] * B[1, c]
            acc = acc + A[r, 2] * B[2, c]
            if c == 3:
                acc = acc + A[r, 3]
            C[r, c] = acc
    return C


def independent_icp_f32(ref, read, ratio, knn=20, max_iter=20, min_rot=0.001, min_trans=0.01, smooth=4):
    """The same algorithm with float32 geometry in the contract's operation order; float64 where the oracle is float64."""
    ref = f32(ref); read = f32(read)
    nrm = f32(normals_knn(ref.astype(np.float64), knn))
    lead = np.argmax(np.abs(nrm), axis=1)                                   # canonical sign: largest component positive
    nrm = nrm * np.where(nrm[np.arange(len(nrm)), lead] < 0, np.float32(-1), np.float32(1))[:, None]
    mu = f32(np.rint(ref.astype(np.float64) * 65536.0).sum(0) / (65536.0 * len(ref)))
    refc = ref - mu
    tree = cKDTree(refc.astype(np.float64))
    M0 = np.eye(4, dtype=np.float32); M0[:3, 3] = f32(0) - mu
    read0 = xform32(M0, read)
    T = np.eye(4, dtype=np.float32)
    hist_R, hist_t = [np.eye(3)], [np.zeros(3)]
    n_used = []
    for it in range(max_iter):
        p = xform32(T, read0)
        _, idx = tree.query(p.astype(np.float64), k=1)
        df = p - refc[idx]
        d2 = (df[:, 0] * df[:, 0] + df[:, 1] * df[:, 1]) + df[:, 2] * df[:, 2]
        valid = d2[(d2 > 0) & np.isfinite(d2)]
        kth = min(int(np.float32(len(valid)) * np.float32(ratio)), len(valid) - 1)
        limit = np.partition(valid, kth)[kth]
        w = d2 <= limit
        n_used.append(int(w.sum()))
        pp, q, n = p[w], refc[idx[w]], nrm[idx[w]]
        c = np.stack([pp[:, 1] * n[:, 2] - pp[:, 2] * n[:, 1], pp[:, 2] * n[:, 0] - pp[:, 0] * n[:, 2],
                      pp[:, 0] * n[:, 1] - pp[:, 1] * n[:, 0]], 1)
        dd = pp - q
        res = (dd[:, 0] * n[:, 0] + dd[:, 1] * n[:, 1]) + dd[:, 2] * n[:, 2]
        F = np.hstack([c, n]).astype(np.float64)                           # float32 factors, float64 sums
        A = F.T @ F
        b = -F.T @ res.astype(np.float64)
        x = np.linalg.solve(A, b)
        dT = np.eye(4); dT[:3, :3] = angle_axis(x[:3]); dT[:3, 3] = x[3:]
        T = matmul32(f32(dT), T)
        hist_R.append(T[:3, :3].astype(np.float64)); hist_t.append(T[:3, 3].astype(np.float64))
        stop = it + 1 >= max_iter
        if len(hist_R) > smooth:
            re = np.mean([quat_angle(hist_R[-1 - j], hist_R[-2 - j]) for j in range(smooth)])
            te = np.mean([np.linalg.norm(hist_t[-1 - j] - hist_t[-2 - j]) for j in range(smooth)])
            if re < min_rot and te < min_trans:
                stop = True
        if stop:
            break
    Tmu = np.eye(4, dtype=np.float32); Tmu[:3, 3] = mu
    return matmul32(matmul32(Tmu, T), M0), it + 1, n_used


@pytest.mark.parametrize("config,trial,n,ratio", [(2, 0, 8192, 0.55), (3, 1, 8192, 0.65), (3, 0, 12000, 0.7), (2, 3, 8192, 0.4)])
def test_oracle_meets_the_1e5_bar_against_the_float32_geometry_rederivation(orc, config, trial, n, ratio):
    """Lidar-shaped clouds (VLP-16 room, HDL-64 street): measured agreement 3e-17 m / 5e-18 rad, i.e. the same float32
    numbers.  The lattice-sampled cube pairs of C5 are left to the float64 test below at 5e-5: they differ from ANY independent
    implementation by 1.5e-5 m, not through rounding but through the two decisions DESIGN.md section 3 lists for inputs a real
    sensor never produces -- exact distance ties (lowest index here, first-visited in scipy / libnabo) and the normal of a
    neighbourhood with two equal eigenvalues along the cube's edges (any vector of a plane is an eigenvector)."""
    pair = synth.make_pair(config, trial, n)
    T_ind, it_ind, used_ind = independent_icp_f32(pair["ref"], pair["read"], ratio)
    o = orc.icp(pair["ref"], pair["read"], orc.default_config(ratio=ratio))
    assert o.rc == 0 and o.iterations == it_ind
    assert [int(t["n_used"]) for t in o.trace] == used_ind
    d = o.T.astype(np.float64) @ np.linalg.inv(T_ind.astype(np.float64))
    assert np.linalg.norm(d[:3, 3]) <= TOL_F32, np.linalg.norm(d[:3, 3])
    assert rot_angle(d[:3, :3]) <= TOL_F32, rot_angle(d[:3, :3])


@pytest.mark.parametrize("config,trial,n,ratio", [(5, 0, 6000, 0.7), (5, 2, 8000, 0.6), (2, 0, 8192, 0.55)])
def test_oracle_matches_independent_float64_icp(orc, config, trial, n, ratio):
    pair = synth.make_pair(config, trial, n)
    T_ind, it_ind, used_ind = independent_icp(pair["ref"], pair["read"], ratio)
    o = orc.icp(pair["ref"], pair["read"], orc.default_config(ratio=ratio))
    assert o.rc == 0 and o.iterations == it_ind
    assert [int(t["n_used"]) for t in o.trace] == used_ind
    d = o.T.astype(np.float64) @ np.linalg.inv(T_ind)
    assert np.linalg.norm(d[:3, 3]) <= TOL, np.linalg.norm(d[:3, 3])
    assert rot_angle(d[:3, :3]) <= TOL, rot_angle(d[:3, :3])


def test_oracle_normals_match_numpy_eigh(orc):
    pair = synth.make_pair(3, 0, 8192)
    n_orc, _ = orc.surface_normals(pair["ref"], 20)
    n_ind = normals_knn(pair["ref"].astype(np.float64), 20)
    cosang = np.abs(np.einsum("ij,ij->i", n_orc[:, :3].astype(np.float64), n_ind))
    # neighbourhoods with two nearly equal small eigenvalues (edges, isolated points) have ill-defined normals
    assert np.mean(cosang > 1 - 1e-6) > 0.97 and np.median(1 - cosang) < 1e-9


def test_oracle_overlap_matches_python_dda(orc):
    """octomap's computeRayKeys restated straight from SURVEY.md A.8 in Python (float/double mix, tie order, exit rule)."""
    res = float(np.float32(0.2))
    inv = 1.0 / res

    def key(c):
        return int(np.floor(inv * c)) + 32768

    def ray(o, e):
        o = np.float32(o); e = np.float32(e)
        ko = [key(float(c)) for c in o]; ke = [key(float(c)) for c in e]
        if ko == ke:
            return set()
        d = (e - o).astype(np.float32)
        length = float(np.float32(np.sqrt(np.float32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))))
        d = (d / np.float32(length)).astype(np.float32)
        step = [int(np.sign(x)) for x in d]
        tmax, tdelta = [0.0] * 3, [0.0] * 3
        cur = list(ko)
        for i in range(3):
            if step[i] != 0:
                border = (float(cur[i] - 32768) + 0.5) * res + float(np.float32(step[i] * res * 0.5))
                tmax[i] = (border - float(o[i])) / float(d[i])
                tdelta[i] = res / abs(float(d[i]))
            else:
                tmax[i] = np.finfo(np.float64).max
        out = {tuple(cur)}
        while True:
            if tmax[0] < tmax[1]:
                dim = 0 if tmax[0] < tmax[2] else 2
            else:
                dim = 1 if tmax[1] < tmax[2] else 2
            cur[dim] += step[dim]; tmax[dim] += tdelta[dim]
            if cur == ke:
                break
            if min(tmax) > length:
                break
            out.add(tuple(cur))
        return out
    rng = np.random.default_rng(5)
    origin = np.array([0.13, -0.27, 1.73])
    pts = rng.uniform(-6, 6, (300, 3)).astype(np.float32)
    free, occ = set(), set()
    for p in pts:
        free |= ray(origin, p)
        occ.add(tuple(key(float(c)) for c in np.float32(p)))
    mine = {(x << 32) | (y << 16) | z for x, y, z in (free | occ)}
    theirs = set(int(k) for k in orc.ray_keys(pts, origin))
    assert len(theirs) == len(mine)
