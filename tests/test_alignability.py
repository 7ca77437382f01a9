"""overlapFilter + alignabilityFilter (aicp_core/src/utils/filteringUtils.cpp:111-576) and App::computeAlignmentRisk
(app.cpp:143-185), SURVEY.md 8(f) rank 2.
not gpu: the oracle (oracle/aicp_oracle_alignability.c) on analytic scenes and against an independent numpy statement of the
         field-of-view test.
gpu    : the CUDA path through the C ABI against the oracle: accepted clouds bit for bit and in order, alignability bit for
         bit, the cluster matching, and the whole computeAlignmentRisk chain."""
import os

import numpy as np
import pytest

import aicp_mapping_b200 as ab
from aicp_mapping_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEFAULT_MODEL = os.path.join(GOLDEN, "svm_models", "svm_1000training_thresh50_cross_validation_opencv3.xml")


def pose(t, yaw=0.0):
    return synth.rigid(t[0], t[1], t[2], 0.0, 0.0, yaw)


def corridor(rng, n=40000, end_wall=False, shift=(0.0, 0.0, 0.0)):
    w1 = np.c_[rng.uniform(0, 20, n), np.full(n, -1.0), rng.uniform(0, 2.5, n)]
    w2 = np.c_[rng.uniform(0, 20, n), np.full(n, 1.0), rng.uniform(0, 2.5, n)]
    fl = np.c_[rng.uniform(0, 20, n), rng.uniform(-1, 1, n), np.zeros(n)]
    parts = [w1, w2, fl]
    if end_wall:
        parts.append(np.c_[np.full(n // 2, 20.0), rng.uniform(-1, 1, n // 2), rng.uniform(0, 2.5, n // 2)])
    pts = np.concatenate(parts, 0)
    return (pts + rng.normal(0.0, 0.01, pts.shape) + np.asarray(shift)).astype(np.float32)      # 1 cm range noise: a noise-free
    # plane has a zero-thickness bounding box that the other cloud's points miss by rounding


def room_box(rng, n=30000, shift=(0.0, 0.0, 0.0)):
    """Three mutually perpendicular planes, seen from inside the corner."""
    a = np.c_[rng.uniform(0, 6, n), rng.uniform(0, 6, n), np.zeros(n)]
    b = np.c_[rng.uniform(0, 6, n), np.zeros(n), rng.uniform(0, 3, n)]
    c = np.c_[np.zeros(n), rng.uniform(0, 6, n), rng.uniform(0, 3, n)]
    pts = np.concatenate([a, b, c], 0)
    return (pts + rng.normal(0.0, 0.01, pts.shape) + np.asarray(shift)).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------------
def test_oracle_fov_overlap_matches_numpy_statement(orc):
    rng = np.random.default_rng(11)
    A = rng.uniform(-40, 40, (20000, 3)).astype(np.float32)
    B = rng.uniform(-40, 40, (15000, 3)).astype(np.float32)
    PA, PB = pose([1.0, -2.0, 0.5], 0.3), pose([-3.0, 1.0, 0.7], -1.1)
    for rng_m, view in ((30.0, 270.0), (100.0, 360.0), (25.0, 180.0)):
        ov, a, b = orc.fov_overlap(A, B, PA, PB, rng_m, view)
        thresh = 180.0 - (360.0 - view) / 2.0

        def accept(cloud, other):
            local = (cloud.astype(np.float64) - other[:3, 3]) @ other[:3, :3]
            r = np.linalg.norm(local, axis=1)
            theta = np.degrees(np.arctan2(local[:, 1], local[:, 0]))
            margin = np.minimum(np.abs(np.abs(theta) - thresh), np.abs(r - rng_m))
            return (np.abs(theta) < thresh) & (r < rng_m), margin < 1e-3
        ka, fa = accept(A, PB)
        kb, fb = accept(B, PA)
        # float32 / float64 can only disagree within rounding distance of the cone or the range sphere
        assert abs(a.shape[0] - ka.sum()) <= fa.sum() and abs(b.shape[0] - kb.sum()) <= fb.sum()
        want = np.float32(np.float32(a.shape[0]) / np.float32(A.shape[0]) * (np.float32(b.shape[0]) / np.float32(B.shape[0])))
        assert ov == np.float32(np.float64(want) * 100.0)
        # accepted points are the input points again, up to the float32 round trip through the other sensor's frame
        idx = np.flatnonzero(ka & ~fa)[:50]
        if view == 360.0 and fa.sum() == 0:
            assert np.abs(a[:, :3] - A[ka]).max() < 2e-5


def test_oracle_fov_overlap_analytic(orc):
    I = np.eye(4)
    pts = np.float32([[1, 0, 0], [0, 1, 0], [-1, 0.5, 0], [-1, 0, 0], [-1, 1.0001, 0], [50, 0, 0], [0, 0, 0]])
    ov, a, b = orc.fov_overlap(pts, pts[:1], I, I, 30.0, 270.0)        # thresh = 135 deg: the rear 90-degree wedge is cut
    assert np.array_equal(a[:, :3], pts[[0, 1, 4, 6]])
    assert ov == np.float32(np.float64(np.float32(4.0 / 7.0) * np.float32(1.0)) * 100.0)
    ov, a, b = orc.fov_overlap(pts, pts[:1], I, I, 100.0, 360.0)       # |theta| < 180: only the exact rear axis is cut
    assert np.array_equal(a[:, :3], pts[[0, 1, 2, 4, 5, 6]])


def test_oracle_alignability_analytic_scenes(orc):
    rng = np.random.default_rng(0)
    P = pose([3.0, 0.2, 1.0])
    cor = corridor(rng)
    al, matching, info = orc.alignability(cor, corridor(rng, shift=(0.1, 0, 0)), P, P, threads=4)
    assert info[2] >= 3 and al < 0.5                         # no constraint along the corridor: degenerate
    Q = pose([3.0, 3.0, 1.5])
    al, matching, info = orc.alignability(room_box(rng), room_box(rng, shift=(0.05, 0.02, 0)), Q, Q, threads=4)
    assert info[2] >= 3 and al > 30.0                        # three perpendicular planes: well constrained
    # every matched pair is (cluster of A, cluster of B) with a B index used once
    assert len(set(m for m in matching if m >= 0)) == info[2]
    # nothing to match: two unrelated scenes far apart
    al, matching, info = orc.alignability(cor, room_box(rng, shift=(200.0, 0, 0)), P, Q, threads=4)
    assert al == 0.0 and info[2] == 0
    # the result does not depend on where the world origin is (shifted covariance in the normals, exact cluster sums)
    A, B = room_box(rng), room_box(rng, shift=(0.05, 0.02, 0))
    near = orc.alignability(A, B, Q, Q, threads=4)
    sh = np.float32([200.0, -150.0, 0.0])
    far = orc.alignability(A + sh, B + sh, pose([203.0, -147.0, 1.5]), pose([203.0, -147.0, 1.5]), threads=4)
    assert near[2] == far[2] and abs(near[0] - far[0]) < 0.05
    # empty and tiny inputs
    al, matching, info = orc.alignability(np.zeros((0, 3), np.float32), cor, P, P)
    assert al == 0.0 and info == (0, info[1], 0)


def numpy_alignability(sa, na, la, sb, nb, lb):
    """Independent float64 statement of filteringUtils.cpp:229-372 written from the reference (no code shared with the oracle):
    cluster OBBs (MomentOfInertiaEstimation::getOBB), CropBox counts through eulerAngles(0,1,2) -> Rz Ry Rx, greedy matching,
    PCA of the matched normals and their mirror images.  Inputs: sampled clouds, normals, cluster labels of both clouds."""
    from scipy.spatial.transform import Rotation

    def boxes(pts, nrm, lab):
        out = []
        for c in range(lab.max() + 1):
            p = pts[lab == c, :3].astype(np.float64)
            mean = p.mean(0)
            w, v = np.linalg.eigh(np.cov((p - mean).T, bias=True))
            axes = v[:, ::-1].copy()                                   # major, middle, minor
            for k in range(3):                                          # canonical sign: largest |component| positive
                if axes[np.argmax(np.abs(axes[:, k])), k] < 0:
                    axes[:, k] = -axes[:, k]
            if np.dot(axes[:, 0], np.cross(axes[:, 1], axes[:, 2])) <= 0:
                axes[:, 0] = -axes[:, 0]
            proj = (p - mean) @ axes
            lo, hi = proj.min(0), proj.max(0)
            shift = (hi + lo) / 2
            lo, hi, pos = lo - shift, hi - shift, mean + axes @ shift
            lo[2], hi[2] = 3 * lo[2], 3 * hi[2]
            a, b, g = Rotation.from_matrix(axes).as_euler("XYZ")        # R = Rx(a) Ry(b) Rz(g), b in [-pi/2, pi/2] (scipy)
            if a < 0:                                                   # Eigen 3.3 documents the ranges [0:pi] x [-pi:pi] x [-pi:pi]:
                a, b, g = a + np.pi, np.pi - b, g + np.pi               # the other triple of the same rotation
                b, g = (b + np.pi) % (2 * np.pi) - np.pi, (g + np.pi) % (2 * np.pi) - np.pi
            # the two triples describe the same Rx Ry Rz but NOT the same Rz Ry Rx: which one Eigen returns decides the box
            Rbox = Rotation.from_euler("xyz", [a, b, g]).as_matrix()    # extrinsic xyz = Rz(g) Ry(b) Rx(a): pcl::getTransformation
            out.append(dict(n=p.shape[0], ncen=nrm[lab == c, :3].astype(np.float64).mean(0), R=Rbox, t=pos, lo=lo, hi=hi,
                            S=nrm[lab == c, :3].astype(np.float64).T @ nrm[lab == c, :3].astype(np.float64)))
        return out

    def count(box, pts):
        loc = (pts[:, :3].astype(np.float64) - box["t"]) @ box["R"]
        return np.all((loc >= box["lo"]) & (loc <= box["hi"]), axis=1)

    A, B = boxes(sa, na, la), boxes(sb, nb, lb)
    mi, mo = [-1] * len(B), [-1.0] * len(B)
    fragile = set()                 # B clusters whose decision hangs on acos(x) with x within rounding of 1
    for i, a in enumerate(A):
        best, mx = -1, 0.0
        for j, b in enumerate(B):
            cosang = np.dot(a["ncen"], b["ncen"]) / (np.linalg.norm(a["ncen"]) * np.linalg.norm(b["ncen"]))
            ov = (count(b, sa[la == i]).sum() / a["n"]) * (count(a, sb[lb == j]).sum() / b["n"]) * 100.0
            if cosang > 1.0 - 1e-6 and ov > 0:
                fragile.add(j)      # the reference takes acos of a float32 ratio that rounding can push above 1: NaN, "dist < 20" false
            dist = np.degrees(np.arccos(np.clip(cosang, -1, 1)))
            if ov > mx and dist < 20:
                best, mx = j, ov
        if mx > 0 and (mi[best] == -1 or mx > mo[best]):
            mi[best], mo[best] = i, mx

    def value(matching):
        S = sum((A[i]["S"] for i in matching if i >= 0), np.zeros((3, 3)))
        if not any(i >= 0 for i in matching):
            return 0.0
        lam = np.sort(np.linalg.eigvalsh(S))[::-1]
        return 100.0 * lam[2] / lam[0]
    return value, mi, fragile


@pytest.mark.parametrize("case", ["room", "corridor", "vlp16"])
def test_oracle_alignability_matches_independent_numpy_statement(orc, pair_cache, case):
    rng = np.random.default_rng(21)
    if case == "room":
        A, B, PA = room_box(rng), room_box(rng, shift=(0.05, 0.02, 0)), pose([3.0, 3.0, 1.5])
        PB = PA
    elif case == "corridor":
        A, B, PA = corridor(rng, end_wall=True), corridor(rng, end_wall=True, shift=(0.1, 0, 0)), pose([3.0, 0.2, 1.0])
        PB = PA
    else:
        p = pair_cache(2, 0, 16384)
        A, B, PA, PB = p["ref"], p["read"], pose(p["ref_origin"]), pose(p["read_origin"])
    al, matching, info = orc.alignability(A, B, PA, PB, threads=4)
    oa = orc.prefilter(A, viewpoint=PA[:3, 3].astype(np.float32), threads=4)
    ob = orc.prefilter(B, viewpoint=PB[:3, 3].astype(np.float32), threads=4)
    value, mi, fragile = numpy_alignability(oa.sampled, oa.normals, oa.labels, ob.sampled, ob.normals, ob.labels)
    # points within float32 rounding of a box face may be counted differently; the matching is robust to that.  What is NOT robust,
    # in the reference itself (filteringUtils.cpp:252): acos(dot / (|a| |b|)) of two mean normals that are parallel to within float32
    # rounding is NaN when the ratio lands on 1.0000001, and NaN < 20 rejects the pair -- such clusters may legitimately differ
    got = [int(m) for m in matching]
    assert all(g == w for j, (g, w) in enumerate(zip(got, mi)) if j not in fragile)
    assert all(g in (-1, w) for j, (g, w) in enumerate(zip(got, mi)) if j in fragile)
    want = value(got)                                              # the PCA stage, from the oracle's own matching
    assert abs(float(al) - want) <= 1e-3 * max(1.0, want)


def test_euler_angles_restatement_matches_python_mirror(orc):
    from aicp_mapping_b200 import filtering
    for seed in range(10):
        rng = np.random.default_rng(seed)
        R = synth.rigid(0, 0, 0, *rng.uniform(-1.4, 1.4, 3))[:3, :3].astype(np.float32)
        got = orc.euler_angles_012(R)
        assert np.abs(got - filtering.euler_angles_xyz(R)).max() < 2e-6


# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def al():
    a = ab.B200Alignability(device=0, svm_model=DEFAULT_MODEL)
    yield a
    a.close()


@pytest.mark.gpu
def test_fov_overlap_parity(orc, al, pair_cache):
    rng = np.random.default_rng(5)
    cases = []
    A = rng.uniform(-40, 40, (70001, 3)).astype(np.float32)
    B = rng.uniform(-40, 40, (4097, 3)).astype(np.float32)
    cases.append((A, B, pose([1.0, -2.0, 0.5], 0.3), pose([-3.0, 1.0, 0.7], -1.1)))
    p = pair_cache(2)
    cases.append((p["ref"], p["read"], pose(p["ref_origin"]), pose(p["read_origin"], 0.05)))
    for A, B, PA, PB in cases:
        for rng_m, view in ((30.0, 270.0), (100.0, 360.0), (5.0, 90.0)):
            ov, a, b = al.overlapFilter(A, B, PA, PB, rng_m, view)
            o_ov, o_a, o_b = orc.fov_overlap(A, B, PA, PB, rng_m, view)
            assert ov == o_ov
            assert np.array_equal(a.view(np.uint32), o_a.view(np.uint32)) and np.array_equal(b.view(np.uint32), o_b.view(np.uint32))
    ov, a, b = al.overlapFilter(A[:1], B[:1], np.eye(4), np.eye(4), 1e-3, 270.0)      # nothing accepted
    assert ov == 0.0 and a.shape == (0, 4) and b.shape == (0, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["corridor", "room", "vlp16_pair", "hdl64_pair", "disjoint"])
def test_alignability_parity(orc, al, pair_cache, case):
    rng = np.random.default_rng(3)
    if case == "corridor":
        A, B = corridor(rng, end_wall=True), corridor(rng, end_wall=True, shift=(0.1, 0, 0))
        PA = PB = pose([3.0, 0.2, 1.0])
    elif case == "room":
        A, B = room_box(rng), room_box(rng, shift=(0.05, 0.02, 0))
        PA = PB = pose([3.0, 3.0, 1.5])
    elif case == "vlp16_pair":
        p = pair_cache(2)
        A, B, PA, PB = p["ref"], p["read"], pose(p["ref_origin"]), pose(p["read_origin"])
    elif case == "hdl64_pair":
        p = pair_cache(3, 0, 60000)
        A, B, PA, PB = p["ref"], p["read"], pose(p["ref_origin"]), pose(p["read_origin"])
    else:
        A, B = corridor(rng), room_box(rng, shift=(200.0, 0, 0))
        PA, PB = pose([3.0, 0.2, 1.0]), pose([203.0, 3.0, 1.5])
    got, matching, info = al.alignabilityFilter(A, B, PA, PB)
    want, o_matching, o_info = orc.alignability(A, B, PA, PB, threads=8)
    assert info == o_info
    assert np.array_equal(matching, o_matching)
    assert got == want, (got, want)
    if case == "disjoint":
        assert got == 0.0 and info[2] == 0


@pytest.mark.gpu
def test_compute_alignment_risk_chain(orc, al, pair_cache):
    """app.cpp:143-185 end to end: FOV overlap -> alignability of the accepted clouds -> SVM(octree overlap, alignability)."""
    from oracle import aicp_oracle_svm as svm_orc
    p = pair_cache(2)
    PA, PB = pose(p["ref_origin"]), pose(p["read_origin"])
    o_ov, counts = orc.overlap(p["ref"], p["ref_origin"], p["read"], p["read_origin"])
    fov, ali, risk = al.computeAlignmentRisk(p["ref"], p["read"], PA, PB, 30.0, 270.0, float(o_ov))
    o_fov, o_a, o_b = orc.fov_overlap(p["ref"], p["read"], PA, PB, 30.0, 270.0)
    o_ali, _, _ = orc.alignability(o_a, o_b, PA, PB, threads=8)
    o_risk = svm_orc.test(svm_orc.load_model(DEFAULT_MODEL), np.array([[float(o_ov), float(o_ali)]]))[0]
    assert fov == o_fov and ali == o_ali
    assert abs(risk - o_risk) <= 1e-6 and 0.0 < risk < 1.0
    # tiny inputs go through without clusters: alignability 0
    fov, ali, risk = al.computeAlignmentRisk(p["ref"][:20], p["read"][:20], PA, PB, 30.0, 270.0, 50.0)
    assert ali == 0.0


@pytest.mark.gpu
def test_pipeline_batch_with_risk_gate(orc, pair_cache):
    """aicp_b200_pipeline_batch = App::runAicpPipeline with failure_prediction_mode per pair (app.cpp:218-247): overlap ->
    alignment risk -> registration only when risk <= threshold; every number against the oracle."""
    from oracle import aicp_oracle_svm as svm_orc
    rng = np.random.default_rng(9)
    pairs, poses = [], []
    for t in range(3):
        p = pair_cache(2, t, 16384)
        pairs.append((p["ref"], p["read"])); poses.append((pose(p["ref_origin"]), pose(p["read_origin"])))
    cor = corridor(rng, 20000, end_wall=True)
    pairs.append((cor, corridor(rng, 20000, end_wall=True, shift=(0.05, 0.01, 0)))); poses.append((pose([3.0, 0.2, 1.0]), pose([3.05, 0.21, 1.0])))
    rb = room_box(rng, 20000)
    pairs.append((rb, room_box(rng, 20000, shift=(0.05, 0.02, 0)))); poses.append((pose([3.0, 3.0, 1.5]), pose([3.05, 3.02, 1.5])))
    model = svm_orc.load_model(DEFAULT_MODEL)
    want = []
    for (a, b), (PA, PB) in zip(pairs, poses):
        ov, _ = orc.overlap(a, PA[:3, 3], b, PB[:3, 3])
        _, fa, fb = orc.fov_overlap(a, b, PA, PB, 30.0, 270.0)
        al, _, _ = orc.alignability(fa, fb, PA, PB, threads=8)
        risk = svm_orc.test(model, np.array([[float(ov), float(al)]]))[0]
        want.append((ov, al, risk))
    thr = float(np.median([w[2] for w in want]))
    reg = ab.B200Registration(device=0)
    try:
        for streams in (1, 3):
            T, ov, al, risk, stats, status, ms = reg.pipelineBatch(pairs, poses, DEFAULT_MODEL, 30.0, 270.0, risk_threshold=thr, streams=streams)
            assert not status.any() and ms > 0
            n_skipped = 0
            for i, ((a, b), w) in enumerate(zip(pairs, want)):
                assert ov[i] == w[0] and al[i] == w[1] and abs(risk[i] - w[2]) <= 1e-6
                if risk[i] > thr:
                    n_skipped += 1
                    assert np.array_equal(T[i], np.eye(4, dtype=np.float32)) and stats[i].iterations == 0
                else:
                    o = orc.icp(a, b, orc.default_config(ratio=ab.autotune_ratio(float(w[0])), threads=8))
                    assert o.rc == 0 and stats[i].iterations == o.iterations and np.array_equal(T[i], o.T)
            assert 0 < n_skipped < len(pairs)
    finally:
        reg.close()


@pytest.mark.gpu
def test_pipeline_batch_from_raw_clouds(orc):
    """prefilter_first: raw accumulated sweeps -> pre-filter x 2 -> overlap -> alignment risk -> registration, per pair, with the
    filtered clouds resident on the device; every number against the oracle chain."""
    from oracle import aicp_oracle_svm as svm_orc
    model = svm_orc.load_model(DEFAULT_MODEL)
    pairs, poses = [], []
    for t in range(3):
        a = synth.raw_sweep(2, 20 + t, n_sweeps=3)
        E = synth.rigid(0.08, -0.05, 0.01, 0.0, 0.0, 0.02)
        b_cloud = synth.apply_T(E, synth.raw_sweep(2, 20 + t, n_sweeps=4)["cloud"][5000:])
        pairs.append((a["cloud"], b_cloud)); poses.append((pose(a["origin"]), pose(a["origin"] + [0.35, 0, 0])))
    reg = ab.B200Registration(device=0)
    try:
        T, ov, al, risk, stats, status, ms = reg.pipelineBatch(pairs, poses, DEFAULT_MODEL, 30.0, 270.0, risk_threshold=1.0, streams=2,
                                                               prefilter_first=True)
        assert not status.any()
        for i, ((a, b), (PA, PB)) in enumerate(zip(pairs, poses)):
            fa, fb = orc.prefilter(a, threads=8).cloud, orc.prefilter(b, threads=8).cloud
            assert list(reg.n_filtered[i]) == [fa.shape[0], fb.shape[0]]
            o_ov, _ = orc.overlap(fa, PA[:3, 3], fb, PB[:3, 3])
            _, ka, kb = orc.fov_overlap(fa, fb, PA, PB, 30.0, 270.0)
            o_al, _, _ = orc.alignability(ka, kb, PA, PB, threads=8)
            o_risk = svm_orc.test(model, np.array([[float(o_ov), float(o_al)]]))[0]
            assert ov[i] == o_ov and al[i] == o_al and abs(risk[i] - o_risk) <= 1e-6
            o = orc.icp(fa, fb, orc.default_config(ratio=ab.autotune_ratio(float(o_ov)), threads=8))
            assert o.rc == 0 and stats[i].iterations == o.iterations and np.array_equal(T[i], o.T)
    finally:
        reg.close()
