"""getPointsInOrientedBox = pcl::CropBox (aicp_core/src/utils/filteringUtils.cpp:621-637), SURVEY.md 8(f) rank 3.
not gpu: the oracle against an independent numpy float64 statement of the box test and analytic cases.
gpu    : the CUDA compaction (csrc/crop.cu) against the oracle, bit for bit and in input order."""
import numpy as np
import pytest

import aicp_mapping_b200 as ab
from aicp_mapping_b200 import filtering, synth


def numpy_crop(pts, bmin, bmax, rpy, t):
    """Independent float64 statement: R = Rz(yaw) Ry(pitch) Rx(roll) (pcl::getTransformation), keep min <= R^T (p - t) <= max."""
    R = synth.rigid(0, 0, 0, float(rpy[0]), float(rpy[1]), float(rpy[2]))[:3, :3]
    local = (pts[:, :3].astype(np.float64) - np.asarray(t, dtype=np.float64)) @ R      # rows of (R^T d)^T = d^T R
    return np.all((local >= bmin) & (local <= bmax), axis=1)


def test_oracle_crop_box_matches_numpy_and_analytic_cases(orc):
    rng = np.random.default_rng(3)
    pts = rng.uniform(-20, 20, (20000, 3)).astype(np.float32)
    rpy, t = np.float32([0.02, -0.03, 0.8]), np.float32([1.5, -2.0, 0.3])
    kept = orc.crop_box(pts, -8.0, 8.0, rpy, t)
    mask = numpy_crop(pts, -8.0, 8.0, rpy, t)
    # float32 vs float64 can only disagree for points within rounding distance of a face
    local = (pts.astype(np.float64) - t.astype(np.float64)) @ synth.rigid(0, 0, 0, *rpy.astype(np.float64))[:3, :3]
    near_face = np.any(np.abs(np.abs(local) - 8.0) < 1e-4, axis=1)
    kept_set = set(map(tuple, kept[:, :3]))
    for p, m, nf in zip(pts, mask, near_face):
        if not nf:
            assert (tuple(p) in kept_set) == bool(m)
    # order preserved, pad column carried through
    idx = [i for i, p in enumerate(pts) if tuple(p) in kept_set]
    assert np.array_equal(kept[:, :3], pts[idx])
    # analytic: axis-aligned unit box, faces inclusive, NaN dropped
    q = np.float32([[0, 0, 0], [1, 1, 1], [1.0000001, 0, 0], [-1, -1, -1], [np.nan, 0, 0], [0, 0, 2]])
    k = orc.crop_box(q, -1.0, 1.0, np.zeros(3, np.float32), np.zeros(3, np.float32))
    assert np.array_equal(k[:, :3], q[[0, 1, 3]])
    # a yaw of 90 degrees maps the box's x axis onto world y
    q = np.float32([[0, 3, 0], [3, 0, 0]])
    k = orc.crop_box(q, -1.0, 4.0, np.float32([0, 0, np.pi / 2]), np.zeros(3, np.float32))
    assert np.array_equal(k[:, :3], q[[0]])


def test_euler_angles_restate_eigen_convention():
    for seed in range(20):
        rng = np.random.default_rng(seed)
        R = synth.rigid(0, 0, 0, *rng.uniform(-1.4, 1.4, 3))[:3, :3]
        a, b, c = filtering.euler_angles_xyz(R).astype(np.float64)
        assert 0.0 <= a <= np.pi + 1e-6                          # Eigen 3.3: first angle in [0, pi]
        back = synth.rigid(0, 0, 0, a, 0, 0)[:3, :3] @ synth.rigid(0, 0, 0, 0, b, 0)[:3, :3] @ synth.rigid(0, 0, 0, 0, 0, c)[:3, :3]
        assert np.abs(back - R).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 31, 2047, 2048, 2049, 4097, 100003])
def test_crop_box_parity_ragged_sizes(orc, n):
    rng = np.random.default_rng(n)
    pts = rng.uniform(-12, 12, (n, 3)).astype(np.float32)
    pts[rng.integers(0, n, max(1, n // 50))] = np.nan              # non-finite points are dropped
    rpy, t = rng.uniform(-0.5, 0.5, 3).astype(np.float32), rng.uniform(-2, 2, 3).astype(np.float32)
    crop = ab.B200CropBox()
    for lo, hi in ((-6.0, 6.0), (-100.0, 100.0), (50.0, 60.0), (-1.0, 0.5)):
        g = crop.filter(pts, lo, hi, rpy, t)
        o = orc.crop_box(pts, lo, hi, rpy, t)
        assert g.shape == o.shape and np.array_equal(g.view(np.uint32), o.view(np.uint32))
    crop.close()


@pytest.mark.gpu
def test_crop_box_full_map_and_registration_against_the_crop(orc):
    """BASELINE config 4 the way App does it (app.cpp:41-69): crop the 10 485 760-point map to +-15 m around the prior pose,
    then register the reading against the CROP.  The crop equals the oracle's; the registration against the device-resident
    crop equals the registration against the same points passed from the host."""
    case = synth.make_map_case(n_map=10_485_760, n_read=122_880, trial=0, n_poses=1)
    mp, rd = case["map"], case["readings"][0]
    prior = np.eye(4, dtype=np.float32)
    prior[:3, 3] = np.asarray(rd["read_origin"], dtype=np.float32)
    prior[:3, :3] = synth.rigid(0, 0, 0, 0.01, -0.02, 0.7)[:3, :3].astype(np.float32)
    rpy = filtering.euler_angles_xyz(prior[:3, :3])
    crop = ab.B200CropBox()
    g = crop.filter(mp, -15.0, 15.0, rpy, prior[:3, 3])
    o = orc.crop_box(mp, -15.0, 15.0, rpy, prior[:3, 3])
    assert g.shape == o.shape and np.array_equal(g.view(np.uint32), o.view(np.uint32))
    assert 50_000 < len(g) < len(mp) // 4
    import torch
    addr, n_kept = crop.filter(torch.from_numpy(ab.capi.to_xyzw(mp)).cuda(), -15.0, 15.0, rpy, prior[:3, 3], keep_on_device=True)
    assert n_kept == len(g)

    class DevView:                                   # the library-owned device buffer as a "device cloud" for ptr_and_count
        def __init__(self, a, n): self._a, self.shape, self.dtype = a, (n, 4), "torch.float32"
        def data_ptr(self): return self._a
        def dim(self): return 2
        def is_contiguous(self): return True
    reg = ab.B200Registration()
    reg.setConfig(ratio=0.5, max_iterations=20)
    T_dev = reg.registerClouds(DevView(addr, n_kept), rd["read"])
    T_host = reg.registerClouds(g, rd["read"])
    assert np.array_equal(T_dev.view(np.uint32), T_host.view(np.uint32))
    reg.close(); crop.close()


@pytest.mark.gpu
def test_device_resident_map_append_and_crop(orc):
    """App's map handling on the GPU: updateCloud, merge of aligned clouds (concatenation, app.cpp:476-480), crop around a
    pose; equal to the oracle's crop of the concatenated host map, and growth keeps the points already stored."""
    rng = np.random.default_rng(11)
    parts = [rng.uniform(-30, 30, (n, 3)).astype(np.float32) for n in (50000, 1, 70001, 300000)]
    m = ab.B200Map()
    m.updateCloud(parts[0])
    for p in parts[1:]:
        m.append(p)
    full = np.concatenate(parts, 0)
    assert m.size() == len(full)
    origin = synth.rigid(3.0, -4.0, 0.5, 0.01, 0.02, -1.1).astype(np.float32)
    view = m.cropAround(15.0, origin)
    got = m.cropToHost()
    want = orc.crop_box(full, -15.0, 15.0, filtering.euler_angles_xyz(origin[:3, :3]), origin[:3, 3])
    assert view.shape[0] == len(want) and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    m.updateCloud(parts[1])                       # replace: a one-point map
    assert m.size() == 1 and m.cropAround(1000.0, np.eye(4)).shape[0] == 1
    m.close()
