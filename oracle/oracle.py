"""ctypes binding of the CPU oracle (oracle/libaicp_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
The product package aicp_mapping_b200 never does.  PARITY UNPINNED -- see oracle/aicp_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libaicp_oracle.so")
_lib = None

ORC_MAX_ITERS = 256
STOP_NONE, STOP_COUNTER, STOP_DIFFERENTIAL = 0, 1, 2
ERR_NAMES = {0: "OK", 1: "BAD_ARG", 2: "KNN_TOO_LARGE", 3: "NO_VALID_MATCH", 4: "NAN", 5: "NONFINITE_INPUT", 6: "EXTENT"}


class IcpConfig(C.Structure):
    _fields_ = [("knn_normals", C.c_int32), ("reading_normals", C.c_int32), ("ratio", C.c_float),
                ("max_iterations", C.c_int32), ("min_diff_rot", C.c_float), ("min_diff_trans", C.c_float),
                ("smooth_length", C.c_int32), ("use_kdtree", C.c_int32), ("threads", C.c_int32)]


class IterTrace(C.Structure):
    _fields_ = [("T_iter", C.c_float * 16), ("limit_d2", C.c_float), ("n_valid", C.c_int64), ("n_used", C.c_int64),
                ("rot_err", C.c_double), ("trans_err", C.c_double)]


class IcpResult(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("stop_reason", C.c_int32), ("weighted_point_used_ratio", C.c_float),
                ("mean_ref", C.c_float * 3), ("trace", IterTrace * ORC_MAX_ITERS)]


def build(force=False):
    """Compile the oracle with oracle/Makefile (gcc)."""
    srcs = [os.path.join(_HERE, f) for f in ("aicp_oracle.c", "aicp_oracle_overlap.c", "aicp_oracle_filters.c", "aicp_oracle_prefilter.c", "aicp_oracle_alignability.c", "aicp_oracle_ingest.c", "aicp_oracle.h", "Makefile")]
    if not force and os.path.exists(_LIB_PATH) and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_atan2_pos.restype = C.c_double
        _lib.orc_atan2_pos.argtypes = [C.c_double, C.c_double]
        _lib.orc_sincos.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib.orc_autotune_ratio.restype = C.c_float
        _lib.orc_autotune_ratio.argtypes = [C.c_float, C.c_char_p]
        _lib.orc_ray_keys.restype = C.c_int64
        _lib.orc_crop_box.restype = C.c_int64
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _ptr(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def to_xyzw(xyz):
    """n x 3 (or n x 4) -> n x 4 float32 with pad = 1 (cloudIO.cpp:81-98)."""
    xyz = np.asarray(xyz, dtype=np.float32)
    if xyz.shape[1] == 4:
        return np.ascontiguousarray(xyz)
    out = np.ones((xyz.shape[0], 4), dtype=np.float32)
    out[:, :3] = xyz
    return out


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("oracle %s failed: %s" % (what, ERR_NAMES.get(rc, rc)))


def default_config(ratio=0.70, **kw):
    cfg = IcpConfig(knn_normals=20, reading_normals=0, ratio=ratio, max_iterations=20, min_diff_rot=0.001,
                    min_diff_trans=0.01, smooth_length=4, use_kdtree=1, threads=1)
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def surface_normals(pts, k=20, use_kdtree=True, threads=1, want_knn=True):
    pts = to_xyzw(pts)
    n = pts.shape[0]
    normals = np.zeros((n, 4), dtype=np.float32)
    knn = np.zeros((n, k), dtype=np.int32) if want_knn else None
    rc = lib().orc_surface_normals(_ptr(pts), C.c_int64(n), C.c_int32(k), int(use_kdtree), int(threads),
                                   _ptr(normals), _ptr(knn, C.c_int32))
    _check(rc, "surface_normals")
    return normals, knn


def match(ref, qry, use_kdtree=True, threads=1):
    ref, qry = to_xyzw(ref), to_xyzw(qry)
    idx = np.zeros(qry.shape[0], dtype=np.int32)
    d2 = np.zeros(qry.shape[0], dtype=np.float32)
    rc = lib().orc_match(_ptr(ref), C.c_int64(ref.shape[0]), _ptr(qry), C.c_int64(qry.shape[0]), int(use_kdtree),
                         int(threads), _ptr(idx, C.c_int32), _ptr(d2))
    _check(rc, "match")
    return idx, d2


def trim_threshold(d2, ratio):
    d2 = _f32(d2)
    limit = C.c_float()
    nv = C.c_int64()
    rc = lib().orc_trim_threshold(_ptr(d2), C.c_int64(d2.shape[0]), C.c_float(ratio), C.byref(limit), C.byref(nv))
    _check(rc, "trim_threshold")
    return np.float32(limit.value), nv.value


def normal_equations(p, ref, normals, idx, d2, limit):
    p, ref, normals = to_xyzw(p), to_xyzw(ref), _f32(normals)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    d2 = _f32(d2)
    hi = np.zeros(27, dtype=np.int64)
    lo = np.zeros(27, dtype=np.uint64)
    used = C.c_int64()
    rc = lib().orc_normal_equations(_ptr(p), C.c_int64(p.shape[0]), _ptr(ref), _ptr(normals), _ptr(idx, C.c_int32),
                                    _ptr(d2), C.c_float(limit), _ptr(hi, C.c_int64), _ptr(lo, C.c_uint64),
                                    C.byref(used))
    _check(rc, "normal_equations")
    return hi, lo, used.value


def solve6(hi, lo):
    hi = np.ascontiguousarray(hi, dtype=np.int64)
    lo = np.ascontiguousarray(lo, dtype=np.uint64)
    x = np.zeros(6, dtype=np.float64)
    path = lib().orc_solve6(_ptr(hi, C.c_int64), _ptr(lo, C.c_uint64), _ptr(x, C.c_double))
    return x, path


def pose_increment(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    dT = np.zeros(16, dtype=np.float32)
    lib().orc_pose_increment(_ptr(x, C.c_double), _ptr(dT))
    return dT.reshape(4, 4).T.copy()       # column-major -> numpy row-major matrix


def sincos(x):
    s, c = C.c_double(), C.c_double()
    lib().orc_sincos(C.c_double(x), C.byref(s), C.byref(c))
    return s.value, c.value


def atan2_pos(y, x):
    return lib().orc_atan2_pos(C.c_double(y), C.c_double(x))


def transform_points(T, pts):
    """T: 4x4 numpy (row-major view of the matrix).  Returns n x 4."""
    pts = to_xyzw(pts)
    Tc = np.ascontiguousarray(np.asarray(T, dtype=np.float32).T).ravel()     # column-major
    out = np.zeros_like(pts)
    lib().orc_transform_points(_ptr(Tc), _ptr(pts), C.c_int64(pts.shape[0]), _ptr(out))
    return out


class IcpOutput:
    pass


def icp(ref, read, cfg=None, init_T=None, want_trace_idx=False, want_reading=True, want_normals=False):
    """Full chain.  Returns an IcpOutput with T (4x4 numpy matrix), iterations, stop_reason, trace..."""
    ref, read = to_xyzw(ref), to_xyzw(read)
    cfg = cfg or default_config()
    n_ref, n_read = ref.shape[0], read.shape[0]
    Tc = np.zeros(16, dtype=np.float32)
    init = None
    if init_T is not None:
        init = np.ascontiguousarray(np.asarray(init_T, dtype=np.float32).T).ravel()
    out_reading = np.zeros((n_read, 4), dtype=np.float32) if want_reading else None
    out_normals = np.zeros((n_ref, 4), dtype=np.float32) if want_normals else None
    tidx = np.full((cfg.max_iterations, n_read), -1, dtype=np.int32) if want_trace_idx else None
    res = IcpResult()
    rc = lib().orc_icp(_ptr(ref), C.c_int64(n_ref), _ptr(read), C.c_int64(n_read), _ptr(init), C.byref(cfg),
                       _ptr(Tc), _ptr(out_reading), _ptr(out_normals), _ptr(tidx, C.c_int32), C.byref(res))
    out = IcpOutput()
    out.rc = rc
    out.error = ERR_NAMES.get(rc, str(rc))
    out.T = Tc.reshape(4, 4).T.copy()
    out.iterations = res.iterations
    out.stop_reason = res.stop_reason
    out.weighted_point_used_ratio = np.float32(res.weighted_point_used_ratio)
    out.mean_ref = np.array(list(res.mean_ref), dtype=np.float32)
    out.reading = out_reading
    out.normals = out_normals
    out.trace_idx = tidx[:res.iterations] if tidx is not None else None
    out.trace = []
    for i in range(res.iterations):
        t = res.trace[i]
        out.trace.append(dict(T_iter=np.array(list(t.T_iter), dtype=np.float32).reshape(4, 4).T.copy(),
                              limit_d2=np.float32(t.limit_d2), n_valid=t.n_valid, n_used=t.n_used,
                              rot_err=t.rot_err, trans_err=t.trans_err))
    return out


def overlap(ref, ref_origin, read, read_origin, resolution=float(np.float32(0.2))):
    """Returns (overlap_pct float32, (n_inter, n_ref_keys, n_read_keys)).  Default resolution is (double)0.2f
    (yaml_configurator.cpp:80-82 reads as<float>() into a double field)."""
    ref, read = to_xyzw(ref), to_xyzw(read)
    ro = np.ascontiguousarray(ref_origin, dtype=np.float64)
    so = np.ascontiguousarray(read_origin, dtype=np.float64)
    ov = C.c_float()
    counts = np.zeros(3, dtype=np.int64)
    rc = lib().orc_overlap(_ptr(ref), C.c_int64(ref.shape[0]), _ptr(ro, C.c_double), _ptr(read),
                           C.c_int64(read.shape[0]), _ptr(so, C.c_double), C.c_double(resolution), C.byref(ov),
                           _ptr(counts, C.c_int64))
    _check(rc, "overlap")
    return np.float32(ov.value), tuple(int(c) for c in counts)


def ray_keys(pts, origin, resolution=float(np.float32(0.2))):
    pts = to_xyzw(pts)
    o = np.ascontiguousarray(origin, dtype=np.float64)
    n = lib().orc_ray_keys(_ptr(pts), C.c_int64(pts.shape[0]), _ptr(o, C.c_double), C.c_double(resolution), None,
                           C.c_int64(0))
    keys = np.zeros(n, dtype=np.uint64)
    lib().orc_ray_keys(_ptr(pts), C.c_int64(pts.shape[0]), _ptr(o, C.c_double), C.c_double(resolution),
                       _ptr(keys, C.c_uint64), C.c_int64(n))
    return keys


def autotune_ratio(overlap_pct):
    buf = C.create_string_buffer(32)
    r = lib().orc_autotune_ratio(C.c_float(overlap_pct), buf)
    return np.float32(r), buf.value.decode()


def crop_box(cloud, bmin, bmax, rpy, translation):
    """getPointsInOrientedBox / pcl::CropBox (filteringUtils.cpp:621-637): points inside the oriented box, input order."""
    pts = to_xyzw(cloud)
    out = np.zeros_like(pts)
    r = np.ascontiguousarray(rpy, dtype=np.float32)
    t = np.ascontiguousarray(translation, dtype=np.float32)
    m = lib().orc_crop_box(_ptr(pts), C.c_int64(pts.shape[0]), C.c_float(bmin), C.c_float(bmax), _ptr(r), _ptr(t), _ptr(out))
    return out[:m].copy()


# ---- pre-filter (aicp_oracle_prefilter.c): regionGrowingUniformPlaneSegmentationFilter, filteringUtils.cpp:5-104 ----
class PrefilterConfig(C.Structure):
    _fields_ = [("leaf_size", C.c_float), ("knn_normals", C.c_int32), ("n_neighbours", C.c_int32),
                ("min_cluster_size", C.c_int32), ("max_cluster_size", C.c_int32), ("smoothness_threshold", C.c_float),
                ("curvature_threshold", C.c_float)]


def prefilter_default_config(**kw):
    cfg = PrefilterConfig()
    lib().orc_prefilter_default_config(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def voxel_grid(cloud, leaf=np.float32(0.08)):
    """pcl::VoxelGrid centroids in ascending voxel-index order (filteringUtils.cpp:10-13)."""
    pts = to_xyzw(cloud)
    out = np.zeros_like(pts)
    L = lib()
    L.orc_voxel_grid.restype = C.c_int64
    m = L.orc_voxel_grid(_ptr(pts), C.c_int64(pts.shape[0]), C.c_float(leaf), _ptr(out))
    if m < 0:
        raise RuntimeError("oracle voxel_grid failed: %s" % ERR_NAMES.get(-m, -m))
    return out[:m].copy()


def region_growing(normals, knn, n_nb=15, min_size=50, max_size=1000000, cos_thr=None, curv_thr=1.0):
    """Sequential pcl::RegionGrowing over given normals (m x 4: nx, ny, nz, curvature) and neighbour lists (m x k)."""
    normals = _f32(normals)
    knn = np.ascontiguousarray(knn, dtype=np.int32)
    m, k = knn.shape
    if cos_thr is None:
        cos_thr = np.cos(np.float32(3.0 / 180.0 * np.pi), dtype=np.float32)
    labels = np.zeros(m, dtype=np.int32)
    L = lib()
    L.orc_region_growing.restype = C.c_int64
    nc = L.orc_region_growing(_ptr(normals), _ptr(knn, C.c_int32), C.c_int64(m), C.c_int32(k), C.c_int32(n_nb),
                              C.c_int32(min_size), C.c_int32(max_size), C.c_float(cos_thr), C.c_float(curv_thr),
                              _ptr(labels, C.c_int32))
    return labels, int(nc)


class PrefilterOutput:
    pass


def prefilter(cloud, cfg=None, viewpoint=None, threads=1):
    """The whole pre-filter.  Returns PrefilterOutput: cloud (n_out x 4), sampled, normals (nx, ny, nz, curvature), labels,
    n_clusters, rc."""
    pts = to_xyzw(cloud)
    n = pts.shape[0]
    cfg = cfg or prefilter_default_config()
    vp = np.ascontiguousarray(viewpoint, dtype=np.float32) if viewpoint is not None else None
    sampled = np.zeros((max(n, 1), 4), dtype=np.float32)
    normals = np.zeros((max(n, 1), 4), dtype=np.float32)
    labels = np.zeros(max(n, 1), dtype=np.int32)
    out = np.zeros((max(n, 1), 4), dtype=np.float32)
    counts = np.zeros(3, dtype=np.int64)
    rc = lib().orc_prefilter(_ptr(pts), C.c_int64(n), C.byref(cfg), _ptr(vp), int(threads), _ptr(sampled), _ptr(normals),
                             _ptr(labels, C.c_int32), _ptr(out), _ptr(counts, C.c_int64))
    o = PrefilterOutput()
    o.rc = rc
    o.error = ERR_NAMES.get(rc, str(rc))
    o.sampled = sampled[:counts[0]].copy()
    o.normals = normals[:counts[0]].copy()
    o.labels = labels[:counts[0]].copy()
    o.n_clusters = int(counts[1])
    o.cloud = out[:counts[2]].copy()
    return o


# ---- FOV overlap + alignability (aicp_oracle_alignability.c): filteringUtils.cpp:111-576 ----
def _pose16(pose):
    """4x4 (numpy, row-major view of the matrix) -> 16 doubles column-major, like Eigen::Isometry3d::matrix().data()."""
    return np.ascontiguousarray(np.asarray(pose, dtype=np.float64).T).ravel()


def fov_overlap(cloudA, cloudB, poseA, poseB, sensor_range, angular_view):
    """overlapFilter: returns (overlap_pct float32, accepted A n x 4, accepted B n x 4)."""
    a, b = to_xyzw(cloudA), to_xyzw(cloudB)
    outa, outb = np.zeros((max(a.shape[0], 1), 4), np.float32), np.zeros((max(b.shape[0], 1), 4), np.float32)
    counts = np.zeros(2, dtype=np.int64)
    pa, pb = _pose16(poseA), _pose16(poseB)
    L = lib()
    L.orc_fov_overlap.restype = C.c_float
    ov = L.orc_fov_overlap(_ptr(a), C.c_int64(a.shape[0]), _ptr(b), C.c_int64(b.shape[0]), _ptr(pa, C.c_double), _ptr(pb, C.c_double),
                           C.c_float(sensor_range), C.c_float(angular_view), _ptr(outa), _ptr(outb), _ptr(counts, C.c_int64))
    return np.float32(ov), outa[:counts[0]].copy(), outb[:counts[1]].copy()


def euler_angles_012(R):
    R = np.ascontiguousarray(R, dtype=np.float32)
    out = np.zeros(3, dtype=np.float32)
    lib().orc_euler_angles_012(_ptr(R), _ptr(out))
    return out


def alignability(cloudA, cloudB, poseA, poseB, cfg=None, threads=1):
    """alignabilityFilter: returns (alignability_pct float32, matching int32[n clusters of B], (clusters A, clusters B, matched))."""
    a, b = to_xyzw(cloudA), to_xyzw(cloudB)
    cfg = cfg or prefilter_default_config()
    pa, pb = _pose16(poseA), _pose16(poseB)
    al = C.c_float()
    matching = np.full(max(b.shape[0], 1), -1, dtype=np.int32)
    info = np.zeros(3, dtype=np.int64)
    rc = lib().orc_alignability(_ptr(a), C.c_int64(a.shape[0]), _ptr(b), C.c_int64(b.shape[0]), _ptr(pa, C.c_double), _ptr(pb, C.c_double),
                                C.byref(cfg), int(threads), C.byref(al), _ptr(matching, C.c_int32), _ptr(info, C.c_int64))
    _check(rc, "alignability")
    return np.float32(al.value), matching[:info[1]].copy(), tuple(int(x) for x in info)


# ---- sweep accumulation (aicp_oracle_ingest.c): velodyne_accumulator.cpp:31-73 ----
def pose_to_float_transform(pose):
    """Translation3f(t) * Quaternionf(R) as a 4x4 float32 matrix (numpy row-major view)."""
    T = np.zeros(16, dtype=np.float32)
    P = _pose16(pose)
    lib().orc_pose_to_float_transform(_ptr(P, C.c_double), _ptr(T))
    return T.reshape(4, 4).T.copy()


def accumulate_sweeps(sweeps, poses, half=30.0):
    """VelodyneAccumulatorROS::processLidar over a batch: returns the accumulated cloud (n x 4 float32)."""
    L = lib()
    L.orc_accumulate_sweep.restype = C.c_int64
    parts = []
    for sw, pose in zip(sweeps, poses):
        a = to_xyzw(sw)
        out = np.zeros((max(a.shape[0], 1), 4), dtype=np.float32)
        P = _pose16(pose)
        m = L.orc_accumulate_sweep(_ptr(a), C.c_int64(a.shape[0]), C.c_float(half), _ptr(P, C.c_double), _ptr(out))
        parts.append(out[:m].copy())
    return np.concatenate(parts, 0) if parts else np.zeros((0, 4), np.float32)
