"""Host-side mirror of the reference's cloud filters that sit next to the registration path (aicp_utils/filteringUtils.hpp):
  getPointsInOrientedBox = pcl::CropBox (filteringUtils.cpp:621-637; App crops the prior / built map with it before every
    registration, app.cpp:41-69)                                                           -> csrc/crop.cu
  regionGrowingUniformPlaneSegmentationFilter = VoxelGrid + NormalEstimation + RegionGrowing (filteringUtils.cpp:5-104;
    App pre-filters every reading, the first cloud and periodically the map with it, app.cpp:102-110,295,486-493)
                                                                                           -> csrc/prefilter.cu
Everything runs on the GPU through the C ABI; this module only prepares the arguments the way the reference does."""
import ctypes as C
import math

import numpy as np

from . import capi


def euler_angles_xyz(R):
    """[UPSTREAM, recalled] Eigen 3.3 `R.eulerAngles(0, 1, 2)` in float32: (a, b, c) with R = Rx(a) * Ry(b) * Rz(c), a in
    [0, pi].  The C++ adapter calls Eigen itself; this restatement serves the Python harness only."""
    R = np.asarray(R, dtype=np.float32)
    f = np.float32
    r0 = f(math.atan2(R[1, 2], R[2, 2]))
    c2 = f(math.hypot(R[0, 0], R[0, 1]))
    if r0 > 0:
        r0 = f(r0 - f(math.pi))
        r1 = f(math.atan2(-R[0, 2], -c2))
    else:
        r1 = f(math.atan2(-R[0, 2], c2))
    s1, c1 = f(math.sin(r0)), f(math.cos(r0))
    r2 = f(math.atan2(s1 * R[2, 0] - c1 * R[1, 0], c1 * R[1, 1] - s1 * R[2, 1]))
    return np.array([-r0, -r1, -r2], dtype=np.float32)


class B200CropBox:
    """pcl::CropBox<pcl::PointXYZ> as the reference configures it (setMin/setMax with one scalar per side, setRotation,
    setTranslation, filter)."""

    def __init__(self, device=-1):
        self._lib = capi.lib()
        h = C.c_void_p()
        rc = self._lib.aicp_b200_create(None, int(device), C.byref(h))
        if rc:
            raise capi.AicpError(rc, self._lib.aicp_b200_last_error(None).decode())
        self._h = h

    def close(self):
        if self._h:
            self._lib.aicp_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def filter(self, cloud, box_min, box_max, rotation_rpy, translation, keep_on_device=False):
        """Returns the kept points (n_kept x 4 float32, input order).  keep_on_device: returns (device address, n_kept) of the
        library-owned result instead, valid until the next call -- pass it to registerClouds as a device cloud."""
        p, n, keep = capi.ptr_and_count(cloud)
        rpy = np.ascontiguousarray(rotation_rpy, dtype=np.float32)
        t = np.ascontiguousarray(translation, dtype=np.float32)
        n_out = C.c_int64()
        out = None if keep_on_device else np.zeros((n, 4), dtype=np.float32)
        rc = self._lib.aicp_b200_crop_box(self._h, p, n, C.c_float(box_min), C.c_float(box_max),
                                          rpy.ctypes.data_as(C.POINTER(C.c_float)), t.ctypes.data_as(C.POINTER(C.c_float)),
                                          C.c_void_p(out.ctypes.data) if out is not None else None, C.byref(n_out))
        if rc:
            raise capi.AicpError(rc, self._lib.aicp_b200_last_error(self._h).decode())
        if keep_on_device:
            return self._lib.aicp_b200_get_cropped(self._h, None), int(n_out.value)
        return out[:n_out.value].copy()


class B200Map(B200CropBox):
    """The prior / built map of App (aligned_map_, prior_map_) kept on the GPU: upload once, append aligned clouds, crop
    around the prior pose before every registration (app.cpp:41-69,469-493)."""

    def _check(self, rc):
        if rc:
            raise capi.AicpError(rc, self._lib.aicp_b200_last_error(self._h).decode())

    def updateCloud(self, cloud):
        p, n, keep = capi.ptr_and_count(cloud)
        self._check(self._lib.aicp_b200_map_append(self._h, p, n, 1))

    def append(self, cloud):
        p, n, keep = capi.ptr_and_count(cloud)
        self._check(self._lib.aicp_b200_map_append(self._h, p, n, 0))

    def size(self):
        return int(self._lib.aicp_b200_map_size(self._h))

    def cropAround(self, half_extent, origin):
        """getPointsInOrientedBox(map copy, -half_extent, +half_extent, origin): returns a device cloud view of the crop
        (valid until the next crop), to be passed to registerClouds / setReference."""
        origin = np.asarray(origin, dtype=np.float32)
        rpy = euler_angles_xyz(origin[:3, :3])
        t = np.ascontiguousarray(origin[:3, 3], dtype=np.float32)
        n_out = C.c_int64()
        self._check(self._lib.aicp_b200_map_crop(self._h, C.c_float(-half_extent), C.c_float(half_extent),
                                                 rpy.ctypes.data_as(C.POINTER(C.c_float)), t.ctypes.data_as(C.POINTER(C.c_float)),
                                                 C.byref(n_out)))
        return DeviceCloudView(self._lib.aicp_b200_get_cropped(self._h, None), int(n_out.value))

    def cropToHost(self):
        """The last crop as an n x 4 float32 array."""
        n = C.c_int64()
        self._lib.aicp_b200_get_cropped(self._h, C.byref(n))
        out = np.zeros((n.value, 4), dtype=np.float32)
        self._check(self._lib.aicp_b200_download_cropped(self._h, C.c_void_p(out.ctypes.data), n.value))
        return out


    def prefilter(self, cfg=None):
        """app.cpp:486-493: prior_map_ <- regionGrowingUniformPlaneSegmentationFilter(prior_map_), on the device.  Returns
        the capi.PrefilterInfo of the run."""
        info = capi.PrefilterInfo()
        n_out = C.c_int64()
        self._check(self._lib.aicp_b200_map_prefilter(self._h, C.byref(cfg) if cfg is not None else None, C.byref(n_out), C.byref(info)))
        return info

    def download(self):
        """The whole map as an n x 4 float32 array (a crop with an all-enclosing box; non-finite points do not survive it)."""
        if self.size() == 0:
            return np.zeros((0, 4), np.float32)
        z = np.zeros(3, dtype=np.float32)
        n_out = C.c_int64()
        self._check(self._lib.aicp_b200_map_crop(self._h, C.c_float(-1.0e30), C.c_float(1.0e30), z.ctypes.data_as(C.POINTER(C.c_float)),
                                                 z.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n_out)))
        return self.cropToHost()


class B200Prefilter:
    """regionGrowingUniformPlaneSegmentationFilter on the GPU.  One instance owns one library handle (device buffers are
    reused across calls, as App reuses its filter for every cloud)."""

    def __init__(self, device=-1, cfg=None):
        self._lib = capi.lib()
        h = C.c_void_p()
        rc = self._lib.aicp_b200_create(None, int(device), C.byref(h))
        if rc:
            raise capi.AicpError(rc, self._lib.aicp_b200_last_error(None).decode())
        self._h = h
        self.cfg = cfg or default_prefilter_config()
        self.info = capi.PrefilterInfo()

    def close(self):
        if self._h:
            self._lib.aicp_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise capi.AicpError(rc, self._lib.aicp_b200_last_error(self._h).decode())

    def voxelGrid(self, cloud, leaf=None):
        """pcl::VoxelGrid::filter (filteringUtils.cpp:10-13): n_voxels x 4 float32, ascending voxel index."""
        p, n, keep = capi.ptr_and_count(cloud)
        out = np.zeros((max(n, 1), 4), dtype=np.float32)
        n_out = C.c_int64()
        self._check(self._lib.aicp_b200_voxel_grid(self._h, p, n, C.c_float(self.cfg.leaf_size if leaf is None else leaf),
                                                   C.c_void_p(out.ctypes.data), C.byref(n_out)))
        return out[:n_out.value].copy()

    def filter(self, cloud_in, view_point=None, keep_on_device=False):
        """First overload (filteringUtils.cpp:5-45): returns cloud_out, the kept clusters concatenated (n_out x 4).
        keep_on_device: returns a DeviceCloudView of the library-owned result instead (valid until the next call)."""
        p, n, keep = capi.ptr_and_count(cloud_in)
        vp = None
        if view_point is not None:
            vp = np.ascontiguousarray(view_point, dtype=np.float32)
        out = None if keep_on_device else np.zeros((max(n, 1), 4), dtype=np.float32)
        n_out = C.c_int64()
        self._check(self._lib.aicp_b200_prefilter(self._h, p, n, C.byref(self.cfg), vp.ctypes.data_as(C.POINTER(C.c_float)) if vp is not None else None,
                                                  C.c_void_p(out.ctypes.data) if out is not None else None, C.byref(n_out), C.byref(self.info)))
        if keep_on_device:
            return DeviceCloudView(self._lib.aicp_b200_get_prefiltered(self._h, None), int(n_out.value))
        return out[:n_out.value].copy()

    def segments(self):
        """By-products of the last filter() call, as the second overload returns them (filteringUtils.cpp:51-104):
        (cloud_sampled n x 4, normals n x 4 = (nx, ny, nz, curvature), labels n int32, clusters = list of index arrays)."""
        n = int(self.info.n_sampled)
        sampled = np.zeros((n, 4), dtype=np.float32)
        normals = np.zeros((n, 4), dtype=np.float32)
        labels = np.full(n, -1, dtype=np.int32)
        if n:
            self._check(self._lib.aicp_b200_prefilter_get_sampled(self._h, C.c_void_p(sampled.ctypes.data), n))
            self._check(self._lib.aicp_b200_prefilter_get_labels(self._h, C.c_void_p(labels.ctypes.data), n))
            if n > self.cfg.knn_normals:
                self._check(self._lib.aicp_b200_prefilter_get_normals(self._h, C.c_void_p(normals.ctypes.data), n))
        order = np.argsort(labels, kind="stable")
        order = order[labels[order] >= 0]
        bounds = np.flatnonzero(np.diff(labels[order])) + 1 if order.size else np.zeros(0, np.int64)
        clusters = np.split(order, bounds) if order.size else []
        return sampled, normals, labels, clusters


def _pose16(pose):
    """4x4 pose (numpy, row-major view of the matrix) -> 16 doubles column-major == Eigen::Isometry3d::matrix().data()."""
    return np.ascontiguousarray(np.asarray(pose, dtype=np.float64).T).ravel()


class B200Alignability:
    """overlapFilter + alignabilityFilter + the SVM, i.e. what App::computeAlignmentRisk runs (app.cpp:143-185), on the GPU."""

    def __init__(self, device=-1, svm_model=None):
        self._lib = capi.lib()
        h = C.c_void_p()
        rc = self._lib.aicp_b200_create(None, int(device), C.byref(h))
        if rc:
            raise capi.AicpError(rc, self._lib.aicp_b200_last_error(None).decode())
        self._h = h
        if svm_model:
            self._check(self._lib.aicp_b200_svm_load(self._h, str(svm_model).encode()))

    def close(self):
        if self._h:
            self._lib.aicp_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise capi.AicpError(rc, self._lib.aicp_b200_last_error(self._h).decode())

    def overlapFilter(self, cloudA, cloudB, poseA, poseB, sensor_range, angular_view):
        """filteringUtils.cpp:111-193: returns (overlap_pct float32, accepted_pointsA, accepted_pointsB)."""
        pa, na, ka = capi.ptr_and_count(cloudA)
        pb, nb, kb = capi.ptr_and_count(cloudB)
        outa, outb = np.zeros((max(na, 1), 4), np.float32), np.zeros((max(nb, 1), 4), np.float32)
        counts = (C.c_int64 * 2)()
        ov = C.c_float()
        PA, PB = _pose16(poseA), _pose16(poseB)
        self._check(self._lib.aicp_b200_fov_overlap(self._h, pa, na, pb, nb, PA.ctypes.data_as(C.POINTER(C.c_double)),
                                                    PB.ctypes.data_as(C.POINTER(C.c_double)), C.c_float(sensor_range), C.c_float(angular_view),
                                                    C.c_void_p(outa.ctypes.data), C.c_void_p(outb.ctypes.data), counts, C.byref(ov)))
        return np.float32(ov.value), outa[:counts[0]].copy(), outb[:counts[1]].copy()

    def alignabilityFilter(self, cloudA, cloudB, poseA, poseB, cfg=None):
        """filteringUtils.cpp:196-430: returns (alignability_pct float32, matching_indeces int32[clusters of B], (clusters A,
        clusters B, matched))."""
        pa, na, ka = capi.ptr_and_count(cloudA)
        pb, nb, kb = capi.ptr_and_count(cloudB)
        al = C.c_float()
        cap = max(nb // 2, 1)
        matching = np.full(cap, -1, dtype=np.int32)
        info = (C.c_int64 * 3)()
        PA, PB = _pose16(poseA), _pose16(poseB)
        self._check(self._lib.aicp_b200_alignability(self._h, pa, na, pb, nb, PA.ctypes.data_as(C.POINTER(C.c_double)),
                                                     PB.ctypes.data_as(C.POINTER(C.c_double)), C.byref(cfg) if cfg is not None else None,
                                                     C.byref(al), matching.ctypes.data_as(C.POINTER(C.c_int32)), cap, info))
        return np.float32(al.value), matching[:info[1]].copy(), (int(info[0]), int(info[1]), int(info[2]))

    def computeAlignmentRisk(self, reference_cloud, reading_cloud, reference_pose, reading_pose, sensor_range, angular_view, octree_overlap):
        """app.cpp:143-185: returns (fov_overlap, alignability, risk_prediction)."""
        pa, na, ka = capi.ptr_and_count(reference_cloud)
        pb, nb, kb = capi.ptr_and_count(reading_cloud)
        fov, al, risk = C.c_float(), C.c_float(), C.c_double()
        PA, PB = _pose16(reference_pose), _pose16(reading_pose)
        self._check(self._lib.aicp_b200_alignment_risk(self._h, pa, na, pb, nb, PA.ctypes.data_as(C.POINTER(C.c_double)),
                                                       PB.ctypes.data_as(C.POINTER(C.c_double)), C.c_float(sensor_range), C.c_float(angular_view),
                                                       C.c_float(octree_overlap), C.byref(fov), C.byref(al), C.byref(risk)))
        return np.float32(fov.value), np.float32(al.value), float(risk.value)


def default_prefilter_config(**kw):
    cfg = capi.PrefilterConfig()
    capi.lib().aicp_b200_prefilter_default_config(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def regionGrowingUniformPlaneSegmentationFilter(cloud_in, view_point=None, prefilter=None):
    """filteringUtils.cpp:5-45 (view_point None) / :51-104 (view_point given: returns (cloud_sampled_out with normals as an
    n x 8 array (x, y, z, 1, nx, ny, nz, curvature), clusters))."""
    own = prefilter is None
    prefilter = prefilter or B200Prefilter()
    try:
        out = prefilter.filter(cloud_in, view_point)
        if view_point is None:
            return out
        sampled, normals, labels, clusters = prefilter.segments()
        return np.concatenate([sampled, normals], axis=1), clusters
    finally:
        if own:
            prefilter.close()


class DeviceCloudView:
    """A library-owned device buffer of n (x, y, z, pad) records, accepted wherever a device cloud is (capi.ptr_and_count)."""

    def __init__(self, address, n):
        self._a, self.shape, self.dtype = address, (n, 4), "torch.float32"

    def data_ptr(self):
        return self._a

    def dim(self):
        return 2

    def is_contiguous(self):
        return True


def getPointsInOrientedBox(cloud, box_min, box_max, origin, cropper=None):
    """filteringUtils.cpp:621-637: crop `cloud` with the box [min, max]^3 placed at the 4x4 pose `origin`."""
    origin = np.asarray(origin, dtype=np.float32)
    own = cropper is None
    cropper = cropper or B200CropBox()
    try:
        return cropper.filter(cloud, box_min, box_max, euler_angles_xyz(origin[:3, :3]), origin[:3, 3])
    finally:
        if own:
            cropper.close()
