"""Host-side mirror of the reference's map-handling filter that sits next to the registration path
(aicp_core/src/utils/filteringUtils.cpp:621-637, called from App on the prior / built map before every registration,
app.cpp:41-69): getPointsInOrientedBox = pcl::CropBox.  The crop itself runs on the GPU (csrc/crop.cu) through the C ABI;
this module only prepares the arguments the way the reference does."""
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
