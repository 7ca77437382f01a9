"""Host-side mirror of the ingest either side of the path:
  VelodyneAccumulatorROS (aicp_ros/src/velodyne_accumulator.cpp:31-81)      -> B200VelodyneAccumulator (csrc/ingest.cu, on the GPU)
  pcl::io::loadPCDFile / PCDWriter as the replay and the tools use them     -> readPCD / writePCD            (library host code)
  PoseFileReader::readPoseFile (aicp_utils/poseFileReader.hpp:46-78)        -> readPoseFile                  (library host code)
  App::processFromFile (aicp_core/src/registration/app.cpp:250-279)         -> processFromFile (the replay loop)"""
import ctypes as C
import os

import numpy as np

from . import capi
from .filtering import DeviceCloudView, _pose16


def _err():
    return C.create_string_buffer(512)


def readPCD(path):
    """pcl::io::loadPCDFile<pcl::PointXYZ>: n x 4 float32 (x, y, z, 1)."""
    L, n, e = capi.lib(), C.c_int64(), _err()
    rc = L.aicp_b200_read_pcd(str(path).encode(), None, 0, C.byref(n), e, 512)
    if rc:
        raise capi.AicpError(rc, e.value.decode(errors="replace"))
    out = np.zeros((max(n.value, 1), 4), dtype=np.float32)
    rc = L.aicp_b200_read_pcd(str(path).encode(), C.c_void_p(out.ctypes.data), n.value, C.byref(n), e, 512)
    if rc:
        raise capi.AicpError(rc, e.value.decode(errors="replace"))
    return out[:n.value].copy()


def readPLY(path):
    """pcl::io::loadPLYFile<pcl::PointXYZ> (the prior map, app_ros.cpp:301): n x 4 float32 (x, y, z, 1)."""
    L, n, e = capi.lib(), C.c_int64(), _err()
    rc = L.aicp_b200_read_ply(str(path).encode(), None, 0, C.byref(n), e, 512)
    if rc:
        raise capi.AicpError(rc, e.value.decode(errors="replace"))
    out = np.zeros((max(n.value, 1), 4), dtype=np.float32)
    rc = L.aicp_b200_read_ply(str(path).encode(), C.c_void_p(out.ctypes.data), n.value, C.byref(n), e, 512)
    if rc:
        raise capi.AicpError(rc, e.value.decode(errors="replace"))
    return out[:n.value].copy()


def writePCD(path, cloud):
    """pcl::PCDWriter::writeBinary of a pcl::PointXYZ cloud (cloudIO.cpp:64, create_cube_cloud.cpp:84)."""
    a = capi.to_xyzw(cloud)
    e = _err()
    rc = capi.lib().aicp_b200_write_pcd(str(path).encode(), C.c_void_p(a.ctypes.data), a.shape[0], e, 512)
    if rc:
        raise capi.AicpError(rc, e.value.decode(errors="replace"))


def readPoseFile(path):
    """PoseFileReader::readPoseFile: returns a list of (counter, sec, nsec, pose 4x4 float64)."""
    L, n, e = capi.lib(), C.c_int64(), _err()
    rc = L.aicp_b200_read_pose_file(str(path).encode(), None, None, 0, C.byref(n), e, 512)
    if rc:
        raise capi.AicpError(rc, e.value.decode(errors="replace"))
    rows = np.zeros((max(n.value, 1), 3), dtype=np.int64)
    poses = np.zeros((max(n.value, 1), 16), dtype=np.float64)
    rc = L.aicp_b200_read_pose_file(str(path).encode(), rows.ctypes.data_as(C.POINTER(C.c_int64)), poses.ctypes.data_as(C.POINTER(C.c_double)),
                                    n.value, C.byref(n), e, 512)
    if rc:
        raise capi.AicpError(rc, e.value.decode(errors="replace"))
    return [(int(rows[i, 0]), int(rows[i, 1]), int(rows[i, 2]), poses[i].reshape(4, 4).T.copy()) for i in range(n.value)]


class B200VelodyneAccumulator:
    """VelodyneAccumulatorROS without ROS: processLidar(sweep, body_pose) crops to +-30 m around the sensor, moves the sweep to
    the inertial frame and appends it to the accumulated cloud, which lives on the GPU."""

    def __init__(self, batch_size=7, box_half=30.0, device=-1):      # batch_size: aicp_ros/launch/aicp.launch:68
        self._lib = capi.lib()
        h = C.c_void_p()
        rc = self._lib.aicp_b200_create(None, int(device), C.byref(h))
        if rc:
            raise capi.AicpError(rc, self._lib.aicp_b200_last_error(None).decode())
        self._h = h
        self.batch_size, self.box_half = int(batch_size), float(box_half)
        self.counter, self.finished, self._fresh = 0, False, True

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

    def processLidar(self, sweep, body_pose):
        if self.finished:                                  # velodyne_accumulator.cpp:33-35
            return 0
        p, n, keep = capi.ptr_and_count(sweep)
        P = _pose16(body_pose)
        added = C.c_int64()
        self._check(self._lib.aicp_b200_accumulate_sweep(self._h, p, n, C.c_float(self.box_half), P.ctypes.data_as(C.POINTER(C.c_double)),
                                                         1 if self._fresh else 0, C.byref(added)))
        self._fresh = False
        self.counter += 1
        if self.counter >= self.batch_size:
            self.finished = True
        return int(added.value)

    def getFinished(self):
        return self.finished

    def clearCloud(self):
        self.counter, self.finished, self._fresh = 0, False, True

    def getCloud(self):
        """The accumulated cloud as a device cloud view (valid until the next processLidar)."""
        n = C.c_int64()
        addr = self._lib.aicp_b200_get_accumulated(self._h, C.byref(n))
        return DeviceCloudView(addr, 0 if self._fresh else int(n.value))

    def download(self):
        v = self.getCloud()
        out = np.zeros((max(v.shape[0], 1), 4), dtype=np.float32)
        if v.shape[0]:
            self._check(self._lib.aicp_b200_download_accumulated(self._h, C.c_void_p(out.ctypes.data), v.shape[0]))
        return out[:v.shape[0]].copy()


def processFromFile(file_path):
    """App::processFromFile (app.cpp:250-279): yields (utime, cloud n x 4, world_to_body 4x4) for every row of
    <file_path>/aicp_input_poses.csv, reading <file_path>/cloud_<counter>_<sec>_<nsec>.pcd; stops at the first unreadable
    cloud like the reference."""
    for counter, sec, nsec, pose in readPoseFile(os.path.join(file_path, "aicp_input_poses.csv")):
        pcd = os.path.join(file_path, "cloud_%d_%d_%d.pcd" % (counter, sec, nsec))
        try:
            cloud = readPCD(pcd)
        except capi.AicpError:
            print("Couldn't read file %s" % pcd)
            return
        yield int(sec * 1e6 + nsec), cloud, pose             # app.cpp:267 (sic: nsec added to microseconds)
