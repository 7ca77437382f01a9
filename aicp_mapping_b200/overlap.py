"""Host-side mirror of the reference's overlap plug-in interface, over the C ABI.

Mirrors:
  aicp::AbstractOverlapper      aicp_core/include/aicp_overlap/abstract_overlapper.hpp:13-19
  aicp::OctreesOverlap          aicp_core/include/aicp_overlap/octrees_overlap.hpp:20-58, src/overlap/octrees_overlap.cpp:29-72
  aicp::create_overlapper       aicp_core/include/aicp_overlap/overlap.hpp:9-19
  OverlapParams                 aicp_core/include/aicp_overlap/common.hpp:7-15
  App::computeOverlap           aicp_core/src/registration/app.cpp:112-141

The reference returns octomap::ColorOcTree pointers that callers use for visualisation only (App ignores them,
app.cpp:132-135); this mirror returns the voxel counts instead.
"""
import ctypes as C
import sys
from dataclasses import dataclass, field

import numpy as np

from . import capi


@dataclass
class OctreeOverlapParams:
    # yaml_configurator.cpp:80-82 reads the value with as<float>() into this double field: 0.2 becomes (double)0.2f
    octomapResolution: float = float(np.float32(0.2))


@dataclass
class OverlapParams:
    type: str = ""
    loadPosesFromFile: str = ""
    octree_based: OctreeOverlapParams = field(default_factory=OctreeOverlapParams)


def _translation(pose):
    """Eigen::Isometry3d stand-in: a 4x4 matrix or a 3-vector; only the translation is used (octrees_overlap.cpp:229-230)."""
    p = np.asarray(pose, dtype=np.float64)
    if p.shape == (4, 4):
        return np.ascontiguousarray(p[:3, 3])
    if p.shape == (3,):
        return np.ascontiguousarray(p)
    raise ValueError("pose must be a 4x4 matrix or a translation 3-vector")


class B200Overlap:
    """AbstractOverlapper implemented by libaicp_b200.so (occupancy bitmaps + popcount intersection on the GPU)."""

    def __init__(self, params=None, device=-1):
        self.params_ = params or OverlapParams(type="B200")
        self._lib = capi.lib()
        self._h = C.c_void_p()
        rc = self._lib.aicp_b200_create(None, int(device), C.byref(self._h))
        if rc:
            raise capi.AicpError(rc, self._lib.aicp_b200_last_error(None).decode())
        self.overlap_ = np.float32(-1.0)
        self.counts = (0, 0, 0)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.aicp_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def computeOverlap(self, ref_cloud, read_cloud, ref_pose, read_pose):
        """octrees_overlap.cpp:29-72.  Returns (n_overlapping, n_ref_nodes, n_read_nodes)."""
        pr, nr, k1 = capi.ptr_and_count(ref_cloud)
        pq, nq, k2 = capi.ptr_and_count(read_cloud)
        ro, so = _translation(ref_pose), _translation(read_pose)
        ov = C.c_float()
        counts = (C.c_int64 * 3)()
        rc = self._lib.aicp_b200_overlap(self._h, pr, nr, ro.ctypes.data_as(C.POINTER(C.c_double)), pq, nq,
                                         so.ctypes.data_as(C.POINTER(C.c_double)),
                                         C.c_double(self.params_.octree_based.octomapResolution), C.byref(ov), counts)
        if rc:
            raise capi.AicpError(rc, self._lib.aicp_b200_last_error(self._h).decode())
        self.overlap_ = np.float32(ov.value)
        self.counts = (int(counts[0]), int(counts[1]), int(counts[2]))
        return self.counts

    def getOverlap(self):
        return self.overlap_


def create_overlapper(parameters, device=-1):
    """overlap.hpp:9-19 with the extra "B200" branch."""
    if parameters.type == "B200":
        return B200Overlap(parameters, device=device)
    sys.stderr.write("Invalid overlap type %s.\n" % parameters.type)
    return None


def computeOverlap(overlapper, reference_cloud, reading_cloud, reference_pose, reading_pose, localize_against_prior_map=False):
    """App::computeOverlap, app.cpp:112-141: 50 % is assumed when localising against a prior map."""
    if localize_against_prior_map:
        return np.float32(50.0)
    overlapper.computeOverlap(reference_cloud, reading_cloud, reference_pose, reading_pose)
    return overlapper.getOverlap()
