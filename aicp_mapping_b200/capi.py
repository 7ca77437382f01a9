"""ctypes binding of libaicp_b200.so (include/aicp_b200.h).  Loads the in-tree library only; there is no fallback:
if the library is missing or no CUDA device is present, calls raise."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# AICP_B200_LIB selects an experiment build of the same library (tools/ab_bench.py); there is still no fallback
LIB_PATH = os.environ.get("AICP_B200_LIB") or os.path.join(HERE, "lib", "libaicp_b200.so")
MAX_ITERS = 256

OK = 0
ERR_NAMES = {0: "OK", 1: "BAD_ARG", 2: "KNN_TOO_LARGE", 3: "NO_VALID_MATCH", 4: "NAN", 5: "NONFINITE_INPUT", 6: "EXTENT",
             7: "CONFIG", 8: "CUDA", 9: "COMM"}
STOP_NONE, STOP_COUNTER, STOP_DIFFERENTIAL = 0, 1, 2

# every symbol declared in include/aicp_b200.h
EXPORTS = ["aicp_b200_create", "aicp_b200_destroy", "aicp_b200_wait_stream", "aicp_b200_last_error", "aicp_b200_version", "aicp_b200_set_config",
           "aicp_b200_set_config_struct", "aicp_b200_get_config", "aicp_b200_parse_icp_yaml", "aicp_b200_register",
           "aicp_b200_set_reference", "aicp_b200_register_to_reference", "aicp_b200_reference_append", "aicp_b200_get_output_reading",
           "aicp_b200_get_initialized_reading", "aicp_b200_get_reference_normals", "aicp_b200_enable_match_trace",
           "aicp_b200_get_trace_matches", "aicp_b200_set_profiling", "aicp_b200_set_knn_schedule", "aicp_b200_set_match_schedule", "aicp_b200_set_loop_schedule", "aicp_b200_surface_normals", "aicp_b200_match", "aicp_b200_trim_threshold",
           "aicp_b200_overlap", "aicp_b200_crop_box", "aicp_b200_get_cropped", "aicp_b200_download_cropped", "aicp_b200_map_append", "aicp_b200_map_size", "aicp_b200_map_crop", "aicp_b200_prefilter_default_config", "aicp_b200_prefilter", "aicp_b200_get_prefiltered",
           "aicp_b200_prefilter_get_sampled", "aicp_b200_prefilter_get_normals", "aicp_b200_prefilter_get_labels", "aicp_b200_voxel_grid",
           "aicp_b200_map_prefilter", "aicp_b200_accumulate_sweep", "aicp_b200_get_accumulated", "aicp_b200_download_accumulated", "aicp_b200_read_pcd", "aicp_b200_read_ply",
           "aicp_b200_write_pcd", "aicp_b200_read_pose_file", "aicp_b200_fov_overlap", "aicp_b200_get_fov_filtered", "aicp_b200_alignability", "aicp_b200_alignment_risk",
           "aicp_b200_svm_parse", "aicp_b200_svm_load", "aicp_b200_svm_info", "aicp_b200_svm_predict", "aicp_b200_autotune_ratio", "aicp_b200_register_batch", "aicp_b200_register_batch_devices", "aicp_b200_aicp_batch", "aicp_b200_pipeline_batch", "aicp_b200_comm_unique_id",
           "aicp_b200_comm_init", "aicp_b200_comm_destroy", "aicp_b200_comm_info"]


class IcpConfig(C.Structure):
    _fields_ = [("knn_normals", C.c_int32), ("reading_normals", C.c_int32), ("ratio", C.c_float),
                ("max_iterations", C.c_int32), ("min_diff_rot", C.c_float), ("min_diff_trans", C.c_float),
                ("smooth_length", C.c_int32), ("matcher_epsilon", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class IterTrace(C.Structure):
    _fields_ = [("T_iter", C.c_float * 16), ("limit_d2", C.c_float), ("n_valid", C.c_int64), ("n_used", C.c_int64),
                ("rot_err", C.c_double), ("trans_err", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("stop_reason", C.c_int32), ("weighted_point_used_ratio", C.c_float),
                ("mean_ref", C.c_float * 3), ("n_ref", C.c_int64), ("n_read", C.c_int64), ("ms_total", C.c_float),
                ("ms_setup", C.c_float), ("ms_iterations", C.c_float), ("gpu_launches", C.c_int32),
                ("profiled", C.c_int32), ("ms_index", C.c_float), ("ms_normals", C.c_float), ("ms_match", C.c_float),
                ("ms_select", C.c_float), ("ms_accumulate", C.c_float), ("ms_tail_pick", C.c_float), ("ms_tail_select", C.c_float),
                ("ms_tail_solve", C.c_float), ("ms_exchange", C.c_float), ("trace", IterTrace * MAX_ITERS)]


class AppendInfo(C.Structure):
    _fields_ = [("n_total", C.c_int64), ("n_recomputed", C.c_int64), ("incremental", C.c_int32), ("ms", C.c_float)]


class PrefilterConfig(C.Structure):
    """aicp_b200_prefilter_config: the PCL parameters hard-coded in filteringUtils.cpp:10-34."""
    _fields_ = [("leaf_size", C.c_float), ("knn_normals", C.c_int32), ("n_neighbours", C.c_int32),
                ("min_cluster_size", C.c_int32), ("max_cluster_size", C.c_int32), ("smoothness_threshold", C.c_float),
                ("curvature_threshold", C.c_float)]


class PrefilterInfo(C.Structure):
    _fields_ = [("n_sampled", C.c_int64), ("n_clusters", C.c_int64), ("n_out", C.c_int64), ("passes", C.c_int32),
                ("gpu_launches", C.c_int32), ("ms_total", C.c_float)]


class SvmSummary(C.Structure):
    _fields_ = [("kernel", C.c_int32), ("dim", C.c_int32), ("sv_total", C.c_int32), ("sv_count", C.c_int32), ("degree", C.c_double),
                ("gamma", C.c_double), ("coef0", C.c_double), ("rho", C.c_double), ("alpha_sum", C.c_double), ("sv_sum", C.c_double),
                ("index_first", C.c_int32), ("index_last", C.c_int32)]


class AicpError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("aicp_b200 error %s: %s" % (ERR_NAMES.get(code, code), message))
        self.code = code
        self.code_name = ERR_NAMES.get(code, str(code))


_lib = None


def lib():
    """The loaded library.  Raises if it has not been built (python -m aicp_mapping_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: build it with `python -m aicp_mapping_b200.build` "
                               "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        fp, ip, i64 = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int64
        L.aicp_b200_create.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        L.aicp_b200_destroy.argtypes = [C.c_void_p]
        L.aicp_b200_wait_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.aicp_b200_last_error.argtypes = [C.c_void_p]
        L.aicp_b200_last_error.restype = C.c_char_p
        L.aicp_b200_version.restype = C.c_char_p
        L.aicp_b200_set_config.argtypes = [C.c_void_p, C.c_char_p]
        L.aicp_b200_set_config_struct.argtypes = [C.c_void_p, C.POINTER(IcpConfig)]
        L.aicp_b200_get_config.argtypes = [C.c_void_p, C.POINTER(IcpConfig)]
        L.aicp_b200_parse_icp_yaml.argtypes = [C.c_char_p, C.POINTER(IcpConfig), C.c_char_p, C.c_int]
        L.aicp_b200_register.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, fp, C.POINTER(Stats)]
        L.aicp_b200_set_reference.argtypes = [C.c_void_p, C.c_void_p, i64]
        L.aicp_b200_register_to_reference.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_void_p, fp, C.POINTER(Stats)]
        L.aicp_b200_reference_append.argtypes = [C.c_void_p, C.c_void_p, i64, C.POINTER(AppendInfo)]
        L.aicp_b200_get_output_reading.argtypes = [C.c_void_p, C.c_void_p, i64]
        L.aicp_b200_get_initialized_reading.argtypes = [C.c_void_p, C.c_void_p, i64]
        L.aicp_b200_get_reference_normals.argtypes = [C.c_void_p, C.c_void_p, i64]
        L.aicp_b200_enable_match_trace.argtypes = [C.c_void_p, C.c_int]
        L.aicp_b200_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.aicp_b200_set_knn_schedule.argtypes = [C.c_void_p, C.c_int]
        L.aicp_b200_set_match_schedule.argtypes = [C.c_void_p, C.c_int]
        L.aicp_b200_set_loop_schedule.argtypes = [C.c_void_p, C.c_int]
        L.aicp_b200_get_trace_matches.argtypes = [C.c_void_p, C.c_void_p, i64, i64]
        L.aicp_b200_surface_normals.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_int32, C.c_void_p, C.c_void_p]
        L.aicp_b200_match.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, C.c_void_p]
        L.aicp_b200_trim_threshold.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_float, fp, C.POINTER(i64)]
        L.aicp_b200_overlap.argtypes = [C.c_void_p, C.c_void_p, i64, C.POINTER(C.c_double), C.c_void_p, i64,
                                        C.POINTER(C.c_double), C.c_double, fp, C.POINTER(i64)]
        L.aicp_b200_crop_box.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_float, C.c_float, fp, fp, C.c_void_p, C.POINTER(i64)]
        L.aicp_b200_get_cropped.argtypes = [C.c_void_p, C.POINTER(i64)]
        L.aicp_b200_get_cropped.restype = C.c_void_p
        L.aicp_b200_download_cropped.argtypes = [C.c_void_p, C.c_void_p, i64]
        L.aicp_b200_map_append.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_int]
        L.aicp_b200_map_size.argtypes = [C.c_void_p]
        L.aicp_b200_map_size.restype = i64
        L.aicp_b200_map_crop.argtypes = [C.c_void_p, C.c_float, C.c_float, fp, fp, C.POINTER(i64)]
        L.aicp_b200_prefilter_default_config.argtypes = [C.POINTER(PrefilterConfig)]
        L.aicp_b200_prefilter.argtypes = [C.c_void_p, C.c_void_p, i64, C.POINTER(PrefilterConfig), fp, C.c_void_p, C.POINTER(i64),
                                          C.POINTER(PrefilterInfo)]
        L.aicp_b200_get_prefiltered.argtypes = [C.c_void_p, C.POINTER(i64)]
        L.aicp_b200_get_prefiltered.restype = C.c_void_p
        L.aicp_b200_prefilter_get_sampled.argtypes = [C.c_void_p, C.c_void_p, i64]
        L.aicp_b200_prefilter_get_normals.argtypes = [C.c_void_p, C.c_void_p, i64]
        L.aicp_b200_prefilter_get_labels.argtypes = [C.c_void_p, C.c_void_p, i64]
        L.aicp_b200_voxel_grid.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_float, C.c_void_p, C.POINTER(i64)]
        L.aicp_b200_map_prefilter.argtypes = [C.c_void_p, C.POINTER(PrefilterConfig), C.POINTER(i64), C.POINTER(PrefilterInfo)]
        dp = C.POINTER(C.c_double)
        L.aicp_b200_accumulate_sweep.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_float, dp, C.c_int, C.POINTER(i64)]
        L.aicp_b200_get_accumulated.argtypes = [C.c_void_p, C.POINTER(i64)]
        L.aicp_b200_get_accumulated.restype = C.c_void_p
        L.aicp_b200_download_accumulated.argtypes = [C.c_void_p, C.c_void_p, i64]
        L.aicp_b200_read_pcd.argtypes = [C.c_char_p, C.c_void_p, i64, C.POINTER(i64), C.c_char_p, C.c_int]
        L.aicp_b200_read_ply.argtypes = [C.c_char_p, C.c_void_p, i64, C.POINTER(i64), C.c_char_p, C.c_int]
        L.aicp_b200_write_pcd.argtypes = [C.c_char_p, C.c_void_p, i64, C.c_char_p, C.c_int]
        L.aicp_b200_read_pose_file.argtypes = [C.c_char_p, C.POINTER(i64), dp, i64, C.POINTER(i64), C.c_char_p, C.c_int]
        L.aicp_b200_fov_overlap.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_void_p, i64, dp, dp, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                            C.POINTER(i64), fp]
        L.aicp_b200_get_fov_filtered.argtypes = [C.c_void_p, C.c_int, C.POINTER(i64)]
        L.aicp_b200_get_fov_filtered.restype = C.c_void_p
        L.aicp_b200_alignability.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_void_p, i64, dp, dp, C.POINTER(PrefilterConfig), fp, ip, i64,
                                             C.POINTER(i64)]
        L.aicp_b200_alignment_risk.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_void_p, i64, dp, dp, C.c_float, C.c_float, C.c_float, fp, fp, dp]
        L.aicp_b200_svm_parse.argtypes = [C.c_char_p, C.POINTER(SvmSummary), C.c_char_p, C.c_int]
        L.aicp_b200_svm_load.argtypes = [C.c_void_p, C.c_char_p]
        L.aicp_b200_svm_info.argtypes = [C.c_void_p, ip, ip]
        L.aicp_b200_svm_predict.argtypes = [C.c_void_p, C.POINTER(C.c_double), i64, C.c_int32, C.POINTER(C.c_double), fp]
        L.aicp_b200_autotune_ratio.argtypes = [C.c_float]
        L.aicp_b200_autotune_ratio.restype = C.c_float
        L.aicp_b200_register_batch.argtypes = [C.c_void_p, i64, C.POINTER(C.c_void_p), C.POINTER(i64), C.POINTER(C.c_void_p),
                                               C.POINTER(i64), fp, C.c_int, fp, C.POINTER(Stats), C.POINTER(C.c_int32), fp]
        L.aicp_b200_register_batch_devices.argtypes = [C.c_void_p, ip, C.c_int32, i64, C.POINTER(C.c_void_p), C.POINTER(i64), C.POINTER(C.c_void_p),
                                                       C.POINTER(i64), fp, C.c_int, fp, C.POINTER(Stats), C.POINTER(C.c_int32), fp]
        L.aicp_b200_aicp_batch.argtypes = [C.c_void_p, i64, C.POINTER(C.c_void_p), C.POINTER(i64), C.POINTER(C.c_double),
                                           C.POINTER(C.c_void_p), C.POINTER(i64), C.POINTER(C.c_double), C.c_double, C.c_int, fp, fp,
                                           C.POINTER(Stats), C.POINTER(C.c_int32), fp]
        L.aicp_b200_pipeline_batch.argtypes = [C.c_void_p, i64, C.POINTER(C.c_void_p), C.POINTER(i64), C.POINTER(C.c_double),
                                               C.POINTER(C.c_void_p), C.POINTER(i64), C.POINTER(C.c_double), C.c_double, C.c_float, C.c_float,
                                               C.c_char_p, C.c_double, C.c_int, C.c_int, fp, fp, fp, C.POINTER(C.c_double), C.POINTER(i64),
                                               C.POINTER(Stats), C.POINTER(C.c_int32), fp]
        L.aicp_b200_comm_unique_id.argtypes = [C.c_void_p]
        L.aicp_b200_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.aicp_b200_comm_destroy.argtypes = [C.c_void_p]
        L.aicp_b200_comm_info.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        _lib = L
    return _lib


def to_xyzw(xyz):
    """n x 3 or n x 4 array-like -> contiguous n x 4 float32, pad = 1 (the pcl::PointXYZ record, cloudIO.cpp:81-98)."""
    a = np.asarray(xyz, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] not in (3, 4):
        raise ValueError("expected an n x 3 or n x 4 array")
    if a.shape[1] == 4:
        return np.ascontiguousarray(a)
    out = np.ones((a.shape[0], 4), dtype=np.float32)
    out[:, :3] = a
    return out


def ptr_and_count(cloud, synced=None):
    """Accepts numpy arrays (host) or torch CUDA tensors of shape n x 4 float32 (device, used in place).  `synced`: a set shared
    by the clouds of ONE call (a batch); torch's stream of a device is synchronised once per call, not once per cloud."""
    if hasattr(cloud, "data_ptr"):      # torch tensor
        if cloud.dim() != 2 or cloud.shape[1] != 4 or str(cloud.dtype) != "torch.float32" or not cloud.is_contiguous():
            raise ValueError("device clouds must be contiguous n x 4 float32 tensors")
        if getattr(cloud, "is_cuda", False):          # duck-typed views of library-owned buffers carry no stream of their own
            # the library reads device inputs on its own streams (aicp_b200.h, "STREAM ORDER"): whatever torch still has in
            # flight on the current stream of the tensor's device -- the op or copy that produces it -- must finish first
            import torch
            key = (cloud.device.index, torch.cuda.current_stream(cloud.device).cuda_stream)
            if synced is None or key not in synced:
                torch.cuda.current_stream(cloud.device).synchronize()
                if synced is not None:
                    synced.add(key)
        return C.c_void_p(cloud.data_ptr()), int(cloud.shape[0]), cloud
    a = to_xyzw(cloud)
    return C.c_void_p(a.ctypes.data), int(a.shape[0]), a


def mat_to_colmajor(T):
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).reshape(4, 4).T).ravel()


def colmajor_to_mat(t16):
    return np.asarray(t16, dtype=np.float32).reshape(4, 4).T.copy()


def colmajor_batch_to_mats(T):
    """n x 16 column-major transforms -> n x 4 x 4 row-major matrices (one vectorised transpose, not n small ones)."""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).reshape(-1, 4, 4).transpose(0, 2, 1))
