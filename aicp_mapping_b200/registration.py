"""Host-side mirror of the reference's registration plug-in interface, over the C ABI.

Mirrors (same names, argument meaning and error behaviour where Python allows):
  aicp::AbstractRegistrator                 aicp_core/include/aicp_registration/abstract_registrator.hpp:8-19
  aicp::PointmatcherRegistration            aicp_core/include/aicp_registration/pointmatcher_registration.hpp:22-67
  aicp::create_registrator                  aicp_core/include/aicp_registration/registration.hpp:9-19
  RegistrationParams                        aicp_core/include/aicp_registration/common.hpp:7-23
  replaceRatioConfigFile                    aicp_core/src/utils/fileIO.cpp:179-214
  App::computeRegistration                  aicp_core/src/registration/app.cpp:187-216
  parseTransformationDeg                    aicp_core/src/utils/cloudIO.cpp:261-302

The C++ adapter a maintainer would compile into aicp_core is include/aicp_b200_adapter.hpp; this module is the same
thing for the Python test and benchmark harness.  All arithmetic happens in libaicp_b200.so on the GPU.
"""
import ctypes as C
import math
import sys
from dataclasses import dataclass, field

import numpy as np

from . import capi


@dataclass
class PointmatcherRegistrationParams:
    configFileName: str = ""
    initialTransform: str = ""      # "x,y,theta" (metres, metres, degrees)
    printOutputStatistics: bool = False


@dataclass
class RegistrationParams:
    type: str = ""
    sensorRange: float = -1.0
    sensorAngularView: float = -1.0
    loadPosesFrom: str = ""
    initialTransform: str = ""
    pointmatcher: PointmatcherRegistrationParams = field(default_factory=PointmatcherRegistrationParams)


def parseTransformationDeg(transform, cloudDimension=3):
    """cloudIO.cpp:261-302: "[x,y,theta_deg]" -> 4x4 float32 (identity, with a message, when unparsable)."""
    T = np.eye(cloudDimension + 1, dtype=np.float32)
    s = transform.replace("[", "").replace("]", "").replace(",", " ").replace(";", " ").split()
    try:
        v = [float(np.float32(float(x))) for x in s[:3]]
        if len(v) < 3:
            raise ValueError
    except ValueError:
        sys.stderr.write("[Cloud IO] An error occured while trying to parse the initial transformation.\n"
                         "No initial transformation will be used\n")
        return T
    th = v[2] * math.pi / 180.0
    T[0, 0] = math.cos(th); T[0, 1] = -math.sin(th)
    T[1, 0] = math.sin(th); T[1, 1] = math.cos(th)
    T[0, cloudDimension] = v[0]
    T[1, cloudDimension] = v[1]
    return T


def replaceRatioConfigFile(in_file, out_file, ratio):
    """fileIO.cpp:179-214, byte for byte: on every line holding "ratio: " the 11 characters from its start are replaced
    by "ratio: " + (ostream << float), and every line (plus one trailing empty line at EOF) is written with '\\n'."""
    try:
        with open(in_file, "r") as f:
            text = f.read()
    except OSError:
        sys.stderr.write("[File IO] Could not open config file for params update.\n")
        text = ""
    lines = text.split("\n")          # getline until eof: a final '\n' yields one extra empty line, as in the reference
    word = "ratio: "
    repl = word + ("%g" % float(np.float32(ratio)))
    out = []
    for line in lines:
        pos = line.find(word)
        if pos != -1:
            line = line[:pos] + repl + line[pos + len(word) + 4:]
        out.append(line + "\n")
    with open(out_file, "w") as f:
        f.write("".join(out))


class B200Registration:
    """AbstractRegistrator implemented by libaicp_b200.so.  One instance == one aicp_b200_handle (one GPU stream)."""

    def __init__(self, params=None, device=-1):
        self.params_ = params or RegistrationParams(type="B200")
        self._lib = capi.lib()
        self._h = C.c_void_p()
        path = self.params_.pointmatcher.configFileName
        rc = self._lib.aicp_b200_create(path.encode() if path else None, int(device), C.byref(self._h))
        if rc:
            raise capi.AicpError(rc, self._lib.aicp_b200_last_error(None).decode())
        self.stats = capi.Stats()
        self._last_T = None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.aicp_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise capi.AicpError(rc, self._lib.aicp_b200_last_error(self._h).decode())

    # ---- AbstractRegistrator ------------------------------------------------------------------------------------
    def updateConfigParams(self, config_name):
        """pointmatcher_registration.hpp:52-54.  The file is re-read at the start of every registerClouds."""
        self.params_.pointmatcher.configFileName = config_name
        self._check(self._lib.aicp_b200_set_config(self._h, config_name.encode() if config_name else None))

    def registerClouds(self, cloud_ref, cloud_read):
        """pointmatcher_registration.cpp:14-23,92-151.  Clouds: n x 3 / n x 4 float32 numpy arrays or n x 4 CUDA
        tensors.  Returns final_transform (4x4 float32) -- the reference writes it through an out-parameter."""
        pr, nr, keep_r = capi.ptr_and_count(cloud_ref)
        pq, nq, keep_q = capi.ptr_and_count(cloud_read)
        init = None
        if self.params_.pointmatcher.initialTransform:
            init = self.applyInitialization()
        return self._register(pr, nr, pq, nq, init)

    def _register(self, pr, nr, pq, nq, init):
        T = np.zeros(16, dtype=np.float32)
        init_c = capi.mat_to_colmajor(init) if init is not None else None
        rc = self._lib.aicp_b200_register(self._h, pr, nr, pq, nq,
                                          C.c_void_p(init_c.ctypes.data) if init_c is not None else None,
                                          T.ctypes.data_as(C.POINTER(C.c_float)), C.byref(self.stats))
        self._check(rc)
        self._last_T = capi.colmajor_to_mat(T)
        return self._last_T

    def applyInitialization(self):
        """pointmatcher_registration.cpp:71-89: parse "x,y,theta", fall back to identity when it is not rigid."""
        T = parseTransformationDeg(self.params_.pointmatcher.initialTransform, 3)
        if abs(1.0 - float(np.linalg.det(T[:3, :3].astype(np.float64)))) > 1e-3:      # RigidTransformation::checkParameters
            sys.stderr.write("\n[Pointmatcher] Initial transformation is not rigid, identity will be used.\n")
            T = np.eye(4, dtype=np.float32)
        return T

    def getOutputReading(self):
        """pointmatcher_registration.hpp:48-50: T * reading, n x 4 float32."""
        n = int(self.stats.n_read)
        out = np.zeros((n, 4), dtype=np.float32)
        self._check(self._lib.aicp_b200_get_output_reading(self._h, C.c_void_p(out.ctypes.data), n))
        return out

    def getInitializedReading(self):
        """pointmatcher_registration.hpp:37-46."""
        n = int(self.stats.n_read)
        out = np.zeros((n, 4), dtype=np.float32)
        self._check(self._lib.aicp_b200_get_initialized_reading(self._h, C.c_void_p(out.ctypes.data), n))
        return out

    # ---- conveniences beyond the reference interface ---------------------------------------------------------------
    def getOutputTransform(self):
        """Named in BASELINE.json; the reference returns the transform only through registerClouds' out-parameter."""
        return self._last_T

    def getWeightedPointUsedRatio(self):
        """icp_.errorMinimizer->getWeightedPointUsedRatio(), printed at pointmatcher_registration.cpp:114."""
        return float(self.stats.weighted_point_used_ratio)

    def setConfig(self, **kw):
        cfg = self.getConfig()
        for k, v in kw.items():
            setattr(cfg, k, v)
        self._check(self._lib.aicp_b200_set_config_struct(self._h, C.byref(cfg)))

    def getConfig(self):
        cfg = capi.IcpConfig()
        self._check(self._lib.aicp_b200_get_config(self._h, C.byref(cfg)))
        return cfg

    def setReference(self, cloud_ref):
        p, n, keep = capi.ptr_and_count(cloud_ref)
        self._check(self._lib.aicp_b200_set_reference(self._h, p, n))

    def appendToReference(self, cloud):
        """aicp_b200_reference_append: merge `cloud` into the reference in place (incremental index + normals update).
        Returns capi.AppendInfo (n_total, n_recomputed, incremental, ms)."""
        p, n, keep = capi.ptr_and_count(cloud)
        info = capi.AppendInfo()
        self._check(self._lib.aicp_b200_reference_append(self._h, p, n, C.byref(info)))
        return info

    def registerToReference(self, cloud_read, init_T=None):
        p, n, keep = capi.ptr_and_count(cloud_read)
        T = np.zeros(16, dtype=np.float32)
        init_c = capi.mat_to_colmajor(init_T) if init_T is not None else None
        rc = self._lib.aicp_b200_register_to_reference(self._h, p, n,
                                                       C.c_void_p(init_c.ctypes.data) if init_c is not None else None,
                                                       T.ctypes.data_as(C.POINTER(C.c_float)), C.byref(self.stats))
        self._check(rc)
        self._last_T = capi.colmajor_to_mat(T)
        return self._last_T

    def registerCloudsInit(self, cloud_ref, cloud_read, init_T):
        """registerClouds with an explicit 4x4 initial guess instead of the "x,y,theta" string."""
        pr, nr, keep_r = capi.ptr_and_count(cloud_ref)
        pq, nq, keep_q = capi.ptr_and_count(cloud_read)
        return self._register(pr, nr, pq, nq, init_T)

    def registerBatch(self, pairs, ratios=None, streams=0, devices=None):
        """aicp_b200_register_batch: `pairs` is a list of (cloud_ref, cloud_read); independent pairs are registered
        concurrently on `streams` CUDA streams.  devices = [d0, d1, ...]: aicp_b200_register_batch_devices, pair i on GPU
        devices[i % len(devices)] with `streams` streams each.  Returns (T [n,4,4], stats list, status array, batch_ms)."""
        n = len(pairs)
        keep, refs, reads = [], (C.c_void_p * n)(), (C.c_void_p * n)()
        n_ref, n_read = (C.c_int64 * n)(), (C.c_int64 * n)()
        synced = set()
        for i, (r, q) in enumerate(pairs):
            pr, nr, kr = capi.ptr_and_count(r, synced)
            pq, nq, kq = capi.ptr_and_count(q, synced)
            refs[i], reads[i], n_ref[i], n_read[i] = pr, pq, nr, nq
            keep.append((kr, kq))
        T = np.zeros((n, 16), dtype=np.float32)
        stats = (capi.Stats * n)()
        status = np.zeros(n, dtype=np.int32)
        ms = C.c_float()
        rat = np.ascontiguousarray(ratios, dtype=np.float32) if ratios is not None else None
        if devices is not None:
            dv = np.ascontiguousarray(devices, dtype=np.int32)
            rc = self._lib.aicp_b200_register_batch_devices(self._h, dv.ctypes.data_as(C.POINTER(C.c_int32)), len(dv), n, refs, n_ref, reads, n_read,
                                                            rat.ctypes.data_as(C.POINTER(C.c_float)) if rat is not None else None,
                                                            int(streams), T.ctypes.data_as(C.POINTER(C.c_float)), stats,
                                                            status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(ms))
        else:
            rc = self._lib.aicp_b200_register_batch(self._h, n, refs, n_ref, reads, n_read,
                                                    rat.ctypes.data_as(C.POINTER(C.c_float)) if rat is not None else None,
                                                    int(streams), T.ctypes.data_as(C.POINTER(C.c_float)), stats,
                                                    status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(ms))
        self._check(rc)
        return capi.colmajor_batch_to_mats(T), stats, status, float(ms.value)

    def aicpBatch(self, pairs, origins, resolution=float(np.float32(0.2)), streams=0):
        """aicp_b200_aicp_batch: one AICP step (overlap -> auto-tuned ratio -> registration, app.cpp:218-247) per pair.
        `pairs`: list of (cloud_ref, cloud_read); `origins`: list of (ref_origin[3], read_origin[3]).
        Returns (T [n,4,4], overlap [n] percent, stats list, status array, batch_ms)."""
        n = len(pairs)
        keep, refs, reads = [], (C.c_void_p * n)(), (C.c_void_p * n)()
        n_ref, n_read = (C.c_int64 * n)(), (C.c_int64 * n)()
        synced = set()
        for i, (r, q) in enumerate(pairs):
            pr, nr, kr = capi.ptr_and_count(r, synced)
            pq, nq, kq = capi.ptr_and_count(q, synced)
            refs[i], reads[i], n_ref[i], n_read[i] = pr, pq, nr, nq
            keep.append((kr, kq))
        ro = np.ascontiguousarray([o[0] for o in origins], dtype=np.float64).reshape(n, 3)
        so = np.ascontiguousarray([o[1] for o in origins], dtype=np.float64).reshape(n, 3)
        T = np.zeros((n, 16), dtype=np.float32)
        ov = np.zeros(n, dtype=np.float32)
        stats = (capi.Stats * n)()
        status = np.zeros(n, dtype=np.int32)
        ms = C.c_float()
        fp = C.POINTER(C.c_float)
        rc = self._lib.aicp_b200_aicp_batch(self._h, n, refs, n_ref, ro.ctypes.data_as(C.POINTER(C.c_double)), reads, n_read,
                                            so.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(resolution), int(streams),
                                            T.ctypes.data_as(fp), ov.ctypes.data_as(fp), stats,
                                            status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(ms))
        self._check(rc)
        return (capi.colmajor_batch_to_mats(T)), ov, stats, status, float(ms.value)

    def pipelineBatch(self, pairs, poses, svm_model, sensor_range, angular_view, risk_threshold=0.5,
                      resolution=float(np.float32(0.2)), streams=0, prefilter_first=False):
        """aicp_b200_pipeline_batch: App::runAicpPipeline with failure_prediction_mode (app.cpp:218-247) per pair: overlap ->
        alignment risk -> registration when risk <= threshold.  `poses`: list of (ref_pose 4x4, read_pose 4x4).
        prefilter_first: the clouds are raw; each is pre-filtered on the device first (app.cpp:77-110); self.n_filtered then holds
        the filtered sizes [n, 2].
        Returns (T [n,4,4], overlap [n], alignability [n], risk [n], stats list, status array, batch_ms)."""
        n = len(pairs)
        keep, refs, reads = [], (C.c_void_p * n)(), (C.c_void_p * n)()
        n_ref, n_read = (C.c_int64 * n)(), (C.c_int64 * n)()
        synced = set()
        for i, (r, q) in enumerate(pairs):
            pr, nr, kr = capi.ptr_and_count(r, synced)
            pq, nq, kq = capi.ptr_and_count(q, synced)
            refs[i], reads[i], n_ref[i], n_read[i] = pr, pq, nr, nq
            keep.append((kr, kq))
        pa = np.ascontiguousarray([np.asarray(p[0], dtype=np.float64).T.ravel() for p in poses], dtype=np.float64).reshape(n, 16)
        pb = np.ascontiguousarray([np.asarray(p[1], dtype=np.float64).T.ravel() for p in poses], dtype=np.float64).reshape(n, 16)
        T = np.zeros((n, 16), dtype=np.float32)
        ov, al = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
        risk = np.zeros(n, dtype=np.float64)
        self.n_filtered = np.zeros((n, 2), dtype=np.int64)
        stats = (capi.Stats * n)()
        status = np.zeros(n, dtype=np.int32)
        ms = C.c_float()
        fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
        rc = self._lib.aicp_b200_pipeline_batch(self._h, n, refs, n_ref, pa.ctypes.data_as(dp), reads, n_read, pb.ctypes.data_as(dp),
                                                C.c_double(resolution), C.c_float(sensor_range), C.c_float(angular_view),
                                                str(svm_model).encode(), C.c_double(risk_threshold), int(bool(prefilter_first)), int(streams),
                                                T.ctypes.data_as(fp), ov.ctypes.data_as(fp), al.ctypes.data_as(fp), risk.ctypes.data_as(dp),
                                                self.n_filtered.ctypes.data_as(C.POINTER(C.c_int64)), stats,
                                                status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(ms))
        self._check(rc)
        return ((capi.colmajor_batch_to_mats(T)), ov, al, risk, stats, status,
                float(ms.value))

    # ---- multi-GPU single registration (reading sharded over ranks) --------------------------------------------------
    def commInit(self, unique_id, rank, n_ranks):
        """aicp_b200_comm_init.  unique_id: the 128 bytes from comm_unique_id() on rank 0, broadcast by the caller."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._check(self._lib.aicp_b200_comm_init(self._h, buf, int(rank), int(n_ranks)))

    def commInfo(self):
        """aicp_b200_comm_info: what the ranks exchange per ICP iteration."""
        buf = C.create_string_buffer(512)
        self._check(self._lib.aicp_b200_comm_info(self._h, buf, 512))
        return buf.value.decode()

    def waitStream(self, cuda_stream=None):
        """aicp_b200_wait_stream: order the handle's work after what is enqueued on `cuda_stream` (a torch.cuda.Stream or a
        raw cudaStream_t; None = torch's current stream)."""
        if cuda_stream is None:
            import torch
            cuda_stream = torch.cuda.current_stream()
        ptr = getattr(cuda_stream, "cuda_stream", cuda_stream)
        self._check(self._lib.aicp_b200_wait_stream(self._h, C.c_void_p(int(ptr))))

    def commDestroy(self):
        self._check(self._lib.aicp_b200_comm_destroy(self._h))

    def getReferenceNormals(self):
        n = int(self.stats.n_ref)
        out = np.zeros((n, 4), dtype=np.float32)
        self._check(self._lib.aicp_b200_get_reference_normals(self._h, C.c_void_p(out.ctypes.data), n))
        return out

    def enableMatchTrace(self, enable=True):
        self._check(self._lib.aicp_b200_enable_match_trace(self._h, int(enable)))

    def setProfiling(self, level=2):
        """0 off, 1 CUDA events around k_match only, 2 around every stage (stats.ms_*)."""
        self._check(self._lib.aicp_b200_set_profiling(self._h, int(level)))

    def setKnnSchedule(self, schedule=0):
        """0 automatic, 1 warp-per-query k-NN kernel, 2 tile kernel (identical results)."""
        self._check(self._lib.aicp_b200_set_knn_schedule(self._h, int(schedule)))

    def setMatchSchedule(self, schedule=0):
        """0 automatic, 1 per-thread correspondence search, 2 tile search (identical results)."""
        self._check(self._lib.aicp_b200_set_match_schedule(self._h, int(schedule)))

    def setLoopSchedule(self, schedule=0):
        """0 automatic (persistent kernel for single and sharded registrations, multi-launch inside batches), 1 three launches
        per iteration with the host staying two iterations ahead, 2 one persistent cooperative kernel for the whole loop
        (identical results)."""
        self._check(self._lib.aicp_b200_set_loop_schedule(self._h, int(schedule)))

    def getTraceMatches(self):
        it, n = int(self.stats.iterations), int(self.stats.n_read)
        out = np.zeros((it, n), dtype=np.int32)
        self._check(self._lib.aicp_b200_get_trace_matches(self._h, C.c_void_p(out.ctypes.data), it, n))
        return out

    def trace(self):
        out = []
        for i in range(int(self.stats.iterations)):
            t = self.stats.trace[i]
            out.append(dict(T_iter=capi.colmajor_to_mat(list(t.T_iter)), limit_d2=np.float32(t.limit_d2), n_valid=t.n_valid,
                            n_used=t.n_used, rot_err=t.rot_err, trans_err=t.trans_err))
        return out

    # ---- stage entry points (parity tests) -----------------------------------------------------------------------
    def surfaceNormals(self, cloud, knn=20, want_knn=True):
        p, n, keep = capi.ptr_and_count(cloud)
        normals = np.zeros((n, 4), dtype=np.float32)
        ids = np.zeros((n, knn), dtype=np.int32) if want_knn else None
        self._check(self._lib.aicp_b200_surface_normals(self._h, p, n, int(knn), C.c_void_p(normals.ctypes.data),
                                                        C.c_void_p(ids.ctypes.data) if ids is not None else None))
        return normals, ids

    def match(self, cloud_ref, cloud_qry):
        pr, nr, k1 = capi.ptr_and_count(cloud_ref)
        pq, nq, k2 = capi.ptr_and_count(cloud_qry)
        idx = np.zeros(nq, dtype=np.int32)
        d2 = np.zeros(nq, dtype=np.float32)
        self._check(self._lib.aicp_b200_match(self._h, pr, nr, pq, nq, C.c_void_p(idx.ctypes.data), C.c_void_p(d2.ctypes.data)))
        return idx, d2

    def trimThreshold(self, d2, ratio):
        d2 = np.ascontiguousarray(d2, dtype=np.float32)
        limit, nv = C.c_float(), C.c_int64()
        self._check(self._lib.aicp_b200_trim_threshold(self._h, C.c_void_p(d2.ctypes.data), d2.shape[0], C.c_float(ratio),
                                                       C.byref(limit), C.byref(nv)))
        return np.float32(limit.value), nv.value


def comm_unique_id():
    """aicp_b200_comm_unique_id: 128-byte ncclUniqueId, to be created on rank 0 and broadcast."""
    buf = (C.c_uint8 * 128)()
    lib = capi.lib()
    rc = lib.aicp_b200_comm_unique_id(buf)
    if rc:
        raise capi.AicpError(rc, lib.aicp_b200_last_error(None).decode())
    return bytes(buf)


def create_registrator(parameters, device=-1):
    """registration.hpp:9-19 with the extra "B200" branch a maintainer adds."""
    if parameters.type == "B200":
        return B200Registration(parameters, device=device)
    sys.stderr.write("Invalid registration type %s.\n" % parameters.type)
    return None


def autotune_ratio(octree_overlap):
    """app.cpp:198-202 clamp followed by the text round trip of fileIO.cpp:194-198 (done in the library)."""
    return float(capi.lib().aicp_b200_autotune_ratio(C.c_float(octree_overlap)))


def computeRegistration(registr, reference, reading, octree_overlap, default_config_file, registration_config_file):
    """App::computeRegistration, app.cpp:187-216: clamp the overlap into the trimmed ratio, rewrite the ICP chain file,
    point the registrator at it, register."""
    current_ratio = np.float32(octree_overlap / 100.0)
    if current_ratio < 0.25:
        current_ratio = np.float32(0.25)
    elif current_ratio > 0.70:
        current_ratio = np.float32(0.70)
    replaceRatioConfigFile(default_config_file, registration_config_file, current_ratio)
    registr.updateConfigParams(registration_config_file)
    return registr.registerClouds(reference, reading)
