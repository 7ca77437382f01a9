/*
 * aicp_b200.h -- C ABI of libaicp_b200.so: the B200-native (sm_100a) replacement for AICP's registration hot path.
 *
 * Every entry point states the reference interface it replaces (paths relative to zbqq/aicp_mapping).
 * The reference has no FFI of its own for this path (it is C++ calling libpointmatcher / octomap in-process), so the
 * binding a maintainer adds is the C++ adapter in include/aicp_b200_adapter.hpp (B200Registration : AbstractRegistrator,
 * B200Overlap : AbstractOverlapper) -- see INTEGRATION.md.
 *
 * Conventions
 *   - points: `n` records of 4 floats (x, y, z, pad), 16-byte stride == sizeof(pcl::PointXYZ); pad is ignored on input
 *     (aicp_core/src/utils/cloudIO.cpp:81-98 writes pad = 1 into the DataPoints feature matrix).
 *   - transforms: 16 floats, column-major 4x4 == Eigen::Matrix4f::data().
 *   - every pointer argument may be a host pointer or a device pointer of the handle's device; the library detects
 *     which (cudaPointerGetAttributes).  Host buffers are copied through the handle's stream.
 *   - STREAM ORDER of device-pointer inputs: the library reads them on its own non-blocking stream(s), which are not
 *     ordered against the caller's streams.  The data must be complete when the call is made (synchronise the producing
 *     stream), or the caller declares the producer once with aicp_b200_wait_stream(h, stream) before the call.  Results
 *     are complete when a call returns (every compute entry point synchronises its stream before returning).
 *   - all functions return AICP_B200_OK (0) or an error code; aicp_b200_last_error() gives the text.  Nothing here
 *     calls exit() (the reference does: pointmatcher_registration.cpp:60-64,96-100).
 *   - a handle is used by one thread at a time (same contract as the reference: app.cpp:528-550).
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with AICP_B200_ERR_CUDA.
 */
#ifndef AICP_B200_H_
#define AICP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AICP_B200_MAX_ITERS 256

enum {
  AICP_B200_OK = 0,
  AICP_B200_ERR_BAD_ARG = 1,
  AICP_B200_ERR_KNN_TOO_LARGE = 2,    /* SurfaceNormalDataPointsFilter needs knn < number of points */
  AICP_B200_ERR_NO_VALID_MATCH = 3,   /* libpointmatcher ConvergenceError: no finite positive distance to trim */
  AICP_B200_ERR_NAN = 4,              /* libpointmatcher ConvergenceError: NaN in the transformation */
  AICP_B200_ERR_NONFINITE_INPUT = 5,
  AICP_B200_ERR_EXTENT = 6,           /* |coordinate| > 1024 m in the reference-centred frame */
  AICP_B200_ERR_CONFIG = 7,           /* unreadable / unsupported libpointmatcher YAML */
  AICP_B200_ERR_CUDA = 8,
  AICP_B200_ERR_COMM = 9
};

enum { AICP_B200_STOP_NONE = 0, AICP_B200_STOP_COUNTER = 1, AICP_B200_STOP_DIFFERENTIAL = 2 };

typedef struct aicp_b200_handle aicp_b200_handle;

/* The libpointmatcher chain parameters of aicp_core/config/icp/icp_autotuned.yaml:9-58 that the path uses. */
typedef struct {
  int32_t knn_normals;        /* referenceDataPointsFilters / SurfaceNormalDataPointsFilter.knn        (20) */
  int32_t reading_normals;    /* 1: also run the reading SurfaceNormal filter (results unused by PointToPlane) (0) */
  float   ratio;              /* outlierFilters / TrimmedDistOutlierFilter.ratio                       (0.70) */
  int32_t max_iterations;     /* CounterTransformationChecker.maxIterationCount                        (20) */
  float   min_diff_rot;       /* DifferentialTransformationChecker.minDiffRotErr                       (0.001) */
  float   min_diff_trans;     /* DifferentialTransformationChecker.minDiffTransErr                     (0.01) */
  int32_t smooth_length;      /* DifferentialTransformationChecker.smoothLength                        (4) */
  float   matcher_epsilon;    /* KDTreeMatcher.epsilon as written in the file; the search is always exact (epsilon 0) */
} aicp_b200_icp_config;

typedef struct {
  float   T_iter[16];         /* accumulated iteration transform in the reference-centred frame */
  float   limit_d2;           /* trimmed squared-distance threshold of this iteration */
  int64_t n_valid;            /* matches with finite, positive distance */
  int64_t n_used;             /* matches with weight 1 */
  double  rot_err, trans_err; /* DifferentialTransformationChecker means (NaN while the history is too short) */
} aicp_b200_iter_trace;

typedef struct {
  int32_t iterations;
  int32_t stop_reason;
  float   weighted_point_used_ratio;  /* icp_.errorMinimizer->getWeightedPointUsedRatio(), pointmatcher_registration.cpp:114 */
  float   mean_ref[3];
  int64_t n_ref, n_read;
  float   ms_total;                   /* device time of the whole call (CUDA events, includes host<->device copies) */
  float   ms_setup;                   /* index + normals + centring */
  float   ms_iterations;              /* ICP loop */
  int32_t gpu_launches;               /* kernels launched by this call */
  /* filled only when aicp_b200_set_profiling(h, level > 0): CUDA-event time of each stage on the handle's stream, summed
   * over the iterations that actually ran (level 1: ms_match only) */
  int32_t profiled;
  float   ms_index, ms_normals;       /* setup: Morton index build, SurfaceNormal filter */
  float   ms_match, ms_select, ms_accumulate;   /* loop: search, trimmed quantile, normal equations + solve.  Persistent loop kernel
                                                 * (the default): always filled, from the kernel's own phase clocks */
  float   ms_tail_pick, ms_tail_select, ms_tail_solve;   /* of which: single-block tails (device globaltimer) */
  float   ms_exchange;                /* sharded registration: time the loop kernel spent waiting for its peers' flags */
  aicp_b200_iter_trace trace[AICP_B200_MAX_ITERS];
} aicp_b200_stats;

/* The pre-filter's PCL parameters as regionGrowingUniformPlaneSegmentationFilter hard-codes them
 * (aicp_core/src/utils/filteringUtils.cpp:10-34). */
typedef struct {
  float   leaf_size;              /* pcl::VoxelGrid::setLeafSize                      (0.08f, :12) */
  int32_t knn_normals;            /* pcl::NormalEstimation::setKSearch                (30, :22); 3..32 */
  int32_t n_neighbours;           /* pcl::RegionGrowing::setNumberOfNeighbours        (15, :30); <= knn_normals */
  int32_t min_cluster_size;       /* setMinClusterSize                                (50, :27) */
  int32_t max_cluster_size;       /* setMaxClusterSize                                (1000000, :28) */
  float   smoothness_threshold;   /* setSmoothnessThreshold, radians                  ((float)(3.0 / 180.0 * M_PI), :33) */
  float   curvature_threshold;    /* setCurvatureThreshold                            (1.0, :34) */
} aicp_b200_prefilter_config;

typedef struct {
  int64_t n_sampled;              /* points after the voxel grid */
  int64_t n_clusters;             /* clusters kept (size within [min, max]) */
  int64_t n_out;                  /* points in the output cloud */
  int32_t passes;                 /* label-propagation passes enqueued by the region growing */
  int32_t gpu_launches;
  float   ms_total;               /* device time of the call (CUDA events) */
} aicp_b200_prefilter_info;

typedef struct {
  int64_t n_total;                /* reference points after the append */
  int64_t n_recomputed;           /* points whose k-NN list and normal were recomputed (the new ones + the old ones they changed) */
  int32_t incremental;            /* 1: merged into the live index; 0: outside the old bounding box, full rebuild at the next call */
  float   ms;                     /* device time of the incremental update (CUDA events) */
} aicp_b200_append_info;

/* ---- lifetime --------------------------------------------------------------------------------------------------
 * replaces: aicp::create_registrator(params) / PointmatcherRegistration(params)
 *           aicp_core/include/aicp_registration/registration.hpp:9-19, pointmatcher_registration.cpp:7-9
 * icp_yaml_path: libpointmatcher chain file, or NULL/"" for the chain defaults above (the reference calls
 * icp_.setDefault() in that case, pointmatcher_registration.cpp:51-55).  device: CUDA ordinal, or -1 for the current one. */
int aicp_b200_create(const char* icp_yaml_path, int device, aicp_b200_handle** out);
int aicp_b200_destroy(aicp_b200_handle* h);
/* Orders all work the handle (and its batch workers) enqueues from now on after everything already enqueued on
 * `cuda_stream` (a cudaStream_t of the handle's device; NULL = the legacy default stream): the call to make between
 * producing a device-resident input on another stream and handing its pointer to this library.  No reference
 * counterpart (the reference is synchronous host code). */
int aicp_b200_wait_stream(aicp_b200_handle* h, void* cuda_stream);
const char* aicp_b200_last_error(const aicp_b200_handle* h);   /* h may be NULL: error of the last failed create */
const char* aicp_b200_version(void);

/* ---- configuration ---------------------------------------------------------------------------------------------
 * replaces: PointmatcherRegistration::updateConfigParams(path)  pointmatcher_registration.hpp:52-54
 *           + applyConfig() -> icp_.loadFromYaml()               pointmatcher_registration.cpp:48-68
 * The file is only remembered here and re-read at the start of every aicp_b200_register call, because
 * App::computeRegistration rewrites it before each registration (app.cpp:204-205, fileIO.cpp:179-214). */
int aicp_b200_set_config(aicp_b200_handle* h, const char* icp_yaml_path);
/* programmatic override (clears the path); used by the tests and the batched sweep */
int aicp_b200_set_config_struct(aicp_b200_handle* h, const aicp_b200_icp_config* cfg);
/* effective configuration (parses the remembered file now) */
int aicp_b200_get_config(aicp_b200_handle* h, aicp_b200_icp_config* cfg);
/* parse a libpointmatcher chain file without a handle (no CUDA needed) */
int aicp_b200_parse_icp_yaml(const char* icp_yaml_path, aicp_b200_icp_config* cfg, char* err, int err_len);

/* ---- registration ----------------------------------------------------------------------------------------------
 * replaces: PointmatcherRegistration::registerClouds(cloud_ref, cloud_read, final_transform)
 *           pointmatcher_registration.cpp:14-23,92-151   (T = icp_(reading, reference, init))
 * init_T: NULL for identity, else the initial guess that applyInitialization() would build (:71-89).
 * out_T: T such that T * reading aligns with the reference ("initialization is already included", :133).
 * stats: nullable. */
int aicp_b200_register(aicp_b200_handle* h, const float* ref_xyzw, int64_t n_ref, const float* read_xyzw,
                       int64_t n_read, const float* init_T, float* out_T, aicp_b200_stats* stats);

/* Fixed-map localisation (BASELINE.json config 4; App with localize_against_prior_map, app.cpp:41-69,123-127):
 * build the reference index + normals once, then register many readings against it. */
int aicp_b200_set_reference(aicp_b200_handle* h, const float* ref_xyzw, int64_t n_ref);
int aicp_b200_register_to_reference(aicp_b200_handle* h, const float* read_xyzw, int64_t n_read, const float* init_T,
                                    float* out_T, aicp_b200_stats* stats);

/* replaces: getOutputReading(out)        pointmatcher_registration.hpp:48-50  (T * reading, unfiltered)
 *           getInitializedReading(out)   pointmatcher_registration.hpp:37-46  (init_T * reading) */
/* replaces: merging an aligned cloud into the reference map between registrations (App::runAicpPipeline's map update,
 * aicp_core/src/registration/app.cpp:476-493, and AlignedCloud merging) for a reference that stays on the GPU.  Appends n
 * points to the reference of aicp_b200_set_reference / aicp_b200_register_to_reference and updates the index and the
 * normals IN PLACE: the state afterwards is bit for bit what set_reference(old + new) and a full rebuild give, but only the
 * new points are sorted and only the neighbourhoods they enter are recomputed (csrc/append.cu).  A cloud that reaches
 * outside the old bounding box falls back to the full rebuild (info->incremental = 0).  Needs a reference that has been
 * registered against at least once; not available on a handle with a communicator.  info: nullable. */
int aicp_b200_reference_append(aicp_b200_handle* h, const float* xyzw, int64_t n, aicp_b200_append_info* info);
int aicp_b200_get_output_reading(aicp_b200_handle* h, float* xyzw, int64_t n);
int aicp_b200_get_initialized_reading(aicp_b200_handle* h, float* xyzw, int64_t n);

/* descriptors of the filtered reference of the last registration, original point order:
 * normals_xyzd = (nx, ny, nz, density) -- SurfaceNormalDataPointsFilter keepNormals / keepDensities */
int aicp_b200_get_reference_normals(aicp_b200_handle* h, float* normals_xyzd, int64_t n);

/* parity instrumentation: when enabled the next registrations record the correspondence (original reference index)
 * of every reading point at every iteration; fetch with get_trace_matches (iters x n_read int32, row per iteration) */
int aicp_b200_enable_match_trace(aicp_b200_handle* h, int enable);
/* measurement instrumentation for the next registrations (see aicp_b200_stats): level 0 none, 1 CUDA events around the
 * dominant kernel (k_match) only, 2 around every stage (costs ~5 % throughput) */
int aicp_b200_set_profiling(aicp_b200_handle* h, int level);
int aicp_b200_get_trace_matches(aicp_b200_handle* h, int32_t* idx, int64_t iters, int64_t n_read);
/* kernel schedule of the SurfaceNormal k-NN search (identical results): 0 automatic (tile kernel for batched
 * registrations and clouds >= 2^20 points, warp-per-query kernel otherwise), 1 warp per query, 2 one tile of 32 queries
 * per warp.  Exposed for the parity tests and benchmarks. */
int aicp_b200_set_knn_schedule(aicp_b200_handle* h, int schedule);
/* kernel schedule of the ICP correspondence search (identical results): 0 automatic (tile kernel inside batched registrations), 1 one query per thread (k_match),
 * 2 one tile of 32 queries per warp with a shared tree walk (k_match_tile) */
int aicp_b200_set_match_schedule(aicp_b200_handle* h, int schedule);
/* how the ICP loop (ICP::compute's while(iterate)) is driven (identical results):
 * 2 ONE persistent cooperative kernel runs every iteration -- grid barriers between the phases, loop control on the device,
 *   the sharded exchange inside the kernel;
 * 1 three launches per iteration with the host staying two iterations ahead of a progress word;
 * 0 automatic: 2 for one registration at a time and for the sharded registration, 1 inside batches (measured) */
int aicp_b200_set_loop_schedule(aicp_b200_handle* h, int schedule);

/* ---- stage entry points (same kernels as aicp_b200_register; exposed for the parity tests) -----------------------
 * SurfaceNormalDataPointsFilter alone: out_normals n x 4, out_knn nullable n x knn (ids sorted by (d2, id)) */
int aicp_b200_surface_normals(aicp_b200_handle* h, const float* xyzw, int64_t n, int32_t knn, float* out_normals,
                              int32_t* out_knn);
/* KDTreeMatcher{knn 1, epsilon 0}::findClosests: out_idx[i] = argmin_j (d2(q_i, ref_j), j), out_d2 = squared distance */
int aicp_b200_match(aicp_b200_handle* h, const float* ref_xyzw, int64_t n_ref, const float* qry_xyzw, int64_t n_qry,
                    int32_t* out_idx, float* out_d2);
/* TrimmedDistOutlierFilter threshold: k-th smallest finite positive d2, k = size_t(float(n_valid) * ratio) */
int aicp_b200_trim_threshold(aicp_b200_handle* h, const float* d2, int64_t n, float ratio, float* out_limit,
                             int64_t* out_n_valid);

/* ---- overlap ---------------------------------------------------------------------------------------------------
 * replaces: OctreesOverlap::computeOverlap(ref_cloud, read_cloud, ref_pose, read_pose, reading_tree) + getOverlap()
 *           aicp_core/src/overlap/octrees_overlap.cpp:29-72 (createTree :153-218, getOverlappingNodes :113-151)
 * origins: translation of the two sensor poses (only the translation is used as ray origin, :229-230).
 * resolution: OverlapParams.octree_based.octomapResolution, i.e. (double)0.2f with the shipped config.
 * counts: nullable, {|A^B|, |A|, |B|} voxel-key counts.  overlap_pct in [0,100]. */
int aicp_b200_overlap(aicp_b200_handle* h, const float* ref_xyzw, int64_t n_ref, const double ref_origin[3],
                      const float* read_xyzw, int64_t n_read, const double read_origin[3], double resolution,
                      float* overlap_pct, int64_t counts[3]);

/* ---- map handling ---------------------------------------------------------------------------
 * replaces: getPointsInOrientedBox(cloud, min, max, origin)   aicp_core/src/utils/filteringUtils.cpp:621-637
 *           = pcl::CropBox{min (m,m,m), max (M,M,M), rotation origin.block<3,3>(0,0).eulerAngles(0,1,2), translation origin.col(3)},
 *           which App runs on the whole prior / built map before every registration against it (app.cpp:41-69).
 * rotation_rpy: the three Euler angles exactly as the reference passes them to CropBox::setRotation (the adapter computes them
 * with Eigen like the reference does); translation: origin.col(3).  A point is kept when min <= R(rpy)^-1 (p - t) <= max on
 * every axis; the output keeps the input order; non-finite points are dropped.
 * out_xyzw: host or device buffer of capacity n records, or NULL to keep the result on the device only -- then
 * aicp_b200_get_cropped() returns its device address, valid until the next crop on this handle, which can be passed straight
 * to aicp_b200_register / aicp_b200_set_reference as the reference cloud. */
int aicp_b200_crop_box(aicp_b200_handle* h, const float* xyzw, int64_t n, float box_min, float box_max,
                       const float rotation_rpy[3], const float translation[3], float* out_xyzw, int64_t* n_out);
const float* aicp_b200_get_cropped(aicp_b200_handle* h, int64_t* n_out);
int aicp_b200_download_cropped(aicp_b200_handle* h, float* xyzw, int64_t n);   /* copy of the last device-resident crop */

/* A device-resident map, so that the (10 M-point, 168 MB) prior map is uploaded once instead of once per registration:
 * replaces the host-side aligned_map_ / prior_map_ clouds of App (app.hpp:140-160).
 *   map_append(replace = 1)  prior_map_->updateCloud(cloud)                              app.cpp:478-493
 *   map_append(replace = 0)  *merged_map = *(prior_map_->getCloud()) + *output           app.cpp:476-480 (concatenation)
 *   map_crop                 getPointsInOrientedBox(copy of the map, -c, +c, prior pose) app.cpp:41-69; result: get_cropped */
int aicp_b200_map_append(aicp_b200_handle* h, const float* xyzw, int64_t n, int replace);
int64_t aicp_b200_map_size(const aicp_b200_handle* h);
int aicp_b200_map_crop(aicp_b200_handle* h, float box_min, float box_max, const float rotation_rpy[3], const float translation[3],
                       int64_t* n_out);

/* ---- pre-filter ----------------------------------------------------------------------------
 * replaces: regionGrowingUniformPlaneSegmentationFilter(cloud_in, cloud_out)   aicp_core/src/utils/filteringUtils.cpp:5-45
 *           = pcl::VoxelGrid{0.08} -> pcl::NormalEstimation{k 30} -> pcl::RegionGrowing{50, 1e6, 15, 3 deg, 1.0} -> the kept
 *           clusters concatenated in cluster order; App runs it on every reading (App::filterCloud, app.cpp:102-110), on the
 *           first cloud (app.cpp:295) and on the merged map every 30 clouds (app.cpp:486-493);
 *           and regionGrowingUniformPlaneSegmentationFilter(cloud_in, cloud_sampled_out, view_point, clusters)  :51-104,
 *           the variant used by the alignability filter (normals flipped towards view_point, clusters returned).
 * cfg: NULL for the reference's hard-coded parameters.  viewpoint: NULL for (0,0,0) (the first overload never sets one).
 * out_xyzw: host or device buffer of capacity n records, or NULL to keep the result on the device only
 * (aicp_b200_get_prefiltered returns its address, valid until the next pre-filter call on this handle, and can be passed
 * to aicp_b200_register / aicp_b200_overlap as a device cloud).  After the call the by-products of the second overload can
 * be fetched: the sampled cloud (voxel-grid output), its normals (nx, ny, nz, curvature) and the cluster ordinal of every
 * sampled point (-1: not in a kept cluster; clusters[c].indices = ascending { i : label[i] == c }). */
int aicp_b200_prefilter_default_config(aicp_b200_prefilter_config* cfg);
int aicp_b200_prefilter(aicp_b200_handle* h, const float* xyzw, int64_t n, const aicp_b200_prefilter_config* cfg,
                        const float viewpoint[3], float* out_xyzw, int64_t* n_out, aicp_b200_prefilter_info* info);
const float* aicp_b200_get_prefiltered(aicp_b200_handle* h, int64_t* n_out);
int aicp_b200_prefilter_get_sampled(aicp_b200_handle* h, float* xyzw, int64_t n_sampled);
int aicp_b200_prefilter_get_normals(aicp_b200_handle* h, float* normals_xyzc, int64_t n_sampled);
int aicp_b200_prefilter_get_labels(aicp_b200_handle* h, int32_t* labels, int64_t n_sampled);
/* pcl::VoxelGrid alone (filteringUtils.cpp:10-13): one centroid per occupied voxel, ascending voxel index; non-finite points
 * are skipped; when leaf is too small for the cloud's extent (more than INT32_MAX voxels) the input is returned unchanged,
 * as PCL does.  out_xyzw: capacity n records, host or device. */
int aicp_b200_voxel_grid(aicp_b200_handle* h, const float* xyzw, int64_t n, float leaf_size, float* out_xyzw, int64_t* n_out);
/* the periodic re-filter of the merged map (app.cpp:486-493): map <- prefilter(map), all on the device */
int aicp_b200_map_prefilter(aicp_b200_handle* h, const aicp_b200_prefilter_config* cfg, int64_t* n_out, aicp_b200_prefilter_info* info);

/* ---- ingest ----------------------------------------------------------------------------------------
 * replaces: VelodyneAccumulatorROS::processLidar   aicp_ros/src/velodyne_accumulator.cpp:31-73: crop the sweep to +-box_half
 *           (30 m, :59-60) around the sensor, transformPointCloud with (body_pose.translation().cast<float>(),
 *           Quaternionf(body_pose.rotation().cast<float>())) (:62-63), append to the accumulated cloud (:66).
 * body_pose: 16 doubles column-major (inertial <- sensor).  clear_first: start a new accumulation (clearCloud, :76-81).
 * The accumulated cloud stays on the device: aicp_b200_get_accumulated returns its address (valid until the next accumulate
 * call), to be passed to aicp_b200_prefilter / aicp_b200_register as a device cloud; download_accumulated copies it out. */
int aicp_b200_accumulate_sweep(aicp_b200_handle* h, const float* sweep_xyzw, int64_t n, float box_half, const double body_pose[16],
                               int clear_first, int64_t* n_added);
const float* aicp_b200_get_accumulated(aicp_b200_handle* h, int64_t* n_out);
int aicp_b200_download_accumulated(aicp_b200_handle* h, float* xyzw, int64_t n);
/* replaces: pcl::io::loadPCDFile<pcl::PointXYZ>(path, cloud) as the replay uses it (app.cpp:269) and the PCD writers of the tools
 * (cloudIO.cpp:64, create_cube_cloud.cpp:84).  PCD v0.7, DATA ascii or binary, float32 x y z fields; no handle, no CUDA.
 * read: out_xyzw NULL returns the point count only. */
int aicp_b200_read_pcd(const char* path, float* out_xyzw, int64_t capacity, int64_t* n_out, char* err, int err_len);
/* replaces: pcl::io::loadPLYFile<pcl::PointXYZ>(path, map)  aicp_ros/src/app_ros.cpp:301 (the prior map).  ascii or
 * binary_little_endian, float / double x y z vertex properties, elements after the vertices ignored. */
int aicp_b200_read_ply(const char* path, float* out_xyzw, int64_t capacity, int64_t* n_out, char* err, int err_len);
int aicp_b200_write_pcd(const char* path, const float* xyzw, int64_t n, char* err, int err_len);
/* replaces: PoseFileReader::readPoseFile   aicp_core/include/aicp_utils/poseFileReader.hpp:46-78 (aicp_input_poses.csv of the
 * replay format, app.cpp:250-279): rows "counter, sec, nsec, x, y, z, qx, qy, qz, qw".  rows: n x 3 (counter, sec, nsec);
 * poses: n x 16 doubles column-major.  rows / poses NULL returns the row count only. */
int aicp_b200_read_pose_file(const char* path, int64_t* rows, double* poses, int64_t capacity, int64_t* n_out, char* err, int err_len);

/* ---- FOV overlap filter + alignability ---------------------------------------------------------------
 * replaces: overlapFilter(cloudA, cloudB, poseA, poseB, range, angularView, accepted_pointsA, accepted_pointsB)
 *           aicp_core/src/utils/filteringUtils.cpp:111-193 (App::computeAlignmentRisk, app.cpp:153-156): the points of each cloud
 *           that lie within the other sensor's range and horizontal field of view; returns 100 * (kept A / |A|) * (kept B / |B|).
 * poses: 16 doubles, column-major == Eigen::Isometry3d::matrix().data().  range / angular_view: RegistrationParams.sensorRange /
 * sensorAngularView (aicp_config.yaml:4-5).  out_a / out_b: nullable buffers of capacity n_a / n_b records (host or device); the
 * accepted clouds also stay on the device (aicp_b200_get_fov_filtered: which = 0 for A, 1 for B) until the next call. */
int aicp_b200_fov_overlap(aicp_b200_handle* h, const float* a_xyzw, int64_t n_a, const float* b_xyzw, int64_t n_b,
                          const double pose_a[16], const double pose_b[16], float range, float angular_view, float* out_a,
                          float* out_b, int64_t counts[2], float* overlap_pct);
const float* aicp_b200_get_fov_filtered(aicp_b200_handle* h, int which, int64_t* n_out);
/* replaces: alignabilityFilter(cloudA, cloudB, poseA, poseB, cloudA_planes, cloudB_planes, eigenvectors)  filteringUtils.cpp:196-430:
 *           both clouds through the pre-filter (normals towards each sensor, plane clusters), clusters matched through their
 *           oriented bounding boxes (:236-282, overlapBoxFilter :507-576), PCA of the matched normals -> 100 * lambda_min / lambda_max.
 * cfg: NULL for the pre-filter's hard-coded parameters.  matching: nullable, receives min(clusters of B, matching_cap) entries
 * (index of the matched cluster of A, or -1: matching_indeces of :229).  info: nullable {clusters A, clusters B, matched}. */
int aicp_b200_alignability(aicp_b200_handle* h, const float* a_xyzw, int64_t n_a, const float* b_xyzw, int64_t n_b,
                           const double pose_a[16], const double pose_b[16], const aicp_b200_prefilter_config* cfg,
                           float* alignability_pct, int32_t* matching, int64_t matching_cap, int64_t info[3]);
/* replaces: App::computeAlignmentRisk (app.cpp:143-185): FOV overlap -> alignability of the two accepted clouds -> SVM probability
 * of (octree_overlap_pct, alignability).  Needs a loaded model (aicp_b200_svm_load).  Everything between the input clouds and the
 * three scalars stays on the device. */
int aicp_b200_alignment_risk(aicp_b200_handle* h, const float* ref_xyzw, int64_t n_ref, const float* read_xyzw, int64_t n_read,
                             const double ref_pose[16], const double read_pose[16], float range, float angular_view,
                             float octree_overlap_pct, float* fov_overlap_pct, float* alignability_pct, double* risk);

/* ---- alignment-risk classifier ------------------------------------------------------------------
 * replaces: aicp::SVM::load(filename)          aicp_core/src/classification/svm.cpp:103-107  (cv::ml::SVM::load)
 *           aicp::SVM::test(data, probs)       svm.cpp:53-101: raw decision value of cv::ml::SVM::predict(sample, out, 1),
 *                                              probability = 1.0 - 1.0 / (1.0 + exp(-raw))  (:82)
 * as App::computeAlignmentRisk uses them on (octree overlap, alignability) (app.cpp:175-181).  model_xml_path: an OpenCV SVM
 * file as shipped in aicp_core/data/classification/ (C_SVC, two classes, POLY or LINEAR kernel; OpenCV 3 or legacy 2.4 layout).
 * features: n x dim doubles, row-major (converted to float32 like svm.cpp:72-74); probabilities: n doubles; raw: nullable.
 * Training (SVM::train) is an offline tool of the reference and is not provided. */
typedef struct {
  int32_t kernel;                 /* 0 LINEAR, 1 POLY */
  int32_t dim, sv_total, sv_count;
  double  degree, gamma, coef0, rho;
  double  alpha_sum, sv_sum;      /* plain sums in file order: a cheap fingerprint of what the reader understood */
  int32_t index_first, index_last;
} aicp_b200_svm_summary;
/* parse a model file without a handle (no CUDA needed) */
int aicp_b200_svm_parse(const char* model_xml_path, aicp_b200_svm_summary* out, char* err, int err_len);
int aicp_b200_svm_load(aicp_b200_handle* h, const char* model_xml_path);
int aicp_b200_svm_info(aicp_b200_handle* h, int32_t* dim, int32_t* sv_total);
int aicp_b200_svm_predict(aicp_b200_handle* h, const double* features, int64_t n, int32_t dim, double* probabilities, float* raw);

/* ---- auto-tune glue --------------------------------------------------------------------------------------------
 * replaces (for callers that do not go through a file): App::computeRegistration's clamp, app.cpp:198-202, followed by
 * the 6-significant-digit text round trip of replaceRatioConfigFile, fileIO.cpp:194-198.  Pure host code. */
float aicp_b200_autotune_ratio(float overlap_pct);

/* ---- batched registration (BASELINE.json config 5: independent pairs, no communication) --------------------------
 * replaces the bash sweeps that call the pairwise tool once per pair (bash/run_registration_validation.sh:7-21,
 * bash/run_registration.sh:8-36) and KITTI-sequence style frame-to-frame runs.
 * The pairs are independent, so they are registered CONCURRENTLY on `streams` CUDA streams of the handle's device (one
 * worker handle + host thread per stream; a single 128k-point registration does not fill a B200).
 * All pairs share the handle's configuration except the trimmed ratio, which may be given per pair (auto-tuned from each
 * pair's overlap).  ref/read arrays hold n_pairs pointers (host or device); out_T: n_pairs x 16; stats: nullable array
 * of n_pairs; status: nullable array of n_pairs per-pair return codes (a failing pair does not stop the batch; the
 * function returns the first non-zero code).  batch_ms (nullable): device time from the first kernel of the batch to
 * the last, measured with CUDA events across all streams.  streams <= 0 selects the default (4). */
int aicp_b200_register_batch(aicp_b200_handle* h, int64_t n_pairs, const float* const* ref_xyzw, const int64_t* n_ref,
                             const float* const* read_xyzw, const int64_t* n_read, const float* ratios /*nullable*/,
                             int streams, float* out_T, aicp_b200_stats* stats, int32_t* status, float* batch_ms);

/* One whole AICP step per pair, batched: replaces App::runAicpPipeline's computeOverlap + computeRegistration
 * (aicp_core/src/registration/app.cpp:218-247, 112-141, 187-216) for many independent pairs -- the registration-validation
 * sweep of BASELINE.json config 5 ("overlap + ..."): per pair the octree overlap, the clamp to [0.25, 0.70] with the 6-digit
 * text round trip, and the registration with that trimmed ratio, all inside the pair's worker stream.
 * ref_origins / read_origins: n_pairs x 3 doubles (sensor pose translations); out_overlap: nullable, n_pairs percentages. */
/* The same over SEVERAL GPUs of the node from one process ("pair i -> GPU i mod G"): pair i is registered
 * on devices[i % n_devices]; every device gets its own worker pool (`streams` streams and host threads) and the results are
 * gathered on the host.  No data-path communication between the devices.  Clouds may be host pointers, or device pointers
 * of any device (copied device to device).  batch_ms: the longest device time over the devices. */
int aicp_b200_register_batch_devices(aicp_b200_handle* h, const int32_t* devices, int32_t n_devices, int64_t n_pairs,
                                     const float* const* ref_xyzw, const int64_t* n_ref, const float* const* read_xyzw,
                                     const int64_t* n_read, const float* ratios, int streams, float* out_T, aicp_b200_stats* stats,
                                     int32_t* status, float* batch_ms);
int aicp_b200_aicp_batch(aicp_b200_handle* h, int64_t n_pairs, const float* const* ref_xyzw, const int64_t* n_ref,
                         const double* ref_origins, const float* const* read_xyzw, const int64_t* n_read,
                         const double* read_origins, double resolution, int streams, float* out_T, float* out_overlap,
                         aicp_b200_stats* stats, int32_t* status, float* batch_ms);

/* The same with the alignment-risk gate: App::runAicpPipeline with failure_prediction_mode (app.cpp:218-247) for many
 * independent pairs -- BASELINE.json config 5 ("overlap + alignment-risk"): per pair the octree overlap, computeAlignmentRisk
 * (FOV overlap -> alignability -> SVM on (overlap, alignability)), and the registration with the auto-tuned ratio ONLY when the
 * risk is at most risk_threshold (:241-243); a skipped pair returns the identity and zeroed stats.
 * ref_poses / read_poses: n_pairs x 16 doubles (column-major sensor poses; their translations are the overlap's ray origins).
 * svm_model_path: OpenCV model file (see aicp_b200_svm_load).  out_overlap / out_alignability / out_risk: nullable.
 * prefilter_first != 0: the clouds are RAW (e.g. accumulated sweeps); every worker first runs the pre-filter on both
 * (App::setAndFilterReading / filterCloud, app.cpp:77-110) and the rest of the step uses the filtered clouds, which never leave
 * the device; out_n_filtered (nullable, n_pairs x 2) receives their sizes. */
int aicp_b200_pipeline_batch(aicp_b200_handle* h, int64_t n_pairs, const float* const* ref_xyzw, const int64_t* n_ref,
                             const double* ref_poses, const float* const* read_xyzw, const int64_t* n_read, const double* read_poses,
                             double resolution, float sensor_range, float angular_view, const char* svm_model_path, double risk_threshold,
                             int prefilter_first, int streams, float* out_T, float* out_overlap, float* out_alignability, double* out_risk,
                             int64_t* out_n_filtered, aicp_b200_stats* stats, int32_t* status, float* batch_ms);

/* ---- multi-GPU single registration (BASELINE.json config 4: reading sharded, reference replicated) ---------------
 * nccl_unique_id: the 128-byte ncclUniqueId obtained on rank 0 with aicp_b200_comm_unique_id and broadcast by the
 * caller (e.g. torch.distributed).  After comm_init, aicp_b200_register* calls on every rank take that rank's SHARD of
 * the reading; the trimmed quantile and the 6x6 normal equations are all-reduced over NCCL every iteration and every
 * rank returns the same transform. */
int aicp_b200_comm_unique_id(uint8_t id_out[128]);
int aicp_b200_comm_init(aicp_b200_handle* h, const uint8_t nccl_unique_id[128], int rank, int n_ranks);
int aicp_b200_comm_destroy(aicp_b200_handle* h);
/* Human-readable description of the exchange the communicator performs per ICP iteration (written to buf, NUL-terminated). */
int aicp_b200_comm_info(aicp_b200_handle* h, char* buf, int len);

#ifdef __cplusplus
}
#endif
#endif
