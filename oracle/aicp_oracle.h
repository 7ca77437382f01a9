/*
 * aicp_oracle.h -- CPU ORACLE for the AICP registration hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the arithmetic that aicp_core runs below
 *   aicp::PointmatcherRegistration::registerClouds()   (aicp_core/src/registration/pointmatcher_registration.cpp:92-151)
 *   aicp::OctreesOverlap::computeOverlap()             (aicp_core/src/overlap/octrees_overlap.cpp:29-72)
 * i.e. the libpointmatcher chain configured by aicp_core/config/icp/icp_autotuned.yaml:9-58 and the
 * octomap ray insertion + leaf-key intersection.  The arithmetic itself lives in libpointmatcher (>=1.3.x),
 * libnabo (1.0.x) and octomap (1.9.x), none of which is vendored under /root/reference nor installed in this
 * image; the reference's only test compares against a golden file outside the repository
 * (aicp_core/test/aicp_test.cpp:50-57).  ==> PARITY UNPINNED: this oracle is checked against analytic
 * known-answer cases and the in-repo text KATs only, not against outputs of the reference itself.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  The product (aicp_mapping_b200/) never links, imports or executes anything under oracle/.
 *
 * Numerical contract restated here (and independently implemented in CUDA by the product):
 *   - float32 geometry with a fixed operation order and NO fused multiply-add
 *     (d2 = ((dx*dx)+(dy*dy))+(dz*dz), libnabo nabo/kdtree_cpu.cpp leaf scan);
 *   - nearest neighbour = exact minimum of (d2, reference index) -- libnabo epsilon = 0, ties to the lowest index;
 *   - k-NN for normals = the k smallest (d2, index) pairs, self included;
 *   - reductions (centroid, 6x6 normal equations) are accumulated EXACTLY as fixed-point integers, so the
 *     result does not depend on summation order, thread count or GPU sharding;
 *   - the 6x6 solve, pose increment and convergence checkers run in float64 using only + - * / sqrt
 *     (own sin/cos/atan series) so that a GPU thread executing the same sequence yields the same bits.
 */
#ifndef AICP_ORACLE_H_
#define AICP_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_ITERS 256
#define ORC_FIXED_SHIFT 30          /* normal-equation terms are rounded to multiples of 2^-30 */
#define ORC_CENTROID_SHIFT 16       /* centroid terms are rounded to multiples of 2^-16 m */

enum {
  ORC_OK = 0,
  ORC_ERR_BAD_ARG = 1,
  ORC_ERR_KNN_TOO_LARGE = 2,      /* libpointmatcher: knn must be < number of points */
  ORC_ERR_NO_VALID_MATCH = 3,     /* TrimmedDist: no finite positive distance -> ConvergenceError */
  ORC_ERR_NAN = 4,                /* checker: NaN in transformation -> ConvergenceError */
  ORC_ERR_NONFINITE_INPUT = 5,
  ORC_ERR_EXTENT = 6              /* |coordinate| in the centred frame exceeds 1024 m */
};

enum { ORC_STOP_NONE = 0, ORC_STOP_COUNTER = 1, ORC_STOP_DIFFERENTIAL = 2 };

/* libpointmatcher chain parameters actually used by icp_autotuned.yaml (aicp_core/config/icp/icp_autotuned.yaml:9-58) */
typedef struct {
  int32_t knn_normals;          /* SurfaceNormalDataPointsFilter.knn (20) */
  int32_t reading_normals;      /* 1: also run the (dead) reading SurfaceNormal filter, as libpointmatcher does */
  float   ratio;                /* TrimmedDistOutlierFilter.ratio */
  int32_t max_iterations;       /* CounterTransformationChecker.maxIterationCount (20) */
  float   min_diff_rot;         /* DifferentialTransformationChecker.minDiffRotErr (0.001) */
  float   min_diff_trans;       /* .minDiffTransErr (0.01) */
  int32_t smooth_length;        /* .smoothLength (4) */
  int32_t use_kdtree;           /* 0: brute force O(N*M) search, 1: sliding-midpoint kd-tree (bucket 8) */
  int32_t threads;              /* OpenMP threads over query points (1 = libpointmatcher's own single-threaded loop) */
} orc_icp_config;

typedef struct {
  float   T_iter[16];           /* column-major 4x4 after this iteration's update (centred frame) */
  float   limit_d2;             /* trimmed squared-distance threshold */
  int64_t n_valid;              /* finite, >0 distances entering the quantile */
  int64_t n_used;               /* sum of weights */
  double  rot_err, trans_err;   /* Differential checker means (NaN until history is long enough) */
} orc_iter_trace;

typedef struct {
  int32_t iterations;
  int32_t stop_reason;
  float   weighted_point_used_ratio;   /* sum(w)/(knn*N) of the last iteration (pointmatcher_registration.cpp:114) */
  float   mean_ref[3];
  orc_iter_trace trace[ORC_MAX_ITERS];
} orc_icp_result;

/* ---- stage functions (each is also used inside orc_icp) ---- */

/* SURVEY.md A.2.  pts: n x 4 floats (x,y,z,pad), stride 16 B like pcl::PointXYZ (cloudIO.cpp:81-98).
 * out_normals: n x 4 floats (nx,ny,nz,density).  out_knn (nullable): n x k int32 neighbour ids sorted by (d2,id). */
int orc_surface_normals(const float* pts, int64_t n, int32_t k, int use_kdtree, int threads,
                        float* out_normals, int32_t* out_knn);

/* SURVEY.md A.3 with epsilon = 0.  out_idx[i] = argmin over ref of (d2, j); out_d2[i] = that d2. */
int orc_match(const float* ref, int64_t n_ref, const float* qry, int64_t n_qry, int use_kdtree, int threads,
              int32_t* out_idx, float* out_d2);

/* SURVEY.md A.4.  Returns the threshold and the count of values entering the quantile. */
int orc_trim_threshold(const float* d2, int64_t n, float ratio, float* out_limit, int64_t* out_n_valid);

/* SURVEY.md A.5 accumulation.  p: step reading (n x 4), ref/normals indexed by idx; weight = d2 <= limit.
 * sums_hi/lo: 27 entries (21 upper-triangle A row-major i<=j, then 6 of g = sum F*(delta.n)), two's complement
 * 128-bit fixed point with ORC_FIXED_SHIFT fractional bits.  */
int orc_normal_equations(const float* p, int64_t n, const float* ref, const float* normals, const int32_t* idx,
                         const float* d2, float limit, int64_t* sums_hi, uint64_t* sums_lo, int64_t* out_n_used);

/* 27 fixed-point sums -> x (6 doubles: rotation vector, translation).  Returns 1 if the Cholesky path was
 * used, 2 for the rank-revealing fallback. */
int orc_solve6(const int64_t* sums_hi, const uint64_t* sums_lo, double* x);

/* x -> float 4x4 increment (column-major) */
void orc_pose_increment(const double* x, float* dT);

/* deterministic libm subset (exposed for tests) */
void   orc_sincos(double x, double* s, double* c);
double orc_atan2_pos(double y, double x);

/* out = T * in (x,y,z rows; pad copied), float, k-ascending, no FMA */
void orc_transform_points(const float* T, const float* in, int64_t n, float* out);

/* full chain, SURVEY.md A.1.  init_T nullable (identity).  out_T column-major.  out_reading nullable (n_read x 4).
 * out_normals nullable (n_ref x 4).  trace_idx nullable: max_iterations x n_read int32 correspondences. */
int orc_icp(const float* ref, int64_t n_ref, const float* read, int64_t n_read, const float* init_T,
            const orc_icp_config* cfg, float* out_T, float* out_reading, float* out_normals,
            int32_t* trace_idx, orc_icp_result* res);

/* ---- overlap (octrees_overlap.cpp:29-72,113-241) ---- */
/* origin: sensor pose translation as double[3] (cast to float like octomap::pose6d).  counts = {n(A^B), n(A), n(B)} */
int orc_overlap(const float* ref, int64_t n_ref, const double* ref_origin, const float* read, int64_t n_read,
                const double* read_origin, double resolution, float* out_overlap_pct, int64_t* counts);
/* sorted unique 48-bit keys of one cloud (x<<32|y<<16|z); returns count; keys may be NULL to only count */
int64_t orc_ray_keys(const float* pts, int64_t n, const double* origin, double resolution, uint64_t* keys, int64_t cap);

/* ---- map handling: getPointsInOrientedBox = pcl::CropBox (filteringUtils.cpp:621-637), see aicp_oracle_filters.c ---- */
void orc_rpy_to_matrix(const float* rpy, float* R_rowmajor9);
/* returns the number of points kept; out (nullable) receives them in input order, capacity n x 4 floats */
int64_t orc_crop_box(const float* xyzw, int64_t n, float bmin, float bmax, const float* rpy, const float* translation, float* out);

/* ---- pre-filter: regionGrowingUniformPlaneSegmentationFilter (filteringUtils.cpp:5-104), see aicp_oracle_prefilter.c ---- */
typedef struct {
  float   leaf_size;              /* VoxelGrid leaf (0.08f) */
  int32_t knn_normals;            /* NormalEstimation.setKSearch (30) */
  int32_t n_neighbours;           /* RegionGrowing.setNumberOfNeighbours (15) */
  int32_t min_cluster_size;       /* 50 */
  int32_t max_cluster_size;       /* 1000000 */
  float   smoothness_threshold;   /* radians, (float)(3/180*pi) */
  float   curvature_threshold;    /* 1.0 */
} orc_prefilter_config;
void orc_prefilter_default_config(orc_prefilter_config* cfg);
/* pcl::VoxelGrid: returns the number of voxels (or the input size in the "leaf too small" case), or -(error code) */
int64_t orc_voxel_grid(const float* xyzw, int64_t n, float leaf, float* out);
void orc_pcl_point_normal(const float* pts, const int32_t* nb, int32_t k, const float* query, const float* viewpoint, float* out4);
int64_t orc_region_growing(const float* normals, const int32_t* knn, int64_t m, int32_t k_stride, int32_t n_nb,
                           int32_t min_size, int32_t max_size, float cos_thr, float curv_thr, int32_t* labels);
/* counts = {n_sampled, n_clusters, n_out}; viewpoint nullable (origin) */
int orc_prefilter(const float* xyzw, int64_t n, const orc_prefilter_config* cfg, const float* viewpoint, int threads,
                  float* sampled, float* normals, int32_t* labels, float* out, int64_t* counts);

/* ---- FOV overlap filter + alignability (filteringUtils.cpp:111-576), see aicp_oracle_alignability.c ---- */
/* poses: 16 doubles, column-major (Eigen::Isometry3d::matrix().data()).  counts = {accepted A, accepted B} */
float orc_fov_overlap(const float* A, int64_t nA, const float* B, int64_t nB, const double* poseA, const double* poseB, float range,
                      float angular_view, float* outA, float* outB, int64_t* counts);
void orc_euler_angles_012(const float* R_rowmajor9, float* rpy);
/* matching: nullable, one entry per kept cluster of B (index of the matched cluster of A or -1); info = {clusters A, clusters B, matched} */
int orc_alignability(const float* A, int64_t nA, const float* B, int64_t nB, const double* poseA, const double* poseB,
                     const orc_prefilter_config* cfg, int threads, float* out_alignability, int32_t* matching, int64_t* info);

/* ---- sweep accumulation (velodyne_accumulator.cpp:31-73), see aicp_oracle_ingest.c ---- */
void orc_pose_to_float_transform(const double* pose_colmajor16, float* T_colmajor16);
int64_t orc_accumulate_sweep(const float* sweep, int64_t n, float half, const double* body_pose, float* out);

/* ---- text glue KATs ---- */
/* app.cpp:198-202 clamp + fileIO.cpp:194-198 "%g"-style 6-digit print + float re-parse */
float orc_autotune_ratio(float overlap_pct, char* text_out /* >=32 bytes, nullable */);

#ifdef __cplusplus
}
#endif
#endif
