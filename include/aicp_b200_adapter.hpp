// aicp_b200_adapter.hpp -- header-only C++ adapters that plug libaicp_b200.so into aicp_core behind the reference's own
// plug-in interfaces.  This is the reference-side binding a maintainer adds (INTEGRATION.md shows the two factory lines):
//
//   aicp::B200Registration : aicp::AbstractRegistrator     aicp_core/include/aicp_registration/abstract_registrator.hpp:8-19
//       stands where aicp::PointmatcherRegistration stands   aicp_core/include/aicp_registration/pointmatcher_registration.hpp:22-67
//   aicp::B200Overlap      : aicp::AbstractOverlapper       aicp_core/include/aicp_overlap/abstract_overlapper.hpp:13-19
//       stands where aicp::OctreesOverlap stands             aicp_core/include/aicp_overlap/octrees_overlap.hpp:20-58
//
// It needs the headers aicp_core already uses (PCL point types, Eigen, octomap) and nothing else; all arithmetic happens
// on the GPU behind the C ABI of aicp_b200.h.  Error convention: the reference exit(1)s on a bad config or cloud
// (pointmatcher_registration.cpp:60-64,96-100); the adapter prints the same kind of message to cerr, leaves
// final_transform = identity and keeps the process alive -- App::runAicpPipeline's own sanity gate on the correction
// magnitude (app.cpp:366-373) then sees a null correction.
#ifndef AICP_B200_ADAPTER_HPP_
#define AICP_B200_ADAPTER_HPP_

#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "aicp_b200.h"

#include "aicp_registration/abstract_registrator.hpp"
#include "aicp_registration/common.hpp"
#include "aicp_overlap/abstract_overlapper.hpp"
#include "aicp_overlap/common.hpp"

namespace aicp {

namespace b200_detail {

// pcl::PointXYZ is a 16-byte (x, y, z, pad) record: it is passed to the C ABI in place.  Wider point types are packed.
template <typename PointT>
inline const float* as_xyzw(const pcl::PointCloud<PointT>& cloud, std::vector<float>& staging, int64_t* n) {
  // the reference takes cloud.width as the point count (cloudIO.cpp:83)
  size_t count = cloud.width;
  if (count > cloud.points.size()) count = cloud.points.size();
  *n = (int64_t)count;
  if (sizeof(PointT) == 16) return reinterpret_cast<const float*>(cloud.points.data());
  staging.resize(4 * count);
  for (size_t i = 0; i < count; ++i) {
    staging[4 * i + 0] = cloud.points[i].x;
    staging[4 * i + 1] = cloud.points[i].y;
    staging[4 * i + 2] = cloud.points[i].z;
    staging[4 * i + 3] = 1.0f;
  }
  return staging.data();
}

inline void to_pcl(const std::vector<float>& xyzw, pcl::PointCloud<pcl::PointXYZ>& out) {
  // fromDataPointsToPCL, cloudIO.cpp:68-79
  size_t n = xyzw.size() / 4;
  out.points.resize(n);
  for (size_t i = 0; i < n; ++i) {
    out.points[i].x = xyzw[4 * i + 0];
    out.points[i].y = xyzw[4 * i + 1];
    out.points[i].z = xyzw[4 * i + 2];
  }
  out.width = (uint32_t)n;
  out.height = 1;
}

}  // namespace b200_detail

class B200Registration : public AbstractRegistrator {
 public:
  B200Registration() : h_(nullptr), n_read_(0) { open(); }
  explicit B200Registration(const RegistrationParams& params) : params_(params), h_(nullptr), n_read_(0) { open(); }
  ~B200Registration() { if (h_) aicp_b200_destroy(h_); }
  B200Registration(const B200Registration&) = delete;
  B200Registration& operator=(const B200Registration&) = delete;

  virtual void registerClouds(pcl::PointCloud<pcl::PointXYZ>& cloud_ref, pcl::PointCloud<pcl::PointXYZ>& cloud_read,
                              Eigen::Matrix4f& final_transform) {
    registerAny(cloud_ref, cloud_read, final_transform);
  }
  virtual void registerClouds(pcl::PointCloud<pcl::PointXYZRGB>& cloud_ref, pcl::PointCloud<pcl::PointXYZRGB>& cloud_read,
                              Eigen::Matrix4f& final_transform) {
    registerAny(cloud_ref, cloud_read, final_transform);
  }
  // an empty no-op in the reference as well (pointmatcher_registration.cpp:36-45)
  virtual void registerClouds(pcl::PointCloud<pcl::PointXYZRGBNormal>&, pcl::PointCloud<pcl::PointXYZRGBNormal>&, Eigen::Matrix4f&) {}

  virtual void getInitializedReading(pcl::PointCloud<pcl::PointXYZ>& initialized_reading) {
    if (params_.pointmatcher.initialTransform.empty())
      std::cout << "[B200] Reading cloud not initialized here." << std::endl;       // pointmatcher_registration.hpp:43
    fetch(&aicp_b200_get_initialized_reading, initialized_reading);
  }
  virtual void getOutputReading(pcl::PointCloud<pcl::PointXYZ>& out_read_cloud) { fetch(&aicp_b200_get_output_reading, out_read_cloud); }

  virtual void updateConfigParams(std::string config_name) {
    params_.pointmatcher.configFileName = config_name;
    if (h_) aicp_b200_set_config(h_, config_name.c_str());
  }

  // Independent pairs at once (bash/run_registration_validation.sh runs the pairwise tool once per pair): pair i is registered
  // on GPU devices[i % devices.size()] (empty: the handle's own device), `streams` concurrent registrations per GPU, every
  // pair with its own trimmed ratio (auto-tuned from its overlap; empty: the configured ratio).  Returns the number of pairs
  // that failed; their transform stays the identity and status[i] holds the error code.
  int registerBatch(const std::vector<pcl::PointCloud<pcl::PointXYZ>*>& refs, const std::vector<pcl::PointCloud<pcl::PointXYZ>*>& reads,
                    const std::vector<float>& ratios, const std::vector<int>& devices, int streams,
                    std::vector<Eigen::Matrix4f>& transforms, std::vector<int>* status = nullptr) {
    const size_t n = refs.size();
    transforms.assign(n, Eigen::Matrix4f::Identity());
    if (status) status->assign(n, AICP_B200_ERR_BAD_ARG);
    if (!h_ || reads.size() != n || (!ratios.empty() && ratios.size() != n)) return (int)n;
    std::vector<std::vector<float> > stage(2 * n);
    std::vector<const float*> pr(n), pq(n);
    std::vector<int64_t> nr(n), nq(n);
    for (size_t i = 0; i < n; ++i) {
      pr[i] = b200_detail::as_xyzw(*refs[i], stage[2 * i], &nr[i]);
      pq[i] = b200_detail::as_xyzw(*reads[i], stage[2 * i + 1], &nq[i]);
    }
    std::vector<float> T(16 * n);
    std::vector<int32_t> st(n, 0), dev(devices.begin(), devices.end());
    const float* rat = ratios.empty() ? nullptr : ratios.data();
    int rc;
    if (dev.empty()) rc = aicp_b200_register_batch(h_, (int64_t)n, pr.data(), nr.data(), pq.data(), nq.data(), rat, streams, T.data(), nullptr, st.data(), nullptr);
    else rc = aicp_b200_register_batch_devices(h_, dev.data(), (int32_t)dev.size(), (int64_t)n, pr.data(), nr.data(), pq.data(), nq.data(), rat, streams,
                                               T.data(), nullptr, st.data(), nullptr);
    if (rc != AICP_B200_OK) std::cerr << "[B200] registerBatch (" << rc << "): " << aicp_b200_last_error(h_) << std::endl;
    int failed = 0;
    for (size_t i = 0; i < n; ++i) {
      if (status) (*status)[i] = st[i];
      if (st[i] == AICP_B200_OK) std::memcpy(transforms[i].data(), T.data() + 16 * i, 16 * sizeof(float));
      else ++failed;
    }
    return failed;
  }

  // conveniences beyond the reference interface
  const Eigen::Matrix4f& getOutputTransform() const { return last_T_; }
  float getWeightedPointUsedRatio() const { return stats_.weighted_point_used_ratio; }
  const aicp_b200_stats& getStats() const { return stats_; }
  aicp_b200_handle* handle() { return h_; }

 private:
  void open() {
    last_T_ = Eigen::Matrix4f::Identity();
    std::memset(&stats_, 0, sizeof(stats_));
    const std::string& cfg = params_.pointmatcher.configFileName;
    if (aicp_b200_create(cfg.empty() ? nullptr : cfg.c_str(), -1, &h_) != AICP_B200_OK) {
      std::cerr << "[B200] " << aicp_b200_last_error(nullptr) << std::endl;
      h_ = nullptr;
    }
  }

  // "x,y,theta_deg" -> column-major 4x4, parseTransformationDeg (cloudIO.cpp:261-302) + the rigidity check of
  // applyInitialization (pointmatcher_registration.cpp:71-89)
  bool initialTransform(float* T) const {
    std::string s = params_.pointmatcher.initialTransform;
    if (s.empty()) return false;
    for (char& c : s) if (c == '[' || c == ']' || c == ',' || c == ';') c = ' ';
    float v[3];
    if (std::sscanf(s.c_str(), "%f %f %f", &v[0], &v[1], &v[2]) != 3) {
      std::cerr << "[Cloud IO] An error occured while trying to parse the initial transformation." << std::endl
                << "No initial transformation will be used" << std::endl;
      return false;
    }
    const double th = v[2] * 3.14159265358979323846 / 180.0;
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
    T[0] = (float)std::cos(th); T[4] = (float)-std::sin(th);
    T[1] = (float)std::sin(th); T[5] = (float)std::cos(th);
    T[12] = v[0]; T[13] = v[1];
    std::cout << "[B200] Initialization: " << params_.pointmatcher.initialTransform << std::endl;
    return true;
  }

  template <typename PointT>
  void registerAny(pcl::PointCloud<PointT>& cloud_ref, pcl::PointCloud<PointT>& cloud_read, Eigen::Matrix4f& final_transform) {
    final_transform = Eigen::Matrix4f::Identity();
    if (!h_) { std::cerr << "[B200] no GPU handle; registration skipped." << std::endl; return; }
    int64_t n_ref = 0, n_read = 0;
    const float* ref = b200_detail::as_xyzw(cloud_ref, stage_ref_, &n_ref);
    const float* read = b200_detail::as_xyzw(cloud_read, stage_read_, &n_read);
    float init[16];
    const bool have_init = initialTransform(init);
    float T[16];
    const int rc = aicp_b200_register(h_, ref, n_ref, read, n_read, have_init ? init : nullptr, T, &stats_);
    if (rc != AICP_B200_OK) {
      std::cerr << "[B200] registerClouds failed (" << rc << "): " << aicp_b200_last_error(h_) << std::endl;
      return;
    }
    n_read_ = n_read;
    std::memcpy(final_transform.data(), T, sizeof(T));       // column-major, like Eigen::Matrix4f
    last_T_ = final_transform;
    std::cout << "[B200] Accepted matches (inliers): " << stats_.weighted_point_used_ratio * 100 << " %" << std::endl;   // :114
  }

  void fetch(int (*getter)(aicp_b200_handle*, float*, int64_t), pcl::PointCloud<pcl::PointXYZ>& out) {
    std::vector<float> xyzw(4 * (size_t)n_read_);
    if (!h_ || n_read_ == 0 || getter(h_, xyzw.data(), n_read_) != AICP_B200_OK) xyzw.clear();
    b200_detail::to_pcl(xyzw, out);
  }

  RegistrationParams params_;
  aicp_b200_handle* h_;
  aicp_b200_stats stats_;
  Eigen::Matrix4f last_T_;
  int64_t n_read_;
  std::vector<float> stage_ref_, stage_read_;
};

class B200Overlap : public AbstractOverlapper {
 public:
  explicit B200Overlap(const OverlapParams& params) : params_(params), h_(nullptr), overlap_(-1.0f) {
    counts_[0] = counts_[1] = counts_[2] = 0;
    // the reference hands ColorOcTree pointers back for visualisation only (App ignores them, app.cpp:132-135); the GPU
    // path does not materialise trees, so callers get an empty tree of the right resolution
    tree_ = new octomap::ColorOcTree(params_.octree_based.octomapResolution);
    if (aicp_b200_create(nullptr, -1, &h_) != AICP_B200_OK) {
      std::cerr << "[B200] " << aicp_b200_last_error(nullptr) << std::endl;
      h_ = nullptr;
    }
  }
  ~B200Overlap() { if (h_) aicp_b200_destroy(h_); delete tree_; }
  B200Overlap(const B200Overlap&) = delete;
  B200Overlap& operator=(const B200Overlap&) = delete;

  virtual octomap::ColorOcTree* computeOverlap(pcl::PointCloud<pcl::PointXYZ>& ref_cloud, pcl::PointCloud<pcl::PointXYZ>& read_cloud,
                                               Eigen::Isometry3d ref_pose, Eigen::Isometry3d read_pose,
                                               octomap::ColorOcTree* /*reading_tree*/) {
    overlap_ = -1.0f;
    if (!h_) return tree_;
    // convertPointCloudToScanGraph iterates cloud.points (octrees_overlap.cpp:232-236); only the pose translation is used
    const double ro[3] = {ref_pose.translation().x(), ref_pose.translation().y(), ref_pose.translation().z()};
    const double so[3] = {read_pose.translation().x(), read_pose.translation().y(), read_pose.translation().z()};
    const int rc = aicp_b200_overlap(h_, reinterpret_cast<const float*>(ref_cloud.points.data()), (int64_t)ref_cloud.points.size(), ro,
                                     reinterpret_cast<const float*>(read_cloud.points.data()), (int64_t)read_cloud.points.size(), so,
                                     params_.octree_based.octomapResolution, &overlap_, counts_);
    if (rc != AICP_B200_OK) {
      std::cerr << "[B200] computeOverlap failed (" << rc << "): " << aicp_b200_last_error(h_) << std::endl;
      overlap_ = -1.0f;
    }
    return tree_;
  }
  virtual float getOverlap() { return overlap_; }
  const int64_t* getCounts() const { return counts_; }      // {overlapping, reference, reading} nodes

 private:
  OverlapParams params_;
  aicp_b200_handle* h_;
  octomap::ColorOcTree* tree_;
  float overlap_;
  int64_t counts_[3];
};

// ---- map handling: stands where getPointsInOrientedBox stands (aicp_core/src/utils/filteringUtils.cpp:621-637) -----------
// App crops the prior / built map around the prior pose before every registration against it (app.cpp:41-69).  Same
// signature as the reference's free function plus the handle; `cloud` is replaced by the points inside the box, in input
// order.  The Euler angles are taken with Eigen exactly as the reference does, so this overload needs the real
// Eigen/Geometry (it is compiled only when that header has been included); rpy-taking overload below has no such need.
inline bool getPointsInOrientedBoxB200(aicp_b200_handle* h, pcl::PointCloud<pcl::PointXYZ>& cloud, float min, float max,
                                       const float rpy[3], const float position[3]) {
  const int64_t n = (int64_t)cloud.points.size();
  std::vector<float> out(4 * (size_t)(n > 0 ? n : 1));
  int64_t kept = 0;
  const int rc = aicp_b200_crop_box(h, reinterpret_cast<const float*>(cloud.points.data()), n, min, max, rpy, position, out.data(), &kept);
  if (rc != AICP_B200_OK) {
    std::cerr << "[B200] getPointsInOrientedBox failed (" << rc << "): " << aicp_b200_last_error(h) << std::endl;
    return false;
  }
  cloud.points.resize((size_t)kept);
  if (kept > 0) std::memcpy(cloud.points.data(), out.data(), sizeof(float) * 4 * (size_t)kept);
  cloud.width = (uint32_t)kept;
  cloud.height = 1;
  return true;
}

#ifdef EIGEN_GEOMETRY_MODULE_H
inline bool getPointsInOrientedBoxB200(aicp_b200_handle* h, pcl::PointCloud<pcl::PointXYZ>::Ptr& cloud, float min, float max,
                                       Eigen::Matrix4f& origin) {
  Eigen::Vector3f orientation = origin.block<3, 3>(0, 0).eulerAngles(0, 1, 2);      // (rx,ry,rz), filteringUtils.cpp:629
  const float rpy[3] = {orientation(0), orientation(1), orientation(2)};
  const float position[3] = {origin(0, 3), origin(1, 3), origin(2, 3)};
  return getPointsInOrientedBoxB200(h, *cloud, min, max, rpy, position);
}
#endif

// ---- pre-filter: stands where regionGrowingUniformPlaneSegmentationFilter stands (filteringUtils.cpp:5-45) -----------------
// Same signature as the reference's free function plus the handle.  Like the reference, the kept clusters are APPENDED to
// *cloud_out ("*cloud_out = *cloud_out + cloud_cluster", :43).  On failure cloud_out is left untouched and false is returned
// (the reference has no failure path here).
inline bool regionGrowingUniformPlaneSegmentationFilterB200(aicp_b200_handle* h, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_in,
                                                            pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_out) {
  const int64_t n = (int64_t)cloud_in->points.size();
  std::vector<float> out(4 * (size_t)(n > 0 ? n : 1));
  int64_t kept = 0;
  const int rc = aicp_b200_prefilter(h, reinterpret_cast<const float*>(cloud_in->points.data()), n, nullptr, nullptr, out.data(), &kept, nullptr);
  if (rc != AICP_B200_OK) {
    std::cerr << "[B200] regionGrowingUniformPlaneSegmentationFilter failed (" << rc << "): " << aicp_b200_last_error(h) << std::endl;
    return false;
  }
  const size_t old = cloud_out->points.size();
  cloud_out->points.resize(old + (size_t)kept);
  if (kept > 0) std::memcpy(reinterpret_cast<float*>(cloud_out->points.data()) + 4 * old, out.data(), sizeof(float) * 4 * (size_t)kept);
  cloud_out->width = (uint32_t)cloud_out->points.size();
  cloud_out->height = 1;
  return true;
}

// Second overload (filteringUtils.cpp:51-104) without the PCL point types it returns: the sampled cloud, its normals
// (nx, ny, nz, curvature) flipped towards view_point, and the clusters as index lists into the sampled cloud.
inline bool regionGrowingUniformPlaneSegmentationFilterB200(aicp_b200_handle* h, pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_in,
                                                            const float view_point[3], std::vector<float>& sampled_xyzw,
                                                            std::vector<float>& normals_xyzc, std::vector<std::vector<int> >& clusters) {
  const int64_t n = (int64_t)cloud_in->points.size();
  int64_t kept = 0;
  aicp_b200_prefilter_info info;
  int rc = aicp_b200_prefilter(h, reinterpret_cast<const float*>(cloud_in->points.data()), n, nullptr, view_point, nullptr, &kept, &info);
  std::vector<int32_t> labels((size_t)info.n_sampled);
  sampled_xyzw.assign(4 * (size_t)info.n_sampled, 0.f);
  normals_xyzc.assign(4 * (size_t)info.n_sampled, 0.f);
  clusters.assign((size_t)info.n_clusters, std::vector<int>());
  if (rc == AICP_B200_OK && info.n_sampled > 0) {
    rc = aicp_b200_prefilter_get_sampled(h, sampled_xyzw.data(), info.n_sampled);
    if (rc == AICP_B200_OK) rc = aicp_b200_prefilter_get_labels(h, labels.data(), info.n_sampled);
    if (rc == AICP_B200_OK && info.n_clusters > 0) rc = aicp_b200_prefilter_get_normals(h, normals_xyzc.data(), info.n_sampled);
  }
  if (rc != AICP_B200_OK) {
    std::cerr << "[B200] regionGrowingUniformPlaneSegmentationFilter failed (" << rc << "): " << aicp_b200_last_error(h) << std::endl;
    return false;
  }
  for (int64_t i = 0; i < info.n_sampled; ++i)
    if (labels[(size_t)i] >= 0) clusters[(size_t)labels[(size_t)i]].push_back((int)i);
  return true;
}

// ---- FOV overlap filter: stands where overlapFilter stands (filteringUtils.cpp:111-193; caller App::computeAlignmentRisk,
// app.cpp:153-156).  Same signature plus the handle: the points of each cloud that the OTHER sensor could have seen (range
// and angular field of view) are appended to accepted_pointsA / accepted_pointsB, the return value is the overlap in
// percent.  -1 on failure (the reference has no failure path).
inline float overlapFilterB200(aicp_b200_handle* h, pcl::PointCloud<pcl::PointXYZ>& cloudA, pcl::PointCloud<pcl::PointXYZ>& cloudB,
                               Eigen::Isometry3d poseA, Eigen::Isometry3d poseB, float range, float angularView,
                               pcl::PointCloud<pcl::PointXYZ>& accepted_pointsA, pcl::PointCloud<pcl::PointXYZ>& accepted_pointsB) {
  const int64_t nA = (int64_t)cloudA.points.size(), nB = (int64_t)cloudB.points.size();
  std::vector<float> outA(4 * (size_t)(nA > 0 ? nA : 1)), outB(4 * (size_t)(nB > 0 ? nB : 1));
  int64_t counts[2] = {0, 0};
  float overlap = -1.f;
  const int rc = aicp_b200_fov_overlap(h, reinterpret_cast<const float*>(cloudA.points.data()), nA, reinterpret_cast<const float*>(cloudB.points.data()), nB,
                                       poseA.matrix().data(), poseB.matrix().data(), range, angularView, outA.data(), outB.data(), counts, &overlap);
  if (rc != AICP_B200_OK) {
    std::cerr << "[B200] overlapFilter failed (" << rc << "): " << aicp_b200_last_error(h) << std::endl;
    return -1.f;
  }
  pcl::PointCloud<pcl::PointXYZ>* dst[2] = {&accepted_pointsA, &accepted_pointsB};
  const std::vector<float>* src[2] = {&outA, &outB};
  for (int c = 0; c < 2; ++c) {
    const size_t old = dst[c]->points.size();
    dst[c]->points.resize(old + (size_t)counts[c]);
    if (counts[c] > 0) std::memcpy(reinterpret_cast<float*>(dst[c]->points.data()) + 4 * old, src[c]->data(), sizeof(float) * 4 * (size_t)counts[c]);
    dst[c]->width = (uint32_t)dst[c]->points.size();
    dst[c]->height = 1;
  }
  return overlap;
}

// ---- alignability: stands where alignabilityFilter stands (filteringUtils.cpp:196-430; caller app.cpp:165-167).  Same
// signature plus the handle.  The three PointXYZRGBNormal clouds of the reference (matched planes of both clouds and the
// eigenvectors) only feed its visualiser (aicp_ros/src/visualizer_ros.cpp) and are left as they come in; the matching
// itself is available through the optional `matching` (for every kept plane of B the index of its plane of A, or -1).
inline float alignabilityFilterB200(aicp_b200_handle* h, pcl::PointCloud<pcl::PointXYZ>& cloudA, pcl::PointCloud<pcl::PointXYZ>& cloudB,
                                    Eigen::Isometry3d poseA, Eigen::Isometry3d poseB,
                                    pcl::PointCloud<pcl::PointXYZRGBNormal>::Ptr /*cloudA_planes*/,
                                    pcl::PointCloud<pcl::PointXYZRGBNormal>::Ptr /*cloudB_planes*/,
                                    pcl::PointCloud<pcl::PointXYZRGBNormal>::Ptr /*eigenvectors*/, std::vector<int32_t>* matching = nullptr) {
  const int64_t nA = (int64_t)cloudA.points.size(), nB = (int64_t)cloudB.points.size();
  float alignability = -1.f;                                                    // the reference's initial value (:201)
  int64_t info[3] = {0, 0, 0};
  std::vector<int32_t> m((size_t)(nB > 0 ? nB : 1), -1);
  const int rc = aicp_b200_alignability(h, reinterpret_cast<const float*>(cloudA.points.data()), nA, reinterpret_cast<const float*>(cloudB.points.data()), nB,
                                        poseA.matrix().data(), poseB.matrix().data(), nullptr, &alignability, m.data(), (int64_t)m.size(), info);
  if (rc != AICP_B200_OK) {
    std::cerr << "[B200] alignabilityFilter failed (" << rc << "): " << aicp_b200_last_error(h) << std::endl;
    return -1.f;
  }
  if (matching) matching->assign(m.begin(), m.begin() + (size_t)info[1]);
  return alignability;
}

// ---- classifier: stands where aicp::SVM stands (aicp_core/include/aicp_classification/svm.hpp:17-34) --------------------------
// Compiled only when aicp_classification/abstract_classification.hpp has been included before this header.  The factory branch:
//   } else if (parameters.type == "B200") { classifier.reset(new B200SVM(parameters)); }      (classification.hpp:12-16)
#ifdef AICP_CLASSIFICATION_ABSTRACT_HPP_
class B200SVM : public AbstractClassification {
 public:
  explicit B200SVM(const ClassificationParams& params) : params_(params), handle_(nullptr) {
    if (aicp_b200_create(nullptr, -1, &handle_) != AICP_B200_OK)
      std::cerr << "[B200] cannot create the classifier handle: " << aicp_b200_last_error(nullptr) << std::endl;
  }
  ~B200SVM() { if (handle_) aicp_b200_destroy(handle_); }

  // cv::ml::SVM::trainAuto (svm.cpp:18-51) is the reference's offline tool; models it wrote are loaded with load()
  virtual void train(const Eigen::MatrixXd&, const Eigen::MatrixXd&) {
    std::cerr << "[B200] SVM::train is not provided: train with the reference's OpenCV tool and load() the saved model." << std::endl;
  }
  virtual void test(const Eigen::MatrixXd& testing_data, Eigen::MatrixXd* probabilities) {       // svm.cpp:46-51
    Eigen::MatrixXd empty_labels = Eigen::MatrixXd::Zero(testing_data.rows(), 1);
    test(testing_data, empty_labels, probabilities);
  }
  virtual void test(const Eigen::MatrixXd& testing_data, const Eigen::MatrixXd& labels, Eigen::MatrixXd* probabilities = NULL) {
    if (params_.svm.saveFile.compare("") != 0) load(params_.svm.saveFile);                         // svm.cpp:61-63
    const int64_t n = (int64_t)testing_data.rows();
    const int32_t dim = (int32_t)testing_data.cols();
    if (probabilities != NULL) probabilities->resize(n, 1);
    if (n == 0 || !handle_) return;
    std::vector<double> x((size_t)(n * dim)), p((size_t)n, 0.0);
    for (int64_t i = 0; i < n; ++i) for (int32_t j = 0; j < dim; ++j) x[(size_t)(i * dim + j)] = testing_data(i, j);   // svm.cpp:72-74
    const int rc = aicp_b200_svm_predict(handle_, x.data(), n, dim, p.data(), nullptr);
    if (rc != AICP_B200_OK) { std::cerr << "[B200] SVM::test failed (" << rc << "): " << aicp_b200_last_error(handle_) << std::endl; return; }
    unsigned int tp = 0u, fp = 0u, tn = 0u, fn = 0u;
    const bool have_labels = !labels.isZero();
    for (int64_t i = 0; i < n; ++i) {                                                               // svm.cpp:84-101
      if (have_labels && p[(size_t)i] >= params_.svm.threshold) { if (labels(i, 0) == 1.0) ++tp; else ++fp; }
      else { if (labels(i, 0) == 0.0) ++tn; else ++fn; }
      if (probabilities != NULL) (*probabilities)(i, 0) = p[(size_t)i];
    }
    if (have_labels && probabilities != NULL && probabilities->rows() > 1) confusionMatrix(tp, tn, fp, fn);
  }
  virtual void save(const std::string&) {
    std::cerr << "[B200] SVM::save is not provided (models are written by the reference's training tool)." << std::endl;
  }
  virtual void load(const std::string& filename) {                                                  // svm.cpp:103-107
    if (!handle_) return;
    const int rc = aicp_b200_svm_load(handle_, filename.c_str());
    if (rc != AICP_B200_OK) std::cerr << "[B200] SVM::load failed (" << rc << "): " << aicp_b200_last_error(handle_) << std::endl;
  }

 private:
  ClassificationParams params_;
  aicp_b200_handle* handle_;
};
#endif

}  // namespace aicp

#endif
