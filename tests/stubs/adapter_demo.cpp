// adapter_demo.cpp -- compiles the B200 adapters against the reference's OWN abstract plug-in headers
// (aicp_core/include/aicp_registration/abstract_registrator.hpp, aicp_overlap/abstract_overlapper.hpp) with stand-in
// PCL/Eigen/octomap headers, then drives them the way App::runAicpPipeline does (app.cpp:218-247):
//   overlap -> clamp to ratio -> rewrite YAML (done by the caller here) -> updateConfigParams -> registerClouds.
// usage: adapter_demo <icp_yaml> <n_points> [svm_model.xml]      prints "OK overlap=<pct> iterations=<n> T=<16 floats>"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "aicp_classification/abstract_classification.hpp"      // the reference's own header: enables B200SVM in the adapter
#include "aicp_b200_adapter.hpp"

// the two factory branches a maintainer adds to registration.hpp:11-17 / overlap.hpp:11
static std::unique_ptr<aicp::AbstractRegistrator> create_registrator(const RegistrationParams& p) {
  std::unique_ptr<aicp::AbstractRegistrator> r;
  if (p.type == "B200") r.reset(new aicp::B200Registration(p));
  else std::cerr << "Invalid registration type " << p.type << "." << std::endl;
  return r;
}
static std::unique_ptr<aicp::AbstractOverlapper> create_overlapper(const OverlapParams& p) {
  std::unique_ptr<aicp::AbstractOverlapper> o;
  if (p.type == "B200") o.reset(new aicp::B200Overlap(p));
  return o;
}

// classification.hpp:8-19 with the extra branch
static std::unique_ptr<aicp::AbstractClassification> create_classifier(const ClassificationParams& p) {
  std::unique_ptr<aicp::AbstractClassification> c;
  if (p.type == "B200") c.reset(new aicp::B200SVM(p));
  return c;
}

int main(int argc, char** argv) {
  const char* yaml = argc > 1 ? argv[1] : "";
  int n_side = argc > 2 ? std::atoi(argv[2]) : 40;
  // cube cloud like aicp_core/src/tools/create_cube_cloud.cpp, reading = reference shifted by a small rigid motion
  pcl::PointCloud<pcl::PointXYZ> ref, read;
  const float step = 4.0f / n_side;
  for (int f = 0; f < 6; ++f)
    for (int i = 0; i < n_side; ++i)
      for (int j = 0; j < n_side; ++j) {
        float a = -2.f + step * i, b = -2.f + step * j, c = (f & 1) ? 2.f : -2.f;
        pcl::PointXYZ p;
        if (f < 2) { p.x = a; p.y = b; p.z = c; } else if (f < 4) { p.x = c; p.y = a; p.z = b; } else { p.x = a; p.y = c; p.z = b; }
        p.pad = 1.f;
        ref.points.push_back(p);
        const float yaw = 0.02f, tx = 0.05f, ty = -0.03f;
        pcl::PointXYZ q;
        q.x = std::cos(yaw) * p.x - std::sin(yaw) * p.y + tx + 0.001f * ((i * 7 + j * 13) % 5);
        q.y = std::sin(yaw) * p.x + std::cos(yaw) * p.y + ty;
        q.z = p.z + 0.0007f * ((i + 3 * j) % 3);
        q.pad = 1.f;
        read.points.push_back(q);
      }
  ref.width = (uint32_t)ref.points.size(); ref.height = 1;
  read.width = (uint32_t)read.points.size(); read.height = 1;

  RegistrationParams rp; rp.type = "B200"; rp.pointmatcher.configFileName = yaml;
  OverlapParams op; op.type = "B200"; op.octree_based.octomapResolution = 0.2f;   // yaml_configurator.cpp:80-82
  auto registr = create_registrator(rp);
  auto overlapper = create_overlapper(op);
  if (!registr || !overlapper) return 2;

  octomap::ColorOcTree read_tree(op.octree_based.octomapResolution);
  overlapper->computeOverlap(ref, read, Eigen::Isometry3d::Identity(), Eigen::Isometry3d::Identity(), &read_tree);
  float overlap = overlapper->getOverlap();
  if (overlap < 0) return 3;

  registr->updateConfigParams(yaml);
  Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
  registr->registerClouds(ref, read, T);
  pcl::PointCloud<pcl::PointXYZ> out;
  registr->getOutputReading(out);
  if (out.points.size() != read.points.size()) return 4;
  // the pre-filter as App::filterCloud calls it (app.cpp:102-110): free function + the library handle
  {
    aicp_b200_handle* fh = nullptr;
    if (aicp_b200_create(nullptr, -1, &fh) != AICP_B200_OK) return 6;
    pcl::PointCloud<pcl::PointXYZ>::Ptr in(new pcl::PointCloud<pcl::PointXYZ>(ref)), filtered(new pcl::PointCloud<pcl::PointXYZ>);
    const bool ok = aicp::regionGrowingUniformPlaneSegmentationFilterB200(fh, in, filtered);
    aicp_b200_destroy(fh);
    if (!ok || filtered->points.empty() || filtered->points.size() > in->points.size()) return 7;
    std::printf("prefilter %zu -> %zu points\n", in->points.size(), filtered->points.size());
  }
  // the alignment-risk inputs as App::computeAlignmentRisk computes them (app.cpp:153-167): FOV overlap, then alignability of the
  // accepted points; and the batched entry point over every GPU of the box
  {
    aicp_b200_handle* fh = nullptr;
    if (aicp_b200_create(nullptr, -1, &fh) != AICP_B200_OK) return 6;
    pcl::PointCloud<pcl::PointXYZ> accA, accB;
    Eigen::Isometry3d poseA = Eigen::Isometry3d::Identity(), poseB = Eigen::Isometry3d::Identity();
    const float fov = aicp::overlapFilterB200(fh, ref, read, poseA, poseB, 30.f, 270.f, accA, accB);
    if (fov < 0.f || accA.points.empty() || accA.points.size() > ref.points.size()) { aicp_b200_destroy(fh); return 10; }
    pcl::PointCloud<pcl::PointXYZRGBNormal>::Ptr none;
    std::vector<int32_t> matching;
    const float al = aicp::alignabilityFilterB200(fh, accA, accB, poseA, poseB, none, none, none, &matching);
    aicp_b200_destroy(fh);
    if (al < 0.f) return 11;
    std::printf("fov_overlap %.4f alignability %.4f planes_B %zu\n", fov, al, matching.size());
    auto* breg = static_cast<aicp::B200Registration*>(registr.get());
    std::vector<pcl::PointCloud<pcl::PointXYZ>*> refs(3, &ref), reads(3, &read);
    std::vector<Eigen::Matrix4f> Ts;
    std::vector<int> st;
    if (breg->registerBatch(refs, reads, std::vector<float>(), std::vector<int>(1, 0), 2, Ts, &st) != 0) return 12;
    for (int i = 0; i < 16; ++i) if (Ts[2].data()[i] != T.data()[i]) return 13;          // same pair, same bits as registerClouds
  }
  // the classifier as App::computeAlignmentRisk calls it (app.cpp:175-181): testing_data << overlap, alignability
  if (argc > 3) {
    ClassificationParams cp; cp.type = "B200"; cp.svm.threshold = 0.5; cp.svm.saveFile = argv[3];
    auto classifier = create_classifier(cp);
    if (!classifier) return 8;
    Eigen::MatrixXd testing_data(1, 2), risk;
    testing_data(0, 0) = 61.63; testing_data(0, 1) = 50.02;
    classifier->test(testing_data, &risk);
    if (risk.rows() != 1) return 9;
    std::printf("risk %.9f\n", risk(0, 0));
  }
  auto* b = static_cast<aicp::B200Registration*>(registr.get());
  std::printf("OK overlap=%.4f iterations=%d T=", overlap, b->getStats().iterations);
  for (int i = 0; i < 16; ++i) std::printf("%.9g ", T.data()[i]);
  std::printf("\n");
  return b->getStats().iterations > 0 ? 0 : 5;
}
