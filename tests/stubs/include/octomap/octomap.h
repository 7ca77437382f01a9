// Minimal stand-in for octomap (TEST INFRASTRUCTURE ONLY): just enough for aicp_overlap/abstract_overlapper.hpp.
#pragma once
#include <cstddef>
namespace octomap {
class ColorOcTree {
 public:
  explicit ColorOcTree(double resolution) : resolution_(resolution) {}
  size_t size() const { return 0; }
  double getResolution() const { return resolution_; }
 private:
  double resolution_;
};
}  // namespace octomap
