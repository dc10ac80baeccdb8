// oracle/ref_shim/pcl/stub.hpp — the two PCL names SampleSet2D (sw_manager.hpp:41-124) touches for its visual-only history
#pragma once
#include <vector>
namespace pcl {
struct PointXYZI { float x, y, z, intensity; };
template <class T> struct PointCloud { std::vector<T> points; };
}  // namespace pcl
