// TEST SCAFFOLDING (oracle build only) -- not product code.
// pcl::PointXYZRGBNormal / pcl::PointCloud as far as FragmentOptimizer/PointCloud.cpp:21-44 and
// OptApp.cpp:902-922 touch them (named members only; no PCL memory layout is relied upon).
#pragma once
#include <memory>
#include <vector>
#include <cstddef>
namespace pcl {
struct PointXYZRGBNormal {
  float x, y, z, rgb, normal_x, normal_y, normal_z, curvature;
  PointXYZRGBNormal() : x(0), y(0), z(0), rgb(0), normal_x(0), normal_y(0), normal_z(0), curvature(0) {}
};
template <typename T>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<T> > Ptr;
  std::vector<T> points;
  void push_back(const T& p) { points.push_back(p); }
  size_t size() const { return points.size(); }
};
}  // namespace pcl
