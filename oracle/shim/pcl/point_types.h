// TEST SCAFFOLDING (oracle build only): pcl::PointXYZI / pcl::PointCloud, as far as
// TSDFVolume::SaveWorld (Integrate/TSDFVolume.cpp:104-132) needs them.
#pragma once
#include <memory>
#include <vector>
namespace pcl {
struct PointXYZI {
  float x, y, z, intensity;
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
