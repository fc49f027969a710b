// TEST SCAFFOLDING (oracle build only): pcl::io::savePCDFile for PointXYZI,
// PCD v0.7 "DATA binary" with packed 16-byte records, the layout PCL writes.
#pragma once
#include <cstdio>
#include <string>
#include <pcl/point_types.h>
namespace pcl { namespace io {
inline int savePCDFile(const std::string& fn, const PointCloud<PointXYZI>& c, bool /*binary*/) {
  std::FILE* f = std::fopen(fn.c_str(), "wb");
  if (!f) return -1;
  std::fprintf(f,
               "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\n"
               "SIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH %zu\nHEIGHT 1\n"
               "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA binary\n",
               c.size(), c.size());
  if (c.size()) std::fwrite(c.points.data(), sizeof(PointXYZI), c.size(), f);
  std::fclose(f);
  return 0;
}
}}  // namespace pcl::io
