// TEST SCAFFOLDING (oracle build only) -- not product code.
// pcl::io::loadPCDFile for PointXYZRGBNormal (PCD v0.7 "DATA ascii" / "DATA binary", fields found by name;
// binary_compressed is not supported by this stand-in and reports failure) and PCDWriter::writeBinaryCompressed
// (writes an uncompressed binary PCD: sample.pcd is a visualisation by-product, FragmentOptimizer/OptApp.cpp:897-923).
#pragma once
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <pcl/point_cloud.h>
#include <pcl/common/time.h>   // PCL_ERROR (real PCL pulls its print macros in through pcd_io.h too)
namespace pcl {
namespace io {
inline int loadPCDFile(const std::string& fn, PointCloud<PointXYZRGBNormal>& out) {
  std::FILE* f = std::fopen(fn.c_str(), "rb");
  if (!f) return -1;
  std::vector<std::string> fields;
  std::vector<int> sizes, counts;
  size_t npoints = 0;
  int mode = -1;   // 0 ascii, 1 binary
  char line[4096];
  while (std::fgets(line, sizeof line, f)) {
    char key[64] = {0};
    std::sscanf(line, "%63s", key);
    if (!std::strcmp(key, "FIELDS") || !std::strcmp(key, "SIZE") || !std::strcmp(key, "COUNT")) {
      char* save = 0;
      char* tok = strtok_r(line, " \t\r\n", &save);
      while ((tok = strtok_r(0, " \t\r\n", &save)) != 0) {
        if (key[0] == 'F') fields.push_back(tok);
        else if (key[0] == 'S') sizes.push_back(std::atoi(tok));
        else counts.push_back(std::atoi(tok));
      }
    } else if (!std::strcmp(key, "POINTS")) {
      npoints = (size_t)std::atoll(line + 6);
    } else if (!std::strcmp(key, "DATA")) {
      if (std::strstr(line, "ascii")) mode = 0;
      else if (std::strstr(line, "binary_compressed")) mode = -1;
      else mode = 1;
      break;
    }
  }
  if (mode < 0 || fields.empty() || sizes.size() != fields.size()) { std::fclose(f); return -1; }
  if (counts.size() != fields.size()) counts.assign(fields.size(), 1);
  size_t rec = 0;
  std::vector<size_t> off(fields.size());
  for (size_t k = 0; k < fields.size(); ++k) { off[k] = rec; rec += (size_t)sizes[k] * counts[k]; }
  const char* names[7] = {"x", "y", "z", "normal_x", "normal_y", "normal_z", "curvature"};
  int where[7];
  for (int t = 0; t < 7; ++t) {
    where[t] = -1;
    for (size_t k = 0; k < fields.size(); ++k) if (fields[k] == names[t]) where[t] = (int)k;
  }
  out.points.resize(npoints);
  std::vector<char> buf(rec);
  for (size_t p = 0; p < npoints; ++p) {
    float v[7] = {0, 0, 0, 0, 0, 0, 0};
    if (mode == 1) {
      if (std::fread(buf.data(), 1, rec, f) != rec) { std::fclose(f); return -1; }
      for (int t = 0; t < 7; ++t) if (where[t] >= 0) std::memcpy(&v[t], buf.data() + off[where[t]], 4);
    } else {
      if (!std::fgets(line, sizeof line, f)) { std::fclose(f); return -1; }
      char* save = 0;
      char* tok = strtok_r(line, " \t\r\n", &save);
      for (size_t k = 0; k < fields.size() && tok; ++k) {
        for (int c = 0; c < counts[k] && tok; ++c) {
          if (c == 0) for (int t = 0; t < 7; ++t) if (where[t] == (int)k) v[t] = std::strtof(tok, 0);
          tok = strtok_r(0, " \t\r\n", &save);
        }
      }
    }
    PointXYZRGBNormal& q = out.points[p];
    q.x = v[0]; q.y = v[1]; q.z = v[2]; q.normal_x = v[3]; q.normal_y = v[4]; q.normal_z = v[5]; q.curvature = v[6];
  }
  std::fclose(f);
  return 0;
}
}  // namespace io
class PCDWriter {
 public:
  int writeBinaryCompressed(const std::string& fn, const PointCloud<PointXYZRGBNormal>& c) {
    std::FILE* f = std::fopen(fn.c_str(), "wb");
    if (!f) return -1;
    std::fprintf(f,
                 "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb normal_x normal_y normal_z curvature\n"
                 "SIZE 4 4 4 4 4 4 4 4\nTYPE F F F F F F F F\nCOUNT 1 1 1 1 1 1 1 1\nWIDTH %zu\nHEIGHT 1\n"
                 "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA binary\n",
                 c.size(), c.size());
    if (c.size()) std::fwrite(c.points.data(), sizeof(PointXYZRGBNormal), c.size(), f);
    std::fclose(f);
    return 0;
  }
};
}  // namespace pcl
