// er_io_tool -- CPU-only helper around apps/er_io.h, used by the format tests (no GPU, no library):
//   er_io_tool log2log  in.log out.log        load a trajectory / pair list and write it back
//   er_io_tool info2info in.info out.info
//   er_io_tool pcd2xyzn in.pcd out.xyzn       load a fragment like BuildCorrespondence does (NaN normals dropped)
//   er_io_tool ctr2txt  in.ctr num res out    load control lattices and dump them as "%.9g %.9g %.9g"
//   er_io_tool camera   in.param              print the six camera parameters
#include <cstdio>
#include <string>

#include "er_io.h"

int main(int argc, char** argv) {
  using namespace er_io;
  if (argc < 3) return 2;
  const std::string cmd = argv[1];
  if (cmd == "log2log" && argc == 4) {
    Trajectory t;
    t.load(argv[2]);
    std::printf("%zu\n", t.data.size());
    return t.save(argv[3]) ? 0 : 1;
  }
  if (cmd == "info2info" && argc == 4) {
    Information t;
    t.load(argv[2]);
    std::printf("%zu\n", t.data.size());
    return t.save(argv[3]) ? 0 : 1;
  }
  if (cmd == "pcd2xyzn" && argc == 4) {
    PointCloudXYZN c;
    if (!load_pcd_xyzn(argv[2], c)) return 1;
    FILE* f = std::fopen(argv[3], "w");
    if (!f) return 1;
    size_t kept = 0;
    for (size_t i = 0; i < c.size(); i++) {
      if (std::isnan(c.nrm[3 * i])) continue;
      std::fprintf(f, "%.9g %.9g %.9g %.9g %.9g %.9g\n", c.xyz[3 * i], c.xyz[3 * i + 1], c.xyz[3 * i + 2], c.nrm[3 * i],
                   c.nrm[3 * i + 1], c.nrm[3 * i + 2]);
      kept++;
    }
    std::fclose(f);
    std::printf("%zu %zu\n", c.size(), kept);
    return 0;
  }
  if (cmd == "ctr2txt" && argc == 6) {
    std::vector<float> g;
    if (!load_ctr(argv[2], std::atoi(argv[3]), std::atoi(argv[4]), g)) return 1;
    FILE* f = std::fopen(argv[5], "w");
    if (!f) return 1;
    for (size_t i = 0; i < g.size() / 3; i++) std::fprintf(f, "%.9g %.9g %.9g\n", g[3 * i], g[3 * i + 1], g[3 * i + 2]);
    std::fclose(f);
    return 0;
  }
  if (cmd == "camera" && argc == 3) {
    CameraParam c;
    c.load(argv[2]);
    std::printf("%.9g %.9g %.9g %.9g %.9g %.9g\n", c.fx, c.fy, c.cx, c.cy, c.icp_trunc, c.integration_trunc);
    return 0;
  }
  return 2;
}
