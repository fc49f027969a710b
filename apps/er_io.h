// er_io.h -- file formats at the drop-in boundary of the two executables.
//
// Every reader/writer states the reference code whose behaviour it reproduces (paths relative to the
// reference checkout).  Header-only, host C++17, no third-party dependencies (the reference used PCL/Boost).
#pragma once
#include <fcntl.h>
#include <unistd.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <sys/stat.h>

namespace er_io {

inline bool file_exists(const std::string& p) {
  struct stat st;
  return !p.empty() && ::stat(p.c_str(), &st) == 0;
}

// ------------------------------------------------------------------------------------------------
// `.log` trajectories / pair lists: RGBDTrajectory (Integrate/TSDFVolumeUnit.h:12-63,
// BuildCorrespondence/Helper.h:9-60).  A record is "id1 id2 frame" followed by four rows of four doubles;
// lines starting with '#' are skipped; missing files leave the trajectory empty (the reference does not
// report an error either).
// ------------------------------------------------------------------------------------------------
struct FramedTransformation {
  int id1 = 0, id2 = 0, frame = 0;
  double T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};   // row-major
};

struct Trajectory {
  std::vector<FramedTransformation> data;

  void load(const std::string& filename) {
    data.clear();
    FILE* f = std::fopen(filename.c_str(), "r");
    if (!f) return;
    char buf[1024];
    FramedTransformation ft;
    while (std::fgets(buf, sizeof buf, f)) {
      if (std::strlen(buf) > 0 && buf[0] != '#') {
        std::sscanf(buf, "%d %d %d", &ft.id1, &ft.id2, &ft.frame);
        for (int r = 0; r < 4; r++) {
          if (!std::fgets(buf, sizeof buf, f)) break;
          std::sscanf(buf, "%lf %lf %lf %lf", &ft.T[r * 4], &ft.T[r * 4 + 1], &ft.T[r * 4 + 2], &ft.T[r * 4 + 3]);
        }
        data.push_back(ft);
      }
    }
    std::fclose(f);
  }

  bool save(const std::string& filename) const {
    FILE* f = std::fopen(filename.c_str(), "w");
    if (!f) return false;
    for (const auto& d : data) {
      std::fprintf(f, "%d\t%d\t%d\n", d.id1, d.id2, d.frame);
      for (int r = 0; r < 4; r++)
        std::fprintf(f, "%.8f %.8f %.8f %.8f\n", d.T[r * 4], d.T[r * 4 + 1], d.T[r * 4 + 2], d.T[r * 4 + 3]);
    }
    std::fclose(f);
    return true;
  }
};

// `.info`: RGBDInformation (BuildCorrespondence/Helper.h:62-119): same header + 6 rows of 6 doubles.
struct FramedInformation {
  int id1 = 0, id2 = 0, frame = 0;
  double I[36] = {0};
};

struct Information {
  std::vector<FramedInformation> data;

  void load(const std::string& filename) {
    data.clear();
    FILE* f = std::fopen(filename.c_str(), "r");
    if (!f) return;
    char buf[1024];
    FramedInformation fi;
    while (std::fgets(buf, sizeof buf, f)) {
      if (std::strlen(buf) > 0 && buf[0] != '#') {
        std::sscanf(buf, "%d %d %d", &fi.id1, &fi.id2, &fi.frame);
        for (int r = 0; r < 6; r++) {
          if (!std::fgets(buf, sizeof buf, f)) break;
          double* p = &fi.I[r * 6];
          std::sscanf(buf, "%lf %lf %lf %lf %lf %lf", p, p + 1, p + 2, p + 3, p + 4, p + 5);
        }
        data.push_back(fi);
      }
    }
    std::fclose(f);
  }

  bool save(const std::string& filename) const {
    FILE* f = std::fopen(filename.c_str(), "w");
    if (!f) return false;
    for (const auto& d : data) {
      std::fprintf(f, "%d\t%d\t%d\n", d.id1, d.id2, d.frame);
      for (int r = 0; r < 6; r++) {
        const double* p = &d.I[r * 6];
        std::fprintf(f, "%.8f %.8f %.8f %.8f %.8f %.8f\n", p[0], p[1], p[2], p[3], p[4], p[5]);
      }
    }
    std::fclose(f);
    return true;
  }
};

// Camera file: CameraParam::LoadFromFile (Integrate/TSDFVolumeUnit.h:65-96): fx fy cx cy ICP_trunc
// integration_trunc, one per line, after optional '#' lines.
struct CameraParam {
  float fx = 525.0f, fy = 525.0f, cx = 319.5f, cy = 239.5f, icp_trunc = 2.5f, integration_trunc = 2.5f;

  void load(const std::string& filename) {
    FILE* f = std::fopen(filename.c_str(), "r");
    if (!f) return;
    char buf[1024];
    while (std::fgets(buf, sizeof buf, f)) {
      if (std::strlen(buf) > 0 && buf[0] != '#') {
        float* dst[6] = {&fx, &fy, &cx, &cy, &icp_trunc, &integration_trunc};
        std::sscanf(buf, "%f", dst[0]);
        for (int i = 1; i < 6; i++) {
          if (!std::fgets(buf, sizeof buf, f)) break;
          std::sscanf(buf, "%f", dst[i]);
        }
      }
    }
    std::fclose(f);
    std::fprintf(stderr, "Camera model set to (fx, fy, cx, cy, icp_trunc, int_trunc):\n\t%.2f, %.2f, %.2f, %.2f, %.2f, %.2f\n",
                 fx, fy, cx, cy, icp_trunc, integration_trunc);
  }
};

// `.ctr`: ControlGrid::Load for `num` fragments (Integrate/ControlGrid.cpp:15-34, IntegrateApp.cpp:49-55):
// num*(res+1)^3 lines "%f %f %f"; a '#' line still consumes a lattice slot (left at zero).
inline bool load_ctr(const std::string& filename, int num, int res, std::vector<float>& out) {
  const size_t per = (size_t)(res + 1) * (res + 1) * (res + 1);
  out.assign(per * num * 3, 0.0f);
  FILE* f = std::fopen(filename.c_str(), "r");
  if (!f) return false;
  char buf[1024];
  for (size_t i = 0; i < per * num; i++) {
    if (std::fgets(buf, sizeof buf, f)) {
      if (std::strlen(buf) > 0 && buf[0] != '#') std::sscanf(buf, "%f %f %f", &out[3 * i], &out[3 * i + 1], &out[3 * i + 2]);
    }
  }
  std::fclose(f);
  return true;
}

// Depth stream.  The reference reads OpenNI .oni files through pcl::ONIGrabber (Integrate/Integrate.cpp:51-54);
// OpenNI is not available here, so "-oni <file>" names a raw container instead: little-endian int32 n, w, h
// followed by n*w*h uint16 millimetres (0 = invalid); FrameID = 1, 2, ... (IntegrateApp.cpp:185).
struct DepthStream {
  int fd = -1;
  int n = 0, w = 0, h = 0, next = 0;
  bool open(const std::string& filename) {
    fd = ::open(filename.c_str(), O_RDONLY);
    if (fd < 0) return false;
    int32_t hdr[3];
    if (::pread(fd, hdr, sizeof hdr, 0) != (ssize_t)sizeof hdr || hdr[0] < 0 || hdr[1] <= 0 || hdr[2] <= 0) { ::close(fd); fd = -1; return false; }
    n = hdr[0]; w = hdr[1]; h = hdr[2];
    return true;
  }
  // reads up to `count` frames; returns how many were read.  The frames of one call are fetched by four threads with
  // pread (the page cache delivers ~7 GB/s per thread; a 32-frame chunk is 20 MB)
  int read(uint16_t* dst, int count) {
    if (fd < 0) return 0;
    const int want = std::min(count, n - next);
    if (want <= 0) return 0;
    const size_t fb = sizeof(uint16_t) * (size_t)w * h;
    int ok = want;
#pragma omp parallel for num_threads(4) schedule(static) reduction(min : ok)
    for (int i = 0; i < want; i++) {
      size_t got = 0;
      const off_t off = (off_t)(12 + fb * (size_t)(next + i));
      while (got < fb) {
        const ssize_t r = ::pread(fd, (char*)dst + fb * (size_t)i + got, fb - got, off + (off_t)got);
        if (r <= 0) break;
        got += (size_t)r;
      }
      if (got != fb) ok = std::min(ok, i);                // a short file: frames before the first short one count
    }
    if (ok < want) { next = n; return ok; }
    next += want;
    return want;
  }
  ~DepthStream() { if (fd >= 0) ::close(fd); }
};

// world.pcd: pcl::io::savePCDFile(filename, PointCloud<PointXYZI>, binary=true) (Integrate/TSDFVolume.cpp:130).
inline bool save_pcd_xyzi(const std::string& filename, const float* xyzi, size_t n) {
  FILE* f = std::fopen(filename.c_str(), "wb");
  if (!f) return false;
  std::fprintf(f,
               "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\n"
               "SIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH %zu\nHEIGHT 1\n"
               "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA binary\n", n, n);
  if (n) std::fwrite(xyzi, sizeof(float) * 4, n, f);
  std::fclose(f);
  return true;
}

// ------------------------------------------------------------------------------------------------
// PCD reader for the fragments (pcl::io::loadPCDFile into PointXYZRGBNormal, BuildCorrespondence/
// CorresApp.cpp:86-97).  Supports DATA ascii | binary | binary_compressed (format mirror:
// Matlab_Toolbox/External/matpcl/loadpcd.m:33-220, LZF: lzfd.m:33-74).  Only x y z normal_x normal_y normal_z
// are extracted; absent fields read as 0 (normals as NaN so that the reference's NaN filter drops the point).
// ------------------------------------------------------------------------------------------------
inline bool lzf_decompress(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
  size_t ip = 0, op = 0;
  while (ip < in_len) {
    unsigned ctrl = in[ip++];
    if (ctrl < 32) {                       // literal run of ctrl + 1 bytes
      const size_t len = ctrl + 1;
      if (op + len > out_len || ip + len > in_len) return false;
      std::memcpy(out + op, in + ip, len);
      op += len; ip += len;
    } else {                               // back reference
      size_t len = ctrl >> 5;
      if (len == 7) { if (ip >= in_len) return false; len += in[ip++]; }
      if (ip >= in_len) return false;
      const size_t off = ((size_t)(ctrl & 0x1f) << 8) + in[ip++] + 1;
      len += 2;
      if (off > op || op + len > out_len) return false;
      for (size_t i = 0; i < len; i++, op++) out[op] = out[op - off];   // may overlap
    }
  }
  return op == out_len;
}

struct PointCloudXYZN {
  std::vector<float> xyz, nrm;           // n*3 each
  size_t size() const { return xyz.size() / 3; }
};

inline bool load_pcd_xyzn(const std::string& filename, PointCloudXYZN& cloud) {
  cloud.xyz.clear(); cloud.nrm.clear();
  FILE* f = std::fopen(filename.c_str(), "rb");
  if (!f) return false;
  std::vector<std::string> fields, types;
  std::vector<int> sizes, counts;
  size_t points = 0, width = 0, height = 1;
  std::string data;
  char line[4096];
  while (std::fgets(line, sizeof line, f)) {
    if (line[0] == '#') continue;
    char key[64] = {0};
    if (std::sscanf(line, "%63s", key) != 1) continue;
    std::vector<std::string> tok;
    char* save = nullptr;      // strtok_r: fragments are loaded from several threads at once
    for (char* p = strtok_r(line, " \t\r\n", &save); p; p = strtok_r(nullptr, " \t\r\n", &save)) tok.push_back(p);
    const std::string k = tok[0];
    if (tok.size() < 2 && k != "DATA") continue;
    if (k == "FIELDS" || k == "COLUMNS") fields.assign(tok.begin() + 1, tok.end());
    else if (k == "SIZE") { sizes.clear(); for (size_t i = 1; i < tok.size(); i++) sizes.push_back(std::atoi(tok[i].c_str())); }
    else if (k == "TYPE") types.assign(tok.begin() + 1, tok.end());
    else if (k == "COUNT") { counts.clear(); for (size_t i = 1; i < tok.size(); i++) counts.push_back(std::atoi(tok[i].c_str())); }
    else if (k == "WIDTH") width = std::strtoull(tok[1].c_str(), nullptr, 10);
    else if (k == "HEIGHT") height = std::strtoull(tok[1].c_str(), nullptr, 10);
    else if (k == "POINTS") points = std::strtoull(tok[1].c_str(), nullptr, 10);
    else if (k == "DATA") { data = tok.size() > 1 ? tok[1] : ""; break; }
  }
  if (points == 0) points = width * height;
  const size_t nf = fields.size();
  if (nf == 0 || data.empty()) { std::fclose(f); return false; }
  if (sizes.size() != nf) sizes.assign(nf, 4);
  if (types.size() != nf) types.assign(nf, "F");
  if (counts.size() != nf) counts.assign(nf, 1);
  std::vector<size_t> offset(nf);
  size_t rec = 0;
  for (size_t i = 0; i < nf; i++) { offset[i] = rec; rec += (size_t)sizes[i] * counts[i]; }
  const char* want[6] = {"x", "y", "z", "normal_x", "normal_y", "normal_z"};
  int col[6];
  for (int q = 0; q < 6; q++) {
    col[q] = -1;
    for (size_t i = 0; i < nf; i++) if (fields[i] == want[q]) col[q] = (int)i;
  }
  auto as_float = [&](const uint8_t* p, int fi) -> float {
    const char t = types[fi][0];
    const int sz = sizes[fi];
    if (t == 'F' && sz == 4) { float v; std::memcpy(&v, p, 4); return v; }
    if (t == 'F' && sz == 8) { double v; std::memcpy(&v, p, 8); return (float)v; }
    if (t == 'U') { uint64_t v = 0; std::memcpy(&v, p, sz); return (float)v; }
    if (t == 'I') {
      if (sz == 1) { int8_t v; std::memcpy(&v, p, 1); return (float)v; }
      if (sz == 2) { int16_t v; std::memcpy(&v, p, 2); return (float)v; }
      if (sz == 4) { int32_t v; std::memcpy(&v, p, 4); return (float)v; }
      int64_t v; std::memcpy(&v, p, 8); return (float)v;
    }
    return 0.0f;
  };
  cloud.xyz.assign(points * 3, 0.0f);
  cloud.nrm.assign(points * 3, std::nanf(""));
  bool ok = true;
  if (data == "ascii") {
    for (size_t i = 0; i < points && ok; i++) {
      if (!std::fgets(line, sizeof line, f)) { ok = false; break; }
      size_t c = 0;
      char* save = nullptr;
      for (char* p = strtok_r(line, " \t\r\n", &save); p; p = strtok_r(nullptr, " \t\r\n", &save), c++) {
        // columns follow FIELDS x COUNT
        size_t fi = 0, acc = 0;
        while (fi < nf && acc + counts[fi] <= c) { acc += counts[fi]; fi++; }
        if (fi >= nf || c != acc) continue;
        for (int q = 0; q < 6; q++)
          if (col[q] == (int)fi) (q < 3 ? cloud.xyz[3 * i + q] : cloud.nrm[3 * i + q - 3]) = (float)std::atof(p);
      }
    }
  } else if (data == "binary") {
    std::vector<uint8_t> buf(rec * points);
    if (points && std::fread(buf.data(), rec, points, f) != points) ok = false;
    for (size_t i = 0; i < points && ok; i++)
      for (int q = 0; q < 6; q++)
        if (col[q] >= 0) (q < 3 ? cloud.xyz[3 * i + q] : cloud.nrm[3 * i + q - 3]) = as_float(&buf[i * rec + offset[col[q]]], col[q]);
  } else if (data == "binary_compressed") {
    uint32_t csize = 0, usize = 0;
    if (std::fread(&csize, 4, 1, f) != 1 || std::fread(&usize, 4, 1, f) != 1) ok = false;
    std::vector<uint8_t> cbuf(csize), ubuf(usize);
    if (ok && csize && std::fread(cbuf.data(), 1, csize, f) != csize) ok = false;
    if (ok && !lzf_decompress(cbuf.data(), csize, ubuf.data(), usize)) ok = false;
    if (ok && usize < rec * points) ok = false;
    // after decompression the layout is field-major (structure of arrays)
    std::vector<size_t> fstart(nf);
    size_t acc = 0;
    for (size_t i = 0; i < nf; i++) { fstart[i] = acc; acc += (size_t)sizes[i] * counts[i] * points; }
    for (size_t i = 0; i < points && ok; i++)
      for (int q = 0; q < 6; q++)
        if (col[q] >= 0)
          (q < 3 ? cloud.xyz[3 * i + q] : cloud.nrm[3 * i + q - 3]) =
              as_float(&ubuf[fstart[col[q]] + i * (size_t)sizes[col[q]] * counts[col[q]]], col[q]);
  } else {
    ok = false;
  }
  std::fclose(f);
  if (!ok) { cloud.xyz.clear(); cloud.nrm.clear(); }
  return ok;
}

// ------------------------------------------------------------------------------------------------
// pcl::console::{parse_argument, find_switch}: "--key value" lookup; > 0 means found
// (Integrate/Integrate.cpp:46-76, BuildCorrespondence/BuildCorrespondence.cpp:34-83).
// ------------------------------------------------------------------------------------------------
inline int find_argument(int argc, char** argv, const char* name) {
  for (int i = 1; i < argc; i++)
    if (std::strcmp(argv[i], name) == 0) return i;
  return -1;
}
inline bool find_switch(int argc, char** argv, const char* name) { return find_argument(argc, argv, name) != -1; }
inline int parse_argument(int argc, char** argv, const char* name, std::string& v) {
  const int i = find_argument(argc, argv, name) + 1;
  if (i > 0 && i < argc) { v = argv[i]; return i - 1; }
  return -1;
}
inline int parse_argument(int argc, char** argv, const char* name, int& v) {
  const int i = find_argument(argc, argv, name) + 1;
  if (i > 0 && i < argc) { v = std::atoi(argv[i]); return i - 1; }
  return -1;
}
inline int parse_argument(int argc, char** argv, const char* name, double& v) {
  const int i = find_argument(argc, argv, name) + 1;
  if (i > 0 && i < argc) { v = std::atof(argv[i]); return i - 1; }
  return -1;
}

}  // namespace er_io
