// FragmentOptimizer -- drop-in for the reference executable of the same name (FragmentOptimizer/FragmentOptimizer.cpp,
// OptApp.{h,cpp}, PointCloud.{h,cpp}): same flags, same input files (.log trajectories, reg_output.log,
// cloud_bin_<i>.pcd / cloud_bin_xyzn_<i>.xyzn, corres_<i>_<j>.txt), same outputs (pose.log, <save_to>.ctr, sample.pcd).
// The per-point work and the assembly of the normal equations over all correspondences run on the GPU through
// the er_fragopt_* calls of include/elastic_b200.h; this file is the control flow of OptimizeSLAC / OptimizeRigid, the
// lattice regulariser (one 3x3 polar rotation per lattice node) and the file formats.
//
// Modes: --slac (CVPR 2014), --rigid and --nonrigid (ICCV 2013; the reference's default when neither flag is given).
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <set>
#include <string>
#include <vector>

#include "elastic_b200.h"
#include "er_io.h"

namespace {

struct Options {
  std::string rgbd_file, reg_file, dir, init_ctr, ctr_file = "output.ctr", sample_file = "sample.pcd", pose_file = "pose.log";
  std::string blacklist_file, ipose_file, bench_json;
  int num = 0, resolution = 8, interval = 50, max_iteration = 5, max_inner_iteration = 10, sample_num = -1, blacklist_pair_num = 10000;
  int device = 0;
  std::vector<int> devices;          // --devices a,b,...: the pair list is dealt over these GPUs (--slac, --rigid)
  double weight = 1.0, length = 3.0;
};

int print_help() {
  printf("\nApplication parameters:\n");
  printf("    --help, -h                      : print this message\n");
  printf("    --rgbdslam <log_file>           : rgbdslam.log/opt_output.log file, get ipose\n");
  printf("    --registration <log_file>       : reg_output.log, invalid pair when frame_ == -1\n");
  printf("    --dir <dir_prefix>              : dir prefix, place to loopup .xyzn files\n");
  printf("    --num <number>                  : number of pieces, important parameter\n");
  printf("    --weight <weight>               : 1.0 for nonrigid, 10000.0 for rigid\n");
  printf("    --resolution <resolution>       : default - 8\n");
  printf("    --length <length>               : default - 3.0\n");
  printf("    --interval <interval>           : default - 50\n");
  printf("    --iteration <max_number>        : default - 5\n");
  printf("    --inner_iteration <max_number>  : default - 10\n");
  printf("    --save_to <ctr_file>            : default - output.ctr\n");
  printf("    --write_xyzn_sample <sample_num>: per <sample_num> write a point\n");
  printf("    --blasklist <blacklist_file>    : each line is the block we want to blacklist\n");
  printf("    --blacklistpair <threshold>     : threshold of accepting pairwise registration, default - 10000\n");
  printf("    --switchable <weight>           : use switchable g2o, default weight 0.1\n");
  printf("    --ipose <log_file>              : get ipose from log file\n");
  printf("    --verbose <verbpse_level>       : set verbosity level, from 0(none) to 5(verbose), default 3\n");
  printf("Optimization options:\n");
  printf("    --nonrigid                      : default, nonrigid alignment published in ICCV 2013\n");
  printf("    --rigid                         : dense rigid optimization\n");
  printf("    --slac                          : simultaneous localization and calibration, published in CVPR 2014\n");
  printf("Added by this implementation:\n");
  printf("    --device <id>                   : CUDA device, default 0\n");
  printf("    --devices <a,b,...>             : several GPUs: every GPU holds all fragments and a share of the pairs (--slac, --rigid)\n");
  printf("    --bench_json <file>             : write phase timings as JSON\n");
  return 0;
}

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------------------------------------ small dense algebra
typedef double Mat4[16];

void mat4_identity(double* m) { for (int e = 0; e < 16; ++e) m[e] = (e % 5 == 0) ? 1.0 : 0.0; }

// R = V U^T of the SVD C = U S V^T with the reference's reflection fix (COptApp::GetRotation, OptApp.cpp:860-882):
// one-sided Jacobi (Hestenes) on the columns of C, singular values sorted descending so that "the third column of U"
// is the one of the smallest singular value, as in Eigen's JacobiSVD.
void rotation_from_correlation(const double C[9], double R[9]) {
  double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  memcpy(A, C, sizeof A);                       // A = C V throughout; at convergence its columns are s_k u_k
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int r = 0; r < 3; ++r) { alpha += A[3 * r + p] * A[3 * r + p]; beta += A[3 * r + q] * A[3 * r + q]; gamma += A[3 * r + p] * A[3 * r + q]; }
        if (gamma == 0.0) continue;
        off = std::max(off, fabs(gamma) / sqrt(alpha * beta + 1e-300));
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int r = 0; r < 3; ++r) {
          const double ap = A[3 * r + p], aq = A[3 * r + q];
          A[3 * r + p] = c * ap - s * aq; A[3 * r + q] = s * ap + c * aq;
          const double vp = V[3 * r + p], vq = V[3 * r + q];
          V[3 * r + p] = c * vp - s * vq; V[3 * r + q] = s * vp + c * vq;
        }
      }
    if (off < 1e-15) break;
  }
  double s[3];
  int order[3] = {0, 1, 2};
  for (int k = 0; k < 3; ++k) s[k] = sqrt(A[k] * A[k] + A[3 + k] * A[3 + k] + A[6 + k] * A[6 + k]);
  std::sort(order, order + 3, [&](int a, int b) { return s[a] > s[b]; });
  double U[9], Vs[9];
  for (int k = 0; k < 3; ++k) {
    const int o = order[k];
    for (int r = 0; r < 3; ++r) { Vs[3 * r + k] = V[3 * r + o]; U[3 * r + k] = s[o] > 0 ? A[3 * r + o] / s[o] : 0.0; }
  }
  if (!(s[order[2]] > 1e-300)) {               // rank deficient: complete U with the cross product of its first two columns
    U[2] = U[3] * U[7] - U[6] * U[4]; U[5] = U[6] * U[1] - U[0] * U[7]; U[8] = U[0] * U[4] - U[3] * U[1];
  }
  auto compose = [&]() {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) R[3 * r + c] = (Vs[3 * r] * U[3 * c] + Vs[3 * r + 1] * U[3 * c + 1]) + Vs[3 * r + 2] * U[3 * c + 2];
  };
  compose();
  const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
  if (det < 0) {
    U[2] = -U[2]; U[5] = -U[5]; U[8] = -U[8];
    compose();
  }
}

// x = A^-1 b for a symmetric positive definite A given by its upper triangle (row-major, n x n)
bool spd_solve_upper(int n, std::vector<double> A, const double* b, double* x) {
  for (int k = 0; k < n; ++k) {
    double d = A[(size_t)k * n + k];
    for (int s = 0; s < k; ++s) d -= A[(size_t)s * n + k] * A[(size_t)s * n + k];
    if (!(d > 0)) return false;
    d = sqrt(d);
    A[(size_t)k * n + k] = d;
    for (int c = k + 1; c < n; ++c) {
      double v = A[(size_t)k * n + c];
      for (int s = 0; s < k; ++s) v -= A[(size_t)s * n + k] * A[(size_t)s * n + c];
      A[(size_t)k * n + c] = v / d;
    }
  }
  std::vector<double> y(n);
  for (int k = 0; k < n; ++k) {
    double v = b[k];
    for (int s = 0; s < k; ++s) v -= A[(size_t)s * n + k] * y[s];
    y[k] = v / A[(size_t)k * n + k];
  }
  for (int k = n - 1; k >= 0; --k) {
    double v = y[k];
    for (int c = k + 1; c < n; ++c) v -= A[(size_t)k * n + c] * x[c];
    x[k] = v / A[(size_t)k * n + k];
  }
  return true;
}

// AngleAxis(rz, Z) * AngleAxis(ry, Y) * AngleAxis(rx, X) with translation t, as a row-major 4x4 (OptApp.cpp:396-400)
void increment_matrix(const double* r6, double* m) {
  const double cx = cos(r6[0]), sx = sin(r6[0]), cy = cos(r6[1]), sy = sin(r6[1]), cz = cos(r6[2]), sz = sin(r6[2]);
  mat4_identity(m);
  m[0] = cz * cy; m[1] = cz * sy * sx - sz * cx; m[2] = cz * sy * cx + sz * sx;
  m[4] = sz * cy; m[5] = sz * sy * sx + cz * cx; m[6] = sz * sy * cx - cz * sx;
  m[8] = -sy;     m[9] = cy * sx;                m[10] = cy * cx;
  m[3] = r6[3]; m[7] = r6[4]; m[11] = r6[5];
}

// ------------------------------------------------------------------------------------------------ the optimiser's state
struct App {
  Options o;
  er_io::Trajectory rgbd, reg;
  std::set<int> blacklist;
  std::vector<int> abs2rel, rel2abs;
  int num = 0, nper = 0, r1 = 0;
  double unit = 0;
  std::vector<double> ipose, pose;              // num x 16
  er_fragopt* h = nullptr;                       // hs[0]: the handle that solves (and the only one in --nonrigid mode)
  std::vector<er_fragopt*> hs;                   // one per device

  // runs f on every handle, one host thread per device; false (after printing the library's message) if any call fails
  bool on_all(const std::function<int(er_fragopt*, int)>& f) {
    int bad = 0;
#pragma omp parallel for num_threads((int)hs.size()) schedule(static, 1) reduction(+ : bad)
    for (int d = 0; d < (int)hs.size(); ++d)
      if (f(hs[d], d) != ER_OK) { fprintf(stderr, "%s\n", er_last_error()); ++bad; }
    return bad == 0;
  }
  std::vector<size_t> cloud_size;
  int npairs = 0;
  double t_load = 0, t_assemble = 0, t_solve = 0, t_update = 0, t_reg = 0;
  double kernel_ms = 0;
  uint64_t rows = 0, flushes = 0;
  int iterations_run = 0;

  int node(int i, int j, int k) const { return i + j * r1 + k * r1 * r1; }

  void init_map() {                              // COptApp::InitMap
    abs2rel.assign(o.num, -1);
    rel2abs.clear();
    for (int i = 0; i < o.num; ++i)
      if (!blacklist.count(i)) { abs2rel[i] = (int)rel2abs.size(); rel2abs.push_back(i); }
    if ((int)rel2abs.size() != o.num) fprintf(stderr, "Blacklisted fragments, number reduced from %d to %d.\n", o.num, (int)rel2abs.size());
    num = (int)rel2abs.size();
  }

  bool init_ipose() {                            // COptApp::InitIPose
    pose.assign((size_t)num * 16, 0.0);
    if (!ipose.empty()) {
      if ((int)ipose.size() < num * 16) { fprintf(stderr, "--ipose holds fewer than %d poses.\n", num); return false; }
      return true;
    }
    double base[16], baseinv[16], t0inv[16], left[16], tmp[16];
    mat4_identity(base);
    base[3] = o.length / 2.0; base[7] = o.length / 2.0; base[11] = -0.3;
    er_mat4_inverse(base, baseinv);
    if (rgbd.data.empty()) { fprintf(stderr, "--rgbdslam trajectory is empty.\n"); return false; }
    er_mat4_inverse(rgbd.data[0].T, t0inv);
    er_mat4_mul(base, t0inv, left);
    ipose.assign((size_t)num * 16, 0.0);
    for (int i = 0; i < num; ++i) {
      const size_t f = (size_t)rel2abs[i] * o.interval;
      if (f >= rgbd.data.size()) { fprintf(stderr, "--rgbdslam trajectory has no frame %zu (fragment %d).\n", f, rel2abs[i]); return false; }
      er_mat4_mul(left, rgbd.data[f].T, tmp);
      er_mat4_mul(tmp, baseinv, &ipose[(size_t)i * 16]);
    }
    printf("IPose initialized.\n");
    return true;
  }

  // PointCloud::LoadFromXYZNFile's parse (PointCloud.cpp:46-63): sscanf into ONE array that lives across lines, so a line
  // that does not scan (blank, too few numbers) repeats what the array holds -- and what it holds has been through
  // GetCoordinate, which divides the normal by unit_length_ IN PLACE (PointCloud.h:136-138) once the point is in bounds.
  static void read_xyzn(const std::string& fn, std::vector<float>& rows, float unit, int res) {
    FILE* f = fopen(fn.c_str(), "r");
    if (!f) return;
    char buf[1024];
    float x[6] = {0, 0, 0, 0, 0, 0};
    while (fgets(buf, sizeof buf, f)) {
      if (strlen(buf) > 0 && buf[0] != '#') {
        sscanf(buf, "%f %f %f %f %f %f", &x[0], &x[1], &x[2], &x[3], &x[4], &x[5]);
        rows.insert(rows.end(), x, x + 6);
        const int c0 = (int)floorf(x[0] / unit), c1 = (int)floorf(x[1] / unit), c2 = (int)floorf(x[2] / unit);
        if (c0 < 0 || c0 >= res || c1 < 0 || c1 >= res || c2 < 0 || c2 >= res) break;     // "Point out of bound": loading stops
        x[3] /= unit; x[4] /= unit; x[5] /= unit;
      }
    }
    fclose(f);
  }

  bool init_pointclouds() {                      // COptApp::InitPointClouds
    cloud_size.assign(num, 0);
    std::vector<std::vector<float>> rows(num);
    std::vector<int> state(num, 0);
#pragma omp parallel for num_threads(8) schedule(dynamic)
    for (int i = 0; i < num; ++i) {
      const int ii = rel2abs[i];
      const std::string pcd = o.dir + "cloud_bin_" + std::to_string(ii) + ".pcd";
      const std::string xyzn = o.dir + "cloud_bin_xyzn_" + std::to_string(ii) + ".xyzn";
      if (er_io::file_exists(pcd)) {
        er_io::PointCloudXYZN c;
        if (!er_io::load_pcd_xyzn(pcd, c)) { state[i] = 2; continue; }
        for (size_t k = 0; k < c.size(); ++k)
          if (!std::isnan(c.nrm[3 * k])) {
            rows[i].insert(rows[i].end(), &c.xyz[3 * k], &c.xyz[3 * k] + 3);
            rows[i].insert(rows[i].end(), &c.nrm[3 * k], &c.nrm[3 * k] + 3);
          }
        state[i] = 1;
      } else if (er_io::file_exists(xyzn)) {
        read_xyzn(xyzn, rows[i], (float)o.length / o.resolution, o.resolution);
        state[i] = 1;
      } else {
        state[i] = 3;
      }
    }
    for (int i = 0; i < num; ++i) {
      const int ii = rel2abs[i];
      if (state[i] == 2) fprintf(stderr, "Error loading file.\n");
      if (state[i] == 3) fprintf(stderr, "File not found ... Check dir and num parameters.\n");
      size_t kept = 0;
      const size_t n = rows[i].size() / 6;
      if (!on_all([&](er_fragopt* g, int) { return er_fragopt_set_cloud(g, i, rows[i].data(), n, &kept); })) return false;
      if (kept != n) fprintf(stderr, "Error!! Point out of bound!!\n");
      else if (state[i] == 1) printf("Read cloud %d ... get %zu points.\n", ii, kept);
      cloud_size[i] = kept;
    }
    return true;
  }

  // Correspondence::LoadFromFile's parse (OptApp.h:22-35), from one read of the whole file
  static void read_corres(const std::string& fn, std::vector<int32_t>& rows) {
    FILE* f = fopen(fn.c_str(), "rb");
    if (!f) return;
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)sz + 1);
    const size_t got = fread(buf.data(), 1, (size_t)sz, f);
    fclose(f);
    buf[got] = 0;
    int a = 0, b = 0;                            // a line that does not scan repeats the previous pair, like sscanf
    const char* p = buf.data();
    const char* end = p + got;
    while (p < end) {
      const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
      const char* stop = nl ? nl : end;
      if (*p != '#') {                           // strlen(buffer) > 0 holds for every line fgets returns
        const char* q = p;
        char* e = nullptr;
        const long v0 = strtol(q, &e, 10);
        if (e != q && e <= stop) {
          a = (int)v0;
          q = e;
          const long v1 = strtol(q, &e, 10);
          if (e != q && e <= stop) b = (int)v1;
        }
        rows.push_back(a);
        rows.push_back(b);
      }
      p = nl ? nl + 1 : end;
    }
  }

  bool init_correspondences() {                  // COptApp::InitCorrespondences
    std::vector<int> idx0, idx1, rec;
    for (size_t r = 0; r < reg.data.size(); ++r) {
      const auto& d = reg.data[r];
      if (blacklist.count(d.id1) || blacklist.count(d.id2) || d.id1 >= num || d.id2 >= num || d.id1 < 0 || d.id2 < 0) continue;
      if (d.frame != -1 && d.frame >= o.blacklist_pair_num) {
        idx0.push_back(abs2rel[d.id1]);
        idx1.push_back(abs2rel[d.id2]);
        rec.push_back((int)r);
      }
    }
    npairs = (int)idx0.size();
    std::vector<std::vector<int32_t>> rows(npairs);
#pragma omp parallel for schedule(dynamic)
    for (int p = 0; p < npairs; ++p) {
      const auto& d = reg.data[rec[p]];
      read_corres(o.dir + "corres_" + std::to_string(d.id1) + "_" + std::to_string(d.id2) + ".txt", rows[p]);
    }
    std::vector<size_t> start(npairs + 1, 0);
    for (int p = 0; p < npairs; ++p) {
      const auto& d = reg.data[rec[p]];
      printf("Read %scorres_%d_%d.txt, get %zu correspondences.\n", o.dir.c_str(), d.id1, d.id2, rows[p].size() / 2);
      start[p + 1] = start[p] + rows[p].size() / 2;
    }
    std::vector<int32_t> all(2 * start[npairs]);
#pragma omp parallel for schedule(dynamic)
    for (int p = 0; p < npairs; ++p)
      if (!rows[p].empty()) memcpy(&all[2 * start[p]], rows[p].data(), rows[p].size() * sizeof(int32_t));
    rows_total = start[npairs];
    if (hs.size() == 1) {
      if (er_fragopt_set_pairs(h, npairs, idx0.data(), idx1.data(), start.data(), all.data()) != ER_OK) {
        fprintf(stderr, "%s\n", er_last_error());
        return false;
      }
      return true;
    }
    // several devices: longest pair first onto the least loaded device; a device keeps its pairs in list order
    std::vector<int> order(npairs), owner(npairs, 0);
    for (int p = 0; p < npairs; ++p) order[p] = p;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return rows[a].size() > rows[b].size(); });
    std::vector<size_t> load(hs.size(), 0);
    for (int p : order) {
      const int d = (int)(std::min_element(load.begin(), load.end()) - load.begin());
      owner[p] = d;
      load[d] += rows[p].size() / 2 + 1;
    }
    return on_all([&](er_fragopt* g, int d) {
      std::vector<int> i0, i1;
      std::vector<size_t> st(1, 0);
      std::vector<int32_t> mine;
      for (int p = 0; p < npairs; ++p)
        if (owner[p] == d) {
          i0.push_back(idx0[p]); i1.push_back(idx1[p]);
          mine.insert(mine.end(), rows[p].begin(), rows[p].end());
          st.push_back(st.back() + rows[p].size() / 2);
        }
      return er_fragopt_set_pairs(g, (int)i0.size(), i0.data(), i1.data(), st.data(), mine.data());
    });
  }
  size_t rows_total = 0;

  void neighbours(int i, int j, int k, int off, int* out, int* n) const {   // the six-neighbourhood, reference order
    const int R = o.resolution;
    *n = 0;
    if (i > 0) out[(*n)++] = node(i - 1, j, k) * 3 + off;
    if (i < R) out[(*n)++] = node(i + 1, j, k) * 3 + off;
    if (j > 0) out[(*n)++] = node(i, j - 1, k) * 3 + off;
    if (j < R) out[(*n)++] = node(i, j + 1, k) * 3 + off;
    if (k > 0) out[(*n)++] = node(i, j, k - 1) * 3 + off;
    if (k < R) out[(*n)++] = node(i, j, k + 1) * 3 + off;
  }

  static void rotation_at(int idx, const int* idxx, int n, const std::vector<double>& ictr, const std::vector<double>& ctr, double R[9]) {
    double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < n; ++t) {
      double a[3], b[3];
      for (int x = 0; x < 3; ++x) { a[x] = ictr[idx + x] - ictr[idxx[t] + x]; b[x] = ctr[idx + x] - ctr[idxx[t] + x]; }
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) C[3 * r + c] += a[r] * b[c];
    }
    rotation_from_correlation(C, R);
  }

  void expand_ctr(const std::vector<double>& ctr, std::vector<double>& out) const {     // COptApp::ExpandCtr
    out.assign((size_t)num * nper, 0.0);
    for (int l = 0; l < num; ++l) {
      const double* P = &pose[(size_t)l * 16];
      for (int i = 0; i < nper / 3; ++i) {
        const double x = ctr[num * 6 + i * 3], y = ctr[num * 6 + i * 3 + 1], z = ctr[num * 6 + i * 3 + 2];
        for (int r = 0; r < 3; ++r) out[(size_t)l * nper + i * 3 + r] = ((P[4 * r] * x + P[4 * r + 1] * y) + P[4 * r + 2] * z) + P[4 * r + 3] * 1.0;
      }
    }
  }

  void pose_to_ctr(std::vector<double>& out) const {                                    // COptApp::Pose2Ctr
    out.assign((size_t)num * nper, 0.0);
    for (int l = 0; l < num; ++l) {
      const double* P = &pose[(size_t)l * 16];
      for (int i = 0; i < r1; ++i)
        for (int j = 0; j < r1; ++j)
          for (int k = 0; k < r1; ++k) {
            const double x = i * unit, y = j * unit, z = k * unit;
            for (int r = 0; r < 3; ++r)
              out[(size_t)l * nper + node(i, j, k) * 3 + r] = ((P[4 * r] * x + P[4 * r + 1] * y) + P[4 * r + 2] * z) + P[4 * r + 3] * 1.0;
          }
    }
  }

  void save_ctr(const std::vector<double>& ctr, const std::string& fn) const {          // COptApp::SaveCtr
    printf("Save ctr to file %s ... ", fn.c_str());
    FILE* f = fopen(fn.c_str(), "w");
    if (f) {
      for (int i = 0; i < num * nper / 3; ++i) fprintf(f, "%.10f %.10f %.10f\n", ctr[(size_t)i * 3], ctr[(size_t)i * 3 + 1], ctr[(size_t)i * 3 + 2]);
      fclose(f);
    }
    printf("Done.\n");
  }

  void save_pose_log() const {
    er_io::Trajectory out;
    for (int i = 0; i < num; ++i) {
      er_io::FramedTransformation ft;
      ft.id1 = i; ft.id2 = i; ft.frame = i + 1;
      memcpy(ft.T, &pose[(size_t)i * 16], sizeof ft.T);
      out.data.push_back(ft);
    }
    out.save(o.pose_file);
  }

  // COptApp::SavePoints: every sample_num-th point of every fragment, warped by `ctr`, normal re-normalised in double.
  // PCD v0.7 "binary_compressed" (what PCDWriter::writeBinaryCompressed produces): fields stored one after the other,
  // LZF stream made of literal runs only (a valid stream; the file is a visualisation by-product).
  bool save_points(const std::vector<double>& ctr) {
    if (o.sample_num <= 0) return true;
    if (er_fragopt_update_normals(h, ctr.data()) != ER_OK) return false;
    std::vector<float> cols[8];
    for (int l = 0; l < num; ++l) {
      std::vector<float> rows(6 * std::max<size_t>(cloud_size[l], 1));
      if (er_fragopt_get_points(h, l, rows.data(), cloud_size[l], ctr.data()) != ER_OK) return false;
      for (size_t i = 0; i < cloud_size[l]; i += (size_t)o.sample_num) {
        const double nx = rows[6 * i + 3], ny = rows[6 * i + 4], nz = rows[6 * i + 5];
        const double len = sqrt(nx * nx + ny * ny + nz * nz);
        cols[0].push_back(rows[6 * i]); cols[1].push_back(rows[6 * i + 1]); cols[2].push_back(rows[6 * i + 2]); cols[3].push_back(0.0f);
        cols[4].push_back((float)(nx / len)); cols[5].push_back((float)(ny / len)); cols[6].push_back((float)(nz / len)); cols[7].push_back(0.0f);
      }
    }
    const size_t n = cols[0].size();
    std::vector<uint8_t> raw;
    for (int c = 0; c < 8; ++c) raw.insert(raw.end(), (const uint8_t*)cols[c].data(), (const uint8_t*)cols[c].data() + 4 * n);
    std::vector<uint8_t> lzf;
    for (size_t p = 0; p < raw.size(); p += 32) {
      const size_t len = std::min<size_t>(32, raw.size() - p);
      lzf.push_back((uint8_t)(len - 1));
      lzf.insert(lzf.end(), raw.begin() + p, raw.begin() + p + len);
    }
    printf("Save sample pcd into %s ... ", o.sample_file.c_str());
    FILE* f = fopen(o.sample_file.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb normal_x normal_y normal_z curvature\n"
               "SIZE 4 4 4 4 4 4 4 4\nTYPE F F F F F F F F\nCOUNT 1 1 1 1 1 1 1 1\nWIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\n"
               "POINTS %zu\nDATA binary_compressed\n", n, n);
    const uint32_t csize = (uint32_t)lzf.size(), usize = (uint32_t)raw.size();
    fwrite(&csize, 4, 1, f);
    fwrite(&usize, 4, 1, f);
    if (!lzf.empty()) fwrite(lzf.data(), 1, lzf.size(), f);
    fclose(f);
    printf("Done.\n");
    return true;
  }

  bool common_init(int matrix_mode) {
    (void)matrix_mode;
    init_map();
    if (num <= 0) { fprintf(stderr, "No fragments left (--num %d).\n", o.num); return false; }
    unit = o.length / o.resolution;
    r1 = o.resolution + 1;
    nper = r1 * r1 * r1 * 3;
    std::vector<int> devs = o.devices.empty() ? std::vector<int>(1, o.device) : o.devices;
    if (matrix_mode == 2 && devs.size() > 1) {
      fprintf(stderr, "--nonrigid keeps its dense system on one GPU: using device %d only.\n", devs[0]);
      devs.resize(1);
    }
    for (int dev : devs) {
      er_fragopt* g = er_fragopt_create(dev, num, o.resolution, o.length);
      if (!g) { fprintf(stderr, "%s\n", er_last_error()); return false; }
      hs.push_back(g);
    }
    h = hs[0];
    const double t0 = now_ms();
    const bool ok = init_ipose() && init_pointclouds() && init_correspondences();
    t_load = now_ms() - t0;
    return ok;
  }

  void apply_increments(const std::vector<double>& result, std::vector<double>* inc_out) {
    if (inc_out) inc_out->assign((size_t)num * 16, 0.0);
    for (int l = 0; l < num; ++l) {
      double inc[16], np[16];
      increment_matrix(&result[(size_t)l * 6], inc);
      er_mat4_mul(inc, &pose[(size_t)l * 16], np);
      memcpy(&pose[(size_t)l * 16], np, sizeof np);
      if (inc_out) memcpy(&(*inc_out)[(size_t)l * 16], inc, sizeof inc);
    }
  }

  // ---------------------------------------------------------------------------------------------- COptApp::OptimizeSLAC
  int optimize_slac() {
    fprintf(stderr, "SLAC optimization.\n");
    printf("Parameters: weight %.5f, resolution %d, piece number %d, max iteration %d\n", o.weight, o.resolution, o.num, o.max_iteration);
    const double default_weight = o.num * o.weight;          // the fragment count BEFORE blacklisting (OptApp.cpp:423)
    if (!common_init(0)) return 1;
    const int M = 6 * num + nper;
    std::vector<double> ictr(M, 0.0), this_ctr(M, 0.0);
    for (int i = 0; i < r1; ++i)
      for (int j = 0; j < r1; ++j)
        for (int k = 0; k < r1; ++k) {
          ictr[num * 6 + node(i, j, k) * 3 + 0] = i * unit;
          ictr[num * 6 + node(i, j, k) * 3 + 1] = j * unit;
          ictr[num * 6 + node(i, j, k) * 3 + 2] = k * unit;
        }
    this_ctr = ictr;
    if (o.init_ctr.length() > 1) {                           // COptApp::InitCtrSLAC
      FILE* f = fopen(o.init_ctr.c_str(), "r");
      if (f) {
        char buf[1024];
        for (int i = 0; i < nper / 3; ++i) {
          if (!fgets(buf, sizeof buf, f)) break;
          sscanf(buf, "%lf %lf %lf", &this_ctr[num * 6 + i * 3], &this_ctr[num * 6 + i * 3 + 1], &this_ctr[num * 6 + i * 3 + 2]);
        }
        fclose(f);
      }
    }
    std::vector<double> rot_t((size_t)num * 9), expand;
    auto refresh_rot = [&]() {
      for (int l = 0; l < num; ++l)
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) rot_t[(size_t)l * 9 + 3 * r + c] = pose[(size_t)l * 16 + 4 * c + r];     // transpose of the 3x3 block
    };
    pose = ipose;
    pose.resize((size_t)num * 16);
    refresh_rot();
    if (!on_all([&](er_fragopt* g, int) { return er_fragopt_update_pose(g, pose.data()); })) return 1;

    const double t_neat = now_ms();
    // baseJJ * default_weight plus the six ones on the first fragment's pose (OptApp.cpp:449-464): every lattice edge
    // is visited from both ends, each visit adding [1 -1; -1 1] per axis
    std::vector<double> base((size_t)M * M, 0.0);
    for (int i = 0; i < r1; ++i)
      for (int j = 0; j < r1; ++j)
        for (int k = 0; k < r1; ++k) {
          int nb[6], n;
          neighbours(i, j, k, 6 * num, nb, &n);
          const int idx = node(i, j, k) * 3 + 6 * num;
          for (int t = 0; t < n; ++t)
            for (int x = 0; x < 3; ++x) {
              const int a = idx + x, b = nb[t] + x;
              base[(size_t)a * M + a] += default_weight;
              base[(size_t)b * M + b] += default_weight;
              base[(size_t)std::min(a, b) * M + std::max(a, b)] -= default_weight;
            }
        }
    {
      const int anchor = node(o.resolution / 2, o.resolution / 2, 0) * 3 + 6 * num;
      for (int x = 0; x < 3; ++x) base[(size_t)(anchor + x) * M + anchor + x] += default_weight;
      for (int k = 0; k < 6; ++k) base[(size_t)k * M + k] += 1.0;
    }
    std::vector<double> Jb(M), base_jb(M), rhs(M), result(M);
    for (int itr = 0; itr < o.max_iteration; ++itr) {
      double t0 = now_ms();
      double score = 0.0;
      printf("Processing %d ... ", npairs);
      {
        std::vector<std::vector<double>> jbs(hs.size(), std::vector<double>(M));
        std::vector<double> scores(hs.size(), 0.0);
        if (!on_all([&](er_fragopt* g, int d) { return er_fragopt_slac_assemble(g, rot_t.data(), jbs[d].data(), &scores[d]); })) return 1;
        Jb = jbs[0];
        score = scores[0];
        for (size_t d = 1; d < hs.size(); ++d) {              // device order: the sums do not depend on thread timing
          if (er_fragopt_slac_accumulate(h, hs[d]) != ER_OK) { fprintf(stderr, "%s\n", er_last_error()); return 1; }
          for (int e = 0; e < M; ++e) Jb[e] += jbs[d][e];
          score += scores[d];
        }
      }
      {
        float ms = 0;
        uint64_t r = 0, fl = 0;
        if (er_fragopt_last_assemble_stats(h, &ms, &r, &fl) == ER_OK) { kernel_ms += ms; rows += r; flushes += fl; }
      }
      t_assemble += now_ms() - t0;
      printf("Done.\n");
      printf("Data error score is : %.2f\n", score);

      t0 = now_ms();
      const std::vector<double>& temp_ctr = this_ctr;        // the lattice before this iteration's update
      std::fill(base_jb.begin(), base_jb.end(), 0.0);
      double regscore = 0.0;
      for (int i = 0; i < r1; ++i)
        for (int j = 0; j < r1; ++j)
          for (int k = 0; k < r1; ++k) {
            int nb[6], n;
            neighbours(i, j, k, 6 * num, nb, &n);
            const int idx = node(i, j, k) * 3 + 6 * num;
            double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            if (!(i == o.resolution / 2 && j == o.resolution / 2 && k == 0)) rotation_at(idx, nb, n, ictr, temp_ctr, R);
            for (int t = 0; t < n; ++t) {
              double d0[3], bx[3];
              for (int x = 0; x < 3; ++x) d0[x] = ictr[idx + x] - ictr[nb[t] + x];
              for (int x = 0; x < 3; ++x)
                bx[x] = (temp_ctr[idx + x] - temp_ctr[nb[t] + x]) - ((R[3 * x] * d0[0] + R[3 * x + 1] * d0[1]) + R[3 * x + 2] * d0[2]);
              regscore += default_weight * ((bx[0] * bx[0] + bx[1] * bx[1]) + bx[2] * bx[2]);
              for (int x = 0; x < 3; ++x) { base_jb[idx + x] += bx[x] * default_weight; base_jb[nb[t] + x] -= bx[x] * default_weight; }
            }
          }
      for (int e = 0; e < M; ++e) rhs[e] = Jb[e] + base_jb[e];
      t_reg += now_ms() - t0;
      printf("Regularization error score is : %.2f\n", regscore);

      t0 = now_ms();
      if (er_fragopt_slac_solve(h, itr == 0 ? base.data() : nullptr, rhs.data(), result.data()) != ER_OK) { fprintf(stderr, "%s\n", er_last_error()); return 1; }
      for (int e = 0; e < M; ++e) result[e] = -result[e];
      t_solve += now_ms() - t0;

      t0 = now_ms();
      for (int l = 0; l < nper; ++l) this_ctr[num * 6 + l] += result[num * 6 + l];
      apply_increments(result, nullptr);
      refresh_rot();
      expand_ctr(this_ctr, expand);
      if (!on_all([&](er_fragopt* g, int) { return er_fragopt_update_points_pn(g, expand.data()); })) return 1;
      t_update += now_ms() - t0;
      ++iterations_run;
    }
    fprintf(stderr, "Neat Optimization took %gms.\n", now_ms() - t_neat);
    save_pose_log();
    expand_ctr(this_ctr, expand);
    save_ctr(expand, o.ctr_file);
    if (!save_points(expand)) { fprintf(stderr, "%s\n", er_last_error()); return 1; }
    return 0;
  }

  // ---------------------------------------------------------------------------------------------- COptApp::OptimizeRigid
  int optimize_rigid() {
    fprintf(stderr, "Rigid optimization.\n");
    printf("Parameters: resolution %d, piece number %d, max iteration %d\n", o.resolution, o.num, o.max_iteration);
    if (!common_init(1)) return 1;
    const int n6 = 6 * num;
    pose = ipose;
    pose.resize((size_t)num * 16);
    if (!on_all([&](er_fragopt* g, int) { return er_fragopt_update_pose(g, pose.data()); })) return 1;
    std::vector<double> JJ((size_t)n6 * n6), Jb(n6), result(n6), inc;
    for (int itr = 0; itr < o.max_iteration; ++itr) {
      double t0 = now_ms();
      double score = 0.0;
      printf("Processing %d ... ", npairs);
      {
        std::vector<std::vector<double>> jjs(hs.size(), std::vector<double>((size_t)n6 * n6)), jbs(hs.size(), std::vector<double>(n6));
        std::vector<double> scores(hs.size(), 0.0);
        if (!on_all([&](er_fragopt* g, int d) { return er_fragopt_rigid_assemble(g, jjs[d].data(), jbs[d].data(), &scores[d]); })) return 1;
        JJ = jjs[0]; Jb = jbs[0]; score = scores[0];
        for (size_t d = 1; d < hs.size(); ++d) {
          for (size_t e = 0; e < JJ.size(); ++e) JJ[e] += jjs[d][e];
          for (int e = 0; e < n6; ++e) Jb[e] += jbs[d][e];
          score += scores[d];
        }
      }
      for (int k = 0; k < 6 && k < n6; ++k) JJ[(size_t)k * n6 + k] += (double)npairs;       // mat_adder.Add(k, k, 1) once per pair
      t_assemble += now_ms() - t0;
      printf("Done.\n");
      printf("Error score is : %.2f\n", score);
      t0 = now_ms();
      if (!spd_solve_upper(n6, JJ, Jb.data(), result.data())) { fprintf(stderr, "The normal equations are not positive definite.\n"); return 1; }
      for (int e = 0; e < n6; ++e) result[e] = -result[e];
      t_solve += now_ms() - t0;
      t0 = now_ms();
      apply_increments(result, &inc);
      if (!on_all([&](er_fragopt* g, int) { return er_fragopt_update_pose(g, inc.data()); })) return 1;
      t_update += now_ms() - t0;
      ++iterations_run;
    }
    save_pose_log();
    std::vector<double> ctr;
    pose_to_ctr(ctr);
    save_ctr(ctr, o.ctr_file);
    if (!save_points(ctr)) { fprintf(stderr, "%s\n", er_last_error()); return 1; }
    return 0;
  }

  // ---------------------------------------------------------------------------------------------- COptApp::OptimizeNonrigid
  int optimize_nonrigid() {
    fprintf(stderr, "Nonrigid optimization.\n");
    printf("Parameters: weight %.5f, resolution %d, piece number %d, max iteration %d\n", o.weight, o.resolution, o.num, o.max_iteration);
    if (!common_init(2)) return 1;
    const size_t N = (size_t)num * nper;
    std::vector<double> ctr(N, 0.0), ictr, oldctr, Ab(N);
    if (o.init_ctr.length() > 1) {                            // COptApp::InitCtr
      FILE* f = fopen(o.init_ctr.c_str(), "r");
      if (f) {
        char buf[1024];
        for (size_t i = 0; i < N / 3; ++i) {
          if (!fgets(buf, sizeof buf, f)) break;
          sscanf(buf, "%lf %lf %lf", &ctr[i * 3], &ctr[i * 3 + 1], &ctr[i * 3 + 2]);
        }
        fclose(f);
      }
    } else {
      pose = ipose;
      pose.resize((size_t)num * 16);
      pose_to_ctr(ctr);
    }
    ictr = ctr;
    oldctr = ctr;
    // COptApp::InitBaseAA: per fragment, every lattice edge visited from both ends adds [1 -1; -1 1] per axis; fragment 0's
    // first node is pinned by three ones.  Upper triangle, as (row, col, value) triplets.
    std::vector<int32_t> br, bc;
    std::vector<double> bv;
    for (int l = 0; l < num; ++l) {
      for (int i = 0; i < r1; ++i)
        for (int j = 0; j < r1; ++j)
          for (int k = 0; k < r1; ++k) {
            int nb[6], n;
            neighbours(i, j, k, l * nper, nb, &n);
            const int idx = node(i, j, k) * 3 + l * nper;
            for (int t = 0; t < n; ++t)
              for (int x = 0; x < 3; ++x) {
                const int a = idx + x, b = nb[t] + x;
                br.push_back(a); bc.push_back(a); bv.push_back(1.0);
                br.push_back(b); bc.push_back(b); bv.push_back(1.0);
                br.push_back(std::min(a, b)); bc.push_back(std::max(a, b)); bv.push_back(-1.0);
              }
          }
      if (l == 0)
        for (int x = 0; x < 3; ++x) { br.push_back(x); bc.push_back(x); bv.push_back(1.0); }
    }
    for (int itr = 0; itr < o.max_iteration; ++itr) {
      double t0 = now_ms();
      if (er_fragopt_update_normals(h, ctr.data()) != ER_OK) { fprintf(stderr, "%s\n", er_last_error()); return 1; }
      t_update += now_ms() - t0;
      t0 = now_ms();
      printf("Processing %d ... ", npairs);
      if (er_fragopt_nonrigid_assemble(h, o.weight) != ER_OK) { fprintf(stderr, "%s\n", er_last_error()); return 1; }
      {
        float ms = 0;
        uint64_t r = 0, fl = 0;
        if (er_fragopt_last_assemble_stats(h, &ms, &r, &fl) == ER_OK) { kernel_ms += ms; rows += r; flushes += fl; }
      }
      t_assemble += now_ms() - t0;
      printf("Done.\n");
      t0 = now_ms();
      if (er_fragopt_nonrigid_factor(h, bv.size(), br.data(), bc.data(), bv.data()) != ER_OK) { fprintf(stderr, "%s\n", er_last_error()); return 1; }
      t_solve += now_ms() - t0;
      if (o.sample_num > 0) save_ctr(oldctr, "itr" + std::to_string(itr) + ".ctr");
      for (int m = 0; m < o.max_inner_iteration; ++m) {
        t0 = now_ms();
        std::fill(Ab.begin(), Ab.end(), 0.0);
#pragma omp parallel for schedule(static)
        for (int l = 0; l < num; ++l)                          // fragments touch disjoint parts of Ab
          for (int i = 0; i < r1; ++i)
            for (int j = 0; j < r1; ++j)
              for (int k = 0; k < r1; ++k) {
                int nb[6], n;
                neighbours(i, j, k, l * nper, nb, &n);
                const int idx = node(i, j, k) * 3 + l * nper;
                double R[9];
                rotation_at(idx, nb, n, ictr, ctr, R);
                for (int t = 0; t < n; ++t) {
                  double d0[3];
                  for (int x = 0; x < 3; ++x) d0[x] = ictr[idx + x] - ictr[nb[t] + x];
                  for (int x = 0; x < 3; ++x) {
                    const double bx = (R[3 * x] * d0[0] + R[3 * x + 1] * d0[1]) + R[3 * x + 2] * d0[2];
                    Ab[idx + x] += bx;
                    Ab[nb[t] + x] -= bx;
                  }
                }
              }
        t_reg += now_ms() - t0;
        t0 = now_ms();
        oldctr = ctr;
        if (er_fragopt_nonrigid_solve(h, Ab.data(), ctr.data()) != ER_OK) { fprintf(stderr, "%s\n", er_last_error()); return 1; }
        t_solve += now_ms() - t0;
        double score = 0.0;
        for (size_t e = 0; e < N; ++e) score += (oldctr[e] - ctr[e]) * (oldctr[e] - ctr[e]);
        fprintf(stderr, "Iteration #%d:%d (%d:%d) : score is %.4f\n", itr + 1, m + 1, o.max_iteration, o.max_inner_iteration, sqrt(score));
        if (o.sample_num > 0) save_ctr(ctr, "itr" + std::to_string(itr) + "_inner" + std::to_string(m) + "_out.ctr");
      }
      ++iterations_run;
    }
    save_ctr(ctr, o.ctr_file);
    if (!save_points(ctr)) { fprintf(stderr, "%s\n", er_last_error()); return 1; }
    return 0;
  }
};

}  // namespace

int main(int argc, char** argv) {
  using namespace er_io;
  if (argc == 1 || find_switch(argc, argv, "--help") || find_switch(argc, argv, "-h")) return print_help();
  App app;
  Options& o = app.o;
  if (parse_argument(argc, argv, "--rgbdslam", o.rgbd_file) > 0) app.rgbd.load(o.rgbd_file);
  if (parse_argument(argc, argv, "--registration", o.reg_file) > 0) app.reg.load(o.reg_file);
  parse_argument(argc, argv, "--dir", o.dir);
  parse_argument(argc, argv, "--init_ctr", o.init_ctr);
  parse_argument(argc, argv, "--num", o.num);
  parse_argument(argc, argv, "--weight", o.weight);
  parse_argument(argc, argv, "--resolution", o.resolution);
  parse_argument(argc, argv, "--length", o.length);
  parse_argument(argc, argv, "--interval", o.interval);
  parse_argument(argc, argv, "--iteration", o.max_iteration);
  parse_argument(argc, argv, "--inner_iteration", o.max_inner_iteration);
  parse_argument(argc, argv, "--save_to", o.ctr_file);
  parse_argument(argc, argv, "--write_xyzn_sample", o.sample_num);
  parse_argument(argc, argv, "--blacklistpair", o.blacklist_pair_num);
  parse_argument(argc, argv, "--device", o.device);
  {
    std::string list;
    if (parse_argument(argc, argv, "--devices", list) > 0) {
      for (size_t p = 0; p < list.size();) {
        const size_t q = list.find(',', p);
        const std::string tok = list.substr(p, q == std::string::npos ? std::string::npos : q - p);
        if (!tok.empty()) o.devices.push_back(atoi(tok.c_str()));
        if (q == std::string::npos) break;
        p = q + 1;
      }
    }
  }
  parse_argument(argc, argv, "--bench_json", o.bench_json);
  if (parse_argument(argc, argv, "--blacklist", o.blacklist_file) > 0) {           // COptApp::Blacklist
    FILE* f = fopen(o.blacklist_file.c_str(), "r");
    if (f) {
      char buf[1024];
      int id = 0;
      while (fgets(buf, sizeof buf, f))
        if (strlen(buf) > 0 && buf[0] != '#') { sscanf(buf, "%d", &id); app.blacklist.insert(id); }
      fclose(f);
    }
  }
  if (parse_argument(argc, argv, "--ipose", o.ipose_file) > 0) {
    Trajectory ip;
    ip.load(o.ipose_file);
    for (const auto& d : ip.data) app.ipose.insert(app.ipose.end(), d.T, d.T + 16);
  }
  const double t0 = now_ms();
  int rc;
  const char* mode;
  if (find_switch(argc, argv, "--slac")) { mode = "slac"; rc = app.optimize_slac(); }
  else if (find_switch(argc, argv, "--rigid")) { mode = "rigid"; rc = app.optimize_rigid(); }
  else { mode = "nonrigid"; rc = app.optimize_nonrigid(); }
  const double total = now_ms() - t0;
  fprintf(stderr, "Optimize All took %gms.\n", total);
  if (!o.bench_json.empty()) {
    FILE* f = fopen(o.bench_json.c_str(), "w");
    if (f) {
      fprintf(f, "{\"mode\": \"%s\", \"rc\": %d, \"fragments\": %d, \"pairs\": %d, \"correspondences\": %zu, \"iterations\": %d, "
                 "\"total_ms\": %.3f, \"load_ms\": %.3f, \"assemble_ms\": %.3f, \"assemble_kernel_ms\": %.3f, \"regulariser_ms\": %.3f, "
                 "\"solve_ms\": %.3f, \"update_ms\": %.3f, \"rows_assembled\": %llu, \"flushes\": %llu, \"launches\": %llu, \"devices\": %zu}\n",
              mode, rc, app.num, app.npairs, app.rows_total, app.iterations_run, total, app.t_load, app.t_assemble, app.kernel_ms, app.t_reg,
              app.t_solve, app.t_update, (unsigned long long)app.rows, (unsigned long long)app.flushes,
              (unsigned long long)(app.h ? er_fragopt_launch_count(app.h) : 0), app.hs.size());
      fclose(f);
    }
  }
  for (er_fragopt* g : app.hs) er_fragopt_destroy(g);
  return rc;
}
