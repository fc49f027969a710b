// BuildCorrespondence -- drop-in for the reference's BuildCorrespondence executable
// (BuildCorrespondence/BuildCorrespondence.cpp, CorresApp.cpp): same command line, same .log / cloud_bin_<i>.pcd
// inputs, same reg_output.log / reg_output.info / corres_<i>_<j>.txt / cloud_bin_xyzn_<i>.xyzn outputs.
// The nearest-neighbour passes and the ICP iterations run on a B200 behind libelastic_b200
// (include/elastic_b200.h); this file restates the host logic of CCorresApp around them.
//
// Differences from the reference, all outside the numerics: the accepted pairs of --registration are aligned in
// one batched call (many pairs in flight on the GPU) instead of one OpenMP thread per pair; extra flags
// --device N, --devices A,B,.. (ICP pairs sharded over several GPUs in-process), --fixed_iters (run exactly 20 ICP
// iterations), --bench_json FILE.
#include <chrono>
#include <cstdio>
#include <iostream>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "elastic_b200.h"
#include "er_io.h"

using namespace er_io;

static int print_help() {
  std::cout << "\nApplication parameters:" << std::endl;
  std::cout << "    --help, -h                      : print this message" << std::endl;
  std::cout << "    --traj <log_file>               : initialization, camera pose trajectory" << std::endl;
  std::cout << "    --num <num_of_fragments>        : use together with --traj" << std::endl;
  std::cout << "    --interval <interval>           : use together with --traj, default : 50" << std::endl;
  std::cout << "    --length <length>               : use together with --traj, default : 3.0" << std::endl;
  std::cout << "    --reg_traj <log_file>           : initialization, registration.log file, will overwrite --traj" << std::endl;
  std::cout << "    --registration                  : registration results are written into reg_output.log file" << std::endl;
  std::cout << "    --reg_dist <dist>               : distance threshold for registration, default 0.03" << std::endl;
  std::cout << "    --reg_ratio <ratio>             : correspondence points are at least <ratio> in each point cloud, default 0.25" << std::endl;
  std::cout << "    --reg_num <number>              : correspondence point number requirement, default 40,000" << std::endl;
  std::cout << "    --blasklist <blacklist_file>    : each line is the block we want to blacklist" << std::endl;
  std::cout << "    --save_xyzn                     : save point cloud into ascii file" << std::endl;
  std::cout << "    --output_information            : output the registration information matrix into reg_output.info" << std::endl;
  std::cout << "    --redux <log_file>              : use transformations in <log_file> as constraints" << std::endl;
  std::cout << "B200 build only:" << std::endl;
  std::cout << "    --device <n> --fixed_iters --bench_json <file>" << std::endl;
  std::cout << "    --devices <a,b,...>             : ICP on several GPUs (fragments replicated, accepted pairs dealt round-robin)" << std::endl;
  return 0;
}

struct CorresApp {
  // defaults: CCorresApp::CCorresApp (CorresApp.cpp:7-24)
  Trajectory corres_traj;
  Information corres_info;
  double dist_thresh = 0.015, normal_thresh = 0.8660;
  bool save_xyzn = false, save_corres = true;
  double reg_dist = 0.03, reg_ratio = 0.25, bbox_length = -1.0;
  int reg_num = 40000;
  bool registration = false, output_information = false;
  std::string dir_name;
  std::set<int> blacklist;
  bool redux = false;
  Trajectory redux_traj;
  std::unordered_map<int, int> redux_map;
  int num = 0, interval = 50;
  double length = 3.0;
  int device = 0;
  std::vector<int> extra_devices;    // --devices a,b,c: further GPUs; accepted pairs are dealt round-robin over all of them
  bool fixed_iters = false;

  std::vector<PointCloudXYZN> pointclouds;
  std::vector<er_cloud*> clouds;                       // on `device`: gate, ICP share, correspondences
  std::vector<std::vector<er_cloud*>> replica;         // replica[k][i]: fragment i on extra_devices[k] (ICP share only)
  long icp_pairs = 0, icp_iterations = 0;
  double icp_ms = 0, gate_ms = 0, corres_ms = 0, write_ms = 0, load_ms = 0, upload_ms = 0;

  // Runs fn(k, share, clouds-on-device-k, error) for device k's round-robin share of a list of n independent items, one
  // host thread per device (the reference's pair loops are OpenMP loops over independent pairs, CorresApp.cpp:121,220).
  template <class Fn>
  bool on_devices(int n, Fn fn, std::string& err) {
    const int ndev = 1 + (int)extra_devices.size();
    std::vector<std::string> errs((size_t)ndev);
#pragma omp parallel for num_threads(ndev) schedule(static, 1)
    for (int k = 0; k < ndev; k++) {
      std::vector<int> sel;
      for (int p = k; p < n; p += ndev) sel.push_back(p);
      if (sel.empty()) continue;
      er_cloud* const* cl = (k == 0) ? clouds.data() : replica[(size_t)k - 1].data();
      fn(k, sel, cl, errs[(size_t)k]);                 // (the library's error text is per host thread: fn copies it)
    }
    for (int k = 0; k < ndev; k++)
      if (!errs[(size_t)k].empty()) {
        err = "device " + std::to_string(k == 0 ? device : extra_devices[(size_t)k - 1]) + ": " + errs[(size_t)k];
        return false;
      }
    return true;
  }

  // "<target index> <source index>" per line (CorresApp.cpp:175-184), formatted by hand into one buffer: fprintf per
  // line dominates the run time of the correspondence pass otherwise
  static void write_corres_file(const char* fn, const int32_t* rows, size_t nrows) {
    FILE* f = std::fopen(fn, "w");
    if (!f) return;
    std::string buf;
    buf.reserve(nrows * 16 + 16);
    char tmp[32];
    auto put = [&](int v) {
      int n = 0;
      unsigned u = v < 0 ? 0u - (unsigned)v : (unsigned)v;
      do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
      if (v < 0) buf.push_back('-');
      while (n) buf.push_back(tmp[--n]);
    };
    for (size_t k = 0; k < 2 * nrows; k += 2) { put(rows[k]); buf.push_back(' '); put(rows[k + 1]); buf.push_back('\n'); }
    std::fwrite(buf.data(), 1, buf.size(), f);
    std::fclose(f);
  }

  ~CorresApp() {
    for (auto* c : clouds) er_cloud_free(c);
    for (auto& r : replica) for (auto* c : r) er_cloud_free(c);
  }

  int redux_index(int i, int j) const { return i + j * num; }        // GetReduxIndex, CorresApp.h:58-60

  // GetVolumeOverlapRatio (CorresApp.h:64-81): share of 20^3 cell centres of one volume that fall inside the other
  double volume_overlap_ratio(const double* T) const {
    const int res = 20;
    const double ul = length / (double)res;
    int s = 0;
    for (int i = 0; i < res; i++)
      for (int j = 0; j < res; j++)
        for (int k = 0; k < res; k++) {
          const double p[4] = {(i + 0.5) * ul, (j + 0.5) * ul, (k + 0.5) * ul, 1.0};
          double q[3];
          for (int r = 0; r < 3; r++) q[r] = ((T[r * 4] * p[0] + T[r * 4 + 1] * p[1]) + T[r * 4 + 2] * p[2]) + T[r * 4 + 3] * p[3];
          if (q[0] >= 0 && q[0] <= length && q[1] >= 0 && q[1] <= length && q[2] >= 0 && q[2] <= length) s++;
        }
    return (double)s / res / res / res;
  }

  // CCorresApp::LoadData (CorresApp.cpp:31-110)
  bool load_data(const std::string& filename, int n) {
    size_t c = filename.rfind('\\');
    if (c == std::string::npos) c = filename.rfind('/');
    dir_name = (c == std::string::npos) ? std::string() : filename.substr(0, c + 1);

    if (n > 0) {
      Trajectory temp;
      temp.load(filename);
      corres_traj.data.clear();
      num = n;
      if ((int)temp.data.size() <= (num - 1) * interval) {
        std::fprintf(stderr, "Trajectory %s has %d poses, need more than %d.\n", filename.c_str(), (int)temp.data.size(), (num - 1) * interval);
        return false;
      }
      double basepose[16] = {1, 0, 0, length / 2.0, 0, 1, 0, length / 2.0, 0, 0, 1, -0.3, 0, 0, 0, 1};
      double baseinverse[16], first_inv[16], leftbase[16];
      er_mat4_inverse(basepose, baseinverse);
      er_mat4_inverse(temp.data[0].T, first_inv);
      er_mat4_mul(basepose, first_inv, leftbase);
      std::vector<std::vector<double>> ipose(num, std::vector<double>(16)), ipose_inv(num, std::vector<double>(16));
      for (int i = 0; i < num; i++) {
        double t[16];
        er_mat4_mul(leftbase, temp.data[i * interval].T, t);
        er_mat4_mul(t, baseinverse, ipose[i].data());
        er_mat4_inverse(ipose[i].data(), ipose_inv[i].data());
      }
      for (int i = 0; i < num - 1; i++) {
        FramedTransformation ft;
        ft.id1 = i; ft.id2 = i + 1; ft.frame = num;
        er_mat4_mul(ipose_inv[i].data(), ipose[i + 1].data(), ft.T);
        corres_traj.data.push_back(ft);
        for (int j = i + 2; j < num; j++) {
          FramedTransformation fj;
          fj.id1 = i; fj.id2 = j; fj.frame = num;
          er_mat4_mul(ipose_inv[i].data(), ipose[j].data(), fj.T);
          if (volume_overlap_ratio(fj.T) > 0.3) corres_traj.data.push_back(fj);
        }
      }
      std::fprintf(stderr, "%d initial matching candidates are created.\n", (int)corres_traj.data.size());
    } else {
      corres_traj.load(filename);
      if (corres_traj.data.empty()) {
        std::fprintf(stderr, "No pairs in %s.\n", filename.c_str());
        return false;
      }
      num = corres_traj.data[0].frame;
    }

    pointclouds.assign(num, PointCloudXYZN());
    clouds.assign(num, nullptr);
    if (bbox_length > 0.0) std::printf("BBox prune enabled : %.6f\n", bbox_length);
    bool ok = true;
    const auto tl0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(8) schedule(dynamic)
    for (int i = 0; i < num; i++) {
      char fn[1024];
      std::snprintf(fn, sizeof fn, "%scloud_bin_%d.pcd", dir_name.c_str(), i);
      std::printf("Load file : %s\n", fn);
      PointCloudXYZN raw;
      if (!load_pcd_xyzn(fn, raw)) std::fprintf(stderr, "Error loading file.\n");
      PointCloudXYZN& pcd = pointclouds[i];
      for (size_t j = 0; j < raw.size(); j++) {
        if (!std::isnan(raw.nrm[3 * j])) {                       // !_isnan(normal_x), CorresApp.cpp:94
          for (int q = 0; q < 3; q++) { pcd.xyz.push_back(raw.xyz[3 * j + q]); pcd.nrm.push_back(raw.nrm[3 * j + q]); }
        }
      }
      if (save_xyzn) {
        std::snprintf(fn, sizeof fn, "%scloud_bin_xyzn_%d.xyzn", dir_name.c_str(), i);
        if (FILE* f = std::fopen(fn, "w")) {
          for (size_t k = 0; k < pcd.size(); k++)
            std::fprintf(f, "%.6f %.6f %.6f %.6f %.6f %.6f\n", pcd.xyz[3 * k], pcd.xyz[3 * k + 1], pcd.xyz[3 * k + 2],
                         pcd.nrm[3 * k], pcd.nrm[3 * k + 1], pcd.nrm[3 * k + 2]);
          std::fclose(f);
        }
      }
    }
    const auto tl1 = std::chrono::steady_clock::now();
    load_ms = std::chrono::duration<double, std::milli>(tl1 - tl0).count();
    const double cell = std::max(reg_dist, dist_thresh);
#pragma omp parallel for num_threads(8) schedule(dynamic)
    for (int i = 0; i < num; i++) {
      clouds[i] = er_cloud_upload(device, pointclouds[i].xyz.data(), pointclouds[i].nrm.data(), pointclouds[i].size(), cell);
      if (!clouds[i]) {
#pragma omp critical
        { std::cerr << "er_cloud_upload: " << er_last_error() << std::endl; ok = false; }
      }
    }
    // fragments are replicated on every further device (4.8 MB each; the pair list is what is sharded)
    replica.assign(extra_devices.size(), std::vector<er_cloud*>((size_t)num, nullptr));
    for (size_t k = 0; k < extra_devices.size() && ok; k++) {
#pragma omp parallel for num_threads(8) schedule(dynamic)
      for (int i = 0; i < num; i++) {
        replica[k][i] = er_cloud_upload(extra_devices[k], pointclouds[i].xyz.data(), pointclouds[i].nrm.data(), pointclouds[i].size(), cell);
        if (!replica[k][i]) {
#pragma omp critical
          { std::cerr << "er_cloud_upload (device " << extra_devices[k] << "): " << er_last_error() << std::endl; ok = false; }
        }
      }
    }
    upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tl1).count();
    return ok;
  }

  // CCorresApp::Blacklist (CorresApp.cpp:330-346)
  void load_blacklist(const std::string& filename) {
    blacklist.clear();
    FILE* f = std::fopen(filename.c_str(), "r");
    if (!f) return;
    char buf[1024];
    int id;
    while (std::fgets(buf, sizeof buf, f))
      if (std::strlen(buf) > 0 && buf[0] != '#' && std::sscanf(buf, "%d", &id) == 1) blacklist.insert(id);
    std::fclose(f);
  }

  // CCorresApp::Redux (CorresApp.cpp:348-359)
  void load_redux(const std::string& filename) {
    redux = true;
    redux_map.clear();
    redux_traj.load(filename);
    for (int i = 0; i < (int)redux_traj.data.size(); i++)
      redux_map.insert({redux_index(redux_traj.data[i].id1, redux_traj.data[i].id2), i});
    std::fprintf(stderr, "%d out of %d pairs are redux pairs.\n", (int)redux_traj.data.size(), (int)corres_traj.data.size());
  }

  bool blacklisted(const FramedTransformation& d) const { return blacklist.count(d.id1) || blacklist.count(d.id2); }

  // CCorresApp::Registration (CorresApp.cpp:212-319)
  bool do_registration() {
    registration = true;
    std::fprintf(stderr, "Registration with dist %.6f, num %d and ratio %.6f\n", reg_dist, reg_num, reg_ratio);
    const int total = (int)corres_traj.data.size();
    int nprocessed = 0;
    std::vector<int> todo;           // accepted pairs that need ICP
    // the gate of every candidate (CorresApp.cpp:250-264) in one batched pass per device instead of one blocking call
    // per pair; the decisions below run in list order on its counts, as the reference's pair loop prints them
    std::vector<int> cand, gate_cnt((size_t)total, 0);
    for (int i = 0; i < total; i++) {
      const FramedTransformation& d = corres_traj.data[i];
      if (!blacklisted(d) && d.frame != -1) cand.push_back(i);
    }
    {
      const auto tg0 = std::chrono::steady_clock::now();
      const int nc = (int)cand.size();
      std::vector<int> ti(nc), si(nc), cnt(nc);
      std::vector<double> Tc(16 * (size_t)nc);
      for (int p = 0; p < nc; p++) {
        const FramedTransformation& d = corres_traj.data[cand[p]];
        ti[p] = d.id1; si[p] = d.id2;
        std::memcpy(&Tc[16 * (size_t)p], d.T, sizeof d.T);
      }
      std::string err;
      if (!on_devices(nc, [&](int k, const std::vector<int>& sel, er_cloud* const* cl, std::string& e) {
            const int m = (int)sel.size();
            std::vector<int> t2(m), s2(m), c2(m);
            std::vector<double> T2(16 * (size_t)m);
            for (int q = 0; q < m; q++) {
              t2[q] = ti[sel[q]]; s2[q] = si[sel[q]];
              std::memcpy(&T2[16 * (size_t)q], &Tc[16 * (size_t)sel[q]], 16 * sizeof(double));
            }
            if (er_icp_count_inliers_batch(cl, num, t2.data(), s2.data(), m, T2.data(), reg_dist, c2.data())) { e = er_last_error(); return; }
            for (int q = 0; q < m; q++) cnt[sel[q]] = c2[q];
            (void)k;
          }, err)) {
        std::cerr << "er_icp_count_inliers_batch: " << err << std::endl;
        return false;
      }
      for (int p = 0; p < nc; p++) gate_cnt[(size_t)cand[p]] = cnt[p];
      gate_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tg0).count();
    }
    for (int i = 0; i < total; i++) {
      FramedTransformation& d = corres_traj.data[i];
      if (blacklisted(d)) {
        nprocessed++;
        d.frame = -1;
        std::printf("Blacklist pair <%d, %d> ... \n", d.id1, d.id2);
        std::fprintf(stderr, "%d / %d\n", nprocessed, total);
        continue;
      }
      if (d.frame == -1) { nprocessed++; continue; }            // means reject
      std::printf("Align pair <%d, %d> ... \n", d.id1, d.id2);
      const int cnt = gate_cnt[(size_t)i];
      const size_t n0 = pointclouds[d.id1].size(), n1 = pointclouds[d.id2].size();
      const double r1 = (double)cnt / (double)n0, r2 = (double)cnt / (double)n1;
      std::printf("    <%d, %d> : %d inliers with ratio %.2f(%d) and %.2f(%d) ... ", d.id1, d.id2, cnt, r1, (int)n0, r2, (int)n1);
      const bool accept = (cnt >= reg_num || (r1 > reg_ratio && r2 > reg_ratio));
      if (accept) {
        std::printf("accept.\n");
        d.frame = cnt;
      } else {
        std::printf("reject.\n");
        d.frame = -1;
        nprocessed++;
        std::fprintf(stderr, "%d / %d\n", nprocessed, total);
        continue;
      }
      if (redux) {
        auto it = redux_map.find(redux_index(d.id1, d.id2));
        if (it != redux_map.end()) {
          std::memcpy(d.T, redux_traj.data[it->second].T, sizeof d.T);
          nprocessed++;
          std::fprintf(stderr, "%d / %d\n", nprocessed, total);
          continue;
        }
      }
      todo.push_back(i);
    }
    if (todo.empty()) return true;
    // pcl::IterativeClosestPoint ... icp.align (CorresApp.cpp:295-312) for every accepted pair, batched
    const int np = (int)todo.size();
    std::vector<int> ti(np), si(np), iters(np);
    std::vector<double> Tin(16 * (size_t)np), Tout(16 * (size_t)np), fit(np);
    for (int p = 0; p < np; p++) {
      const FramedTransformation& d = corres_traj.data[todo[p]];
      ti[p] = d.id1; si[p] = d.id2;
      std::memcpy(&Tin[16 * (size_t)p], d.T, sizeof d.T);
    }
    const auto t0 = std::chrono::steady_clock::now();
    // pairs are independent (CorresApp.cpp:220 runs them from an OpenMP loop): share k of the list goes to device k,
    // one host thread per device, results land in the slots of the original list -- no exchange between devices
    {
      std::string err;
      if (!on_devices(np, [&](int, const std::vector<int>& sel, er_cloud* const* cl, std::string& e) {
            const int m = (int)sel.size();
            std::vector<int> t2(m), s2(m), it2(m);
            std::vector<double> Ti2(16 * (size_t)m), To2(16 * (size_t)m), f2(m);
            for (int q = 0; q < m; q++) {
              t2[q] = ti[sel[q]]; s2[q] = si[sel[q]];
              std::memcpy(&Ti2[16 * (size_t)q], &Tin[16 * (size_t)sel[q]], 16 * sizeof(double));
            }
            if (er_icp_align_batch(cl, num, t2.data(), s2.data(), m, Ti2.data(), reg_dist, 20, 1e-6, fixed_iters ? 1 : 0, To2.data(),
                                   it2.data(), f2.data())) { e = er_last_error(); return; }
            for (int q = 0; q < m; q++) {
              std::memcpy(&Tout[16 * (size_t)sel[q]], &To2[16 * (size_t)q], 16 * sizeof(double));
              iters[sel[q]] = it2[q];
              fit[sel[q]] = f2[q];
            }
          }, err)) {
        std::cerr << "er_icp_align_batch: " << err << std::endl;
        return false;
      }
    }
    icp_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    for (int p = 0; p < np; p++) {
      FramedTransformation& d = corres_traj.data[todo[p]];
      // (the reference prints pcl's getFitnessScore() here -- mean NN distance^2 of ALL source points under the final pose;
      // this is the mean over the last iteration's correspondences, hence the different label)
      std::printf("    <%d, %d> : ICP correspondence MSE is %.6f\n", d.id1, d.id2, fit[p]);
      std::printf("    Matrix from : \n");
      for (int r = 0; r < 4; r++) std::printf("%g %g %g %g\n", d.T[r * 4], d.T[r * 4 + 1], d.T[r * 4 + 2], d.T[r * 4 + 3]);
      std::memcpy(d.T, &Tout[16 * (size_t)p], sizeof d.T);
      std::printf("    To : \n");
      for (int r = 0; r < 4; r++) std::printf("%g %g %g %g\n", d.T[r * 4], d.T[r * 4 + 1], d.T[r * 4 + 2], d.T[r * 4 + 3]);
      nprocessed++;
      icp_pairs++;
      icp_iterations += iters[p];
      std::fprintf(stderr, "%d / %d\n", nprocessed, total);
    }
    return true;
  }

  // CCorresApp::FindCorrespondence (CorresApp.cpp:112-210)
  bool find_correspondence() {
    if (output_information) {
      corres_info.data.clear();
      for (const auto& d : corres_traj.data) {
        FramedInformation fi;
        fi.id1 = d.id1; fi.id2 = d.id2; fi.frame = d.frame;
        corres_info.data.push_back(fi);
      }
    }
    // the pair loop of CorresApp.cpp:121 in batches: one launch sequence and one synchronisation per batch on each
    // device, then the corres_<i>_<j>.txt files of the batch written from 8 host threads
    std::vector<int> cand;
    for (int i = 0; i < (int)corres_traj.data.size(); i++) {
      const FramedTransformation& d = corres_traj.data[i];
      if (!blacklisted(d) && d.frame != -1) cand.push_back(i);              // -1 means reject
    }
    const size_t kBatchRows = 24u << 20;                                    // rows (source points) per device and batch: 192 MB of pairs
    const int ndev = 1 + (int)extra_devices.size();
    // one page-locked result buffer per device, reused by every batch; the corres files are written straight from it
    std::vector<int32_t*> dev_buf((size_t)ndev, nullptr);
    std::vector<size_t> dev_cap((size_t)ndev, 0);
    struct FreeBufs { std::vector<int32_t*>& b; ~FreeBufs() { for (auto* p : b) er_host_free(p); } } free_bufs{dev_buf};
    for (size_t c0 = 0; c0 < cand.size();) {
      size_t c1 = c0, rows = 0;
      while (c1 < cand.size() && (c1 == c0 || rows + pointclouds[corres_traj.data[cand[c1]].id2].size() <= kBatchRows * (size_t)ndev)) {
        rows += pointclouds[corres_traj.data[cand[c1]].id2].size();
        c1++;
      }
      const int nb = (int)(c1 - c0);
      std::vector<int> ti(nb), si(nb);
      std::vector<double> Tc(16 * (size_t)nb), info(output_information ? 36 * (size_t)nb : 0);
      std::vector<const int32_t*> rows_ptr((size_t)nb, nullptr);
      std::vector<size_t> rows_n((size_t)nb, 0);
      for (int p = 0; p < nb; p++) {
        const FramedTransformation& d = corres_traj.data[cand[c0 + p]];
        ti[p] = d.id1; si[p] = d.id2;
        std::memcpy(&Tc[16 * (size_t)p], d.T, sizeof d.T);
      }
      const auto tc0 = std::chrono::steady_clock::now();
      std::string err;
      if (!on_devices(nb, [&](int k, const std::vector<int>& sel, er_cloud* const* cl, std::string& e) {
            const int m = (int)sel.size();
            std::vector<int> t2(m), s2(m);
            std::vector<double> T2(16 * (size_t)m), I2(output_information ? 36 * (size_t)m : 0);
            size_t cap = 0;
            for (int q = 0; q < m; q++) {
              t2[q] = ti[sel[q]]; s2[q] = si[sel[q]];
              std::memcpy(&T2[16 * (size_t)q], &Tc[16 * (size_t)sel[q]], 16 * sizeof(double));
              cap += pointclouds[si[sel[q]]].size();
            }
            // (a single pair larger than the batch size gets a buffer of its own size)
            const size_t want = std::max(cap, (size_t)1);
            if (!dev_buf[(size_t)k] || want > dev_cap[(size_t)k]) {
              er_host_free(dev_buf[(size_t)k]);
              dev_cap[(size_t)k] = std::max(want, kBatchRows);
              dev_buf[(size_t)k] = (int32_t*)er_host_alloc(sizeof(int32_t) * 2 * dev_cap[(size_t)k]);
              if (!dev_buf[(size_t)k]) { e = er_last_error(); dev_cap[(size_t)k] = 0; return; }
            }
            int32_t* buf = dev_buf[(size_t)k];
            std::vector<size_t> start((size_t)m + 1);
            if (er_icp_correspondences_batch(cl, num, t2.data(), s2.data(), m, T2.data(), dist_thresh, normal_thresh, buf, cap,
                                             start.data(), output_information ? I2.data() : nullptr)) { e = er_last_error(); return; }
            for (int q = 0; q < m; q++) {
              rows_ptr[(size_t)sel[q]] = buf + 2 * start[q];
              rows_n[(size_t)sel[q]] = start[q + 1] - start[q];
              if (output_information) std::memcpy(&info[36 * (size_t)sel[q]], &I2[36 * (size_t)q], 36 * sizeof(double));
            }
          }, err)) {
        std::cerr << "er_icp_correspondences_batch: " << err << std::endl;
        return false;
      }
      corres_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count();
      for (int p = 0; p < nb; p++) {
        const int i = cand[c0 + p];
        FramedTransformation& d = corres_traj.data[i];
        const size_t n = rows_n[(size_t)p];
        std::printf("Processing pair <%d, %d>\n", d.id1, d.id2);
        std::printf("    <%d, %d> : Corresponce number is %d, ratio is %.2f(%d)\n", d.id1, d.id2, (int)n, (double)n / (double)d.frame, d.frame);
        if ((double)n / (double)d.frame < 0.5) {
          std::fprintf(stderr, "    <%d, %d> : Reduced too much!!\n", d.id1, d.id2);
          d.frame = (reg_num > 0) ? -1 : (int)n;
        } else {
          d.frame = (int)n;
        }
        if (output_information) {
          corres_info.data[i].frame = d.frame;
          std::memcpy(corres_info.data[i].I, &info[36 * (size_t)p], 36 * sizeof(double));
        }
      }
      if (save_corres) {
        const auto tw0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(8) schedule(dynamic)
        for (int p = 0; p < nb; p++) {
          const FramedTransformation& d = corres_traj.data[cand[c0 + p]];
          char fn[1024];
          std::snprintf(fn, sizeof fn, "%scorres_%d_%d.txt", dir_name.c_str(), d.id1, d.id2);
          write_corres_file(fn, rows_ptr[(size_t)p], rows_n[(size_t)p]);
        }
        write_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw0).count();
      }
      c0 = c1;
    }
    return true;
  }

  // CCorresApp::Finalize (CorresApp.cpp:321-328): written to the current working directory
  void finalize() {
    corres_traj.save("reg_output.log");
    if (output_information) corres_info.save("reg_output.info");
  }
};

int main(int argc, char* argv[]) {
  if (argc == 1 || find_switch(argc, argv, "--help") || find_switch(argc, argv, "-h")) return print_help();

  CorresApp app;
  if (find_switch(argc, argv, "--save_xyzn")) app.save_xyzn = true;
  std::string log_file, reg_log_file, blacklist_file, redux_file, bench_json;
  double reg_dist = 0, reg_ratio = 0;
  int reg_num = 0, num = 0;
  parse_argument(argc, argv, "--device", app.device);
  std::string devices;
  if (parse_argument(argc, argv, "--devices", devices) > 0) {          // "0,1,2": the first is the main device
    std::vector<int> all;
    for (size_t pos = 0; pos < devices.size();) {
      size_t end = devices.find(',', pos);
      if (end == std::string::npos) end = devices.size();
      if (end > pos) all.push_back(std::atoi(devices.substr(pos, end - pos).c_str()));
      pos = end + 1;
    }
    if (!all.empty()) {
      app.device = all[0];
      app.extra_devices.assign(all.begin() + 1, all.end());
    }
  }
  parse_argument(argc, argv, "--bench_json", bench_json);
  app.fixed_iters = find_switch(argc, argv, "--fixed_iters");

  if ((parse_argument(argc, argv, "--traj", log_file) > 0 && parse_argument(argc, argv, "--num", num) > 0) ||
      parse_argument(argc, argv, "--reg_traj", reg_log_file) > 0) {
    const auto t0 = std::chrono::steady_clock::now();
    if (parse_argument(argc, argv, "--reg_dist", reg_dist) > 0) {
      app.reg_dist = reg_dist;
      app.dist_thresh = reg_dist / 2.0;
    }
    if (parse_argument(argc, argv, "--reg_ratio", reg_ratio) > 0) app.reg_ratio = reg_ratio;
    if (parse_argument(argc, argv, "--reg_num", reg_num) > 0) app.reg_num = reg_num;
    parse_argument(argc, argv, "--length", app.length);
    parse_argument(argc, argv, "--interval", app.interval);

    bool ok = reg_log_file.length() > 0 ? app.load_data(reg_log_file, -1) : app.load_data(log_file, num);
    if (parse_argument(argc, argv, "--blacklist", blacklist_file) > 0) app.load_blacklist(blacklist_file);
    if (parse_argument(argc, argv, "--redux", redux_file) > 0) app.load_redux(redux_file);
    if (find_switch(argc, argv, "--output_information")) app.output_information = true;
    const auto t1 = std::chrono::steady_clock::now();
    if (ok && find_switch(argc, argv, "--registration")) {
      ok = app.do_registration();
      std::cerr << "Neat Registration took "
                << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count() << "ms." << std::endl;
    }
    if (ok) ok = app.find_correspondence();
    if (ok) app.finalize();
    const auto t2 = std::chrono::steady_clock::now();
    std::cerr << "Registration All took " << std::chrono::duration<double, std::milli>(t2 - t0).count() << "ms." << std::endl;
    if (!bench_json.empty()) {
      if (FILE* f = std::fopen(bench_json.c_str(), "w")) {
        std::fprintf(f, "{\"pairs\": %d, \"icp_pairs\": %ld, \"icp_iterations\": %ld, \"load_ms\": %.3f, \"upload_ms\": %.3f, "
                        "\"gate_ms\": %.3f, \"icp_ms\": %.3f, \"corres_ms\": %.3f, \"write_ms\": %.3f, \"total_ms\": %.3f, \"devices\": %d}\n",
                     (int)app.corres_traj.data.size(), app.icp_pairs, app.icp_iterations, app.load_ms, app.upload_ms, app.gate_ms,
                     app.icp_ms, app.corres_ms, app.write_ms, std::chrono::duration<double, std::milli>(t2 - t0).count(),
                     1 + (int)app.extra_devices.size());
        std::fclose(f);
      }
    }
  }
  return 0;
}
