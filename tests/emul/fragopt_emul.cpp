// fragopt_emul.cpp -- TEST ONLY: a CPU emulation of the er_fragopt_* part of libelastic_b200's C ABI, built from the
// product's own arithmetic header elasticreconstruction_b200/csrc/er_fragopt_math.h.  The CUDA kernels of
// csrc/er_fragopt.cu are re-enacted thread by thread: the same row formation, the same 4x4 tile ownership, the same
// batch / break / flush walk over the sorted correspondences, with the CTA barriers turned into loops over tid.
// tests/test_fragopt_cpu.py (a) calls it through ctypes against a numpy restatement of the reference's formulas and
// (b) links apps/FragmentOptimizer.cpp against it and compares the executable's pose.log / output.ctr with the
// VERBATIM reference binary's (oracle/_ref/FragmentOptimizer_ref).  Never built into the product.
//
// What it cannot cover is the CUDA glue of er_fragopt.cu (launch geometry, cub radix sort, cuSOLVER): that is the
// business of tests/test_z_fragopt_gpu.py.
#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <vector>

#include "elastic_b200.h"
#include "er_fragopt_math.h"

using namespace er_fo;

namespace er { void set_error(const char* fmt, ...); }   // csrc/er_host_math.cpp (er_last_error lives there too)
#define set_err er::set_error

namespace {
struct Corr { int gi, gj, pair; };
struct HostAdd { void operator()(double* p, double v) const { *p += v; } };
constexpr int kThreads = 128;
}  // namespace

struct er_fragopt {
  int num = 0, res = 0, M = 0;
  Lattice L;
  std::vector<std::vector<StaticPoint>> stat;     // per fragment until frozen
  std::vector<std::vector<float>> p3, n3;
  bool frozen = false;
  std::vector<int> cloud_off, frag_of;
  std::vector<StaticPoint> S;
  std::vector<float> P, N;                        // 3 floats per point
  int npairs = 0;
  std::vector<int> pair_i, pair_j;
  std::vector<long long> pair_start;
  std::vector<Corr> corr;                         // sorted
  std::vector<double> JJ, base;
  std::vector<double> AA;                         // nonrigid: dense N x N (upper), Cholesky factor after nr_factor
  bool nr_factored = false;
  bool have_base = false;
  int emul_ctas = 7;                              // ranges of the persistent CTAs (any value must give the same sums)
  uint64_t launches = 0, flushes = 0;
};

namespace {
void freeze(er_fragopt* h) {
  if (h->frozen) return;
  h->cloud_off.assign(h->num + 1, 0);
  for (int l = 0; l < h->num; ++l) h->cloud_off[l + 1] = h->cloud_off[l] + (int)h->stat[l].size();
  for (int l = 0; l < h->num; ++l) {
    h->S.insert(h->S.end(), h->stat[l].begin(), h->stat[l].end());
    h->P.insert(h->P.end(), h->p3[l].begin(), h->p3[l].end());
    h->N.insert(h->N.end(), h->n3[l].begin(), h->n3[l].end());
    h->frag_of.insert(h->frag_of.end(), h->stat[l].size(), l);
  }
  h->frozen = true;
}
}  // namespace

extern "C" {

er_fragopt* er_fragopt_create(int, int num, int resolution, double length) {
  if (num <= 0 || resolution <= 0 || !(length > 0)) { set_err("er_fragopt_create: bad arguments"); return nullptr; }
  er_fragopt* h = new er_fragopt();
  h->num = num; h->res = resolution;
  h->L = fo_lattice(resolution, (float)length);
  h->M = 6 * num + h->L.nper;
  h->stat.resize(num); h->p3.resize(num); h->n3.resize(num);
  if (const char* e = getenv("ER_FRAGOPT_EMUL_CTAS")) h->emul_ctas = std::max(1, atoi(e));
  return h;
}
void er_fragopt_destroy(er_fragopt* h) { delete h; }

int er_fragopt_set_cloud(er_fragopt* h, int l, const float* xyzn, size_t n, size_t* n_kept) {
  if (!h || l < 0 || l >= h->num || h->frozen) { set_err("er_fragopt_set_cloud: bad arguments"); return ER_ERR_ARG; }
  h->stat[l].clear(); h->p3[l].clear(); h->n3[l].clear();
  for (size_t k = 0; k < n; ++k) {
    StaticPoint s;
    float p[3], nn[3];
    const bool ok = fo_get_coordinate(h->L, xyzn + 6 * k, &s, p, nn);
    h->stat[l].push_back(s);
    h->p3[l].insert(h->p3[l].end(), p, p + 3);
    h->n3[l].insert(h->n3[l].end(), nn, nn + 3);
    if (!ok) break;
  }
  if (n_kept) *n_kept = h->stat[l].size();
  return ER_OK;
}

int er_fragopt_set_pairs(er_fragopt* h, int npairs, const int* idx0, const int* idx1, const size_t* row_start, const int32_t* rows) {
  if (!h) return ER_ERR_ARG;
  freeze(h);
  h->npairs = npairs;
  h->pair_i.assign(idx0, idx0 + npairs);
  h->pair_j.assign(idx1, idx1 + npairs);
  h->pair_start.assign(npairs + 1, 0);
  std::vector<Corr> c;
  for (int p = 0; p < npairs; ++p) {
    h->pair_start[p] = (long long)row_start[p];
    const int o0 = h->cloud_off[idx0[p]], n0 = h->cloud_off[idx0[p] + 1] - o0;
    const int o1 = h->cloud_off[idx1[p]], n1 = h->cloud_off[idx1[p] + 1] - o1;
    for (size_t k = row_start[p]; k < row_start[p + 1]; ++k) {
      const int a = rows[2 * k], b = rows[2 * k + 1];
      if (a < 0 || a >= n0 || b < 0 || b >= n1) { set_err("er_fragopt_set_pairs: row outside its clouds"); return ER_ERR_ARG; }
      c.push_back(Corr{o0 + a, o1 + b, p});
    }
  }
  h->pair_start[npairs] = npairs ? (long long)row_start[npairs] : 0;
  std::stable_sort(c.begin(), c.end(), [&](const Corr& x, const Corr& y) {
    if (x.pair != y.pair) return x.pair < y.pair;
    if (h->S[x.gi].base != h->S[y.gi].base) return h->S[x.gi].base < h->S[y.gi].base;
    return h->S[x.gj].base < h->S[y.gj].base;
  });
  h->corr.swap(c);
  h->launches += 3;
  return ER_OK;
}

int er_fragopt_update_pose(er_fragopt* h, const double* T) {
  freeze(h);
  for (size_t g = 0; g < h->S.size(); ++g) {
    float m[16];
    for (int e = 0; e < 16; ++e) m[e] = (float)T[16 * h->frag_of[g] + e];
    fo_update_pose(m, &h->P[3 * g], &h->N[3 * g]);
  }
  ++h->launches;
  return ER_OK;
}

static int update_lattice(er_fragopt* h, const double* ctr, int mode) {
  freeze(h);
  for (size_t g = 0; g < h->S.size(); ++g) {
    const double* c = ctr + (size_t)h->frag_of[g] * h->L.nper;
    fo_update_normal(h->L, h->S[g], c, &h->N[3 * g]);
    if (mode == 1) {
      double pos[3];
      fo_update_point(h->L, h->S[g], c, pos);
      for (int i = 0; i < 3; ++i) h->P[3 * g + i] = (float)pos[i];
    }
  }
  ++h->launches;
  return ER_OK;
}
int er_fragopt_update_normals(er_fragopt* h, const double* ctr) { return update_lattice(h, ctr, 0); }
int er_fragopt_update_points_pn(er_fragopt* h, const double* ctr) { return update_lattice(h, ctr, 1); }

int er_fragopt_get_points(er_fragopt* h, int l, float* xyzn, size_t cap_rows, const double* ctr) {
  freeze(h);
  const int first = h->cloud_off[l], n = h->cloud_off[l + 1] - first;
  if ((size_t)n > cap_rows) { set_err("er_fragopt_get_points: buffer too small"); return ER_ERR_ARG; }
  for (int t = 0; t < n; ++t) {
    const int g = first + t;
    float p[3] = {h->P[3 * g], h->P[3 * g + 1], h->P[3 * g + 2]};
    if (ctr) {
      double pos[3];
      fo_update_point(h->L, h->S[g], ctr + (size_t)l * h->L.nper, pos);
      for (int i = 0; i < 3; ++i) p[i] = (float)pos[i];
    }
    for (int i = 0; i < 3; ++i) { xyzn[6 * (size_t)t + i] = p[i]; xyzn[6 * (size_t)t + 3 + i] = h->N[3 * g + i]; }
  }
  return ER_OK;
}

// slac_assemble_kernel, one CTA after the other, its barriers turned into loops over tid
int er_fragopt_slac_assemble(er_fragopt* h, const double* rot_t, double* Jb, double* score) {
  const int M = h->M;
  h->JJ.assign((size_t)M * M, 0.0);
  std::fill(Jb, Jb + M, 0.0);
  *score = 0.0;
  h->flushes = 0;
  const long long total = (long long)h->corr.size();
  const long long batches = (total + kThreads - 1) / kThreads;
  const int grid = (int)std::min<long long>(batches, h->emul_ctas);
  std::vector<double> rows((size_t)kThreads * kRowStride);
  std::vector<SlacKey> keys(kThreads);
  std::vector<int> brk(kThreads);
  for (int cta = 0; cta < grid; ++cta) {
    const long long begin = total * (long long)cta / grid, end = total * (long long)(cta + 1) / grid;
    double acc[kThreads][16];
    int ty[kThreads], tx[kThreads];
    for (int tid = 0; tid < kThreads; ++tid) {
      fo_tile_of(tid < 120 ? tid : 0, &ty[tid], &tx[tid]);
      for (int e = 0; e < 16; ++e) acc[tid][e] = 0.0;
    }
    SlacKey cur;
    cur.pair = -1; cur.bi = -1; cur.bj = -1;
    auto flush = [&]() {
      for (int tid = 0; tid < kThreads; ++tid)
        fo_slac_flush(h->L, h->num, M, h->pair_i[cur.pair], h->pair_j[cur.pair], cur, tid, ty[tid], tx[tid], acc[tid], h->JJ.data(), Jb,
                      score, HostAdd());
      ++h->flushes;
    };
    for (long long base = begin; base < end; base += kThreads) {
      const int cnt = (int)std::min<long long>(kThreads, end - base);
      for (int tid = 0; tid < cnt; ++tid) {
        const Corr& c = h->corr[base + tid];
        const StaticPoint &si = h->S[c.gi], &sj = h->S[c.gj];
        fo_slac_row(&h->P[3 * c.gi], &h->N[3 * c.gi], si.r, &h->P[3 * c.gj], sj.r, rot_t + 9 * h->pair_i[c.pair], rot_t + 9 * h->pair_j[c.pair],
                    &rows[(size_t)tid * kRowStride], tid & 7);
        keys[tid].pair = c.pair; keys[tid].bi = si.base; keys[tid].bj = sj.base;
      }
      for (int tid = 0; tid < cnt; ++tid) brk[tid] = !(keys[tid] == (tid == 0 ? cur : keys[tid - 1]));
      for (int c = 0; c < cnt; ++c) {
        if (brk[c]) {
          if (cur.pair >= 0) flush();
          for (int tid = 0; tid < kThreads; ++tid) for (int e = 0; e < 16; ++e) acc[tid][e] = 0.0;
          cur = keys[c];
        }
        for (int tid = 0; tid < kThreads; ++tid) fo_slac_mac(tid, ty[tid], tx[tid], &rows[(size_t)c * kRowStride], c & 7, acc[tid]);
      }
    }
    if (cur.pair >= 0) flush();
  }
  ++h->launches;
  return ER_OK;
}

int er_fragopt_slac_accumulate(er_fragopt* dst, er_fragopt* src) {
  if (!dst || !src || dst == src || dst->JJ.empty() || dst->JJ.size() != src->JJ.size()) { set_err("er_fragopt_slac_accumulate: bad handles"); return ER_ERR_ARG; }
  for (size_t k = 0; k < dst->JJ.size(); ++k) dst->JJ[k] += src->JJ[k];
  return ER_OK;
}

int er_fragopt_slac_get_jj(er_fragopt* h, double* jj) {
  if (h->JJ.empty()) { set_err("er_fragopt_slac_get_jj: nothing assembled yet"); return ER_ERR_ARG; }
  memcpy(jj, h->JJ.data(), h->JJ.size() * sizeof(double));
  return ER_OK;
}

// dense Cholesky of the upper triangle (row-major), A = U^T U
static bool chol_solve(int n, std::vector<double>& A, const double* rhs, double* x) {
  for (int k = 0; k < n; ++k) {
    double d = A[(size_t)k * n + k];
    for (int s = 0; s < k; ++s) d -= A[(size_t)s * n + k] * A[(size_t)s * n + k];
    if (!(d > 0)) return false;
    d = sqrt(d);
    A[(size_t)k * n + k] = d;
#pragma omp parallel for schedule(static)
    for (int c = k + 1; c < n; ++c) {
      double v = A[(size_t)k * n + c];
      for (int s = 0; s < k; ++s) v -= A[(size_t)s * n + k] * A[(size_t)s * n + c];
      A[(size_t)k * n + c] = v / d;
    }
  }
  std::vector<double> y(n);
  for (int k = 0; k < n; ++k) {             // U^T y = rhs
    double v = rhs[k];
    for (int s = 0; s < k; ++s) v -= A[(size_t)s * n + k] * y[s];
    y[k] = v / A[(size_t)k * n + k];
  }
  for (int k = n - 1; k >= 0; --k) {        // U x = y
    double v = y[k];
    for (int c = k + 1; c < n; ++c) v -= A[(size_t)k * n + c] * x[c];
    x[k] = v / A[(size_t)k * n + k];
  }
  return true;
}

int er_fragopt_slac_solve(er_fragopt* h, const double* base, const double* rhs, double* x) {
  const int M = h->M;
  if (base) { h->base.assign(base, base + (size_t)M * M); h->have_base = true; }
  if (!h->have_base || h->JJ.empty()) { set_err("er_fragopt_slac_solve: no base / nothing assembled"); return ER_ERR_ARG; }
  std::vector<double> A((size_t)M * M, 0.0);
  for (int r = 0; r < M; ++r)
    for (int c = r; c < M; ++c) A[(size_t)r * M + c] = h->base[(size_t)r * M + c] + h->JJ[(size_t)r * M + c];
  if (!chol_solve(M, A, rhs, x)) { set_err("er_fragopt_slac_solve: not positive definite"); return ER_ERR_ARG; }
  return ER_OK;
}

// rigid_assemble_kernel: per pair 91 sums, then the host scatter of er_fragopt_rigid_assemble
int er_fragopt_rigid_assemble(er_fragopt* h, double* JJ, double* Jb, double* score) {
  const int n6 = 6 * h->num;
  std::fill(JJ, JJ + (size_t)n6 * n6, 0.0);
  std::fill(Jb, Jb + n6, 0.0);
  *score = 0.0;
  std::vector<double> sums((size_t)std::max(h->npairs, 1) * 91, 0.0);
  for (const Corr& c : h->corr) {
    double val[12], b;
    fo_rigid_row(&h->P[3 * c.gi], &h->N[3 * c.gi], &h->P[3 * c.gj], val, &b);
    double* s = &sums[(size_t)c.pair * 91];
    int e = 0;
    for (int r = 0; r < 12; ++r)
      for (int q = r; q < 12; ++q) s[e++] += val[r] * val[q];
    for (int r = 0; r < 12; ++r) s[78 + r] += b * val[r];
    s[90] += b * b;
  }
  for (int p = 0; p < h->npairs; ++p) {
    const double* s = &sums[(size_t)p * 91];
    int idx[12];
    for (int r = 0; r < 6; ++r) { idx[r] = h->pair_i[p] * 6 + r; idx[6 + r] = h->pair_j[p] * 6 + r; }
    int e = 0;
    for (int r = 0; r < 12; ++r)
      for (int c = r; c < 12; ++c, ++e) {
        const int a = idx[r], b = idx[c];
        if (r == c) JJ[(size_t)a * n6 + a] += s[e];
        else if (a == b) JJ[(size_t)a * n6 + a] += 2 * s[e];
        else JJ[(size_t)std::min(a, b) * n6 + std::max(a, b)] += s[e];
      }
    for (int r = 0; r < 12; ++r) Jb[idx[r]] += s[78 + r];
    *score += s[90];
  }
  ++h->launches;
  return ER_OK;
}

// nonrigid_assemble_kernel, CTA after CTA
int er_fragopt_nonrigid_assemble(er_fragopt* h, double weight) {
  const long long N = (long long)h->num * h->L.nper;
  h->AA.assign((size_t)N * N, 0.0);
  h->nr_factored = false;
  h->flushes = 0;
  const long long total = (long long)h->corr.size();
  const long long batches = (total + kNrThreads - 1) / kNrThreads;
  const int grid = (int)std::min<long long>(batches, h->emul_ctas);
  std::vector<double> rows((size_t)kNrThreads * kNrStride);
  std::vector<SlacKey> keys(kNrThreads);
  std::vector<int> brk(kNrThreads);
  for (int cta = 0; cta < grid; ++cta) {
    const long long begin = total * (long long)cta / grid, end = total * (long long)(cta + 1) / grid;
    double acc[kNrThreads][16];
    int ty[kNrThreads], tx[kNrThreads];
    for (int tid = 0; tid < kNrThreads; ++tid) {
      fo_nr_tile_of(tid < kNrTiles ? tid : 0, &ty[tid], &tx[tid]);
      for (int e = 0; e < 16; ++e) acc[tid][e] = 0.0;
    }
    SlacKey cur;
    cur.pair = -1; cur.bi = -1; cur.bj = -1;
    auto flush = [&]() {
      for (int tid = 0; tid < kNrThreads; ++tid)
        fo_nr_flush(h->L, N, h->pair_i[cur.pair], h->pair_j[cur.pair], cur, tid, ty[tid], tx[tid], acc[tid], h->AA.data(), HostAdd());
      ++h->flushes;
    };
    for (long long base = begin; base < end; base += kNrThreads) {
      const int cnt = (int)std::min<long long>(kNrThreads, end - base);
      for (int tid = 0; tid < cnt; ++tid) {
        const Corr& c = h->corr[base + tid];
        const StaticPoint &si = h->S[c.gi], &sj = h->S[c.gj];
        fo_nonrigid_row(si.r, &h->N[3 * c.gi], sj.r, weight, &rows[(size_t)tid * kNrStride], tid & 7);
        keys[tid].pair = c.pair; keys[tid].bi = si.base; keys[tid].bj = sj.base;
      }
      for (int tid = 0; tid < cnt; ++tid) brk[tid] = !(keys[tid] == (tid == 0 ? cur : keys[tid - 1]));
      for (int c = 0; c < cnt; ++c) {
        if (brk[c]) {
          if (cur.pair >= 0) flush();
          for (int tid = 0; tid < kNrThreads; ++tid) for (int e = 0; e < 16; ++e) acc[tid][e] = 0.0;
          cur = keys[c];
        }
        for (int tid = 0; tid < kNrTiles; ++tid) fo_tile_mac<kNrStride>(ty[tid], tx[tid], &rows[(size_t)c * kNrStride], c & 7, acc[tid]);
      }
    }
    if (cur.pair >= 0) flush();
  }
  ++h->launches;
  return ER_OK;
}

int er_fragopt_nonrigid_get_aa(er_fragopt* h, double* aa) {
  if (h->AA.empty() || h->nr_factored) { set_err("er_fragopt_nonrigid_get_aa: assemble first (and before factoring)"); return ER_ERR_ARG; }
  memcpy(aa, h->AA.data(), h->AA.size() * sizeof(double));
  return ER_OK;
}

// in-place Cholesky of the upper triangle (A = U^T U), then solves against the stored factor
static bool chol_factor(int n, std::vector<double>& A) {
  for (int k = 0; k < n; ++k) {
    double d = A[(size_t)k * n + k];
    for (int s = 0; s < k; ++s) d -= A[(size_t)s * n + k] * A[(size_t)s * n + k];
    if (!(d > 0)) return false;
    d = sqrt(d);
    A[(size_t)k * n + k] = d;
#pragma omp parallel for schedule(static)
    for (int c = k + 1; c < n; ++c) {
      double v = A[(size_t)k * n + c];
      for (int s = 0; s < k; ++s) v -= A[(size_t)s * n + k] * A[(size_t)s * n + c];
      A[(size_t)k * n + c] = v / d;
    }
  }
  return true;
}

int er_fragopt_nonrigid_factor(er_fragopt* h, size_t nnz, const int32_t* rows, const int32_t* cols, const double* vals) {
  if (h->AA.empty() || h->nr_factored) { set_err("er_fragopt_nonrigid_factor: assemble first"); return ER_ERR_ARG; }
  const long long N = (long long)h->num * h->L.nper;
  for (size_t k = 0; k < nnz; ++k) {
    if (rows[k] < 0 || rows[k] > cols[k] || cols[k] >= N) { set_err("er_fragopt_nonrigid_factor: triplets must lie in the upper triangle"); return ER_ERR_ARG; }
    h->AA[(size_t)rows[k] * N + cols[k]] += vals[k];
  }
  if (!chol_factor((int)N, h->AA)) { set_err("er_fragopt_nonrigid_factor: not positive definite"); return ER_ERR_ARG; }
  h->nr_factored = true;
  return ER_OK;
}

int er_fragopt_nonrigid_solve(er_fragopt* h, const double* rhs, double* x) {
  if (!h->nr_factored) { set_err("er_fragopt_nonrigid_solve: factor first"); return ER_ERR_ARG; }
  const int n = h->num * h->L.nper;
  const std::vector<double>& A = h->AA;
  std::vector<double> y(n);
  for (int k = 0; k < n; ++k) {
    double v = rhs[k];
    for (int s = 0; s < k; ++s) v -= A[(size_t)s * n + k] * y[s];
    y[k] = v / A[(size_t)k * n + k];
  }
  for (int k = n - 1; k >= 0; --k) {
    double v = y[k];
    for (int c = k + 1; c < n; ++c) v -= A[(size_t)k * n + c] * x[c];
    x[k] = v / A[(size_t)k * n + k];
  }
  return ER_OK;
}

int er_fragopt_event_record(er_fragopt*, int) { return ER_OK; }
int er_fragopt_event_elapsed_ms(er_fragopt*, int, int, float* ms) { if (ms) *ms = 0.0f; return ER_OK; }
uint64_t er_fragopt_launch_count(const er_fragopt* h) { return h ? h->launches : 0; }
int er_fragopt_sync(er_fragopt*) { return ER_OK; }
int er_fragopt_last_assemble_stats(er_fragopt* h, float* ms, uint64_t* rows, uint64_t* flushes) {
  if (ms) *ms = 0.0f;
  if (rows) *rows = h->corr.size();
  if (flushes) *flushes = h->flushes;
  return ER_OK;
}

}  // extern "C"
