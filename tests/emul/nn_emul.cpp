// nn_emul.cpp -- TEST ONLY: runs the ICP path's nearest-neighbour search on the CPU.
//
// Grid geometry, cell assignment, candidate comparison and the pruned neighbourhood walk are the product's own
// source, elasticreconstruction_b200/csrc/er_nn_math.h, compiled here by g++ (-ffp-contract=off).  This file adds
//   * the counting sort of the points by cell (on the device: grid_count_kernel / scan / grid_scatter_kernel);
//   * a sequential restatement of the warp-cooperative search of icp_iter_kernel<COOP = true> for ONE query: own
//     cell -> bound, cell columns nearest first, each column's z run derived from the bound at that moment, scan,
//     merge -- the same inequalities through the same header functions (slab_gap, nn_scan_key, pack_key), without
//     the warp-level list handling.
// tests/test_nn_math.py compares both against a brute-force search.
#include <cstddef>
#include <vector>

#include "er_nn_math.h"

int er_voxel_rcp_bias = 0;   // (er_emul.h knob, unused here)

namespace {
struct Grid {
  GridDev g{};
  std::vector<er_float4> pts;
  std::vector<int> cell_start;
  double radius = 0;
};
constexpr int kNoMatch = 0x7fffffff;
}  // namespace

extern "C" {

void* emul_grid_build(const float* xyz, int n, double radius, int rings) {
  Grid* G = new Grid();
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = 0; i < n; i++)
    for (int a = 0; a < 3; a++) {
      const float v = xyz[3 * (size_t)i + a];
      if (v < lo[a]) lo[a] = v;
      if (v > hi[a]) hi[a] = v;
    }
  for (int a = 0; a < 3; a++)
    if (n == 0 || !(lo[a] <= hi[a])) lo[a] = hi[a] = 0.f;
  G->radius = er_grid_geometry(lo, hi, radius, rings, G->g);
  const GridDev& g = G->g;
  const size_t ncell = (size_t)g.dim[0] * g.dim[1] * g.dim[2];
  std::vector<int> cid((size_t)n);
  G->cell_start.assign(ncell + 1, 0);
  for (int i = 0; i < n; i++) {
    cid[i] = (grid_cell_axis(xyz[3 * (size_t)i], g.org[0], g.inv_cell, g.dim[0]) * g.dim[1] +
              grid_cell_axis(xyz[3 * (size_t)i + 1], g.org[1], g.inv_cell, g.dim[1])) * g.dim[2] +
             grid_cell_axis(xyz[3 * (size_t)i + 2], g.org[2], g.inv_cell, g.dim[2]);
    G->cell_start[(size_t)cid[i] + 1]++;
  }
  for (size_t c = 0; c < ncell; c++) G->cell_start[c + 1] += G->cell_start[c];
  std::vector<int> fill(G->cell_start.begin(), G->cell_start.end() - 1);
  G->pts.resize(n ? n : 1);
  // points of one cell in DEcreasing index order: the result must not depend on the order inside a cell
  for (int i = n - 1; i >= 0; i--)
    G->pts[(size_t)fill[cid[i]]++] = er_float4{xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], __int_as_float(i)};
  G->g.pts = G->pts.data();
  G->g.cell_start = G->cell_start.data();
  return G;
}

void emul_grid_free(void* h) { delete (Grid*)h; }
double emul_grid_radius(void* h) { return ((Grid*)h)->radius; }
int emul_grid_rings(void* h) { return ((Grid*)h)->g.rings; }

// nn_search (the per-thread walk used by the gate / correspondence kernels and icp_iter_kernel<COOP = false>)
void emul_nn_search(void* h, const float* q, int m, int* idx, float* d2) {
  const GridDev& g = ((Grid*)h)->g;
#pragma omp parallel for
  for (int i = 0; i < m; i++) {
    float bd;
    const int s = nn_search(g, q[3 * (size_t)i], q[3 * (size_t)i + 1], q[3 * (size_t)i + 2], bd);
    idx[i] = s >= 0 ? __float_as_int(g.pts[s].w) : -1;
    d2[i] = bd;
  }
}

// exploration counters (not product code): candidate points compared and cell columns examined by emul_nn_coop
static long long g_points = 0, g_items = 0, g_runs = 0;
void emul_nn_counters(long long* out, int reset) {
  out[0] = g_points; out[1] = g_items; out[2] = g_runs;
  if (reset) g_points = g_items = g_runs = 0;
}
// exploration: seed[i] >= 0 names a target point (original index) whose distance bounds the search from the start
static int* g_trace = nullptr;                 // per query: centre-cell points, xlo, xhi, ylo, yhi, runs scanned, points in runs
void emul_nn_set_trace(int* trace7) { g_trace = trace7; }
static const int* g_seed = nullptr;
static const float* g_seed_xyz = nullptr;      // target points in original order
void emul_nn_set_seed(const int* seed, const float* target_xyz) { g_seed = seed; g_seed_xyz = target_xyz; }

// the cooperative search of icp_iter_kernel<COOP = true>, one query at a time; bd_init as in er_icp_align_batch
void emul_nn_coop(void* h, const float* q, int m, float bd_init, int* idx, float* d2) {
  const GridDev& g = ((Grid*)h)->g;
  const int K = g.rings;
  const float cell = g.cell, slack = 1e-5f * g.cell;
#pragma omp parallel for
  for (int i = 0; i < m; i++) {
    const float sx = q[3 * (size_t)i], sy = q[3 * (size_t)i + 1], sz = q[3 * (size_t)i + 2];
    unsigned long long key = pack_key(bd_init, kNoMatch);
    long long n_pts = 0, n_items = 0, n_runs = 0;
    if (g_seed && g_seed[i] >= 0) {
      const float* t = g_seed_xyz + 3 * (size_t)g_seed[i];
      const unsigned long long k2 = nn_key(er_float4{t[0], t[1], t[2], __int_as_float(g_seed[i])}, sx, sy, sz);
      if (k2 < key) key = k2;
      n_pts++;
    }
    const float lx = __fmul_rn(__fsub_rn(sx, g.org[0]), g.inv_cell);
    const float ly = __fmul_rn(__fsub_rn(sy, g.org[1]), g.inv_cell);
    const float lz = __fmul_rn(__fsub_rn(sz, g.org[2]), g.inv_cell);
    const float flx = floorf(lx), fly = floorf(ly), flz = floorf(lz);
    const bool valid = flx >= -(float)(K + 1) && flx <= (float)(g.dim[0] + K) && fly >= -(float)(K + 1) && fly <= (float)(g.dim[1] + K) &&
                       flz >= -(float)(K + 1) && flz <= (float)(g.dim[2] + K);
    if (valid) {
      const int cx = (int)flx, cy = (int)fly, cz = (int)flz;
      const float fx = lx - flx, fy = ly - fly, fz = lz - flz;
      float bd = __uint_as_float((unsigned)(key >> 32));
      if (cx >= 0 && cx < g.dim[0] && cy >= 0 && cy < g.dim[1] && cz >= 0 && cz < g.dim[2]) {
        const int c = (cx * g.dim[1] + cy) * g.dim[2] + cz;
        nn_scan_key(g, g.cell_start[c], g.cell_start[c + 1], sx, sy, sz, key);
        n_pts += g.cell_start[c + 1] - g.cell_start[c];
        if (g_trace) g_trace[7 * (size_t)i] = g.cell_start[c + 1] - g.cell_start[c];
        bd = __uint_as_float((unsigned)(key >> 32));
      }
      int xlo = 0, xhi = 0, ylo = 0, yhi = 0;
      for (int o = 1; o <= K; o++) { const float a = slab_gap(o, fx, cell, slack); if (a * a > bd) break; xhi = o; }
      for (int o = 1; o <= K; o++) { const float a = slab_gap(-o, fx, cell, slack); if (a * a > bd) break; xlo = -o; }
      for (int o = 1; o <= K; o++) { const float a = slab_gap(o, fy, cell, slack); if (a * a > bd) break; yhi = o; }
      for (int o = 1; o <= K; o++) { const float a = slab_gap(-o, fy, cell, slack); if (a * a > bd) break; ylo = -o; }
      if (g_trace) { int* tr = g_trace + 7 * (size_t)i; tr[1] = xlo; tr[2] = xhi; tr[3] = ylo; tr[4] = yhi; }
      for (int t = 0; t < 25; t++) {
        const int ox = kColumnOrder[t][0], oy = kColumnOrder[t][1];
        if (!(ox >= xlo && ox <= xhi && oy >= ylo && oy <= yhi)) continue;
        for (int upper = 0; upper < (t == 0 ? 2 : 1); upper++) {
          const int ix = cx + ox, iy = cy + oy;
          n_items++;
          if (ix < 0 || ix >= g.dim[0] || iy < 0 || iy >= g.dim[1]) continue;
          const float bl = __uint_as_float((unsigned)(key >> 32));
          const float ax = slab_gap(ox, fx, cell, slack), ay = slab_gap(oy, fy, cell, slack);
          const float axy2 = ax * ax + ay * ay;
          if (axy2 > bl) continue;
          int zhi = cz, zlo = cz;
          for (int oz = 1; oz <= K; oz++) {
            const float az = slab_gap(oz, fz, cell, slack);
            if (axy2 + az * az > bl) break;
            zhi = cz + oz;
          }
          for (int oz = 1; oz <= K; oz++) {
            const float az = slab_gap(-oz, fz, cell, slack);
            if (axy2 + az * az > bl) break;
            zlo = cz - oz;
          }
          int a0 = max(zlo, 0), a1 = min(zhi, g.dim[2] - 1);
          if (t == 0) {
            if (upper) a0 = max(cz + 1, 0); else a1 = min(cz - 1, g.dim[2] - 1);
          }
          if (a0 > a1) continue;
          const int row = (ix * g.dim[1] + iy) * g.dim[2];
          nn_scan_key(g, g.cell_start[row + a0], g.cell_start[row + a1 + 1], sx, sy, sz, key);
          n_pts += g.cell_start[row + a1 + 1] - g.cell_start[row + a0];
          n_runs += g.cell_start[row + a1 + 1] > g.cell_start[row + a0];
          if (g_trace && g.cell_start[row + a1 + 1] > g.cell_start[row + a0]) {
            g_trace[7 * (size_t)i + 5]++;
            g_trace[7 * (size_t)i + 6] += g.cell_start[row + a1 + 1] - g.cell_start[row + a0];
          }
        }
      }
    }
#pragma omp atomic
    g_points += n_pts;
#pragma omp atomic
    g_items += n_items;
#pragma omp atomic
    g_runs += n_runs;
    const int best = (int)(unsigned)(key & 0xffffffffu);
    idx[i] = best == kNoMatch ? -1 : best;
    d2[i] = __uint_as_float((unsigned)(key >> 32));
  }
}

}  // extern "C"
