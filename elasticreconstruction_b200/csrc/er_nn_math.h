// er_nn_math.h -- the nearest-neighbour search of the ICP path (stand-in for pcl::KdTreeFLANN::nearestKSearch with
// k = 1 as CorresApp.cpp:259,153 and PCL's CorrespondenceEstimation use it): uniform grid geometry, cell assignment,
// the pruned neighbourhood walk and the candidate comparison.  Shared by the kernels of er_icp.cu (nvcc, device code)
// and by the CPU test that checks this very text against a brute-force search (g++, tests/emul/nn_emul.cpp,
// tests/test_nn_math.py); see er_emul.h for how the same source compiles on both sides.
#pragma once

#include "er_emul.h"

#ifndef ER_ICP_SCANW
#define ER_ICP_SCANW 4                 // candidate loads in flight per scan step (4 or 8; compile-time experiment knob)
#endif

#if defined(__CUDACC__)
typedef float4 er_float4;
#else
struct er_float4 { float x, y, z, w; };
#endif

// the 25 cell columns around a query's own, nearest first (cooperative search, rings <= 2)
#if defined(__CUDACC__)
__constant__
#else
static const
#endif
signed char kColumnOrder[25][2] = {
    {0, 0},  {-1, 0}, {1, 0},  {0, -1}, {0, 1},  {-1, -1}, {-1, 1}, {1, -1}, {1, 1},  {-2, 0}, {2, 0},  {0, -2}, {0, 2},
    {-2, -1}, {-2, 1}, {2, -1}, {2, 1}, {-1, -2}, {-1, 2}, {1, -2}, {1, 2},  {-2, -2}, {-2, 2}, {2, -2}, {2, 2}};

struct GridDev {
  const er_float4* pts;      // sorted by cell: x, y, z, original index (int bits)
  const er_float4* nrm;      // same order
  const int* cell_start;  // ncell + 1
  const float* xyz_orig;  // the same points / normals in original order (3 floats each)
  const float* nrm_orig;
  float org[3];
  float inv_cell;
  float cell;
  int rings;              // search neighbourhood in cells: rings * cell >= the largest admissible radius
  int dim[3];
};


// Grid geometry for a cloud with bounding box [lo, hi] that has to answer searches up to `radius`: cells of
// radius * 1.001 / rings, searched `rings` cells deep (rings * cell >= radius); fewer rings / larger cells when the
// box would need more than 64 M cells.  Fills org, dim, cell, inv_cell, rings of `g`; returns the radius the grid
// supports.  (Host side, once per fragment.)
static inline double er_grid_geometry(const float lo[3], const float hi[3], double radius, int rings, GridDev& g) {
  double cell = radius * 1.001 / rings;
  int dim[3];
  for (;;) {
    double cells = 1;
    for (int a = 0; a < 3; a++) { dim[a] = (int)std::floor((hi[a] - lo[a]) / cell) + 1; cells *= dim[a]; }
    if (cells <= 64e6) break;
    if (rings > 1) { rings--; cell = radius * 1.001 / rings; } else cell *= 2;
  }
  g.inv_cell = (float)(1.0 / cell);
  g.cell = (float)cell;
  g.rings = rings;
  for (int a = 0; a < 3; a++) { g.org[a] = lo[a]; g.dim[a] = dim[a]; }
  return cell * rings;
}

// cell of a coordinate along one axis (grid_count_kernel; the search computes the same product for its queries)
ER_VOXEL_FN int grid_cell_axis(float v, float org, float inv_cell, int dim) {
  const int i = __float2int_rd(__fmul_rn(__fsub_rn(v, org), inv_cell));       // floor; NaN -> 0
  return min(max(i, 0), dim - 1);
}


// offset bound: distance from a query at fraction f of its cell to the slab `o` cells away (0 for its own slab)
ER_VOXEL_FN float slab_gap(int o, float f, float cell, float slack) {
  return fmaxf(0.0f, (o > 0 ? ((float)o - f) : (o < 0 ? (f - (float)(o + 1)) : 0.0f)) * cell - slack);
}

// exact 1-NN among all target points within `rings` cells of q (rings * cell >= the search radius); returns the
// sorted position or -1.  Distance = ((dx*dx + dy*dy) + dz*dz) in float (flann::L2_Simple order); ties go to the
// smallest original index.  The query's own cell is scanned first; every other cell column is visited only if
// its lower distance bound (computed from the query's position inside its cell, minus a slack that covers the
// float rounding of the cell assignment) does not exceed the best distance so far -- so a converged pair scans
// a handful of candidates instead of the full neighbourhood, with the same result as the exhaustive scan.
// A candidate as one 64-bit key: float bits of d^2 (non-negative, so they order like the values) << 32 | original
// index.  The smallest key is the nearest neighbour with ties going to the smallest original index; a NaN distance
// has larger bits than any finite one and never wins.
ER_VOXEL_FN unsigned long long pack_key(float d2, int idx) {
  return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)idx;
}

ER_VOXEL_FN unsigned long long nn_key(const er_float4 p, float qx, float qy, float qz) {
  const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
  return pack_key(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)), __float_as_int(p.w));
}

// key = min(key, keys of the candidates s0..s1-1); four loads in flight (slots past the end re-read the last point)
ER_VOXEL_FN void nn_scan_key(const GridDev& g, int s0, int s1, float qx, float qy, float qz, unsigned long long& key) {
  const int last = s1 - 1;
  for (int s = s0; s < s1; s += ER_ICP_SCANW) {
    const er_float4 p0 = __ldg(&g.pts[s]);
    const er_float4 p1 = __ldg(&g.pts[min(s + 1, last)]);
    const er_float4 p2 = __ldg(&g.pts[min(s + 2, last)]);
    const er_float4 p3 = __ldg(&g.pts[min(s + 3, last)]);
#if ER_ICP_SCANW == 8
    const er_float4 p4 = __ldg(&g.pts[min(s + 4, last)]);
    const er_float4 p5 = __ldg(&g.pts[min(s + 5, last)]);
    const er_float4 p6 = __ldg(&g.pts[min(s + 6, last)]);
    const er_float4 p7 = __ldg(&g.pts[min(s + 7, last)]);
#endif
    key = min(min(key, nn_key(p0, qx, qy, qz)), min(nn_key(p1, qx, qy, qz), min(nn_key(p2, qx, qy, qz), nn_key(p3, qx, qy, qz))));
#if ER_ICP_SCANW == 8
    if (s + 4 <= last)
      key = min(min(key, nn_key(p4, qx, qy, qz)), min(nn_key(p5, qx, qy, qz), min(nn_key(p6, qx, qy, qz), nn_key(p7, qx, qy, qz))));
#endif
  }
}

ER_VOXEL_FN void nn_consider(const er_float4 p, int s, float qx, float qy, float qz, float& bd, int& best, int& best_idx) {
  const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  const int idx = __float_as_int(p.w);
  if (d2 < bd || (d2 == bd && idx < best_idx)) { bd = d2; best = s; best_idx = idx; }
}

// candidates s0..s1-1, four loads in flight at a time (the scan is latency-bound: one cell holds only a handful of
// points).  Slots past the end re-read the run's last point, which cannot change the result.
ER_VOXEL_FN void nn_scan(const GridDev& g, int s0, int s1, float qx, float qy, float qz, float& bd, int& best,
                                        int& best_idx) {
  const int last = s1 - 1;
  for (int s = s0; s < s1; s += 4) {
    const int t1 = min(s + 1, last), t2 = min(s + 2, last), t3 = min(s + 3, last);
    const er_float4 p0 = __ldg(&g.pts[s]);
    const er_float4 p1 = __ldg(&g.pts[t1]);
    const er_float4 p2 = __ldg(&g.pts[t2]);
    const er_float4 p3 = __ldg(&g.pts[t3]);
    nn_consider(p0, s, qx, qy, qz, bd, best, best_idx);
    nn_consider(p1, t1, qx, qy, qz, bd, best, best_idx);
    nn_consider(p2, t2, qx, qy, qz, bd, best, best_idx);
    nn_consider(p3, t3, qx, qy, qz, bd, best, best_idx);
  }
}

ER_VOXEL_FN int nn_search(const GridDev& g, float qx, float qy, float qz, float& best_d2) {
  const float lx = __fmul_rn(__fsub_rn(qx, g.org[0]), g.inv_cell);
  const float ly = __fmul_rn(__fsub_rn(qy, g.org[1]), g.inv_cell);
  const float lz = __fmul_rn(__fsub_rn(qz, g.org[2]), g.inv_cell);
  const float flx = floorf(lx), fly = floorf(ly), flz = floorf(lz);
  int best = -1, best_idx = 0x7fffffff;
  float bd = FLT_MAX;
  const int K = g.rings;
  // queries far outside the grid cannot have a neighbour within the radius
  if (!(flx >= -(float)(K + 1) && flx <= (float)(g.dim[0] + K) && fly >= -(float)(K + 1) && fly <= (float)(g.dim[1] + K) &&
        flz >= -(float)(K + 1) && flz <= (float)(g.dim[2] + K))) {
    best_d2 = bd;
    return -1;
  }
  const int cx = (int)flx, cy = (int)fly, cz = (int)flz;
  const float fx = lx - flx, fy = ly - fly, fz = lz - flz;        // position inside the cell, [0, 1)
  const float cell = g.cell, slack = 1e-5f * g.cell;
  if (cx >= 0 && cx < g.dim[0] && cy >= 0 && cy < g.dim[1] && cz >= 0 && cz < g.dim[2]) {
    const int c = (cx * g.dim[1] + cy) * g.dim[2] + cz;
    nn_scan(g, __ldg(&g.cell_start[c]), __ldg(&g.cell_start[c + 1]), qx, qy, qz, bd, best, best_idx);
  }
  // per axis, the cell offsets whose slab can still hold something closer than the best so far (after the centre
  // cell that is usually just {0} or {0, +-1}); an empty centre cell leaves bd = FLT_MAX and hence the full range
  int xlo = 0, xhi = 0, ylo = 0, yhi = 0;
  for (int o = 1; o <= K; o++) { const float a = fmaxf(0.0f, ((float)o - fx) * cell - slack); if (a * a > bd) break; xhi = o; }
  for (int o = 1; o <= K; o++) { const float a = fmaxf(0.0f, (fx + (float)(o - 1)) * cell - slack); if (a * a > bd) break; xlo = -o; }
  for (int o = 1; o <= K; o++) { const float a = fmaxf(0.0f, ((float)o - fy) * cell - slack); if (a * a > bd) break; yhi = o; }
  for (int o = 1; o <= K; o++) { const float a = fmaxf(0.0f, (fy + (float)(o - 1)) * cell - slack); if (a * a > bd) break; ylo = -o; }
  for (int ox = xlo; ox <= xhi; ox++) {
    const int ix = cx + ox;
    if (ix < 0 || ix >= g.dim[0]) continue;
    const float ax = fmaxf(0.0f, (ox > 0 ? ((float)ox - fx) : (ox < 0 ? (fx - (float)(ox + 1)) : 0.0f)) * cell - slack);
    for (int oy = ylo; oy <= yhi; oy++) {
      const int iy = cy + oy;
      const float ay = fmaxf(0.0f, (oy > 0 ? ((float)oy - fy) : (oy < 0 ? (fy - (float)(oy + 1)) : 0.0f)) * cell - slack);
      const float axy2 = ax * ax + ay * ay;
      if (axy2 > bd || iy < 0 || iy >= g.dim[1]) continue;
      // z cells of one (x, y) column are contiguous in memory: shrink the range to the cells that can still matter
      int zhi = cz, zlo = cz;
      for (int oz = 1; oz <= K; oz++) {
        const float az = fmaxf(0.0f, ((float)oz - fz) * cell - slack);
        if (axy2 + az * az > bd) break;
        zhi = cz + oz;
      }
      for (int oz = 1; oz <= K; oz++) {
        const float az = fmaxf(0.0f, (fz + (float)(oz - 1)) * cell - slack);
        if (axy2 + az * az > bd) break;
        zlo = cz - oz;
      }
      const int row = (ix * g.dim[1] + iy) * g.dim[2];
      if (ox == 0 && oy == 0) {
        // two separate runs around the already scanned centre cell
        const int a0 = max(zlo, 0), a1 = min(cz - 1, g.dim[2] - 1);
        if (a0 <= a1) nn_scan(g, __ldg(&g.cell_start[row + a0]), __ldg(&g.cell_start[row + a1 + 1]), qx, qy, qz, bd, best, best_idx);
        const int b0 = max(cz + 1, 0), b1 = min(zhi, g.dim[2] - 1);
        if (b0 <= b1) nn_scan(g, __ldg(&g.cell_start[row + b0]), __ldg(&g.cell_start[row + b1 + 1]), qx, qy, qz, bd, best, best_idx);
      } else {
        const int a0 = max(zlo, 0), a1 = min(zhi, g.dim[2] - 1);
        if (a0 <= a1) nn_scan(g, __ldg(&g.cell_start[row + a0]), __ldg(&g.cell_start[row + a1 + 1]), qx, qy, qz, bd, best, best_idx);
      }
    }
  }
  best_d2 = bd;
  return best;
}

