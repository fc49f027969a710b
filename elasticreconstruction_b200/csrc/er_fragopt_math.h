// er_fragopt_math.h -- per-thread arithmetic of the FragmentOptimizer path, shared between
// the CUDA kernels of er_fragopt.cu and the CPU emulation of tests/emul/fragopt_emul.cpp (the same text compiled by
// g++ and run thread by thread; tests/test_fragopt_cpu.py).  Plain IEEE operators only: the library is built with
// --fmad=false and the host side with -ffp-contract=off, so float expressions round exactly like the reference's
// (g++ -O2 -ffp-contract=off build of FragmentOptimizer/PointCloud.h, OptApp.cpp).
//
// Reference semantics restated here (paths relative to the reference checkout):
//   fo_get_coordinate      PointCloud::GetCoordinate          FragmentOptimizer/PointCloud.h:96-177
//   fo_vals / fo_nvals     the val_[8] / nval_[8] tables       PointCloud.h:127-170
//   fo_update_normal       PointCloud::UpdateNormal            PointCloud.h:57-68
//   fo_update_point        PointCloud::UpdatePoint             PointCloud.h:85-94
//   fo_update_pose         PointCloud::UpdatePose              PointCloud.h:70-83
//   fo_slac_row            one row of the SLAC Jacobian        OptApp.cpp:487-545
//   fo_slac_flush_*        its normal-equation update          OptApp.cpp:547-559
//   fo_rigid_row           one row of the rigid Jacobian       OptApp.cpp:332-369
//   fo_nonrigid_row / fo_nr_flush   one row of the nonrigid system and its update   OptApp.cpp:171-195
#pragma once

#include <cmath>

#if defined(__CUDACC__)
#define ER_FO_FN __host__ __device__ __forceinline__
#else
#define ER_FO_FN static inline
#endif

namespace er_fo {

// A point as the kernels keep it: what GetCoordinate derives once (lattice cell + trilinear residuals + the normal
// divided by the cell edge) and what the iterations update (position, normal).  The reference's Point additionally
// stores idx_[8], val_[8], nval_[8]; they are pure functions of (base, r, n0) and are recomputed where needed.
struct StaticPoint {   // 32 bytes, two 16-byte loads
  int base;            // GetIndex(corner): idx_[t] = (base + dx + dy*R1 + dz*R1*R1) * 3, t = dx*4 + dy*2 + dz
  float r[3];          // residual[3]
  float n0[3];         // pt[3..5] / unit_length_   (PointCloud.h:136-138)
  int valid;           // 0 for the zero-filled record the reference leaves behind an out-of-bound point
};

struct Lattice {
  int res, r1, r1sq, nper;   // resolution, resolution + 1, (resolution + 1)^2, 3 * (resolution + 1)^3
  float unit;                // unit_length_ as PointCloud keeps it: float(length) / resolution in float
};

ER_FO_FN Lattice fo_lattice(int resolution, float length) {
  Lattice L;
  L.res = resolution;
  L.r1 = resolution + 1;
  L.r1sq = L.r1 * L.r1;
  L.nper = L.r1 * L.r1 * L.r1 * 3;
  L.unit = length / resolution;            // PointCloud.cpp:10: float / int -> float
  return L;
}

ER_FO_FN int fo_corner_offset(const Lattice& L, int t) {   // t = dx*4 + dy*2 + dz (PointCloud.h:118-125)
  return ((t >> 2) & 1) + ((t >> 1) & 1) * L.r1 + (t & 1) * L.r1sq;
}

// PointCloud::GetCoordinate.  pt = x y z nx ny nz.  Returns false for "Point out of bound".
ER_FO_FN bool fo_get_coordinate(const Lattice& L, const float pt[6], StaticPoint* s, float p[3], float n[3]) {
  p[0] = pt[0]; p[1] = pt[1]; p[2] = pt[2];
  const float q0 = pt[0] / L.unit, q1 = pt[1] / L.unit, q2 = pt[2] / L.unit;
  const int c0 = (int)floorf(q0), c1 = (int)floorf(q1), c2 = (int)floorf(q2);
  s->base = 0; s->valid = 0;
  s->r[0] = s->r[1] = s->r[2] = 0.0f;
  s->n0[0] = s->n0[1] = s->n0[2] = 0.0f;
  n[0] = n[1] = n[2] = 0.0f;
  if (c0 < 0 || c0 >= L.res || c1 < 0 || c1 >= L.res || c2 < 0 || c2 >= L.res) return false;
  s->r[0] = q0 - (float)c0;
  s->r[1] = q1 - (float)c1;
  s->r[2] = q2 - (float)c2;
  s->base = c0 + c1 * L.r1 + c2 * L.r1sq;
  s->n0[0] = pt[3] / L.unit;
  s->n0[1] = pt[4] / L.unit;
  s->n0[2] = pt[5] / L.unit;
  s->valid = 1;
  n[0] = s->n0[0]; n[1] = s->n0[1]; n[2] = s->n0[2];     // PointCloud.h:172-174 (the divided normal)
  return true;
}

ER_FO_FN void fo_vals(const float r[3], float v[8]) {
  const float x0 = 1 - r[0], y0 = 1 - r[1], z0 = 1 - r[2];
  v[0] = x0 * y0 * z0;
  v[1] = x0 * y0 * r[2];
  v[2] = x0 * r[1] * z0;
  v[3] = x0 * r[1] * r[2];
  v[4] = r[0] * y0 * z0;
  v[5] = r[0] * y0 * r[2];
  v[6] = r[0] * r[1] * z0;
  v[7] = r[0] * r[1] * r[2];
}

ER_FO_FN void fo_nvals(const float r[3], const float n0[3], float v[8]) {
  const float x0 = 1 - r[0], y0 = 1 - r[1], z0 = 1 - r[2];
  const float x1 = r[0], y1 = r[1], z1 = r[2];
  const float a = n0[0], b = n0[1], c = n0[2];
  v[0] = -a * y0 * z0 - b * x0 * z0 - c * x0 * y0;
  v[1] = -a * y0 * z1 - b * x0 * z1 + c * x0 * y0;
  v[2] = -a * y1 * z0 + b * x0 * z0 - c * x0 * y1;
  v[3] = -a * y1 * z1 + b * x0 * z1 + c * x0 * y1;
  v[4] = a * y0 * z0 - b * x1 * z0 - c * x1 * y0;
  v[5] = a * y0 * z1 - b * x1 * z1 + c * x1 * y0;
  v[6] = a * y1 * z0 + b * x1 * z0 - c * x1 * y1;
  v[7] = a * y1 * z1 + b * x1 * z1 + c * x1 * y1;
}

// ctr = the lattice block of this point's fragment (ctr + offset_), doubles.
ER_FO_FN void fo_update_normal(const Lattice& L, const StaticPoint& s, const double* ctr, float n[3]) {
  float nv[8];
  fo_nvals(s.r, s.n0, nv);
  for (int i = 0; i < 3; ++i) {
    float a = 0.0f;
    for (int j = 0; j < 8; ++j) a += nv[j] * (float)ctr[(s.base + fo_corner_offset(L, j)) * 3 + i];
    n[i] = a;
  }
  const float len = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  n[0] /= len;
  n[1] /= len;
  n[2] /= len;
}

ER_FO_FN void fo_update_point(const Lattice& L, const StaticPoint& s, const double* ctr, double pos[3]) {
  float v[8];
  fo_vals(s.r, v);
  for (int i = 0; i < 3; ++i) {
    double a = 0.0;
    for (int j = 0; j < 8; ++j) a += v[j] * (float)ctr[(s.base + fo_corner_offset(L, j)) * 3 + i];   // float product, double sum
    pos[i] = a;
  }
}

// Matrix4f * (x, y, z, w) with Eigen's fixed-size evaluation order ((m0*x + m1*y) + m2*z) + m3*w; m row-major.
ER_FO_FN void fo_update_pose(const float m[16], float p[3], float n[3]) {
  const float x = p[0], y = p[1], z = p[2];
  const float a = n[0], b = n[1], c = n[2];
  for (int r = 0; r < 3; ++r) {
    p[r] = ((m[4 * r] * x + m[4 * r + 1] * y) + m[4 * r + 2] * z) + m[4 * r + 3] * 1.0f;
    n[r] = ((m[4 * r] * a + m[4 * r + 1] * b) + m[4 * r + 2] * c) + m[4 * r + 3] * 0.0f;
  }
}

// ------------------------------------------------------------------------------------------------ SLAC
// One correspondence (pi in fragment i, pj in fragment j) gives a Jacobian row with 60 entries
//   [ 6 pose i | 6 pose j | 8 corners x 3 of pi's lattice cell | 8 x 3 of pj's ]   (OptApp.cpp:500-545)
// and the residual b.  Element 60 receives b, 61..63 zeros (the kernel's shared-memory rows are 64 doubles; fo_pos maps
// element numbers to positions).
// rot_ti / rot_tj = pose_rot_t_[i], [j] row-major.
constexpr int kSlacRow = 60;
constexpr int kRowStride = 64;

// Where element e of a staged row lives in shared memory.  A thread of the accumulation loop needs the four values
// 4t .. 4t+3 of two tile indices t; with the natural order (e -> e) neighbouring tiles are 32 bytes apart and the 16-byte
// loads of a quarter-warp hit every bank twice (measured: 53 % of the kernel's shared-memory wavefronts were bank conflicts,
// the shared-memory pipe busy 84 % of the cycles, profiles/r2_slac_assemble_kernel_v1.txt).  Two planes instead: plane 0 holds the
// pairs (4t, 4t+1) of all tiles back to back, plane 1 the pairs (4t+2, 4t+3), so that consecutive tiles are consecutive
// 16-byte words and a quarter-warp reads one conflict-free 128-byte line.
#ifndef ER_FO_PLANES
#define ER_FO_PLANES 1
#endif
template <int STRIDE>
ER_FO_FN int fo_pos(int e) {
#if ER_FO_PLANES
  return ((e >> 1) & 1) * (STRIDE / 2) + (e >> 2) * 2 + (e & 1);
#else
  return e;
#endif
}
// Thread t of the staging phase writes row t, STRIDE doubles after row t - 1: with 64 doubles per row every lane of a store
// hits the same bank (measured: the staging stores were 26 % of all shared-memory wavefronts, profiles/r2_slac_assemble_kernel_v2.txt).
// The 16-byte words of row t are therefore permuted by XOR with (t & 7): eight consecutive rows spread over all 32 banks
// (a 4-way conflict instead of a 32-way one); the accumulation loop reads row c with the same, CTA-uniform, key c & 7.
#ifndef ER_FO_SWIZZLE
#define ER_FO_SWIZZLE 1
#endif
template <int STRIDE>
ER_FO_FN int fo_spos(int e, int sw) {
#if ER_FO_SWIZZLE && ER_FO_PLANES
  const int p = fo_pos<STRIDE>(e);
  return (((p >> 1) ^ sw) << 1) | (p & 1);
#else
  (void)sw;
  return fo_pos<STRIDE>(e);
#endif
}

ER_FO_FN void fo_slac_row(const float pi[3], const float ni[3], const float ri[3], const float pj[3], const float rj[3],
                          const double* rot_ti, const double* rot_tj, double* row, int sw) {
  const double ppi[3] = {pi[0], pi[1], pi[2]}, ppj[3] = {pj[0], pj[1], pj[2]}, npi[3] = {ni[0], ni[1], ni[2]};
  const double d0 = ppi[0] - ppj[0], d1 = ppi[1] - ppj[1], d2 = ppi[2] - ppj[2];
  const double b = (d0 * npi[0] + d1 * npi[1]) + d2 * npi[2];
  const double t0 = ppj[1] * npi[2] - ppj[2] * npi[1];       // ppj.cross(npi)
  const double t1 = ppj[2] * npi[0] - ppj[0] * npi[2];
  const double t2 = ppj[0] * npi[1] - ppj[1] * npi[0];
  constexpr int S = kRowStride;
  row[fo_spos<S>(0, sw)] = t0; row[fo_spos<S>(1, sw)] = t1; row[fo_spos<S>(2, sw)] = t2;
  row[fo_spos<S>(3, sw)] = npi[0]; row[fo_spos<S>(4, sw)] = npi[1]; row[fo_spos<S>(5, sw)] = npi[2];
  row[fo_spos<S>(6, sw)] = -t0; row[fo_spos<S>(7, sw)] = -t1; row[fo_spos<S>(8, sw)] = -t2;
  row[fo_spos<S>(9, sw)] = -npi[0]; row[fo_spos<S>(10, sw)] = -npi[1]; row[fo_spos<S>(11, sw)] = -npi[2];
  double dTi[3], dTj[3];
  for (int r = 0; r < 3; ++r) {
    dTi[r] = (rot_ti[3 * r] * npi[0] + rot_ti[3 * r + 1] * npi[1]) + rot_ti[3 * r + 2] * npi[2];
    dTj[r] = (-rot_tj[3 * r] * npi[0] + -rot_tj[3 * r + 1] * npi[1]) + -rot_tj[3 * r + 2] * npi[2];
  }
  float a[8], c[8];
  fo_vals(ri, a);
  fo_vals(rj, c);
  for (int ll = 0; ll < 8; ++ll)
    for (int xyz = 0; xyz < 3; ++xyz) {
      row[fo_spos<S>(12 + ll * 3 + xyz, sw)] = a[ll] * dTi[xyz];
      row[fo_spos<S>(36 + ll * 3 + xyz, sw)] = c[ll] * dTj[xyz];
    }
  row[fo_spos<S>(60, sw)] = b;
  row[fo_spos<S>(61, sw)] = 0.0; row[fo_spos<S>(62, sw)] = 0.0; row[fo_spos<S>(63, sw)] = 0.0;
}

// The segment key: all correspondences with the same key share the 60 global column indices of their rows.
struct SlacKey {
  int pair, bi, bj;     // pair number, base of pi, base of pj
};
ER_FO_FN bool operator==(const SlacKey& a, const SlacKey& b) { return a.pair == b.pair && a.bi == b.bi && a.bj == b.bj; }

// global index of row position q (0..59) for fragments (i, j) of the pair and the two lattice cells
ER_FO_FN int fo_slac_index(const Lattice& L, int num, int i, int j, int bi, int bj, int q) {
  if (q < 6) return i * 6 + q;
  if (q < 12) return j * 6 + (q - 6);
  if (q < 36) { const int ll = (q - 12) / 3; return 6 * num + (bi + fo_corner_offset(L, ll)) * 3 + (q - 12) % 3; }
  const int ll = (q - 36) / 3;
  return 6 * num + (bj + fo_corner_offset(L, ll)) * 3 + (q - 36) % 3;
}

// Thread `tid` of a 128-thread CTA owns one 4x4 tile of the upper triangle of the 60x60 product row^T row
// (15 x 15 tiles, 120 of them on or above the diagonal: tid 0..119), or -- tid 120..127 -- eight entries of
// b * row (positions 8u .. 8u+7; position 60 holds b itself, so thread 127 also sums b*b, the score).
ER_FO_FN void fo_tile_of(int tid, int* ty, int* tx) {
  int t = tid, y = 0;
  while (t >= 15 - y) { t -= 15 - y; ++y; }      // row y of the tile triangle has 15 - y tiles
  *ty = y;
  *tx = y + t;
}

struct alignas(16) D2 { double x, y; };

// ER_FO_JB_DIAG (compile-time option, default off: found after the round's last GPU session, checked on the CPU emulation
// only -- tests/test_fragopt_cpu.py::test_emulated_kernels_with_the_diagonal_jb_variant).  The eight b * row threads share a
// warp with 24 tile threads, which therefore runs two paths per row (19.7 % of the kernel's warp instructions,
// profiles/r2_slac_assemble_lines.txt).  With this option nobody owns b * row: a diagonal tile (ty == tx) needs only the ten
// products on or above its diagonal, so four of its six idle accumulators -- (1,0), (2,0), (2,1), (3,0) -- sum b * a0..a3
// (the Jb entries of positions 4t .. 4t+3; its row operands are already in registers) and, for tile (0,0), (3,1) sums b*b.
// Operand selects instead of a second path.  Arithmetic verified on the emulation; NOT tuned: as compiled today (nvcc 12.9, 128-register
// cap) the selects push the accumulators into local memory inside the row loop (STL.128 per row in the SASS), so it needs a pass over
// the register budget before it can be expected to win.
#ifndef ER_FO_JB_DIAG
#define ER_FO_JB_DIAG 0
#endif

// acc[16] += a (x) b for the 4x4 tile (ty, tx) of row^T row
template <int STRIDE, bool JB = false>
ER_FO_FN void fo_tile_mac(int ty, int tx, const double* row, int sw, double* acc) {
  // four 128-bit shared-memory loads (rows are 16-byte aligned)
#if ER_FO_PLANES && ER_FO_SWIZZLE
  const D2 p0 = *reinterpret_cast<const D2*>(row + 2 * (ty ^ sw)), p1 = *reinterpret_cast<const D2*>(row + 2 * ((STRIDE / 4 + ty) ^ sw));
  const D2 q0 = *reinterpret_cast<const D2*>(row + 2 * (tx ^ sw)), q1 = *reinterpret_cast<const D2*>(row + 2 * ((STRIDE / 4 + tx) ^ sw));
#elif ER_FO_PLANES
  (void)sw;
  const D2 p0 = *reinterpret_cast<const D2*>(row + 2 * ty), p1 = *reinterpret_cast<const D2*>(row + STRIDE / 2 + 2 * ty);
  const D2 q0 = *reinterpret_cast<const D2*>(row + 2 * tx), q1 = *reinterpret_cast<const D2*>(row + STRIDE / 2 + 2 * tx);
#else
  (void)sw;
  const D2 p0 = *reinterpret_cast<const D2*>(row + 4 * ty), p1 = *reinterpret_cast<const D2*>(row + 4 * ty + 2);
  const D2 q0 = *reinterpret_cast<const D2*>(row + 4 * tx), q1 = *reinterpret_cast<const D2*>(row + 4 * tx + 2);
#endif
  const double a0 = p0.x, a1 = p0.y, a2 = p1.x, a3 = p1.y;
  const double b0 = q0.x, b1 = q0.y, b2 = q1.x, b3 = q1.y;
  // fused multiply-adds: the order of the sums differs from the reference's anyway (per-pair dense matrices added
  // under an OpenMP critical section, OptApp.cpp:562-569), so only the tolerance matters here, and DFMA is the roof
  acc[0] = fma(a0, b0, acc[0]); acc[1] = fma(a0, b1, acc[1]); acc[2] = fma(a0, b2, acc[2]); acc[3] = fma(a0, b3, acc[3]);
#if ER_FO_JB_DIAG
  if (JB) {
    const bool diag = ty == tx;
    const double bv = row[fo_spos<STRIDE>(60, sw)];
    acc[4] = fma(diag ? bv : a1, diag ? a0 : b0, acc[4]); acc[5] = fma(a1, b1, acc[5]); acc[6] = fma(a1, b2, acc[6]); acc[7] = fma(a1, b3, acc[7]);
    acc[8] = fma(diag ? bv : a2, diag ? a1 : b0, acc[8]); acc[9] = fma(diag ? bv : a2, diag ? a2 : b1, acc[9]);
    acc[10] = fma(a2, b2, acc[10]); acc[11] = fma(a2, b3, acc[11]);
    acc[12] = fma(diag ? bv : a3, diag ? a3 : b0, acc[12]); acc[13] = fma(diag ? bv : a3, diag ? bv : b1, acc[13]);
    acc[14] = fma(a3, b2, acc[14]); acc[15] = fma(a3, b3, acc[15]);
    return;
  }
#endif
  acc[4] = fma(a1, b0, acc[4]); acc[5] = fma(a1, b1, acc[5]); acc[6] = fma(a1, b2, acc[6]); acc[7] = fma(a1, b3, acc[7]);
  acc[8] = fma(a2, b0, acc[8]); acc[9] = fma(a2, b1, acc[9]); acc[10] = fma(a2, b2, acc[10]); acc[11] = fma(a2, b3, acc[11]);
  acc[12] = fma(a3, b0, acc[12]); acc[13] = fma(a3, b1, acc[13]); acc[14] = fma(a3, b2, acc[14]); acc[15] = fma(a3, b3, acc[15]);
}

// one staged row: the tile threads add their tile, the eight vector threads acc[0..7] += b * row[8u..]
ER_FO_FN void fo_slac_mac(int tid, int ty, int tx, const double* row, int sw, double* acc) {
#if ER_FO_JB_DIAG
  if (tid < 120) fo_tile_mac<kRowStride, true>(ty, tx, row, sw, acc);
  return;
#endif
  if (tid < 120) {
    fo_tile_mac<kRowStride>(ty, tx, row, sw, acc);
  } else {
    const double b = row[fo_spos<kRowStride>(60, sw)];
    const int u = tid - 120;
    for (int e = 0; e < 8; ++e) acc[e] = fma(b, row[fo_spos<kRowStride>(8 * u + e, sw)], acc[e]);
  }
}

// Adds a finished segment's sums into the dense upper-triangular JJ (row-major M x M), Jb and the score, with the
// reference's folding rule (OptApp.cpp:548-558): equal indices inside one row add 2*v_p*v_q to the diagonal, otherwise
// the product goes to (min, max).  ADD(ptr, value) is atomicAdd on the device and += in the emulation.
template <class Add>
ER_FO_FN void fo_slac_flush(const Lattice& L, int num, int M, int i, int j, const SlacKey& key, int tid, int ty, int tx,
                            const double* acc, double* JJ, double* Jb, double* score, Add ADD) {
  if (tid < 120) {
    int gp[4], gq[4];
    for (int r = 0; r < 4; ++r) {
      gp[r] = fo_slac_index(L, num, i, j, key.bi, key.bj, 4 * ty + r);
      gq[r] = fo_slac_index(L, num, i, j, key.bi, key.bj, 4 * tx + r);
    }
    for (int r = 0; r < 4; ++r)
      for (int s = 0; s < 4; ++s) {
        const int p = 4 * ty + r, q = 4 * tx + s;
        if (p > q) {                                       // lower half of a diagonal tile
#if ER_FO_JB_DIAG
          const double w = acc[4 * r + s];                 // (1,0) (2,0) (2,1) (3,0): b * a0..a3; (3,1) of tile (0,0): b * b
          if (w != 0.0) {
            const int e = r == 1 ? 0 : r == 2 ? 1 + s : (s == 0 ? 3 : (s == 1 ? 4 : 5));
            if (e < 4) ADD(&Jb[gp[e]], w);
            else if (e == 4 && ty == 0) ADD(score, w);
          }
#endif
          continue;
        }
        const double v = acc[4 * r + s];
        if (v == 0.0) continue;
        const int a = gp[r], b = gq[s];
        if (p == q) ADD(&JJ[(size_t)a * M + a], v);
        else if (a == b) ADD(&JJ[(size_t)a * M + a], 2 * v);
        else if (a < b) ADD(&JJ[(size_t)a * M + b], v);
        else ADD(&JJ[(size_t)b * M + a], v);
      }
  } else if (!ER_FO_JB_DIAG) {
    const int u = tid - 120;
    for (int e = 0; e < 8; ++e) {
      const int q = 8 * u + e;
      if (q < 60) { if (acc[e] != 0.0) ADD(&Jb[fo_slac_index(L, num, i, j, key.bi, key.bj, q)], acc[e]); }
      else if (q == 60) ADD(score, acc[e]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ rigid
// OptApp.cpp:337-369: val[12] over (6 of fragment i, 6 of fragment j) and the residual b.
ER_FO_FN void fo_rigid_row(const float pi[3], const float ni[3], const float pj[3], double val[12], double* b_out) {
  const double ppi[4] = {pi[0], pi[1], pi[2], 1.0}, ppj[4] = {pj[0], pj[1], pj[2], 1.0}, npi[4] = {ni[0], ni[1], ni[2], 0.0};
  const double d[4] = {ppi[0] - ppj[0], ppi[1] - ppj[1], ppi[2] - ppj[2], ppi[3] - ppj[3]};
  // Vector4d::dot with SSE2 packets of two doubles: (a0*b0 + a2*b2) + (a1*b1 + a3*b3)
#define ER_FO_DOT4(a0, a1, a2, a3, v) (((a0) * (v)[0] + (a2) * (v)[2]) + ((a1) * (v)[1] + (a3) * (v)[3]))
  *b_out = ER_FO_DOT4(d[0], d[1], d[2], d[3], npi);
  val[0] = ER_FO_DOT4(0.0, -ppi[2], ppi[1], 1.0, npi) + ER_FO_DOT4(0.0, -npi[2], npi[1], 0.0, d);
  val[1] = ER_FO_DOT4(ppi[2], 0.0, -ppi[0], 1.0, npi) + ER_FO_DOT4(npi[2], 0.0, -npi[0], 0.0, d);
  val[2] = ER_FO_DOT4(-ppi[1], ppi[0], 0.0, 1.0, npi) + ER_FO_DOT4(-npi[1], npi[0], 0.0, 0.0, d);
  val[3] = npi[0]; val[4] = npi[1]; val[5] = npi[2];
  val[6] = -ER_FO_DOT4(0.0, -ppj[2], ppj[1], 1.0, npi);
  val[7] = -ER_FO_DOT4(ppj[2], 0.0, -ppj[0], 1.0, npi);
  val[8] = -ER_FO_DOT4(-ppj[1], ppj[0], 0.0, 1.0, npi);
  val[9] = -npi[0]; val[10] = -npi[1]; val[11] = -npi[2];
#undef ER_FO_DOT4
}

// ------------------------------------------------------------------------------------------------ nonrigid
// OptApp.cpp:171-195.  One correspondence gives a 48-entry row [24 on fragment i's lattice | 24 on fragment j's]:
// position t*8 + ll (axis t, corner ll) holds pi.val_[ll] * weight * pi.n_[t], position 24 + t*8 + ll holds
// -pj.val_[ll] * weight * pi.n_[t] (float val, double weight, float normal), at the unknowns idx_[ll] + t of the two
// fragments' lattices.  The system has no data right-hand side: mati / matj / matij add row^T row only.
constexpr int kNrRow = 48;
constexpr int kNrStride = 48;
constexpr int kNrThreads = 96;     // 78 of them own a 4x4 tile of the 12 x 12-tile upper triangle
constexpr int kNrTiles = 78;

ER_FO_FN void fo_nonrigid_row(const float ri[3], const float ni[3], const float rj[3], double weight, double* row, int sw) {
  float a[8], c[8];
  fo_vals(ri, a);
  fo_vals(rj, c);
  for (int t = 0; t < 3; ++t)
    for (int ll = 0; ll < 8; ++ll) {
      row[fo_spos<kNrStride>(t * 8 + ll, sw)] = a[ll] * weight * ni[t];
      row[fo_spos<kNrStride>(24 + t * 8 + ll, sw)] = -c[ll] * weight * ni[t];
    }
}

ER_FO_FN void fo_nr_tile_of(int tid, int* ty, int* tx) {
  int t = tid, y = 0;
  while (t >= 12 - y) { t -= 12 - y; ++y; }
  *ty = y;
  *tx = y + t;
}

// unknown of row position q (0..47): fragment i's / j's lattice block, node of the cell's corner ll, axis t
ER_FO_FN long long fo_nr_index(const Lattice& L, int i, int j, int bi, int bj, int q) {
  const int half = q < 24 ? 0 : 1, r = q - 24 * half, t = r / 8, ll = r % 8;
  return (long long)(half ? j : i) * L.nper + (long long)((half ? bj : bi) + fo_corner_offset(L, ll)) * 3 + t;
}

// A (dense, row-major N x N, N = num * nper) += the segment's tile, upper triangle only: the reference's solver reads
// thisAA.selfadjointView<Upper>() (OptApp.cpp:209-211), so of mati / matj (full symmetric blocks) the upper halves
// count, and matij -- rows of fragment i, columns of fragment j, never its transpose (OptApp.cpp:194) -- counts only
// when i < j: for a pair listed with idx0 > idx1 the cross term lies in the lower triangle and is never read.
template <class Add>
ER_FO_FN void fo_nr_flush(const Lattice& L, long long N, int i, int j, const SlacKey& key, int tid, int ty, int tx, const double* acc,
                          double* A, Add ADD) {
  if (tid >= kNrTiles) return;
  long long gp[4], gq[4];
  for (int r = 0; r < 4; ++r) {
    gp[r] = fo_nr_index(L, i, j, key.bi, key.bj, 4 * ty + r);
    gq[r] = fo_nr_index(L, i, j, key.bi, key.bj, 4 * tx + r);
  }
  for (int r = 0; r < 4; ++r)
    for (int s = 0; s < 4; ++s) {
      const int p = 4 * ty + r, q = 4 * tx + s;
      if (p > q) continue;
      const bool cross = (p < 24) != (q < 24);
      if (cross && !(i < j)) continue;
      const double v = acc[4 * r + s];
      if (v == 0.0) continue;
      const long long a = gp[r], b = gq[s];
      if (p == q) ADD(&A[a * N + a], v);
      else if (a == b) ADD(&A[a * N + a], 2 * v);
      else if (a < b) ADD(&A[a * N + b], v);
      else ADD(&A[b * N + a], v);
    }
}

}  // namespace er_fo
