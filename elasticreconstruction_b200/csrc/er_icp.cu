// er_icp.cu -- pairwise point-to-plane ICP refinement on B200 (sm_100a): kernels + the er_cloud_*/er_icp_* C ABI.
//
// Replaces, behind include/elastic_b200.h, the inner loops of BuildCorrespondence:
//   CCorresApp::Registration      BuildCorrespondence/CorresApp.cpp:212-319  (gate :250-264, ICP :295-312)
//   CCorresApp::FindCorrespondence BuildCorrespondence/CorresApp.cpp:112-210 (:145-161 search, :186-208 information)
// and the PCL 1.7 pieces those lines instantiate (not in the reference tree; semantics as restated in
// oracle/icp_oracle.c, which is the definition this file is tested against):
//   pcl::KdTreeFLANN::{setInputCloud,nearestKSearch}      -> uniform grid over the target + pruned 5x5x5 cell search
//   CorrespondenceEstimation::determineCorrespondences     -> icp_iter_kernel (search + gate)
//   TransformationEstimationPointToPlaneLLS                -> icp_iter_kernel (rows, 6x6 normal equations, solve)
//   IterativeClosestPoint::computeTransformation           -> one launch of icp_iter_kernel per iteration, no host sync
//   DefaultConvergenceCriteria                             -> last CTA of each pair
//   transformPointCloudWithNormals                         -> fused into the gate / correspondence kernels
//
// Data layout in HBM: per fragment the target role uses points sorted by grid cell as float4 (x, y, z, original
// index) plus float4 normals in the same order and an int cell_start[] table; the source role uses the original
// order (xyz, normals as float3 arrays).  One pair's working set is ~7 MB, far below the 126 MB L2, so the search
// is latency/L2-bound, not HBM-bound; throughput comes from running many pairs per launch (blockIdx.y = pair).
//
// Numerics: squared distances and the incremental float transform use explicit round-to-nearest intrinsics in
// the restated PCL/FLANN/Eigen order; the normal equations are accumulated in double and reduced in a fixed
// order (per-thread -> warp shuffle tree -> per-CTA partial -> last CTA sums partials by index), so results are
// reproducible run to run.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <vector>

#include "er_common.h"
#include "er_nn_math.h"

namespace {

#ifndef ER_ICP_THREADS
#define ER_ICP_THREADS 128             // CTA size of the ICP kernels (compile-time experiment knob: 64 or 128)
#endif
constexpr int kThreads = ER_ICP_THREADS;
#ifndef ER_ICP_PTS
#define ER_ICP_PTS 8                   // source points per thread (compile-time experiment knob)
#endif
constexpr int kPtsPerThread = ER_ICP_PTS;
constexpr int kPtsPerCta = kThreads * kPtsPerThread;
constexpr int kAcc = 29;               // 21 (upper ATA) + 6 (ATb) + sum d^2 + count
constexpr int kMaxEvents = 16;
constexpr int kMaxRings = 2;           // cooperative search: at most (2*2+1)^2 + 1 = 26 cell-column items per query
constexpr int kMaxItems = 32 * ((2 * kMaxRings + 1) * (2 * kMaxRings + 1) + 1);
constexpr int kRunCap = 64;            // per-warp list of point runs waiting to be scanned (drained at 32)
constexpr int kNoMatch = 0x7fffffff;

// (evict-first hints on the streamed working copy of the source points were measured in round 2: 381.9 vs 382.1 ms per
// 1225-pair pass -- no effect, not kept)
#define ER_LDCUR(p) (*(p))
#define ER_STCUR(p, v) (*(p) = (v))

struct PairState {
  GridDev tgt;
  float* cur;             // source points, transformed in place iteration by iteration (n_src * 3)
  const er_float4* src4;  // first iteration reads the fragment itself instead of a copy: cell-sorted points (x, y, z, index)
  const float* src3;      // ... or, when src4 is null, the points in file order
  int n_src;
  int nblocks;
  double* partials;       // nblocks * kAcc
  float Tinc[16];         // transform applied at the start of the next iteration
  float fin[16];          // final_transformation_
  double prev_mse, mse;
  int iters, done, converged, cnt;
  unsigned ticket;
};

// (GridDev, pack_key, nn_key, nn_scan_key, nn_scan, nn_search, slab_gap, grid_cell_axis, er_grid_geometry:
// er_nn_math.h, shared with the CPU search test)

// 6x6 solve by Gaussian elimination with partial pivoting (same algorithm as the oracle's solve6)
__device__ bool solve6(double* A, double* b, double* x) {
  for (int col = 0; col < 6; col++) {
    int piv = col;
    double big = fabs(A[col * 6 + col]);
    for (int r = col + 1; r < 6; r++)
      if (fabs(A[r * 6 + col]) > big) { big = fabs(A[r * 6 + col]); piv = r; }
    if (big == 0.0) return false;
    if (piv != col) {
      for (int c = 0; c < 6; c++) { const double t = A[col * 6 + c]; A[col * 6 + c] = A[piv * 6 + c]; A[piv * 6 + c] = t; }
      const double t = b[col]; b[col] = b[piv]; b[piv] = t;
    }
    for (int r = col + 1; r < 6; r++) {
      const double f = __ddiv_rn(A[r * 6 + col], A[col * 6 + col]);
      for (int c = col; c < 6; c++) A[r * 6 + c] = __dsub_rn(A[r * 6 + c], __dmul_rn(f, A[col * 6 + c]));
      b[r] = __dsub_rn(b[r], __dmul_rn(f, b[col]));
    }
  }
  for (int r = 5; r >= 0; r--) {
    double s = b[r];
    for (int c = r + 1; c < 6; c++) s = __dsub_rn(s, __dmul_rn(A[r * 6 + c], x[c]));
    x[r] = __ddiv_rn(s, A[r * 6 + r]);
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// icp_iter_kernel: one ICP iteration for every pair of the batch (grid = (CTAs per pair, pairs)).
//   1. input_transformed = T_inc * input_transformed   (float, Eigen order; T_inc of the previous iteration)
//   2. 1-NN in the target grid, keep unless dist^2 > max_dist^2   (determineCorrespondences)
//   3. point-to-plane row [s x n, n], d = n.(t - s), accumulate A^T A / A^T b in double   (LLS)
//   4. CTA partial -> global; the pair's last CTA sums the partials in index order, solves the 6x6 system,
//      builds T_inc = Rz Ry Rx + t, updates final = T_inc * final and evaluates the convergence criteria.
//
// Step 2 comes in two forms with identical results (the exact nearest neighbour, ties to the smallest original
// index, is independent of the visiting order):
//   COOP = false  every thread walks the pruned neighbourhood of its own query (nn_search); lanes of a warp have
//                 very different trip counts (measured 11.7 of 32 lanes active on average);
//   COOP = true   warp-cooperative: each lane scans only its query's own cell, which gives a bound bd, and lists
//                 the cell columns (x, y offsets) that can still hold something as close as bd -- usually one to
//                 four -- as items (lane, ox, oy) in a per-warp list in shared memory.  The items are then dealt
//                 round-robin to the lanes regardless of the query they belong to: a lane derives the column's run of
//                 z cells from the owning query's freshest bound (most items are pruned here); the surviving point
//                 runs are compacted into a second list, dealt round-robin again and scanned, the best candidate of
//                 a run merged into the query's 64-bit key (float bits of d^2 << 32 | original index) with atomicMin
//                 in shared memory.  The warp's work is thereby balanced over its lanes whatever the spread between
//                 its queries.
// A match is kept as that 64-bit key; kNoMatch in the low word means none.  Candidates with d^2 above bd_init (the
// float just above max_dist^2) can never be kept, so the search starts from that bound.
// ---------------------------------------------------------------------------------------------
// The rows of one thread's accepted correspondences (keys[it * kThreads], source point k0 + it * kThreads), summed in
// double and reduced over the warp by a fixed shuffle tree into out[].  PASS 0: the 15 sums of a, b, c against
// (a, b, c, nx, ny, nz); PASS 1: the 6 n.n sums, the 6 right-hand sides, sum of d^2 and the count.
template <int PASS>
__device__ __forceinline__ void icp_rows(const PairState& S, const GridDev& g, const unsigned long long* keys, int k0,
                                         double max_d2, double* out, int lane) {
  constexpr int N = PASS == 0 ? 15 : 14;
  double acc[N];
#pragma unroll
  for (int i = 0; i < N; i++) acc[i] = 0.0;
#pragma unroll 1
  for (int it = 0; it < kPtsPerThread; it++) {
    const unsigned long long key = keys[it * kThreads];
    const int idx = (int)(unsigned)(key & 0xffffffffu);
    const float md2 = __uint_as_float((unsigned)(key >> 32));
    if (idx == kNoMatch || (double)md2 > max_d2) continue;
    const float* cp = S.cur + 3 * (size_t)(k0 + it * kThreads);
    const float sx = ER_LDCUR(cp), sy = ER_LDCUR(cp + 1), sz = ER_LDCUR(cp + 2);
    const float* tp = g.xyz_orig + 3 * (size_t)idx;
    const float* np = g.nrm_orig + 3 * (size_t)idx;
    const float dx = __ldg(tp), dy = __ldg(tp + 1), dz = __ldg(tp + 2), nx = __ldg(np), ny = __ldg(np + 1), nz = __ldg(np + 2);
    if (PASS == 1) { acc[12] += (double)md2; acc[13] += 1.0; }
    if (isfinite(sx) && isfinite(sy) && isfinite(sz) && isfinite(dx) && isfinite(dy) && isfinite(dz) && isfinite(nx) &&
        isfinite(ny) && isfinite(nz)) {
      // float expressions widened afterwards, as in PCL's estimateRigidTransformation (float operands)
      const double a = (double)__fsub_rn(__fmul_rn(nz, sy), __fmul_rn(ny, sz));
      const double b = (double)__fsub_rn(__fmul_rn(nx, sz), __fmul_rn(nz, sx));
      const double c = (double)__fsub_rn(__fmul_rn(ny, sx), __fmul_rn(nx, sy));
      const double dnx = (double)nx, dny = (double)ny, dnz = (double)nz;
      if (PASS == 0) {
        acc[0] += a * a;   acc[1] += a * b;   acc[2] += a * c;   acc[3] += a * dnx;  acc[4] += a * dny;  acc[5] += a * dnz;
        acc[6] += b * b;   acc[7] += b * c;   acc[8] += b * dnx; acc[9] += b * dny;  acc[10] += b * dnz;
        acc[11] += c * c;  acc[12] += c * dnx; acc[13] += c * dny; acc[14] += c * dnz;
      } else {
        float df = __fadd_rn(__fadd_rn(__fmul_rn(nx, dx), __fmul_rn(ny, dy)), __fmul_rn(nz, dz));
        df = __fsub_rn(__fsub_rn(__fsub_rn(df, __fmul_rn(nx, sx)), __fmul_rn(ny, sy)), __fmul_rn(nz, sz));
        const double d = (double)df;
        acc[0] += (double)__fmul_rn(nx, nx); acc[1] += (double)__fmul_rn(nx, ny); acc[2] += (double)__fmul_rn(nx, nz);
        acc[3] += (double)__fmul_rn(ny, ny); acc[4] += (double)__fmul_rn(ny, nz);
        acc[5] += (double)__fmul_rn(nz, nz);
        acc[6] += a * d;  acc[7] += b * d;  acc[8] += c * d;  acc[9] += dnx * d;  acc[10] += dny * d;  acc[11] += dnz * d;
      }
    }
  }
  // warp reduction (fixed tree); out[] is laid out as the 29 sums of the normal equations
#pragma unroll
  for (int i = 0; i < N; i++) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) out[PASS == 0 ? i : 15 + i] = v;
  }
}

template <bool COOP, int MINB>
__global__ void __launch_bounds__(kThreads, MINB * 128 / kThreads) icp_iter_kernel(PairState* states, double max_d2, float bd_init, int max_iter,
                                                            double eps, int fixed_iters, int first) {
  PairState& S = states[blockIdx.y];
  if (blockIdx.x >= S.nblocks || S.done) return;
  __shared__ double s_part[kThreads / 32][kAcc];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float M[12];
#pragma unroll
  for (int i = 0; i < 12; i++) M[i] = S.Tinc[i];
  const GridDev g = S.tgt;

  const int base = blockIdx.x * kPtsPerCta;
  // phase 1: move the points and search (the double accumulators are not live yet: lower register pressure)
  __shared__ unsigned long long s_key[kPtsPerThread][kThreads];
  __shared__ unsigned short s_items[COOP ? kThreads / 32 : 1][COOP ? kMaxItems : 1];   // lane | ox+8 | oy+8 | upper
  __shared__ int2 s_runs[COOP ? kThreads / 32 : 1][COOP ? kRunCap : 1];                      // start, length << 5 | lane
  __shared__ int s_qi[COOP ? kThreads / 32 : 1][3][32];
  __shared__ float s_qf[COOP ? kThreads / 32 : 1][6][32];
#pragma unroll 1
  for (int it = 0; it < kPtsPerThread; it++) {
    const int k = base + it * kThreads + tid;
    unsigned long long key = pack_key(bd_init, kNoMatch);
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    if (k < S.n_src) {
      float* cp = S.cur + 3 * (size_t)k;
      float x, y, z;
      if (first) {                      // the working copy is created by the first iteration (no separate copy pass)
        if (S.src4) { const er_float4 p = __ldg(&S.src4[k]); x = p.x; y = p.y; z = p.z; }
        else { const float* sp = S.src3 + 3 * (size_t)k; x = __ldg(sp); y = __ldg(sp + 1); z = __ldg(sp + 2); }
      } else {
        x = ER_LDCUR(cp); y = ER_LDCUR(cp + 1); z = ER_LDCUR(cp + 2);
      }
      // Matrix4f * (x, y, z, 1): ((m0*x + m1*y) + m2*z) + m3*1
      sx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[0], x), __fmul_rn(M[1], y)), __fmul_rn(M[2], z)), M[3]);
      sy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[4], x), __fmul_rn(M[5], y)), __fmul_rn(M[6], z)), M[7]);
      sz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[8], x), __fmul_rn(M[9], y)), __fmul_rn(M[10], z)), M[11]);
      ER_STCUR(cp, sx); ER_STCUR(cp + 1, sy); ER_STCUR(cp + 2, sz);
    }
    if (!COOP) {
      if (k < S.n_src) {
        float d2;
        const int m = nn_search(g, sx, sy, sz, d2);
        if (m >= 0) key = pack_key(d2, __float_as_int(__ldg(&g.pts[m]).w));
      }
      s_key[it][tid] = key;
      continue;
    }
    // ---- cooperative search
    const int K = g.rings;
    const float cell = g.cell, slack = 1e-5f * g.cell;
    bool valid = false;
    int cx = 0, cy = 0, cz = 0;
    float fx = 0.0f, fy = 0.0f, fz = 0.0f, bd = bd_init;
    if (k < S.n_src) {
      const float lx = __fmul_rn(__fsub_rn(sx, g.org[0]), g.inv_cell);
      const float ly = __fmul_rn(__fsub_rn(sy, g.org[1]), g.inv_cell);
      const float lz = __fmul_rn(__fsub_rn(sz, g.org[2]), g.inv_cell);
      const float flx = floorf(lx), fly = floorf(ly), flz = floorf(lz);
      // queries far outside the grid (or not finite) cannot have a neighbour within the radius
      valid = flx >= -(float)(K + 1) && flx <= (float)(g.dim[0] + K) && fly >= -(float)(K + 1) && fly <= (float)(g.dim[1] + K) &&
              flz >= -(float)(K + 1) && flz <= (float)(g.dim[2] + K);
      if (valid) {
        cx = (int)flx; cy = (int)fly; cz = (int)flz;
        fx = lx - flx; fy = ly - fly; fz = lz - flz;
        if (cx >= 0 && cx < g.dim[0] && cy >= 0 && cy < g.dim[1] && cz >= 0 && cz < g.dim[2]) {
          const int c = (cx * g.dim[1] + cy) * g.dim[2] + cz;
          nn_scan_key(g, __ldg(&g.cell_start[c]), __ldg(&g.cell_start[c + 1]), sx, sy, sz, key);
          bd = __uint_as_float((unsigned)(key >> 32));
        }
      }
    }
    s_key[it][tid] = key;
    s_qi[warp][0][lane] = cx; s_qi[warp][1][lane] = cy; s_qi[warp][2][lane] = cz;
    s_qf[warp][0][lane] = fx; s_qf[warp][1][lane] = fy; s_qf[warp][2][lane] = fz;
    s_qf[warp][3][lane] = sx; s_qf[warp][4][lane] = sy; s_qf[warp][5][lane] = sz;
    // per axis, the cell offsets whose slab can still hold something as close as bd
    int xlo = 0, xhi = 0, ylo = 0, yhi = 0;
    if (valid) {
      for (int o = 1; o <= K; o++) { const float a = slab_gap(o, fx, cell, slack); if (a * a > bd) break; xhi = o; }
      for (int o = 1; o <= K; o++) { const float a = slab_gap(-o, fx, cell, slack); if (a * a > bd) break; xlo = -o; }
      for (int o = 1; o <= K; o++) { const float a = slab_gap(o, fy, cell, slack); if (a * a > bd) break; yhi = o; }
      for (int o = 1; o <= K; o++) { const float a = slab_gap(-o, fy, cell, slack); if (a * a > bd) break; ylo = -o; }
    }
    // every (query, cell column) that survives becomes one item of the warp's list; the query's own column gives two
    // (the runs below and above the cell already scanned).  The list is ordered by the column's distance from the
    // query's column, nearest first for all queries of the warp, so that the bounds tighten before the far columns
    // are looked at.
    const unsigned lt_mask = (1u << lane) - 1u;
    int total = 0;
    {
      const unsigned bal = __ballot_sync(0xffffffffu, valid);
      if (valid) {
        const int code = lane | (8 << 5) | (8 << 9), at = 2 * __popc(bal & lt_mask);
        s_items[warp][at] = (unsigned short)code;
        s_items[warp][at + 1] = (unsigned short)(code | (1 << 13));
      }
      total = 2 * __popc(bal);
    }
    const int reach = __reduce_max_sync(0xffffffffu, max(max(-xlo, xhi), max(-ylo, yhi)));   // 0 for idle lanes
    const int ncol = reach == 0 ? 1 : (reach == 1 ? 9 : 25);
    for (int t = 1; t < ncol; t++) {
      const int ox = kColumnOrder[t][0], oy = kColumnOrder[t][1];
      const bool in = valid && ox >= xlo && ox <= xhi && oy >= ylo && oy <= yhi;
      const unsigned bal = __ballot_sync(0xffffffffu, in);
      if (in) s_items[warp][total + __popc(bal & lt_mask)] = (unsigned short)(lane | ((ox + 8) << 5) | ((oy + 8) << 9));
      total += __popc(bal);
    }
    __syncwarp();
    // The items are dealt round-robin to the lanes whatever query they belong to.  First every lane derives its
    // column's run of z cells from the freshest bound of the owning query (most items end here) and the surviving
    // point runs are compacted into a second list; then the runs are dealt round-robin and scanned, each run's best
    // candidate merged into its query's key.  Both steps keep most lanes of the warp busy.
    int rn = 0;                                                   // runs waiting in s_runs (warp-uniform)
    auto drain = [&]() {
      __syncwarp();
      for (int r = lane; r < rn; r += 32) {
        const int2 run = s_runs[warp][r];
        const int q = run.y & 31;
        unsigned long long* kp = &s_key[it][warp * 32 + q];
        const unsigned long long cur = *(volatile unsigned long long*)kp;   // any earlier value is a valid bound
        unsigned long long best = cur;
        nn_scan_key(g, run.x, run.x + (run.y >> 5), s_qf[warp][3][q], s_qf[warp][4][q], s_qf[warp][5][q], best);
        if (best < cur) atomicMin(kp, best);
      }
      __syncwarp();
      rn = 0;
    };
    for (int j0 = 0; j0 < total; j0 += 32) {
      const int j = j0 + lane;
      int s0 = 0, s1 = 0, q = 0;
      do {
        if (j >= total) break;
        const int code = (int)s_items[warp][j];
        q = code & 31;
        const int ox = ((code >> 5) & 15) - 8, oy = ((code >> 9) & 15) - 8;
        const bool upper = (code >> 13) & 1;
        const int qcx = s_qi[warp][0][q], qcy = s_qi[warp][1][q], qcz = s_qi[warp][2][q];
        const int ix = qcx + ox, iy = qcy + oy;
        if (ix < 0 || ix >= g.dim[0] || iy < 0 || iy >= g.dim[1]) break;
        const unsigned long long cur = *(volatile unsigned long long*)&s_key[it][warp * 32 + q];
        const float bl = __uint_as_float((unsigned)(cur >> 32));
        const float ax = slab_gap(ox, s_qf[warp][0][q], cell, slack), ay = slab_gap(oy, s_qf[warp][1][q], cell, slack);
        const float axy2 = ax * ax + ay * ay;
        if (axy2 > bl) break;
        const float qfz = s_qf[warp][2][q];
        // z cells of one (x, y) column are contiguous in memory: one run over the cells that can still matter
        int zhi = qcz, zlo = qcz;
        for (int oz = 1; oz <= K; oz++) {
          const float az = slab_gap(oz, qfz, cell, slack);
          if (axy2 + az * az > bl) break;
          zhi = qcz + oz;
        }
        for (int oz = 1; oz <= K; oz++) {
          const float az = slab_gap(-oz, qfz, cell, slack);
          if (axy2 + az * az > bl) break;
          zlo = qcz - oz;
        }
        int a0 = max(zlo, 0), a1 = min(zhi, g.dim[2] - 1);
        if (ox == 0 && oy == 0) {
          if (upper) a0 = max(qcz + 1, 0); else a1 = min(qcz - 1, g.dim[2] - 1);
        }
        if (a0 > a1) break;
        const int row = (ix * g.dim[1] + iy) * g.dim[2];
        s0 = __ldg(&g.cell_start[row + a0]);
        s1 = __ldg(&g.cell_start[row + a1 + 1]);
      } while (false);
      const bool has = s1 > s0;
      const unsigned bal = __ballot_sync(0xffffffffu, has);
      if (has) s_runs[warp][rn + __popc(bal & lt_mask)] = make_int2(s0, ((s1 - s0) << 5) | q);
      rn += __popc(bal);
      if (rn >= 32) drain();                                      // a full round of runs: scan now, tighten the bounds
    }
    drain();
  }
  // phase 2: point-to-plane rows of the accepted correspondences, accumulated in two passes (15 + 14 double sums) so
  // that the accumulators of one pass fit the register budget of the search
  __syncwarp();
  icp_rows<0>(S, g, &s_key[0][tid], base + tid, max_d2, s_part[warp], lane);
  icp_rows<1>(S, g, &s_key[0][tid], base + tid, max_d2, s_part[warp], lane);
  __syncthreads();
  if (tid < kAcc) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) v += s_part[w][tid];
    S.partials[(size_t)blockIdx.x * kAcc + tid] = v;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&S.ticket, 1u) == (unsigned)(S.nblocks - 1));
  __syncthreads();
  if (!s_last) return;
  // ---- the pair's last CTA: finish the iteration
  __shared__ double s_sum[kAcc];
  if (tid < kAcc) {
    double v = 0.0;
    for (int b = 0; b < S.nblocks; b++) v += __ldcg(&S.partials[(size_t)b * kAcc + tid]);
    s_sum[tid] = v;
  }
  __syncthreads();
  if (tid != 0) return;
  S.ticket = 0;
  const int cnt = (int)s_sum[28];
  S.cnt = cnt;
  if (cnt < 3) { S.done = 1; S.converged = 0; return; }          // min_number_correspondences_
  const double mse = s_sum[27] / (double)cnt;
  S.mse = mse;
  double A[36], rhs[6], x[6];
  {
    int q = 0;
    for (int r = 0; r < 6; r++)
      for (int c = r; c < 6; c++) { A[r * 6 + c] = s_sum[q]; A[c * 6 + r] = s_sum[q]; q++; }
    for (int r = 0; r < 6; r++) rhs[r] = s_sum[21 + r];
  }
  if (!solve6(A, rhs, x)) { S.done = 1; S.converged = 0; return; }
  // constructTransformationMatrix(alpha, beta, gamma, tx, ty, tz)
  const double al = x[0], be = x[1], ga = x[2];
  const double sa = sin(al), ca = cos(al), sb = sin(be), cb = cos(be), sg = sin(ga), cg = cos(ga);
  float T[16];
  T[0] = (float)__dmul_rn(cg, cb);
  T[1] = (float)__dadd_rn(__dmul_rn(-sg, ca), __dmul_rn(__dmul_rn(cg, sb), sa));
  T[2] = (float)__dadd_rn(__dmul_rn(sg, sa), __dmul_rn(__dmul_rn(cg, sb), ca));
  T[4] = (float)__dmul_rn(sg, cb);
  T[5] = (float)__dadd_rn(__dmul_rn(cg, ca), __dmul_rn(__dmul_rn(sg, sb), sa));
  T[6] = (float)__dadd_rn(__dmul_rn(-cg, sa), __dmul_rn(__dmul_rn(sg, sb), ca));
  T[8] = (float)(-sb);
  T[9] = (float)__dmul_rn(cb, sa);
  T[10] = (float)__dmul_rn(cb, ca);
  T[3] = (float)x[3]; T[7] = (float)x[4]; T[11] = (float)x[5];
  T[12] = T[13] = T[14] = 0.0f; T[15] = 1.0f;
  float F[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++)
      F[r * 4 + c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[r * 4], S.fin[c]), __fmul_rn(T[r * 4 + 1], S.fin[4 + c])),
                                         __fmul_rn(T[r * 4 + 2], S.fin[8 + c])), __fmul_rn(T[r * 4 + 3], S.fin[12 + c]));
  for (int i = 0; i < 16; i++) { S.fin[i] = F[i]; S.Tinc[i] = T[i]; }
  const int iters = ++S.iters;
  // DefaultConvergenceCriteria::hasConverged
  if (iters >= max_iter) { S.done = 1; S.converged = 1; return; }
  if (!fixed_iters) {
    const double cos_angle = 0.5 * ((double)T[0] + (double)T[5] + (double)T[10] - 1.0);
    const double tsq = (double)T[3] * (double)T[3] + (double)T[7] * (double)T[7] + (double)T[11] * (double)T[11];
    if (cos_angle >= 1.0 - eps && tsq <= eps) { S.done = 1; S.converged = 1; return; }
    if (fabs(mse - S.prev_mse) < 1e-12) { S.done = 1; S.converged = 1; return; }
    S.prev_mse = mse;
  }
}

// pcl::transformPointCloudWithNormals in double, narrowed to float (CorresApp.cpp:145,250)
struct XformD { double T[12]; };

__device__ __forceinline__ void xform_point(const XformD& X, const float* p, float& ox, float& oy, float& oz) {
  const double x = p[0], y = p[1], z = p[2];
  ox = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(X.T[0], x), __dmul_rn(X.T[1], y)), __dmul_rn(X.T[2], z)), X.T[3]);
  oy = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(X.T[4], x), __dmul_rn(X.T[5], y)), __dmul_rn(X.T[6], z)), X.T[7]);
  oz = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(X.T[8], x), __dmul_rn(X.T[9], y)), __dmul_rn(X.T[10], z)), X.T[11]);
}
__device__ __forceinline__ void xform_normal(const XformD& X, const float* n, float& ox, float& oy, float& oz) {
  const double x = n[0], y = n[1], z = n[2];
  ox = (float)__dadd_rn(__dadd_rn(__dmul_rn(X.T[0], x), __dmul_rn(X.T[1], y)), __dmul_rn(X.T[2], z));
  oy = (float)__dadd_rn(__dadd_rn(__dmul_rn(X.T[4], x), __dmul_rn(X.T[5], y)), __dmul_rn(X.T[6], z));
  oz = (float)__dadd_rn(__dadd_rn(__dmul_rn(X.T[8], x), __dmul_rn(X.T[9], y)), __dmul_rn(X.T[10], z));
}

// ---- batched gate / correspondence passes: one job per pair, blockIdx.y = pair, so that a whole candidate list costs
// one launch sequence and one synchronisation (the reference runs these pair-parallel from OpenMP threads,
// CorresApp.cpp:121,220)
struct GateJob {
  GridDev tgt;
  const float* src_xyz;
  int n;
  XformD X;
};
struct CorrJob {
  GridDev tgt;
  const float* src_xyz;
  const float* src_nrm;
  int n, nblk;
  XformD X;
  int* match;        // n
  int* cta_count;    // nblk
  int* offsets;      // nblk + 1
  double* info;      // 13 sums or null
};

// gate of CCorresApp::Registration (CorresApp.cpp:257-264)
__global__ void __launch_bounds__(kThreads) count_inliers_kernel(const GateJob* __restrict__ jobs, double d2max,
                                                                 int* __restrict__ counts) {
  const GateJob& J = jobs[blockIdx.y];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if ((int)(blockIdx.x * blockDim.x) >= J.n) return;
  bool in = false;
  if (k < J.n) {
    float qx, qy, qz, d2;
    xform_point(J.X, J.src_xyz + 3 * (size_t)k, qx, qy, qz);
    in = nn_search(J.tgt, qx, qy, qz, d2) >= 0 && (double)d2 < d2max;
  }
  const unsigned m = __ballot_sync(0xffffffffu, in);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counts[blockIdx.y], __popc(m));
}

// FindCorrespondence search (CorresApp.cpp:145-161): match[k] = original target index or -1
__global__ void __launch_bounds__(kThreads) correspond_kernel(const CorrJob* __restrict__ jobs, double d2max, double ncos) {
  const CorrJob& J = jobs[blockIdx.y];
  if ((int)blockIdx.x >= J.nblk) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  int mt = -1;
  if (k < J.n) {
    float qx, qy, qz, d2;
    xform_point(J.X, J.src_xyz + 3 * (size_t)k, qx, qy, qz);
    const int m = nn_search(J.tgt, qx, qy, qz, d2);
    if (m >= 0 && (double)d2 < d2max) {
      float nx, ny, nz;
      xform_normal(J.X, J.src_nrm + 3 * (size_t)k, nx, ny, nz);
      const float4 tn = __ldg(&J.tgt.nrm[m]);
      const float dot = __fadd_rn(__fadd_rn(__fmul_rn(tn.x, nx), __fmul_rn(tn.y, ny)), __fmul_rn(tn.z, nz));   // NormalDot
      if ((double)dot > ncos) mt = __float_as_int(__ldg(&J.tgt.pts[m]).w);
    }
    J.match[k] = mt;
  }
  const int c = __syncthreads_count(mt >= 0);
  if (threadIdx.x == 0) J.cta_count[blockIdx.x] = c;
}

// exclusive scan of one pair's per-CTA counts (a few hundred to a few thousand entries) by one CTA; total -> offsets[nblk]
// and totals[pair]
__global__ void scan_counts_kernel(const CorrJob* __restrict__ jobs, long long* __restrict__ totals) {
  const CorrJob& J = jobs[blockIdx.x];
  const int nblk = J.nblk;
  const int* counts = J.cta_count;
  int* offsets = J.offsets;
  __shared__ int s_carry;
  __shared__ int s_w[32];
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = (i < nblk) ? counts[i] : 0;
    int x = v;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_w[w] = x;
    __syncthreads();
    if (w == 0) {
      int t = (lane < (int)(blockDim.x >> 5)) ? s_w[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
      s_w[lane] = t;
    }
    __syncthreads();
    const int incl = x + (w > 0 ? s_w[w - 1] : 0) + s_carry;
    if (i < nblk) offsets[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) { offsets[nblk] = s_carry; totals[blockIdx.x] = s_carry; }
}

// first output row of every pair: exclusive prefix of the totals (bases[npairs] = grand total)
__global__ void pair_base_kernel(const long long* __restrict__ totals, long long* __restrict__ bases, int npairs) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  long long acc = 0;
  for (int p = 0; p < npairs; p++) { bases[p] = acc; acc += totals[p]; }
  bases[npairs] = acc;
}

// ordered compaction (ascending source index, CorresApp.cpp:157) + information matrix sums (CorresApp.cpp:186-208)
__global__ void __launch_bounds__(kThreads) compact_kernel(const CorrJob* __restrict__ jobs, const long long* __restrict__ bases,
                                                           int32_t* __restrict__ pairs, long long cap) {
  const CorrJob& J = jobs[blockIdx.y];
  if ((int)blockIdx.x >= J.nblk) return;
  const int n = J.n;
  const float* src_xyz = J.src_xyz;
  double* info_acc = J.info;
  __shared__ int s_w[kThreads / 32];
  __shared__ double s_acc[kThreads / 32][12];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int mt = (k < n) ? J.match[k] : -1;
  const unsigned bal = __ballot_sync(0xffffffffu, mt >= 0);
  if (lane == 0) s_w[w] = __popc(bal);
  __syncthreads();
  int before = 0;
  for (int q = 0; q < w; q++) before += s_w[q];
  double acc[12];
#pragma unroll
  for (int i = 0; i < 12; i++) acc[i] = 0.0;
  if (mt >= 0) {
    const long long pos = bases[blockIdx.y] + (long long)J.offsets[blockIdx.x] + before + __popc(bal & ((1u << lane) - 1));
    if (pairs && pos < cap) { pairs[2 * pos] = mt; pairs[2 * pos + 1] = k; }
    if (info_acc) {
      // A = [I | 0 2sz -2sy; -2sz 0 2sx; 2sy -2sx 0]; entry (r,c) of A^T A = (A0r*A0c + A1r*A1c) + A2r*A2c in double
      const double sx2 = (double)(2.0f * src_xyz[3 * (size_t)k]), sy2 = (double)(2.0f * src_xyz[3 * (size_t)k + 1]),
                   sz2 = (double)(2.0f * src_xyz[3 * (size_t)k + 2]);
      acc[0] = 1.0;                                  // (0,0) = (1,1) = (2,2)
      acc[1] = sz2;  acc[2] = -sy2;                  // (0,4), (0,5)
      acc[3] = -sz2; acc[4] = sx2;                   // (1,3), (1,5)
      acc[5] = sy2;  acc[6] = -sx2;                  // (2,3), (2,4)
      acc[7] = __dadd_rn(__dmul_rn(-sz2, -sz2), __dmul_rn(sy2, sy2));    // (3,3) = (0 + (-2sz)^2) + (2sy)^2
      acc[8] = __dmul_rn(sy2, -sx2);                                      // (3,4) = (0 + 0) + (2sy)(-2sx)
      acc[9] = __dmul_rn(-sz2, sx2);                                      // (3,5) = (0 + (-2sz)(2sx)) + 0
      acc[10] = __dadd_rn(__dmul_rn(sz2, sz2), __dmul_rn(-sx2, -sx2));   // (4,4) = ((2sz)^2 + 0) + (-2sx)^2
      acc[11] = __dmul_rn(sz2, -sy2);                                     // (4,5) = ((2sz)(-2sy) + 0) + 0
      // (5,5) = ((-2sy)^2 + (2sx)^2) + 0 is accumulated from the same operands below
    }
  }
  if (info_acc) {
    double a55 = 0.0;
    if (mt >= 0) {
      const double sx2 = (double)(2.0f * src_xyz[3 * (size_t)k]), sy2 = (double)(2.0f * src_xyz[3 * (size_t)k + 1]);
      a55 = __dadd_rn(__dmul_rn(-sy2, -sy2), __dmul_rn(sx2, sx2));
    }
    // reduce 13 values: reuse acc[0..11] + a55
#pragma unroll
    for (int i = 0; i < 12; i++) {
      double v = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) s_acc[w][i] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a55 += __shfl_down_sync(0xffffffffu, a55, o);
    __syncthreads();
    if (threadIdx.x < 12) {
      double v = 0.0;
      for (int q = 0; q < kThreads / 32; q++) v += s_acc[q][threadIdx.x];
      if (v != 0.0) atomicAdd(&info_acc[threadIdx.x], v);
    }
    if (lane == 0 && a55 != 0.0) atomicAdd(&info_acc[12], a55);
  }
}

// ---------------------------------------------------------------------------------------------
// Grid build (once per fragment, on the device): cell of every point -> counting sort by cell.
//   grid_count_kernel    cid[i] = cell of point i; histogram into A[cell + 2]
//   (inclusive scan of A in place: A[c + 1] = number of points in the cells before c)
//   grid_scatter_kernel  pos = atomicAdd(&A[cid + 1], 1): the cursor of cell c starts at its first slot and ends at
//                        the first slot of cell c + 1, so after the scatter A[0..ncell] is the cell_start table
// The order of the points inside a cell depends on the atomics; the search results do not (nearest neighbour with
// ties to the smallest ORIGINAL index, which travels in pts[].w).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grid_count_kernel(GridDev g, const float* __restrict__ xyz, int n, int* __restrict__ cid,
                                                         int* __restrict__ A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (grid_cell_axis(xyz[3 * (size_t)i], g.org[0], g.inv_cell, g.dim[0]) * g.dim[1] +
                 grid_cell_axis(xyz[3 * (size_t)i + 1], g.org[1], g.inv_cell, g.dim[1])) * g.dim[2] +
                grid_cell_axis(xyz[3 * (size_t)i + 2], g.org[2], g.inv_cell, g.dim[2]);
  cid[i] = c;
  atomicAdd(&A[c + 2], 1);
}

__global__ void __launch_bounds__(256) grid_scatter_kernel(const float* __restrict__ xyz, const float* __restrict__ nrm, int n,
                                                           const int* __restrict__ cid, int* __restrict__ A,
                                                           float4* __restrict__ pts, float4* __restrict__ snrm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pos = atomicAdd(&A[cid[i] + 1], 1);
  pts[pos] = make_float4(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1], xyz[3 * (size_t)i + 2], __int_as_float(i));
  snrm[pos] = make_float4(nrm[3 * (size_t)i], nrm[3 * (size_t)i + 1], nrm[3 * (size_t)i + 2], 0.0f);
}


// puts the points of every cell in ascending original index (insertion sort; a cell holds a handful of points), so that
// the sorted layout -- which is also the order in which the source role feeds the normal equations -- does not depend
// on the order in which the scatter's atomics were served: results are reproducible run to run
__global__ void __launch_bounds__(256) grid_order_cells_kernel(const int* __restrict__ cell_start, size_t ncell,
                                                               float4* __restrict__ pts, float4* __restrict__ nrm) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const int s0 = cell_start[c], s1 = cell_start[c + 1];
  for (int i = s0 + 1; i < s1; i++) {
    const float4 p = pts[i], q = nrm[i];
    const int key = __float_as_int(p.w);
    int j = i - 1;
    while (j >= s0 && __float_as_int(pts[j].w) > key) { pts[j + 1] = pts[j]; nrm[j + 1] = nrm[j]; j--; }
    pts[j + 1] = p; nrm[j + 1] = q;
  }
}

// ---------------------------------------------------------------------------------------------
// Per-device context: the ICP stream and an arena of device / pinned-host buffers that persists across calls (the
// working copies of one 1225-pair batch are 2.9 GB: allocating and freeing them inside every call cost tens of
// milliseconds of host time inside the caller's timed region and made the placement of that memory -- and with it
// the throughput of a latency-bound gather kernel -- differ from call to call).
// ---------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const size_t want = (bytes + (bytes >> 3) + 4095) & ~(size_t)4095;       // 12 % head room: sizes vary a little
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, bytes); if (e == cudaSuccess) cap = bytes; return e; }
    cap = want;
    return cudaSuccess;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct HostBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    const size_t want = (bytes + (bytes >> 2) + 4095) & ~(size_t)4095;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct IcpCtx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t events[kMaxEvents] = {};
  uint64_t launches = 0;
  std::mutex op_mu;                 // one batched operation at a time per device (they share the arena)
  DevBuf cur, part, state;          // align: working copies, CTA partials, PairState[]
  DevBuf jobs, match, counts, offsets, pairs, info, totals;   // gate / correspondence passes
  HostBuf h_state, h_jobs, h_small;
  void release() {
    cur.release(); part.release(); state.release(); jobs.release(); match.release(); counts.release(); offsets.release();
    pairs.release(); info.release(); totals.release(); h_state.release(); h_jobs.release(); h_small.release();
  }
};
std::mutex g_ctx_mu;
std::map<int, IcpCtx*> g_ctx;

IcpCtx* get_ctx(int device) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  auto it = g_ctx.find(device);
  if (it != g_ctx.end()) return it->second;
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  IcpCtx* c = new IcpCtx();
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return nullptr; }
  for (int i = 0; i < kMaxEvents; i++) cudaEventCreate(&c->events[i]);
  g_ctx[device] = c;
  return c;
}

int cuda_fail(const char* what, cudaError_t e) {
  er::set_error("%s: %s", what, cudaGetErrorString(e));
  return ER_ERR_CUDA;
}

}  // namespace

struct er_cloud {
  int device = 0;
  size_t n = 0;
  size_t block_bytes = 0;      // the part of d_block a copy of the cloud needs (everything but the build's scratch)
  size_t off[5] = {};          // byte offsets of pts, nrm, xyz, nrm_orig, cell_start inside d_block
  void* d_block = nullptr;     // one allocation holding everything below
  float4* d_pts = nullptr;
  float4* d_nrm = nullptr;
  int* d_cell_start = nullptr;
  float* d_xyz = nullptr;      // original order
  float* d_nrm_orig = nullptr;
  GridDev grid{};
  double cell = 0;
};

extern "C" {

er_cloud* er_cloud_upload(int device, const float* xyz, const float* normals, size_t n, double cell_size) {
  // an empty fragment (a file that failed to load, or whose normals are all NaN: CorresApp.cpp:93-98 keeps going) is a
  // valid cloud; its buffers may be null
  if ((n > 0 && (!xyz || !normals)) || !(cell_size > 0)) { er::set_error("er_cloud_upload: bad arguments"); return nullptr; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    er::set_error("er_cloud_upload: CUDA device %d not available (%d devices); this library has no CPU path", device, ndev);
    return nullptr;
  }
  if (!get_ctx(device)) { er::set_error("er_cloud_upload: cannot create a stream on device %d", device); return nullptr; }
  er_cloud* c = new (std::nothrow) er_cloud();
  if (!c) return nullptr;
  c->device = device;
  c->n = n;
  // ---- uniform grid over the bounding box (the box on the host, the counting sort on the device; once per fragment)
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (size_t i = 0; i < n; i++)
    for (int a = 0; a < 3; a++) {
      const float v = xyz[3 * i + a];
      if (v < lo[a]) lo[a] = v;
      if (v > hi[a]) hi[a] = v;
    }
  for (int a = 0; a < 3; a++)
    if (n == 0 || !(lo[a] <= hi[a])) lo[a] = hi[a] = 0.f;       // empty cloud, or no finite coordinate on this axis
  // cells of half the admissible radius (a hair more), searched two rings deep: fewer candidates per cell
  int rings = 2;
  if (const char* e = std::getenv("ER_ICP_RINGS")) rings = std::max(1, std::min(6, std::atoi(e)));   // experiment knob
  GridDev& g = c->grid;
  c->cell = er_grid_geometry(lo, hi, cell_size, rings, g);    // the radius this grid supports
  const int* dim = g.dim;
  const size_t ncell = (size_t)dim[0] * dim[1] * dim[2];
  // one device allocation: sorted points and normals, the original-order copies, the cell table (+1 cursor slot),
  // and the build's scratch (cell of every point, scan workspace).  Uploads may run from several host threads at once
  // (CorresApp.cpp:82 loads the fragments from an OpenMP loop): each builds on a stream of its own.
  cudaError_t e = cudaSetDevice(device);
  cudaStream_t st = nullptr;
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  size_t scan_bytes = 0;
  if (e == cudaSuccess) e = cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (int*)nullptr, (int*)nullptr, (int)(ncell + 2), st);
  auto up256 = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t o_pts = 0, o_nrm = o_pts + up256(sizeof(float4) * n), o_xyz = o_nrm + up256(sizeof(float4) * n),
               o_no = o_xyz + up256(sizeof(float) * 3 * n), o_A = o_no + up256(sizeof(float) * 3 * n),
               o_cid = o_A + up256(sizeof(int) * (ncell + 2)), o_scan = o_cid + up256(sizeof(int) * n),
               total = o_scan + up256(scan_bytes) + 256;
  c->off[0] = o_pts; c->off[1] = o_nrm; c->off[2] = o_xyz; c->off[3] = o_no; c->off[4] = o_A;
  c->block_bytes = o_cid;
  if (e == cudaSuccess) e = cudaMalloc(&c->d_block, total);
  if (e == cudaSuccess) {
    char* base = (char*)c->d_block;
    c->d_pts = (float4*)(base + o_pts); c->d_nrm = (float4*)(base + o_nrm);
    c->d_xyz = (float*)(base + o_xyz); c->d_nrm_orig = (float*)(base + o_no);
    c->d_cell_start = (int*)(base + o_A);
    int* d_cid = (int*)(base + o_cid);
    void* d_scan = base + o_scan;
    g.pts = c->d_pts; g.nrm = c->d_nrm; g.cell_start = c->d_cell_start;
    g.xyz_orig = c->d_xyz; g.nrm_orig = c->d_nrm_orig;
    if (n) e = cudaMemcpyAsync(c->d_xyz, xyz, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(c->d_nrm_orig, normals, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_cell_start, 0, sizeof(int) * (ncell + 2), st);
    if (e == cudaSuccess && n) {
      const int nb = (int)((n + 255) / 256);
      grid_count_kernel<<<nb, 256, 0, st>>>(g, c->d_xyz, (int)n, d_cid, c->d_cell_start);
      e = cub::DeviceScan::InclusiveSum(d_scan, scan_bytes, c->d_cell_start, c->d_cell_start, (int)(ncell + 2), st);
      if (e == cudaSuccess) {
        grid_scatter_kernel<<<nb, 256, 0, st>>>(c->d_xyz, c->d_nrm_orig, (int)n, d_cid, c->d_cell_start, c->d_pts, c->d_nrm);
        grid_order_cells_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, st>>>(c->d_cell_start, ncell, c->d_pts, c->d_nrm);
        e = cudaGetLastError();
      }
    }
    // the caller's buffers may be reused after return, and build errors belong to this call
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  if (st) cudaStreamDestroy(st);
  if (e != cudaSuccess) {
    er::set_error("er_cloud_upload: %s", cudaGetErrorString(e));
    er_cloud_free(c);
    return nullptr;
  }
  return c;
}

void er_cloud_free(er_cloud* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaFree(c->d_block);
  delete c;
}

size_t er_cloud_size(const er_cloud* c) { return c ? c->n : 0; }

// ---- a built cloud as a self-contained blob: lets one rank build a fragment's grid and the others receive it over
// NVLink (NCCL broadcast / all-gather of the blob) instead of every rank uploading and sorting every fragment
int er_cloud_get_meta(const er_cloud* c, er_cloud_meta* m, const void** d_blob) {
  ER_ARG(c && m, "cloud_get_meta: null argument");
  std::memset(m, 0, sizeof *m);
  m->n = c->n;
  m->blob_bytes = c->block_bytes;
  for (int i = 0; i < 5; i++) m->offset[i] = c->off[i];
  m->cell = c->cell;
  for (int a = 0; a < 3; a++) { m->org[a] = c->grid.org[a]; m->dim[a] = c->grid.dim[a]; }
  m->inv_cell = c->grid.inv_cell;
  m->cell_f = c->grid.cell;
  m->rings = c->grid.rings;
  if (d_blob) *d_blob = c->d_block;
  return ER_OK;
}

er_cloud* er_cloud_from_blob(int device, const er_cloud_meta* m, const void* d_blob) {
  if (!m || !d_blob || m->blob_bytes == 0) { er::set_error("er_cloud_from_blob: bad arguments"); return nullptr; }
  if (!get_ctx(device)) { er::set_error("er_cloud_from_blob: cannot create a stream on device %d", device); return nullptr; }
  er_cloud* c = new (std::nothrow) er_cloud();
  if (!c) return nullptr;
  c->device = device;
  c->n = (size_t)m->n;
  c->block_bytes = (size_t)m->blob_bytes;
  for (int i = 0; i < 5; i++) c->off[i] = (size_t)m->offset[i];
  c->cell = m->cell;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_block, c->block_bytes);
  if (e == cudaSuccess) e = cudaMemcpy(c->d_block, d_blob, c->block_bytes, cudaMemcpyDeviceToDevice);
  if (e != cudaSuccess) {
    er::set_error("er_cloud_from_blob: %s", cudaGetErrorString(e));
    er_cloud_free(c);
    return nullptr;
  }
  char* base = (char*)c->d_block;
  c->d_pts = (float4*)(base + c->off[0]); c->d_nrm = (float4*)(base + c->off[1]);
  c->d_xyz = (float*)(base + c->off[2]); c->d_nrm_orig = (float*)(base + c->off[3]);
  c->d_cell_start = (int*)(base + c->off[4]);
  GridDev& g = c->grid;
  g.pts = c->d_pts; g.nrm = c->d_nrm; g.cell_start = c->d_cell_start; g.xyz_orig = c->d_xyz; g.nrm_orig = c->d_nrm_orig;
  for (int a = 0; a < 3; a++) { g.org[a] = m->org[a]; g.dim[a] = m->dim[a]; }
  g.inv_cell = m->inv_cell; g.cell = m->cell_f; g.rings = m->rings;
  return c;
}

int er_icp_release(int device) {
  IcpCtx* ctx = get_ctx(device);
  ER_ARG(ctx, "no CUDA context");
  std::lock_guard<std::mutex> lk(ctx->op_mu);
  ER_CUDA(cudaSetDevice(device));
  ER_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->release();
  return ER_OK;
}

// ---- gate
int er_icp_count_inliers_batch(er_cloud* const* clouds, int nclouds, const int* tgt_idx, const int* src_idx, int npairs,
                               const double* T, double dist, int* cnt) {
  ER_ARG(clouds && nclouds > 0 && tgt_idx && src_idx && npairs >= 0 && T && cnt, "count_inliers_batch: bad arguments");
  if (npairs == 0) return ER_OK;
  ER_ARG(clouds[0], "count_inliers_batch: null cloud");
  const int device = clouds[0]->device;
  IcpCtx* ctx = get_ctx(device);
  ER_ARG(ctx, "no CUDA context");
  std::lock_guard<std::mutex> lk(ctx->op_mu);
  ER_CUDA(cudaSetDevice(device));
  cudaError_t e = ctx->h_jobs.ensure(sizeof(GateJob) * (size_t)npairs);
  if (e == cudaSuccess) e = ctx->jobs.ensure(sizeof(GateJob) * (size_t)npairs);
  if (e == cudaSuccess) e = ctx->counts.ensure(sizeof(int) * (size_t)npairs);
  if (e != cudaSuccess) return cuda_fail("count_inliers_batch", e);
  GateJob* hj = (GateJob*)ctx->h_jobs.p;
  int max_blk = 0;
  for (int p = 0; p < npairs; p++) {
    ER_ARG(tgt_idx[p] >= 0 && tgt_idx[p] < nclouds && src_idx[p] >= 0 && src_idx[p] < nclouds, "count_inliers_batch: cloud index out of range");
    const er_cloud* t = clouds[tgt_idx[p]];
    const er_cloud* s = clouds[src_idx[p]];
    ER_ARG(t && s && t->device == device && s->device == device, "count_inliers_batch: clouds must live on one device");
    ER_ARG(dist <= t->cell, "count_inliers: dist exceeds the target cloud's grid cell size");
    hj[p].tgt = t->grid;
    hj[p].src_xyz = s->d_xyz;
    hj[p].n = (int)s->n;
    for (int i = 0; i < 12; i++) hj[p].X.T[i] = T[16 * (size_t)p + i];
    max_blk = std::max(max_blk, (int)((s->n + kThreads - 1) / kThreads));
  }
  e = cudaMemcpyAsync(ctx->jobs.p, hj, sizeof(GateJob) * (size_t)npairs, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(ctx->counts.p, 0, sizeof(int) * (size_t)npairs, ctx->stream);
  if (e == cudaSuccess && max_blk > 0) {
    // blockIdx.y is limited to 65535 pairs per launch
    for (int p0 = 0; p0 < npairs && e == cudaSuccess; p0 += 65535) {
      const int np = std::min(65535, npairs - p0);
      count_inliers_kernel<<<dim3((unsigned)max_blk, (unsigned)np), kThreads, 0, ctx->stream>>>((const GateJob*)ctx->jobs.p + p0, dist * dist,
                                                                                              (int*)ctx->counts.p + p0);
      ctx->launches++;
      e = cudaGetLastError();
    }
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(cnt, ctx->counts.p, sizeof(int) * (size_t)npairs, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail("count_inliers_batch", e);
  return ER_OK;
}

int er_icp_count_inliers(er_cloud* tgt, er_cloud* src, const double T[16], double dist, int* cnt) {
  ER_ARG(tgt && src && T && cnt && tgt->device == src->device, "count_inliers: bad arguments");
  er_cloud* cl[2] = {tgt, src};
  const int ti = 0, si = 1;
  return er_icp_count_inliers_batch(cl, 2, &ti, &si, 1, T, dist, cnt);
}

// ---- ICP
static int align_chunk(IcpCtx* ctx, er_cloud* const* clouds, int nclouds, const int* tgt_idx, const int* src_idx, int npairs,
                       const double* Tinit, double max_dist, int max_iter, double eps, int fixed_iters, double* Tout,
                       int* iters, double* fitness);

int er_icp_align_batch(er_cloud* const* clouds, int nclouds, const int* tgt_idx, const int* src_idx, int npairs,
                       const double* Tinit, double max_dist, int max_iter, double eps, int fixed_iters, double* Tout,
                       int* iters, double* fitness) {
  ER_ARG(clouds && nclouds > 0 && tgt_idx && src_idx && npairs >= 0 && Tinit && Tout && max_iter >= 1, "align_batch: bad arguments");
  if (npairs == 0) return ER_OK;
  ER_ARG(clouds[0], "align_batch: null cloud");
  IcpCtx* ctx = get_ctx(clouds[0]->device);
  ER_ARG(ctx, "no CUDA context");
  std::lock_guard<std::mutex> lk(ctx->op_mu);
  // every pair in flight owns a working copy of its source cloud (12 B per point); ER_ICP_CHUNK bounds that memory
  // by aligning the list in chunks (each chunk pays a synchronisation)
  int chunk = npairs;
  if (const char* ev = std::getenv("ER_ICP_CHUNK")) chunk = std::max(1, std::atoi(ev));             // experiment knob
  chunk = std::min(chunk, 65535);                                                                    // blockIdx.y
  for (int p0 = 0; p0 < npairs; p0 += chunk) {
    const int n = std::min(chunk, npairs - p0);
    const int rc = align_chunk(ctx, clouds, nclouds, tgt_idx + p0, src_idx + p0, n, Tinit + 16 * (size_t)p0, max_dist, max_iter, eps,
                               fixed_iters, Tout + 16 * (size_t)p0, iters ? iters + p0 : nullptr, fitness ? fitness + p0 : nullptr);
    if (rc != ER_OK) return rc;
  }
  return ER_OK;
}

static int align_chunk(IcpCtx* ctx, er_cloud* const* clouds, int nclouds, const int* tgt_idx, const int* src_idx, int npairs,
                       const double* Tinit, double max_dist, int max_iter, double eps, int fixed_iters, double* Tout,
                       int* iters, double* fitness) {
  const int device = ctx->device;
  ER_CUDA(cudaSetDevice(device));
  size_t cur_floats = 0, part_doubles = 0;
  int max_blocks = 1;
  for (int p = 0; p < npairs; p++) {
    ER_ARG(tgt_idx[p] >= 0 && tgt_idx[p] < nclouds && src_idx[p] >= 0 && src_idx[p] < nclouds, "align_batch: cloud index out of range");
    const er_cloud* t = clouds[tgt_idx[p]];
    const er_cloud* s = clouds[src_idx[p]];
    ER_ARG(t && s && t->device == device && s->device == device, "align_batch: clouds must live on one device");
    ER_ARG(max_dist <= t->cell, "align_batch: max_dist exceeds the target cloud's grid cell size");
    const int nb = std::max(1, (int)((s->n + kPtsPerCta - 1) / kPtsPerCta));
    max_blocks = std::max(max_blocks, nb);
    cur_floats += (3 * s->n + 3) & ~(size_t)3;          // every working copy starts 16-byte aligned
    part_doubles += (size_t)nb * kAcc;
  }
  cudaError_t e = ctx->cur.ensure(sizeof(float) * std::max<size_t>(cur_floats, 4));
  if (e == cudaSuccess) e = ctx->part.ensure(sizeof(double) * part_doubles);
  if (e == cudaSuccess) e = ctx->state.ensure(sizeof(PairState) * (size_t)npairs);
  if (e == cudaSuccess) e = ctx->h_state.ensure(sizeof(PairState) * (size_t)npairs);
  if (e != cudaSuccess) return cuda_fail("align_batch (arena)", e);
  // the order in which the source points enter the search and the sums: the fragment's cell-sorted order (queries of a
  // warp are neighbours in space: coherent gathers) unless ER_ICP_SRC_ORDER=file asks for the file order
  bool sorted_src = true;
  if (const char* ev = std::getenv("ER_ICP_SRC_ORDER")) sorted_src = std::strcmp(ev, "file") != 0;    // experiment knob
  PairState* st = (PairState*)ctx->h_state.p;
  float* d_cur = (float*)ctx->cur.p;
  double* d_part = (double*)ctx->part.p;
  PairState* d_st = (PairState*)ctx->state.p;
  size_t co = 0, po = 0;
  int max_rings = 1;
  for (int p = 0; p < npairs; p++) {
    const er_cloud* t = clouds[tgt_idx[p]];
    const er_cloud* s = clouds[src_idx[p]];
    PairState& S = st[(size_t)p];
    std::memset(&S, 0, sizeof S);
    S.tgt = t->grid;
    S.cur = d_cur + co;
    S.src4 = sorted_src ? (const er_float4*)s->d_pts : nullptr;
    S.src3 = s->d_xyz;
    S.n_src = (int)s->n;
    S.nblocks = std::max(1, (int)((s->n + kPtsPerCta - 1) / kPtsPerCta));
    S.partials = d_part + po;
    for (int i = 0; i < 16; i++) S.Tinc[i] = S.fin[i] = (float)Tinit[16 * (size_t)p + i];   // guess.cast<float>()
    S.prev_mse = DBL_MAX;
    co += (3 * s->n + 3) & ~(size_t)3;
    po += (size_t)S.nblocks * kAcc;
    max_rings = std::max(max_rings, t->grid.rings);
  }
  e = cudaMemcpyAsync(d_st, st, sizeof(PairState) * (size_t)npairs, cudaMemcpyHostToDevice, ctx->stream);
  const double max_d2 = max_dist * max_dist;
  // the smallest float that is >= max_d2: no candidate with a larger float distance can pass `(double)d2 > max_d2`
  float bd_init = (float)max_d2;
  if ((double)bd_init < max_d2) bd_init = std::nextafterf(bd_init, FLT_MAX);
  bool coop = true;
  if (const char* ev = std::getenv("ER_ICP_COOP")) coop = std::atoi(ev) != 0;                        // experiment knob
  if (max_rings > kMaxRings) coop = false;              // the warp's item list is sized for kMaxRings
  int minb = 8;                                         // resident CTAs per SM the kernel is compiled for (register budget)
  if (const char* ev = std::getenv("ER_ICP_MINB")) minb = std::atoi(ev) == 6 ? 6 : 8;              // experiment knob
  const dim3 grid((unsigned)max_blocks, (unsigned)npairs);
  for (int it = 0; it < max_iter && e == cudaSuccess; it++) {
    const int first = it == 0;
    if (coop && minb == 8) icp_iter_kernel<true, 8><<<grid, kThreads, 0, ctx->stream>>>(d_st, max_d2, bd_init, max_iter, eps, fixed_iters, first);
    else if (coop) icp_iter_kernel<true, 6><<<grid, kThreads, 0, ctx->stream>>>(d_st, max_d2, bd_init, max_iter, eps, fixed_iters, first);
    else if (minb == 8) icp_iter_kernel<false, 8><<<grid, kThreads, 0, ctx->stream>>>(d_st, max_d2, bd_init, max_iter, eps, fixed_iters, first);
    else icp_iter_kernel<false, 6><<<grid, kThreads, 0, ctx->stream>>>(d_st, max_d2, bd_init, max_iter, eps, fixed_iters, first);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(st, d_st, sizeof(PairState) * (size_t)npairs, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail("align_batch", e);
  for (int p = 0; p < npairs; p++) {
    for (int i = 0; i < 16; i++) Tout[16 * (size_t)p + i] = (double)st[(size_t)p].fin[i];    // .cast<double>(), CorresApp.cpp:312
    if (iters) iters[p] = st[(size_t)p].iters;
    if (fitness) fitness[p] = st[(size_t)p].mse;
  }
  return ER_OK;
}

int er_icp_align(er_cloud* tgt, er_cloud* src, const double Tinit[16], double max_dist, int max_iter, double eps,
                 int fixed_iters, double Tout[16], int* iters, double* fitness) {
  ER_ARG(tgt && src, "align: null cloud");
  er_cloud* cl[2] = {tgt, src};
  const int ti = 0, si = 1;
  return er_icp_align_batch(cl, 2, &ti, &si, 1, Tinit, max_dist, max_iter, eps, fixed_iters, Tout, iters, fitness);
}

// ---- correspondences
int er_icp_correspondences_batch(er_cloud* const* clouds, int nclouds, const int* tgt_idx, const int* src_idx, int npairs,
                                 const double* T, double dist, double normal_cos, int32_t* pairs, size_t cap,
                                 size_t* row_start, double* info36) {
  ER_ARG(clouds && nclouds > 0 && tgt_idx && src_idx && npairs >= 0 && T && row_start, "correspondences_batch: bad arguments");
  row_start[0] = 0;
  if (npairs == 0) return ER_OK;
  ER_ARG(clouds[0], "correspondences_batch: null cloud");
  const int device = clouds[0]->device;
  IcpCtx* ctx = get_ctx(device);
  ER_ARG(ctx, "no CUDA context");
  std::lock_guard<std::mutex> lk(ctx->op_mu);
  ER_CUDA(cudaSetDevice(device));
  size_t n_total = 0, blk_total = 0;
  int max_blk = 0;
  for (int p = 0; p < npairs; p++) {
    ER_ARG(tgt_idx[p] >= 0 && tgt_idx[p] < nclouds && src_idx[p] >= 0 && src_idx[p] < nclouds, "correspondences_batch: cloud index out of range");
    const er_cloud* t = clouds[tgt_idx[p]];
    const er_cloud* s = clouds[src_idx[p]];
    ER_ARG(t && s && t->device == device && s->device == device, "correspondences_batch: clouds must live on one device");
    ER_ARG(dist <= t->cell, "correspondences: dist exceeds the target cloud's grid cell size");
    const int nblk = (int)((s->n + kThreads - 1) / kThreads);
    n_total += s->n;
    blk_total += (size_t)nblk + 1;
    max_blk = std::max(max_blk, nblk);
  }
  const size_t pcap = pairs ? std::min(cap, n_total) : 0;
  cudaError_t e = ctx->h_jobs.ensure(sizeof(CorrJob) * (size_t)npairs);
  if (e == cudaSuccess) e = ctx->jobs.ensure(sizeof(CorrJob) * (size_t)npairs);
  if (e == cudaSuccess) e = ctx->match.ensure(sizeof(int) * std::max<size_t>(n_total, 1));
  if (e == cudaSuccess) e = ctx->counts.ensure(sizeof(int) * blk_total);
  if (e == cudaSuccess) e = ctx->offsets.ensure(sizeof(int) * blk_total);
  if (e == cudaSuccess) e = ctx->totals.ensure(sizeof(long long) * (2 * (size_t)npairs + 1));
  if (e == cudaSuccess) e = ctx->info.ensure(sizeof(double) * 13 * (size_t)npairs);
  if (e == cudaSuccess && pcap) e = ctx->pairs.ensure(sizeof(int32_t) * 2 * pcap);
  if (e == cudaSuccess) e = ctx->h_small.ensure((sizeof(long long) * (2 * (size_t)npairs + 1)) + sizeof(double) * 13 * (size_t)npairs);
  if (e != cudaSuccess) return cuda_fail("correspondences_batch (arena)", e);
  CorrJob* hj = (CorrJob*)ctx->h_jobs.p;
  size_t no = 0, bo = 0;
  for (int p = 0; p < npairs; p++) {
    const er_cloud* t = clouds[tgt_idx[p]];
    const er_cloud* s = clouds[src_idx[p]];
    CorrJob& J = hj[p];
    J.tgt = t->grid;
    J.src_xyz = s->d_xyz;
    J.src_nrm = s->d_nrm_orig;
    J.n = (int)s->n;
    J.nblk = (int)((s->n + kThreads - 1) / kThreads);
    for (int i = 0; i < 12; i++) J.X.T[i] = T[16 * (size_t)p + i];
    J.match = (int*)ctx->match.p + no;
    J.cta_count = (int*)ctx->counts.p + bo;
    J.offsets = (int*)ctx->offsets.p + bo;
    J.info = info36 ? (double*)ctx->info.p + 13 * (size_t)p : nullptr;
    no += s->n;
    bo += (size_t)J.nblk + 1;
  }
  long long* d_totals = (long long*)ctx->totals.p;
  long long* d_bases = d_totals + npairs;
  e = cudaMemcpyAsync(ctx->jobs.p, hj, sizeof(CorrJob) * (size_t)npairs, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && info36) e = cudaMemsetAsync(ctx->info.p, 0, sizeof(double) * 13 * (size_t)npairs, ctx->stream);
  const CorrJob* dj = (const CorrJob*)ctx->jobs.p;
  if (e == cudaSuccess) {
    if (max_blk > 0)
      for (int p0 = 0; p0 < npairs; p0 += 65535)
        correspond_kernel<<<dim3((unsigned)max_blk, (unsigned)std::min(65535, npairs - p0)), kThreads, 0, ctx->stream>>>(dj + p0, dist * dist, normal_cos);
    scan_counts_kernel<<<(unsigned)npairs, 256, 0, ctx->stream>>>(dj, d_totals);
    pair_base_kernel<<<1, 32, 0, ctx->stream>>>(d_totals, d_bases, npairs);
    if (max_blk > 0)
      for (int p0 = 0; p0 < npairs; p0 += 65535)
        compact_kernel<<<dim3((unsigned)max_blk, (unsigned)std::min(65535, npairs - p0)), kThreads, 0, ctx->stream>>>(
            dj + p0, d_bases + p0, (int32_t*)(pcap ? ctx->pairs.p : nullptr), (long long)pcap);
    ctx->launches += 4;
    e = cudaGetLastError();
  }
  long long* h_bases = (long long*)ctx->h_small.p;
  double* h_info = (double*)(h_bases + 2 * (size_t)npairs + 1);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_bases, d_bases, sizeof(long long) * ((size_t)npairs + 1), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && info36) e = cudaMemcpyAsync(h_info, ctx->info.p, sizeof(double) * 13 * (size_t)npairs, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail("correspondences_batch", e);
  for (int p = 0; p <= npairs; p++) row_start[p] = (size_t)h_bases[p];
  const size_t total = (size_t)h_bases[npairs];
  if (pairs && total > 0) {
    const size_t m = std::min(total, pcap);
    if (m) ER_CUDA(cudaMemcpy(pairs, ctx->pairs.p, sizeof(int32_t) * 2 * m, cudaMemcpyDeviceToHost));
  }
  if (info36) {
    for (int p = 0; p < npairs; p++) {
      const double* acc = h_info + 13 * (size_t)p;
      double* I = info36 + 36 * (size_t)p;
      std::memset(I, 0, sizeof(double) * 36);
      I[0] = I[7] = I[14] = acc[0];
      I[0 * 6 + 4] = I[4 * 6 + 0] = acc[1];  I[0 * 6 + 5] = I[5 * 6 + 0] = acc[2];
      I[1 * 6 + 3] = I[3 * 6 + 1] = acc[3];  I[1 * 6 + 5] = I[5 * 6 + 1] = acc[4];
      I[2 * 6 + 3] = I[3 * 6 + 2] = acc[5];  I[2 * 6 + 4] = I[4 * 6 + 2] = acc[6];
      I[3 * 6 + 3] = acc[7];  I[3 * 6 + 4] = I[4 * 6 + 3] = acc[8];  I[3 * 6 + 5] = I[5 * 6 + 3] = acc[9];
      I[4 * 6 + 4] = acc[10]; I[4 * 6 + 5] = I[5 * 6 + 4] = acc[11];
      I[5 * 6 + 5] = acc[12];
    }
  }
  if (pairs && total > cap) {
    er::set_error("correspondences_batch: %zu rows do not fit the caller's %zu (row_start is filled; retry with a larger buffer)", total, cap);
    return ER_ERR_CAPACITY;
  }
  return ER_OK;
}

int er_icp_correspondences(er_cloud* tgt, er_cloud* src, const double T[16], double dist, double normal_cos, int32_t* pairs,
                           size_t cap, size_t* n_out, double info36[36]) {
  ER_ARG(tgt && src && T && n_out && tgt->device == src->device, "correspondences: bad arguments");
  er_cloud* cl[2] = {tgt, src};
  const int ti = 0, si = 1;
  size_t rs[2] = {0, 0};
  *n_out = 0;
  const int rc = er_icp_correspondences_batch(cl, 2, &ti, &si, 1, T, dist, normal_cos, pairs, cap, rs, info36);
  *n_out = rs[1];
  // the single-pair call keeps its old contract: more rows than `cap` is not an error, n reports how many there are
  return rc == ER_ERR_CAPACITY ? ER_OK : rc;
}

int er_icp_event_record(int device, int slot) {
  IcpCtx* ctx = get_ctx(device);
  ER_ARG(ctx && slot >= 0 && slot < kMaxEvents, "icp event: bad device or slot");
  ER_CUDA(cudaSetDevice(device));
  ER_CUDA(cudaEventRecord(ctx->events[slot], ctx->stream));
  return ER_OK;
}

int er_icp_event_elapsed_ms(int device, int a, int b, float* ms) {
  IcpCtx* ctx = get_ctx(device);
  ER_ARG(ctx && ms && a >= 0 && a < kMaxEvents && b >= 0 && b < kMaxEvents, "icp event: bad device or slot");
  ER_CUDA(cudaSetDevice(device));
  ER_CUDA(cudaEventSynchronize(ctx->events[b]));
  ER_CUDA(cudaEventElapsedTime(ms, ctx->events[a], ctx->events[b]));
  return ER_OK;
}

uint64_t er_icp_launch_count(int device) {
  IcpCtx* ctx = get_ctx(device);
  return ctx ? ctx->launches : 0;
}

int er_icp_sync(int device) {
  IcpCtx* ctx = get_ctx(device);
  ER_ARG(ctx, "no CUDA context");
  ER_CUDA(cudaSetDevice(device));
  ER_CUDA(cudaStreamSynchronize(ctx->stream));
  return ER_OK;
}

}  // extern "C"
