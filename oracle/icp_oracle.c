/*
 * oracle/icp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the pairwise ICP refinement path of BuildCorrespondence, used only by tests/,
 * __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) as the checker for the CUDA path.
 *
 * PARITY UNPINNED.  The control flow and thresholds follow the reference's own code
 * (BuildCorrespondence/CorresApp.cpp:112-319, cited below), but the arithmetic of the inner loop lives
 * in a third-party dependency that is NOT in /root/reference and not installed here: PCL, include dir
 * "pcl-1.7" (BuildCorrespondence/BuildCorrespondence.vcxproj:109,128; README.txt:7-8 names the authors'
 * fork qianyizh/StanfordPCL "development version", no version pin) with FLANN (unpinned).  No test, golden
 * vector or fixture of the reference pins ICP results.  What follows restates the published PCL 1.7
 * algorithm (registration/impl/icp.hpp computeTransformation, correspondence_estimation.hpp
 * determineCorrespondences, transformation_estimation_point_to_plane_lls.hpp estimateRigidTransformation +
 * constructTransformationMatrix, default_convergence_criteria.hpp hasConverged, kdtree_flann.hpp with
 * flann::L2_Simple<float>, common/impl/transforms.hpp transformPointCloudWithNormals) from knowledge of those
 * sources; this file is therefore the DEFINITION the CUDA path is held to (tests cross-check its nearest
 * neighbour search against brute force and scipy.spatial.cKDTree and its 6x6 solve against numpy).  The FLANN half IS
 * pinned against real FLANN code since round 2: OpenCV bundles FLANN, cv2.flann_Index(algorithm=4) is the
 * flann::KDTreeSingleIndex that pcl::KdTreeFLANN instantiates, and tests/test_oracle_icp.py
 * (test_search_and_gate_against_a_real_flann_build) finds this file's search distances bit-equal to its exact search and
 * the gate count of CorresApp.cpp:257-264 equal.  The Eigen
 * half of PCL's arithmetic IS pinned: oracle/eigen_check.cpp runs the Eigen 3.1.2 the reference vendors, and
 * tests/test_oracle_icp.py finds the Matrix4f products restated below bit-identical to it and the 6x6 solve within
 * 1.6e-10 relative of ATA.inverse() * ATb (no difference left after the narrowing to float).
 *
 * Deliberate choices where PCL/FLANN leave freedom:
 *   - exact 1-NN on xyz with squared distance ((dx*dx + dy*dy) + dz*dz) in float (L2_Simple order); among
 *     equal distances the smallest target index wins (FLANN's tie order is unspecified);
 *   - the 6x6 system is solved by Gaussian elimination with partial pivoting (Eigen's inverse() of a 6x6 is
 *     a partial-pivot LU); agreement with an explicit inverse is ~1e-12 relative.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int n;
  float *xyz, *nrm;          /* n*3 each, caller's order */
  /* uniform grid over the bounding box, cell edge >= the largest search radius */
  double cell;
  float inv_cell;
  float org[3];
  int dim[3];
  int *cell_start;           /* dim0*dim1*dim2 + 1 */
  int *order;                /* point indices sorted by cell */
} oracle_cloud;

static int cell_of(const oracle_cloud *c, float v, int a) {
  int i = (int)floorf((v - c->org[a]) * c->inv_cell);
  if (i < 0) i = 0;
  if (i >= c->dim[a]) i = c->dim[a] - 1;
  return i;
}

oracle_cloud *oracle_cloud_create(const float *xyz, const float *nrm, int n, double cell) {
  oracle_cloud *c = (oracle_cloud *)calloc(1, sizeof *c);
  c->n = n;
  c->xyz = (float *)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
  c->nrm = (float *)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
  memcpy(c->xyz, xyz, sizeof(float) * 3 * (size_t)n);
  memcpy(c->nrm, nrm, sizeof(float) * 3 * (size_t)n);
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = 0; i < n; i++)
    for (int a = 0; a < 3; a++) {
      float v = xyz[3 * i + a];
      if (v < lo[a]) lo[a] = v;
      if (v > hi[a]) hi[a] = v;
    }
  if (n == 0) for (int a = 0; a < 3; a++) { lo[a] = 0; hi[a] = 0; }
  c->cell = cell * 1.001;      /* a hair larger than the radius so float cell arithmetic cannot lose a neighbour */
  for (;;) {
    double cells = 1;
    for (int a = 0; a < 3; a++) {
      c->dim[a] = (int)floor((hi[a] - lo[a]) / c->cell) + 1;
      cells *= c->dim[a];
    }
    if (cells <= 64e6) break;
    c->cell *= 2;
  }
  c->inv_cell = (float)(1.0 / c->cell);
  for (int a = 0; a < 3; a++) c->org[a] = lo[a];
  size_t ncell = (size_t)c->dim[0] * c->dim[1] * c->dim[2];
  c->cell_start = (int *)calloc(ncell + 1, sizeof(int));
  c->order = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int *cid = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) {
    int cx = cell_of(c, xyz[3 * i], 0), cy = cell_of(c, xyz[3 * i + 1], 1), cz = cell_of(c, xyz[3 * i + 2], 2);
    cid[i] = (cx * c->dim[1] + cy) * c->dim[2] + cz;
    c->cell_start[cid[i] + 1]++;
  }
  for (size_t k = 0; k < ncell; k++) c->cell_start[k + 1] += c->cell_start[k];
  int *fill = (int *)malloc(sizeof(int) * ncell);
  memcpy(fill, c->cell_start, sizeof(int) * ncell);
  for (int i = 0; i < n; i++) c->order[fill[cid[i]]++] = i;   /* ascending index inside a cell */
  free(fill);
  free(cid);
  return c;
}

void oracle_cloud_free(oracle_cloud *c) {
  if (!c) return;
  free(c->xyz); free(c->nrm); free(c->cell_start); free(c->order); free(c);
}
int oracle_cloud_size(const oracle_cloud *c) { return c->n; }

/* pcl::KdTreeFLANN::nearestKSearch(q, 1, ...) restricted to candidates within one cell ring: returns the
 * index of the nearest target point and its squared float distance, or -1 if the 27 surrounding cells are
 * empty.  Every target point within c->cell of q lies in those cells, so for any radius <= cell the result
 * is the exact nearest neighbour whenever that neighbour is within the radius. */
static int nn_search(const oracle_cloud *c, const float *q, float *d2_out) {
  int best = -1;
  float bd = FLT_MAX;
  /* the query may lie outside the target's bounding box */
  int qc[3];
  for (int a = 0; a < 3; a++) qc[a] = (int)floorf((q[a] - c->org[a]) * c->inv_cell);
  for (int ix = qc[0] - 1; ix <= qc[0] + 1; ix++) {
    if (ix < 0 || ix >= c->dim[0]) continue;
    for (int iy = qc[1] - 1; iy <= qc[1] + 1; iy++) {
      if (iy < 0 || iy >= c->dim[1]) continue;
      for (int iz = qc[2] - 1; iz <= qc[2] + 1; iz++) {
        if (iz < 0 || iz >= c->dim[2]) continue;
        int cell = (ix * c->dim[1] + iy) * c->dim[2] + iz;
        for (int s = c->cell_start[cell]; s < c->cell_start[cell + 1]; s++) {
          int p = c->order[s];
          float dx = q[0] - c->xyz[3 * p], dy = q[1] - c->xyz[3 * p + 1], dz = q[2] - c->xyz[3 * p + 2];
          float d2 = dx * dx + dy * dy + dz * dz;      /* flann::L2_Simple<float>: running float sum in x,y,z order */
          if (d2 < bd || (d2 == bd && p < best)) { bd = d2; best = p; }
        }
      }
    }
  }
  *d2_out = bd;
  return best;
}

/* test hook: brute-force 1-NN with the same distance and tie rule */
int oracle_nn_brute(const oracle_cloud *c, const float *q, float *d2_out) {
  int best = -1;
  float bd = FLT_MAX;
  for (int p = 0; p < c->n; p++) {
    float dx = q[0] - c->xyz[3 * p], dy = q[1] - c->xyz[3 * p + 1], dz = q[2] - c->xyz[3 * p + 2];
    float d2 = dx * dx + dy * dy + dz * dz;
    if (d2 < bd) { bd = d2; best = p; }
  }
  *d2_out = bd;
  return best;
}
int oracle_nn_grid(const oracle_cloud *c, const float *q, float *d2_out) { return nn_search(c, q, d2_out); }

/* pcl::transformPointCloudWithNormals(cloud, out, Matrix4d) (used at CorresApp.cpp:145,250): per point in
 * double, each component narrowed to float; T row-major double[16]. */
static void transform_with_normals(const oracle_cloud *s, const double *T, float *xyz, float *nrm) {
  for (int i = 0; i < s->n; i++) {
    const float *p = s->xyz + 3 * i, *n = s->nrm + 3 * i;
    for (int r = 0; r < 3; r++) {
      xyz[3 * i + r] = (float)(T[r * 4 + 0] * p[0] + T[r * 4 + 1] * p[1] + T[r * 4 + 2] * p[2] + T[r * 4 + 3]);
      if (nrm) nrm[3 * i + r] = (float)(T[r * 4 + 0] * n[0] + T[r * 4 + 1] * n[1] + T[r * 4 + 2] * n[2]);
    }
  }
}

/* ------------------------------------------------------------------ b2 gate: CorresApp.cpp:250-264 */
int oracle_icp_count_inliers(const oracle_cloud *tgt, const oracle_cloud *src, const double *T, double dist, int threads) {
  float *xyz = (float *)malloc(sizeof(float) * 3 * (size_t)(src->n > 0 ? src->n : 1));
  transform_with_normals(src, T, xyz, NULL);
  int cnt = 0;
  const double d2max = dist * dist;
#pragma omp parallel for num_threads(threads) reduction(+ : cnt) schedule(static)
  for (int k = 0; k < src->n; k++) {
    float d2;
    if (nn_search(tgt, xyz + 3 * k, &d2) >= 0 && (double)d2 < d2max) cnt++;     /* CorresApp.cpp:259-261 */
  }
  free(xyz);
  return cnt;
}

/* 6x6 solve, partial pivoting; returns 0 on success */
static int solve6(double A[36], double b[6], double x[6]) {
  int n = 6;
  for (int col = 0; col < n; col++) {
    int piv = col;
    double big = fabs(A[col * 6 + col]);
    for (int r = col + 1; r < n; r++)
      if (fabs(A[r * 6 + col]) > big) { big = fabs(A[r * 6 + col]); piv = r; }
    if (big == 0.0) return -1;
    if (piv != col) {
      for (int c = 0; c < n; c++) { double t = A[col * 6 + c]; A[col * 6 + c] = A[piv * 6 + c]; A[piv * 6 + c] = t; }
      double t = b[col]; b[col] = b[piv]; b[piv] = t;
    }
    for (int r = col + 1; r < n; r++) {
      double f = A[r * 6 + col] / A[col * 6 + col];
      for (int c = col; c < n; c++) A[r * 6 + c] -= f * A[col * 6 + c];
      b[r] -= f * b[col];
    }
  }
  for (int r = n - 1; r >= 0; r--) {
    double s = b[r];
    for (int c = r + 1; c < n; c++) s -= A[r * 6 + c] * x[c];
    x[r] = s / A[r * 6 + r];
  }
  return 0;
}
int oracle_solve6(const double *A, const double *b, double *x) {
  double a[36], bb[6];
  memcpy(a, A, sizeof a); memcpy(bb, b, sizeof bb);
  return solve6(a, bb, x);
}

/* Matrix4f product / point transform in Eigen's order: ((a0*b0 + a1*b1) + a2*b2) + a3*b3, float */
static void mat4f_mul(const float *a, const float *b, float *o) {
  float t[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++)
      t[r * 4 + c] = ((a[r * 4] * b[c] + a[r * 4 + 1] * b[4 + c]) + a[r * 4 + 2] * b[8 + c]) + a[r * 4 + 3] * b[12 + c];
  memcpy(o, t, sizeof t);
}
static void transform_points_f(const float *M, float *xyz, int n) {   /* IterativeClosestPoint::transformCloud */
  for (int i = 0; i < n; i++) {
    float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    for (int r = 0; r < 3; r++) xyz[3 * i + r] = ((M[r * 4] * x + M[r * 4 + 1] * y) + M[r * 4 + 2] * z) + M[r * 4 + 3] * 1.0f;
  }
}

/* exported for tests/test_oracle_icp.py, which compares them with the vendored Eigen 3.1.2 (oracle/eigen_check.cpp) */
void oracle_mat4f_mul(const float *a, const float *b, float *o) { mat4f_mul(a, b, o); }
void oracle_transform_point_f(const float *M, const float *p, float *o) {
  float q[3] = {p[0], p[1], p[2]};
  transform_points_f(M, q, 1);
  o[0] = q[0]; o[1] = q[1]; o[2] = q[2];
}

/* ------------------------------------------------------------------ b2/b3: CorresApp.cpp:295-312
 * pcl::IterativeClosestPoint<PointXYZRGBNormal,PointXYZRGBNormal> with TransformationEstimationPointToPlaneLLS,
 * source = src, target = tgt, setMaxCorrespondenceDistance(max_dist), setMaximumIterations(max_iter),
 * setTransformationEpsilon(eps), align(out, Tinit.cast<float>()).
 * Returns 1 if PCL's hasConverged() fired, 0 if the loop broke on < 3 correspondences. */
int oracle_icp_align(const oracle_cloud *tgt, const oracle_cloud *src, const double *Tinit, double max_dist, int max_iter,
                     double eps, int fixed_iters, double *Tout, int *iters_out, double *fitness_out, int threads) {
  const int n = src->n;
  float final[16], guess[16];
  for (int i = 0; i < 16; i++) guess[i] = final[i] = (float)Tinit[i];
  float *cur = (float *)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
  memcpy(cur, src->xyz, sizeof(float) * 3 * (size_t)n);
  int is_identity = 1;
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) if (guess[r * 4 + c] != (r == c ? 1.0f : 0.0f)) is_identity = 0;
  if (!is_identity) transform_points_f(guess, cur, n);        /* icp.hpp: input_transformed = guess * input */
  int *match = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  float *md2 = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  const double max_d2 = max_dist * max_dist;
  const double rot_thr = 1.0 - eps, trans_thr = eps;          /* setRotationThreshold / setTranslationThreshold */
  double prev_mse = DBL_MAX, mse = 0;
  int iters = 0, converged = 0;
  for (;;) {
    /* determineCorrespondences: every source point in order, keep unless dist^2 > max_dist^2 */
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int k = 0; k < n; k++) {
      float d2;
      int m = nn_search(tgt, cur + 3 * k, &d2);
      if (m >= 0 && (double)d2 > max_d2) m = -1;
      match[k] = m;
      md2[k] = d2;
    }
    int cnt = 0;
    double ATA[36], ATb[6], sumd = 0;
    memset(ATA, 0, sizeof ATA); memset(ATb, 0, sizeof ATb);
    for (int k = 0; k < n; k++) {
      if (match[k] < 0) continue;
      cnt++;
      sumd += md2[k];
      /* estimateRigidTransformation: the operands are floats, so a, b, c, d and the n*n products are float
       * expressions that are widened afterwards */
      const float sx = cur[3 * k], sy = cur[3 * k + 1], sz = cur[3 * k + 2];
      const float *t = tgt->xyz + 3 * match[k], *nn = tgt->nrm + 3 * match[k];
      const float dx = t[0], dy = t[1], dz = t[2], nx = nn[0], ny = nn[1], nz = nn[2];
      if (!isfinite(sx) || !isfinite(sy) || !isfinite(sz) || !isfinite(dx) || !isfinite(dy) || !isfinite(dz) ||
          !isfinite(nx) || !isfinite(ny) || !isfinite(nz)) continue;
      double a = nz * sy - ny * sz;
      double b = nx * sz - nz * sx;
      double c = ny * sx - nx * sy;
      ATA[0] += a * a;  ATA[1] += a * b;  ATA[2] += a * c;  ATA[3] += a * nx;  ATA[4] += a * ny;  ATA[5] += a * nz;
      ATA[7] += b * b;  ATA[8] += b * c;  ATA[9] += b * nx; ATA[10] += b * ny; ATA[11] += b * nz;
      ATA[14] += c * c; ATA[15] += c * nx; ATA[16] += c * ny; ATA[17] += c * nz;
      ATA[21] += nx * nx; ATA[22] += nx * ny; ATA[23] += nx * nz;
      ATA[28] += ny * ny; ATA[29] += ny * nz;
      ATA[35] += nz * nz;
      double d = nx * dx + ny * dy + nz * dz - nx * sx - ny * sy - nz * sz;
      ATb[0] += a * d; ATb[1] += b * d; ATb[2] += c * d; ATb[3] += nx * d; ATb[4] += ny * d; ATb[5] += nz * d;
    }
    if (cnt < 3) { converged = 0; break; }                    /* min_number_correspondences_ */
    mse = sumd / (double)cnt;
    for (int r = 1; r < 6; r++) for (int c = 0; c < r; c++) ATA[r * 6 + c] = ATA[c * 6 + r];
    double x[6];
    if (solve6(ATA, ATb, x) != 0) { converged = 0; break; }
    /* constructTransformationMatrix(alpha, beta, gamma, tx, ty, tz): R = Rz(gamma) Ry(beta) Rx(alpha) */
    const double al = x[0], be = x[1], ga = x[2];
    float Tinc[16];
    Tinc[0] = (float)(cos(ga) * cos(be));
    Tinc[1] = (float)(-sin(ga) * cos(al) + cos(ga) * sin(be) * sin(al));
    Tinc[2] = (float)(sin(ga) * sin(al) + cos(ga) * sin(be) * cos(al));
    Tinc[4] = (float)(sin(ga) * cos(be));
    Tinc[5] = (float)(cos(ga) * cos(al) + sin(ga) * sin(be) * sin(al));
    Tinc[6] = (float)(-cos(ga) * sin(al) + sin(ga) * sin(be) * cos(al));
    Tinc[8] = (float)(-sin(be));
    Tinc[9] = (float)(cos(be) * sin(al));
    Tinc[10] = (float)(cos(be) * cos(al));
    Tinc[3] = (float)x[3]; Tinc[7] = (float)x[4]; Tinc[11] = (float)x[5];
    Tinc[12] = Tinc[13] = Tinc[14] = 0.0f; Tinc[15] = 1.0f;
    transform_points_f(Tinc, cur, n);                         /* input_transformed = T_inc * input_transformed */
    mat4f_mul(Tinc, final, final);                            /* final = T_inc * final */
    iters++;
    /* DefaultConvergenceCriteria::hasConverged */
    if (iters >= max_iter) { converged = 1; break; }
    if (!fixed_iters) {
      double cos_angle = 0.5 * ((double)Tinc[0] + (double)Tinc[5] + (double)Tinc[10] - 1.0);
      double tsq = (double)Tinc[3] * Tinc[3] + (double)Tinc[7] * Tinc[7] + (double)Tinc[11] * Tinc[11];
      if (cos_angle >= rot_thr && tsq <= trans_thr) { converged = 1; break; }
      if (fabs(mse - prev_mse) < 1e-12) { converged = 1; break; }
      prev_mse = mse;
    }
  }
  for (int i = 0; i < 16; i++) Tout[i] = (double)final[i];    /* getFinalTransformation().cast<double>(), :312 */
  if (iters_out) *iters_out = iters;
  if (fitness_out) *fitness_out = mse;
  free(cur); free(match); free(md2);
  return converged;
}

/* Pair loop of CCorresApp::Registration (CorresApp.cpp:220): omp parallel for num_threads(8) schedule(dynamic)
 * over pairs, every pair running the single-threaded alignment above. */
void oracle_icp_align_batch(oracle_cloud *const *clouds, const int *tgt_idx, const int *src_idx, int npairs,
                            const double *Tinit, double max_dist, int max_iter, double eps, int fixed_iters,
                            double *Tout, int *iters, double *fitness, int threads) {
#pragma omp parallel for num_threads(threads) schedule(dynamic)
  for (int p = 0; p < npairs; p++) {
    int it = 0;
    double fit = 0;
    oracle_icp_align(clouds[tgt_idx[p]], clouds[src_idx[p]], Tinit + 16 * (size_t)p, max_dist, max_iter, eps, fixed_iters,
                     Tout + 16 * (size_t)p, &it, &fit, 1);
    if (iters) iters[p] = it;
    if (fitness) fitness[p] = fit;
  }
}

/* ------------------------------------------------------------------ b4: CorresApp.cpp:145-161, 186-208
 * pairs = (target index, source index) in ascending source order; info36 = sum of A^T A (row-major) with
 * A = [I | 2[s]x-like] on the UNTRANSFORMED source points.  Returns the number of correspondences. */
int oracle_icp_correspondences(const oracle_cloud *tgt, const oracle_cloud *src, const double *T, double dist,
                               double normal_cos, int32_t *pairs, int cap, double *info36, int threads) {
  const int n = src->n;
  float *xyz = (float *)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
  float *nrm = (float *)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
  int *match = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  transform_with_normals(src, T, xyz, nrm);
  const double d2max = dist * dist;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int k = 0; k < n; k++) {
    float d2;
    int m = nn_search(tgt, xyz + 3 * k, &d2);
    match[k] = -1;
    if (m >= 0 && (double)d2 < d2max) {
      const float *a = tgt->nrm + 3 * m, *b = nrm + 3 * k;
      float dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];          /* NormalDot, CorresApp.h:53-55 (float expr) */
      if ((double)dot > normal_cos) match[k] = m;
    }
  }
  int cnt = 0;
  if (info36) memset(info36, 0, sizeof(double) * 36);
  for (int k = 0; k < n; k++) {
    if (match[k] < 0) continue;
    if (cnt < cap && pairs) { pairs[2 * cnt] = match[k]; pairs[2 * cnt + 1] = k; }
    cnt++;
    if (info36) {
      const float sx = src->xyz[3 * k], sy = src->xyz[3 * k + 1], sz = src->xyz[3 * k + 2];
      double A[3][6] = {{1, 0, 0, 0, 2 * sz, -2 * sy}, {0, 1, 0, -2 * sz, 0, 2 * sx}, {0, 0, 1, 2 * sy, -2 * sx, 0}};
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++)
          info36[r * 6 + c] += (A[0][r] * A[0][c] + A[1][r] * A[1][c]) + A[2][r] * A[2][c];
    }
  }
  free(xyz); free(nrm); free(match);
  return cnt;
}
