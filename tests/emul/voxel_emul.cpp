// voxel_emul.cpp -- TEST ONLY: runs the integrate kernel's per-voxel arithmetic on the CPU.
//
// The arithmetic itself is the product's own source, elasticreconstruction_b200/csrc/er_voxel_math.h (div2_rn,
// voxel_pixel, voxel_finish, voxel_update, er_voxel_thresholds), compiled here by g++ against that header's
// emulation layer (-ffp-contract=off).  What this file adds is the loop around it, which restates the few lines
// of integrate_kernel (er_tsdf.cu) that feed those functions: block origin and voxel centre (double products
// narrowed to float, as I2F / gridv in TSDFVolume.h:66-68, TSDFVolume.cpp:75), tgv = float(T^-1) * gridv in Eigen
// order, and the squared camera distance.  tests/test_voxel_math.py compares the result, bit for bit, with the
// oracle's IntegrateVolumeUnit on the same block, frame and scaled depth.
#include <cstddef>
#include <vector>

#include "er_voxel_math.h"

int er_voxel_rcp_bias = 0;

namespace {
constexpr int kRes = 64, kGridHalf = 256, kGridDim = 512;      // er_tsdf.cu:43-46
struct VoxelConsts {
  float fx, fy, cx, cy;
  int cols, rows;
  float trunc_eps, thr_neg, thr_one;
  double trunc_d;
};
}  // namespace

extern "C" {

void emul_set_rcp_bias(int ulps) { er_voxel_rcp_bias = ulps; }

static int g_tight = 0;
void emul_set_tight_classification(int on) { g_tight = on; }

static int g_pixel_form = 1;
void emul_set_pixel_form(int form) { g_pixel_form = form; }     // 1 = voxel_pixel, 2 = voxel_pixel_v2

// the two forms on one projected voxel (for randomised edge-case sweeps)
int emul_pixel(int form, float fx, float fy, float cx, float cy, int cols, int rows, float t0, float t1, float t2) {
  VoxelConsts v{};
  v.fx = fx; v.fy = fy; v.cx = cx; v.cy = cy; v.cols = cols; v.rows = rows;
  return form == 2 ? voxel_pixel_v2(v, t0, t1, t2) : voxel_pixel(v, t0, t1, t2);
}

void emul_thresholds(double trunc, float* out3) { er_voxel_thresholds(trunc, &out3[0], &out3[1], &out3[2]); }

// One IntegrateVolumeUnit call: frame (scaled depth, T camera->world, Tinv = its inverse, both row-major double[16])
// applied to block `key`; sdf / weight are the block's 64^3 arrays in the reference layout l = (i*64 + j)*64 + k.
// Returns the number of voxels updated.
long long emul_integrate_block(const float* scaled, int cols, int rows, float fx, float fy, float cx, float cy, double ul,
                               double trunc, const double* T, const double* Tinv, int key, float* sdf, float* weight) {
  VoxelConsts v;
  v.fx = fx; v.fy = fy; v.cx = cx; v.cy = cy; v.cols = cols; v.rows = rows; v.trunc_d = trunc;
  er_voxel_thresholds(trunc, &v.thr_neg, &v.thr_one, &v.trunc_eps);
  float tinv[12];
  for (int i = 0; i < 12; i++) tinv[i] = (float)Tinv[i];                        // fill_poses, er_tsdf.cu
  const float tx = (float)T[3], ty = (float)T[7], tz = (float)T[11];
  const int xi = key / (kGridDim * kGridDim), yi = (key / kGridDim) % kGridDim, zi = key % kGridDim;
  const float xs = (float)((double)((xi - kGridHalf) * kRes) * ul);
  const float ys = (float)((double)((yi - kGridHalf) * kRes) * ul);
  const float zs = (float)((double)((zi - kGridHalf) * kRes) * ul);
  long long updated = 0;
#pragma omp parallel for reduction(+ : updated)
  for (int i = 0; i < kRes; i++)
    for (int j = 0; j < kRes; j++)
      for (int k = 0; k < kRes; k++) {
        const float gx = (float)((double)i * ul + (double)xs);
        const float gy = (float)((double)j * ul + (double)ys);
        const float gz = (float)((double)k * ul + (double)zs);
        const float s0 = __fadd_rn(__fmul_rn(tinv[0], gx), __fmul_rn(tinv[1], gy));
        const float s1 = __fadd_rn(__fmul_rn(tinv[4], gx), __fmul_rn(tinv[5], gy));
        const float s2 = __fadd_rn(__fmul_rn(tinv[8], gx), __fmul_rn(tinv[9], gy));
        const float rx = __fsub_rn(gx, tx), ry = __fsub_rn(gy, ty), rz = __fsub_rn(gz, tz);
        const float rxy2 = __fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry));
        const float t0 = __fadd_rn(__fadd_rn(s0, __fmul_rn(tinv[2], gz)), tinv[3]);
        const float t1 = __fadd_rn(__fadd_rn(s1, __fmul_rn(tinv[6], gz)), tinv[7]);
        const float t2 = __fadd_rn(__fadd_rn(s2, __fmul_rn(tinv[10], gz)), tinv[11]);
        const int p = g_pixel_form == 2 ? voxel_pixel_v2(v, t0, t1, t2) : voxel_pixel(v, t0, t1, t2);
        const float dp = (p >= 0) ? scaled[p] : 0.0f;
        const float sq = __fadd_rn(rxy2, __fmul_rn(rz, rz));
        const size_t l = ((size_t)i * kRes + j) * kRes + k;
        if (voxel_finish(v, dp, sq, sdf[l], weight[l])) updated++;
      }
  return updated;
}

// ---------------------------------------------------------------------------------------------------------------
// Property check of the conservative classification (classify_ball, er_voxel_math.h) on one block and one frame:
// every 4^3 sub-brick is classified the way classify_kernel does it (8^3 brick ball first, then the eight sub-brick
// balls) and then evaluated voxel by voxel with the exact path on fresh voxels.
//   SKIP must mean: no voxel is updated;   ONE must mean: every voxel is updated with tsdf == 1.
// The depth pyramids are rebuilt here from `scaled` with the semantics of touch_kernel: (min, max) over aligned
// 16 / 8 / 4 pixel tiles, pixels outside the image counting as 0.
// out[0..2] = sub-bricks classified SKIP / ONE / FULL, out[3] = violations, out[4] = FULL sub-bricks in which no
// voxel was updated, out[5] = FULL sub-bricks in which every voxel got tsdf == 1 (the slack of the classification).
// ---------------------------------------------------------------------------------------------------------------
struct ClassifyConsts : VoxelConsts {
  int ntx;
  const er_tile_t *tiles, *tiles8, *tiles4;
};

static void build_level(const float* scaled, int cols, int rows, int ntx, int nty, int edge, std::vector<er_tile_t>& out) {
  const int per = kErTile / edge, w = ntx * per, h = nty * per;
  out.assign((size_t)w * h * kErMaxFusion, er_tile_t{0.0f, 0.0f});
  for (int ty = 0; ty < h; ty++)
    for (int tx = 0; tx < w; tx++) {
      float mn = 1e30f, mx = 0.0f;
      for (int y = ty * edge; y < (ty + 1) * edge; y++)
        for (int x = tx * edge; x < (tx + 1) * edge; x++) {
          const float s = (x < cols && y < rows) ? scaled[(size_t)y * cols + x] : 0.0f;
          mn = fminf(mn, s);
          mx = fmaxf(mx, s);
        }
      out[((size_t)ty * w + tx) * kErMaxFusion] = er_tile_t{mn, mx};             // frame slot 0
    }
}

// ---- exploration only (not product code): what would a classifier gain from exact footprint statistics?
// mode 1: min/max of `scaled` over the exact pixel rectangle of classify_ball's footprint (no tile alignment slack)
// mode 2: as 1, with the footprint taken from the projections of the 8 corners of the voxel-centre box (+1.5 px)
// mode 3: as 2, plus the TIGHT distance bound
// mode 4: footprint from the closed-form bound |u(c + o) - u(c)| <= fx * half * sum_i |a_i zc - pcx b_i| / (zc (zc - hz)),
//         hz = half * sum_i |b_i| (a, b = rows 0 and 2 of the rotation), + 1.5 px, plus the TIGHT distance bound
// mode 5: mode 4 on the tile pyramids (4 / 8 / 16 px levels as classify_ball) instead of exact rectangle statistics
static int g_explore = 0;
void emul_set_explore(int mode) { g_explore = mode; }

// scale_depth_div1000 against the division for every float in [lo, hi): number of mismatches
long long emul_check_div1000(float lo, float hi) {
  unsigned a, b;
  std::memcpy(&a, &lo, 4);
  std::memcpy(&b, &hi, 4);
  long long bad = 0;
  for (unsigned u = a; u < b; u++) {
    float x;
    std::memcpy(&x, &u, 4);
    if (scale_depth_div1000(x) != x / 1000.0f) bad++;
  }
  return bad;
}

static int classify_explore(const ClassifyConsts& v, const float* scaled, const float* P, float wx, float wy, float wz, float R,
                            float half) {
  const float eps = 1e-3f, trunc = (float)v.trunc_d;
  const float pcx = P[0] * wx + P[1] * wy + P[2] * wz + P[3];
  const float pcy = P[4] * wx + P[5] * wy + P[6] * wz + P[7];
  const float zc = P[8] * wx + P[9] * wy + P[10] * wz + P[11];
  if (!(zc + R > 0.0f)) return 0;
  if (!(zc - R > 0.05f)) return 2;
  const float dxw = wx - P[12], dyw = wy - P[13], dzw = wz - P[14];
  const float dc = sqrtf(dxw * dxw + dyw * dyw + dzw * dzw);
  float u0, u1, v0, v1;
  if (g_explore >= 4) {
    const float hz = half * (fabsf(P[8]) + fabsf(P[9]) + fabsf(P[10]));
    const float den = 1.0f / (zc * (zc - hz));
    const float su = fabsf(P[0] * zc - pcx * P[8]) + fabsf(P[1] * zc - pcx * P[9]) + fabsf(P[2] * zc - pcx * P[10]);
    const float sv = fabsf(P[4] * zc - pcy * P[8]) + fabsf(P[5] * zc - pcy * P[9]) + fabsf(P[6] * zc - pcy * P[10]);
    const float ru = v.fx * half * su * den * 1.001f + 1.5f, rv = v.fy * half * sv * den * 1.001f + 1.5f;
    const float uc = v.fx * pcx / zc + v.cx, vc = v.fy * pcy / zc + v.cy;
    u0 = uc - ru; u1 = uc + ru; v0 = vc - rv; v1 = vc + rv;
  } else if (g_explore == 1) {
    const float inv = 1.0f / (zc - R), izc = 1.0f / zc;
    const float ru = v.fx * R * (1.0f + fabsf(pcx) * izc) * inv + 1.5f;
    const float rv = v.fy * R * (1.0f + fabsf(pcy) * izc) * inv + 1.5f;
    const float uc = v.fx * pcx * izc + v.cx, vc = v.fy * pcy * izc + v.cy;
    u0 = uc - ru; u1 = uc + ru; v0 = vc - rv; v1 = vc + rv;
  } else {
    u0 = v0 = 1e30f; u1 = v1 = -1e30f;
    for (int c = 0; c < 8; c++) {
      const float ox = (c & 1) ? half : -half, oy = (c & 2) ? half : -half, oz = (c & 4) ? half : -half;
      const float x = pcx + P[0] * ox + P[1] * oy + P[2] * oz, y = pcy + P[4] * ox + P[5] * oy + P[6] * oz;
      const float z = zc + P[8] * ox + P[9] * oy + P[10] * oz;
      const float u = v.fx * x / z + v.cx, w = v.fy * y / z + v.cy;
      u0 = fminf(u0, u); u1 = fmaxf(u1, u); v0 = fminf(v0, w); v1 = fmaxf(v1, w);
    }
    u0 -= 1.5f; u1 += 1.5f; v0 -= 1.5f; v1 += 1.5f;
  }
  if (u1 < -0.5f || u0 > (float)v.cols - 0.5f || v1 < -0.5f || v0 > (float)v.rows - 0.5f) return 0;
  const bool inside = u0 >= 0.0f && u1 <= (float)(v.cols - 1) && v0 >= 0.0f && v1 <= (float)(v.rows - 1);
  const int px0 = max(0, (int)floorf(u0)), px1 = min(v.cols - 1, (int)ceilf(u1));
  const int py0 = max(0, (int)floorf(v0)), py1 = min(v.rows - 1, (int)ceilf(v1));
  float tmin = 1e30f, tmax = 0.0f;
  if (g_explore == 5) {
    const float wpx = fmaxf(u1 - u0, v1 - v0);
    const int sh = wpx <= 20.0f ? 2 : (wpx <= 40.0f ? 3 : 4);
    const int tx0 = px0 >> sh, tx1 = px1 >> sh, ty0 = py0 >> sh, ty1 = py1 >> sh;
    if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > 64) return 2;
    const er_tile_t* lvl = (sh == 4 ? v.tiles : (sh == 3 ? v.tiles8 : v.tiles4));
    const int pitch = v.ntx * (kErTile >> sh);
    for (int ty = ty0; ty <= ty1; ty++)
      for (int tx = tx0; tx <= tx1; tx++) {
        const er_tile_t t = lvl[(size_t)(ty * pitch + tx) * kErMaxFusion];
        tmin = fminf(tmin, t.x); tmax = fmaxf(tmax, t.y);
      }
  } else {
    for (int y = py0; y <= py1; y++)
      for (int x = px0; x <= px1; x++) { const float sv = scaled[(size_t)y * v.cols + x]; tmin = fminf(tmin, sv); tmax = fmaxf(tmax, sv); }
  }
  float near = R, far = R;
  if (g_explore >= 3) {
    const float h = half * (fabsf(dxw) + fabsf(dyw) + fabsf(dzw)) / dc * 1.0001f + 1e-5f;
    near = fminf(R, h);
    far = fminf(R, h + R * R / (2.0f * (dc - R)));
  }
  if (dc - near > tmax + trunc + eps) return 0;
  if (inside && tmin > 0.002f && tmin - trunc - eps >= dc + far) return 1;
  return 2;
}

void emul_check_classification(const float* scaled, int cols, int rows, float fx, float fy, float cx, float cy, double ul_d,
                               double trunc, const double* T, const double* Tinv, int key, long long* out) {
  ClassifyConsts v;
  v.fx = fx; v.fy = fy; v.cx = cx; v.cy = cy; v.cols = cols; v.rows = rows; v.trunc_d = trunc;
  er_voxel_thresholds(trunc, &v.thr_neg, &v.thr_one, &v.trunc_eps);
  v.ntx = (cols + kErTile - 1) / kErTile;
  const int nty = (rows + kErTile - 1) / kErTile;
  std::vector<er_tile_t> l16, l8, l4;
  build_level(scaled, cols, rows, v.ntx, nty, 16, l16);
  build_level(scaled, cols, rows, v.ntx, nty, 8, l8);
  build_level(scaled, cols, rows, v.ntx, nty, 4, l4);
  v.tiles = l16.data(); v.tiles8 = l8.data(); v.tiles4 = l4.data();
  float P[15];                                                    // s_pose row of classify_kernel
  for (int i = 0; i < 12; i++) P[i] = (float)Tinv[i];
  P[12] = (float)T[3]; P[13] = (float)T[7]; P[14] = (float)T[11];
  const float ul = (float)ul_d;
  const float R8 = (float)(1.7320508 * 3.5 * ul_d) + 1e-4f;
  const float R4 = (float)(1.7320508 * 1.5 * ul_d) + 1e-4f;
  const int xi = key / (kGridDim * kGridDim), yi = (key / kGridDim) % kGridDim, zi = key % kGridDim;
  const float xs = (float)((double)((xi - kGridHalf) * kRes) * ul_d);
  const float ys = (float)((double)((yi - kGridHalf) * kRes) * ul_d);
  const float zs = (float)((double)((zi - kGridHalf) * kRes) * ul_d);
  long long n_cls[3] = {0, 0, 0}, bad = 0, full_idle = 0, full_one = 0;
#pragma omp parallel for reduction(+ : bad, full_idle, full_one) reduction(+ : n_cls[:3])
  for (int brick = 0; brick < 512; brick++) {
    const float bx = (float)((xi - kGridHalf) * kRes + (brick >> 6) * 8);
    const float by = (float)((yi - kGridHalf) * kRes + ((brick >> 3) & 7) * 8);
    const float bz = (float)((zi - kGridHalf) * kRes + (brick & 7) * 8);
    int cls[8];
    const int c8 = g_explore ? classify_explore(v, scaled, P, (bx + 3.5f) * ul, (by + 3.5f) * ul, (bz + 3.5f) * ul, R8, 3.5f * ul)
                   : g_tight ? classify_ball<true>(v, P, 0, (bx + 3.5f) * ul, (by + 3.5f) * ul, (bz + 3.5f) * ul, R8, 3.5f * ul)
                             : classify_ball(v, P, 0, (bx + 3.5f) * ul, (by + 3.5f) * ul, (bz + 3.5f) * ul, R8);
    for (int q = 0; q < 8; q++) {
      const float qx = (bx + 1.5f + 4.0f * ((q >> 2) & 1)) * ul, qy = (by + 1.5f + 4.0f * ((q >> 1) & 1)) * ul,
                  qz = (bz + 1.5f + 4.0f * (q & 1)) * ul;
      cls[q] = (c8 != 2) ? c8
               : g_explore ? classify_explore(v, scaled, P, qx, qy, qz, R4, 1.5f * ul)
               : (g_tight ? classify_ball<true>(v, P, 0, qx, qy, qz, R4, 1.5f * ul) : classify_ball(v, P, 0, qx, qy, qz, R4));
    }
    for (int q = 0; q < 8; q++) {
      int n_upd = 0, n_one = 0;
      for (int l = 0; l < 64; l++) {
        const int i = (brick >> 6) * 8 + ((q >> 2) & 1) * 4 + (l >> 4);
        const int j = ((brick >> 3) & 7) * 8 + ((q >> 1) & 1) * 4 + ((l >> 2) & 3);
        const int k = (brick & 7) * 8 + (q & 1) * 4 + (l & 3);
        const float gx = (float)((double)i * ul_d + (double)xs);
        const float gy = (float)((double)j * ul_d + (double)ys);
        const float gz = (float)((double)k * ul_d + (double)zs);
        const float t0 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[0], gx), __fmul_rn(P[1], gy)), __fmul_rn(P[2], gz)), P[3]);
        const float t1 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[4], gx), __fmul_rn(P[5], gy)), __fmul_rn(P[6], gz)), P[7]);
        const float t2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P[8], gx), __fmul_rn(P[9], gy)), __fmul_rn(P[10], gz)), P[11]);
        const int p = voxel_pixel(v, t0, t1, t2);
        const float dp = (p >= 0) ? scaled[p] : 0.0f;
        const float rx = __fsub_rn(gx, P[12]), ry = __fsub_rn(gy, P[13]), rz = __fsub_rn(gz, P[14]);
        const float sq = __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
        float sdf = 0.0f, w = 0.0f;
        const int u = voxel_finish(v, dp, sq, sdf, w);
        n_upd += (u != 0);
        n_one += (u == 2);
      }
      n_cls[cls[q]]++;
      if (cls[q] == 0 && n_upd != 0) bad++;
      if (cls[q] == 1 && n_one != 64) bad++;
      if (cls[q] == 2 && n_upd == 0) full_idle++;
      if (cls[q] == 2 && n_one == 64) full_one++;
    }
  }
  out[0] = n_cls[0]; out[1] = n_cls[1]; out[2] = n_cls[2]; out[3] = bad; out[4] = full_idle; out[5] = full_one;
}

// ---------------------------------------------------------------------------------------------------------------
// Property check of the touch kernel's block keys on one frame (estimate_block / touch_pixel_index, er_voxel_math.h,
// and the tile shortcut of touch_kernel restated from er_tsdf.cu): against the reference's double sequence
// (UVD2XYZ TSDFVolume.h:40-49, TSDFVolume.cpp:49-53),
//   * every valid pixel's block from touch_pixel_index must be the reference's block;
//   * whenever the 8 frustum corners of a 16x16 tile (corner pixels x {min, max} raw depth) agree on one block
//     without raising `risky`, every valid pixel of the tile must lie in that block.
// out[0] = valid pixels, out[1] = pixels that took the exact fallback, out[2] = tiles with a valid pixel,
// out[3] = tiles the shortcut declared uniform, out[4] = violations.
// ---------------------------------------------------------------------------------------------------------------
struct TouchConsts {
  float fx, fy, cx, cy;
  double ul, inv_ul, inv_fx, inv_fy;
};

static void reference_block(const TouchConsts& v, const double* T, int u, int vr, unsigned d, int blk[3]) {
  const double z = d / 1000.0;
  const double x = ((float)u - v.cx) * z / v.fx;
  const double y = ((float)vr - v.cy) * z / v.fy;
  for (int r = 0; r < 3; r++) {
    const double w = ((T[r * 4 + 0] * x + T[r * 4 + 1] * y) + T[r * 4 + 2] * z) + T[r * 4 + 3];
    blk[r] = ((int)floor(w / v.ul + 0.5) + kGridHalf * kRes) / kRes;
  }
}

void emul_check_touch(const unsigned short* depth, int cols, int rows, float fx, float fy, float cx, float cy, double ul,
                      const double* T, long long* out) {
  TouchConsts v;
  v.fx = fx; v.fy = fy; v.cx = cx; v.cy = cy; v.ul = ul;
  v.inv_ul = 1.0 / ul; v.inv_fx = 1.0 / (double)fx; v.inv_fy = 1.0 / (double)fy;      // er_tsdf_set_camera / create
  float Tf[12];
  for (int i = 0; i < 12; i++) Tf[i] = (float)T[i];
  long long n_valid = 0, n_exact = 0, n_tiles = 0, n_uniform = 0, bad = 0;
  const int ntx = (cols + kErTile - 1) / kErTile, nty = (rows + kErTile - 1) / kErTile;
#pragma omp parallel for reduction(+ : n_valid, n_exact, n_tiles, n_uniform, bad)
  for (int t = 0; t < ntx * nty; t++) {
    const int tx0 = (t % ntx) * kErTile, ty0 = (t / ntx) * kErTile;
    unsigned lo = 0xFFFFFFFFu, hi = 0u;
    for (int y = ty0; y < ty0 + kErTile && y < rows; y++)
      for (int x = tx0; x < tx0 + kErTile && x < cols; x++) {
        const unsigned d = depth[(size_t)y * cols + x];
        if (d) { lo = d < lo ? d : lo; hi = d > hi ? d : hi; }
      }
    if (hi == 0) continue;
    n_tiles++;
    // tile shortcut: lanes 0..7 of warp 0 in touch_kernel
    int first[3] = {0, 0, 0};
    bool uniform = true;
    for (int c = 0; c < 8; c++) {
      const int cu = (c & 1) ? min(tx0 + kErTile - 1, cols - 1) : tx0;
      const int cv = (c & 2) ? min(ty0 + kErTile - 1, rows - 1) : ty0;
      const float cd = (c & 4) ? (float)hi : (float)lo;
      int bi[3];
      bool risky = false;
      estimate_block(v, Tf, __fsub_rn((float)cu, v.cx), __fsub_rn((float)cv, v.cy), cd, bi, risky);
      if (c == 0) { first[0] = bi[0]; first[1] = bi[1]; first[2] = bi[2]; }
      const bool same = (bi[0] == first[0]) & (bi[1] == first[1]) & (bi[2] == first[2]);
      const bool okc = same && !risky && bi[0] >= 0 && bi[0] < kGridDim && bi[1] >= 0 && bi[1] < kGridDim && bi[2] >= 0 && bi[2] < kGridDim;
      uniform = uniform && okc;
    }
    if (uniform) n_uniform++;
    for (int y = ty0; y < ty0 + kErTile && y < rows; y++)
      for (int x = tx0; x < tx0 + kErTile && x < cols; x++) {
        const unsigned d = depth[(size_t)y * cols + x];
        if (!d) continue;
        n_valid++;
        int ref[3];
        reference_block(v, T, x, y, d, ref);
        const bool ref_in = ref[0] >= 0 && ref[0] < kGridDim && ref[1] >= 0 && ref[1] < kGridDim && ref[2] >= 0 && ref[2] < kGridDim;
        // per-pixel path
        int idx[3];
        bool risky = false;
        {
          int tmp[3];
          estimate_block(v, Tf, __fsub_rn((float)x, v.cx), __fsub_rn((float)y, v.cy), (float)d, tmp, risky);
          if (risky) n_exact++;
        }
        touch_pixel_index(v, Tf, T, __fsub_rn((float)x, v.cx), __fsub_rn((float)y, v.cy), d, idx);
        const int lim = kGridDim * kRes;
        const bool in = idx[0] >= 0 && idx[0] < lim && idx[1] >= 0 && idx[1] < lim && idx[2] >= 0 && idx[2] < lim;
        if (in != ref_in) bad++;
        else if (in && (idx[0] / kRes != ref[0] || idx[1] / kRes != ref[1] || idx[2] / kRes != ref[2])) bad++;
        if (uniform && ref_in && (ref[0] != first[0] || ref[1] != first[1] || ref[2] != first[2])) bad++;
        if (uniform && !ref_in) bad++;
      }
  }
  out[0] = n_valid; out[1] = n_exact; out[2] = n_tiles; out[3] = n_uniform; out[4] = bad;
}

// ---------------------------------------------------------------------------------------------------------------
// CIntegrateApp::Reproject for a whole image through reproject_pixel (er_voxel_math.h): z-buffer minimum over the
// pixels of `depth`, result in `out` (0 where nothing landed), as reproject_kernel + zbuf_to_u16_kernel do.
// adj = Ti^-1 * T0 * A0^-1 is passed in (the product computes it with er_mat4_inverse / er_mat4_mul on the host).
// ---------------------------------------------------------------------------------------------------------------
void emul_reproject(const unsigned short* depth, int cols, int rows, float fx, float fy, float cx, float cy, const double* Ai,
                    const double* adj, const float* ctr, int res, float length, unsigned short* out) {
  struct { float fx, fy, cx, cy; int cols, rows; } v{fx, fy, cx, cy, cols, rows};
  ReprojParams R;
  for (int i = 0; i < 12; i++) { R.Ai[i] = Ai[i]; R.adj[i] = adj[i]; }
  R.res = res;
  R.unit_length = length / (float)res;                           // ControlGrid.cpp:19, er_tsdf.cu reproject_on_device
  std::vector<unsigned> zbuf((size_t)cols * rows, 0xFFFFFFFFu);
  for (int pix = 0; pix < cols * rows; pix++) {
    const unsigned d = depth[pix];
    if (d == 0) continue;
    unsigned dd;
    const int target = reproject_pixel(v, R, ctr, pix % cols, pix / cols, d, dd);
    if (target >= 0 && dd < zbuf[target]) zbuf[target] = dd;
  }
  for (int i = 0; i < cols * rows; i++) out[i] = (zbuf[i] == 0xFFFFFFFFu) ? (unsigned short)0 : (unsigned short)zbuf[i];
}

// exploration helper: classes (0 SKIP, 1 ONE, 2 FULL) of the 4096 sub-bricks of one block for one frame, as
// classify_kernel assigns them (index = brick * 8 + q); tight = 1 uses classify_ball<true>
void emul_classify_block(const float* scaled, int cols, int rows, float fx, float fy, float cx, float cy, double ul_d, double trunc,
                         const double* T, const double* Tinv, int key, int tight, unsigned char* cls_out) {
  ClassifyConsts v;
  v.fx = fx; v.fy = fy; v.cx = cx; v.cy = cy; v.cols = cols; v.rows = rows; v.trunc_d = trunc;
  er_voxel_thresholds(trunc, &v.thr_neg, &v.thr_one, &v.trunc_eps);
  v.ntx = (cols + kErTile - 1) / kErTile;
  const int nty = (rows + kErTile - 1) / kErTile;
  static thread_local const float* cached = nullptr;
  static thread_local std::vector<er_tile_t> l16, l8, l4;
  if (cached != scaled) {
    build_level(scaled, cols, rows, v.ntx, nty, 16, l16);
    build_level(scaled, cols, rows, v.ntx, nty, 8, l8);
    build_level(scaled, cols, rows, v.ntx, nty, 4, l4);
    cached = scaled;
  }
  v.tiles = l16.data(); v.tiles8 = l8.data(); v.tiles4 = l4.data();
  float P[15];
  for (int i = 0; i < 12; i++) P[i] = (float)Tinv[i];
  P[12] = (float)T[3]; P[13] = (float)T[7]; P[14] = (float)T[11];
  const float ul = (float)ul_d;
  const float R8 = (float)(1.7320508 * 3.5 * ul_d) + 1e-4f, R4 = (float)(1.7320508 * 1.5 * ul_d) + 1e-4f;
  const int xi = key / (kGridDim * kGridDim), yi = (key / kGridDim) % kGridDim, zi = key % kGridDim;
  for (int brick = 0; brick < 512; brick++) {
    const float bx = (float)((xi - kGridHalf) * kRes + (brick >> 6) * 8);
    const float by = (float)((yi - kGridHalf) * kRes + ((brick >> 3) & 7) * 8);
    const float bz = (float)((zi - kGridHalf) * kRes + (brick & 7) * 8);
    const int c8 = tight ? classify_ball<true>(v, P, 0, (bx + 3.5f) * ul, (by + 3.5f) * ul, (bz + 3.5f) * ul, R8, 3.5f * ul)
                         : classify_ball(v, P, 0, (bx + 3.5f) * ul, (by + 3.5f) * ul, (bz + 3.5f) * ul, R8);
    for (int q = 0; q < 8; q++) {
      const float qx = (bx + 1.5f + 4.0f * ((q >> 2) & 1)) * ul, qy = (by + 1.5f + 4.0f * ((q >> 1) & 1)) * ul,
                  qz = (bz + 1.5f + 4.0f * (q & 1)) * ul;
      cls_out[brick * 8 + q] = (unsigned char)((c8 != 2) ? c8
                               : (tight ? classify_ball<true>(v, P, 0, qx, qy, qz, R4, 1.5f * ul) : classify_ball(v, P, 0, qx, qy, qz, R4)));
    }
  }
}

}  // extern "C"
