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

}  // extern "C"
