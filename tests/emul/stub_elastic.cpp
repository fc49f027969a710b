// stub_elastic.cpp -- TEST ONLY: a CPU stand-in for libelastic_b200 that implements the part of the C ABI the two
// executables call (include/elastic_b200.h) on top of the ORACLE (oracle/tsdf_oracle.c, oracle/icp_oracle.c).
// tests/test_apps_host_logic.py links apps/Integrate.cpp and apps/BuildCorrespondence.cpp against it to exercise
// their HOST logic without a GPU -- argument handling, frame gating, batching, ControlGrid mode, --devices sharding
// and merging, file formats -- and compares their output files with the verbatim reference binary's.  It says nothing
// about the CUDA path (tests/test_apps_gpu.py does that on a B200) and is never built into the product.
//
// "Devices" are just integers here.  A partitioned volume (er_tsdf_set_partition) integrates everything and emits only
// the blocks it owns at extraction time -- equivalent, since a voxel update reads nothing but depth and pose.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "elastic_b200.h"

extern "C" {
// oracle/tsdf_oracle.c
typedef struct oracle_tsdf oracle_tsdf;
oracle_tsdf* oracle_tsdf_create(int cols, int rows);
void oracle_tsdf_destroy(oracle_tsdf* o);
void oracle_tsdf_set_camera(oracle_tsdf* o, float fx, float fy, float cx, float cy, float integration_trunc);
void oracle_tsdf_set_voxel(oracle_tsdf* o, double unit_length, double tsdf_trunc);
int oracle_tsdf_integrate(oracle_tsdf* o, const uint16_t* depth, const double* T);
size_t oracle_tsdf_extract(const oracle_tsdf* o, float* xyzi, size_t cap);
void oracle_reproject(const oracle_tsdf* o, uint16_t* depth, const double* Ti, const double* T0, const double* A0, const double* Ai,
                      const float* ctr, int res, float length);
// oracle/icp_oracle.c
typedef struct oracle_cloud oracle_cloud;
oracle_cloud* oracle_cloud_create(const float* xyz, const float* nrm, int n, double cell);
void oracle_cloud_free(oracle_cloud* c);
int oracle_cloud_size(const oracle_cloud* c);
int oracle_icp_count_inliers(const oracle_cloud* tgt, const oracle_cloud* src, const double* T, double dist, int threads);
void oracle_icp_align_batch(oracle_cloud* const* clouds, const int* tgt_idx, const int* src_idx, int npairs, const double* Tinit,
                            double max_dist, int max_iter, double eps, int fixed_iters, double* Tout, int* iters, double* fitness,
                            int threads);
int oracle_icp_correspondences(const oracle_cloud* tgt, const oracle_cloud* src, const double* T, double dist, double normal_cos,
                               int32_t* pairs, int cap, double* info36, int threads);
}

struct er_tsdf {
  oracle_tsdf* o = nullptr;
  int cols = 0, rows = 0, device = 0;
  std::vector<float> ctr;
  int ctr_num = 0, ctr_res = 0;
  float ctr_length = 0;
  int rank = 0, world = 1;
};
struct er_cloud {
  oracle_cloud* c = nullptr;
  int device = 0;
};

extern "C" {

er_tsdf* er_tsdf_create(int device, int cols, int rows, double voxel, int block_res, double trunc, size_t) {
  if (block_res != 0 && block_res != 64) return nullptr;
  er_tsdf* v = new er_tsdf();
  v->o = oracle_tsdf_create(cols, rows);
  v->cols = cols; v->rows = rows; v->device = device;
  if (voxel > 0 || trunc > 0) oracle_tsdf_set_voxel(v->o, voxel > 0 ? voxel : 3.0 / 512.0, trunc > 0 ? trunc : 0.03);
  return v;
}
void er_tsdf_destroy(er_tsdf* v) { if (v) { oracle_tsdf_destroy(v->o); delete v; } }
int er_tsdf_set_camera(er_tsdf* v, float fx, float fy, float cx, float cy, float integration_trunc) {
  oracle_tsdf_set_camera(v->o, fx, fy, cx, cy, integration_trunc);
  return 0;
}
int er_tsdf_set_fusion(er_tsdf*, int f) { return (f >= 1 && f <= 32) ? 0 : -1; }
int er_tsdf_set_control_grids(er_tsdf* v, const float* ctr_xyz, int num, int res, float length) {
  const size_t per = (size_t)(res + 1) * (res + 1) * (res + 1) * 3;
  v->ctr.assign(ctr_xyz, ctr_xyz + per * num);
  v->ctr_num = num; v->ctr_res = res; v->ctr_length = length;
  return 0;
}
int er_tsdf_set_partition(er_tsdf* v, int rank, int world) {
  if (world < 1 || rank < 0 || rank >= world) return -1;
  v->rank = rank; v->world = world;
  return 0;
}
int er_tsdf_integrate_batch(er_tsdf* v, const uint16_t* depth, const double* T, int nframes) {
  const size_t npix = (size_t)v->cols * v->rows;
  for (int f = 0; f < nframes; f++) oracle_tsdf_integrate(v->o, depth + npix * f, T + 16 * (size_t)f);
  return 0;
}
int er_tsdf_integrate_ctr_batch(er_tsdf* v, const uint16_t* depth, const double* Ti, const double T0[16], const double A0[16],
                                const double* Ai, const int* chunk, int nframes) {
  const size_t npix = (size_t)v->cols * v->rows, per = (size_t)(v->ctr_res + 1) * (v->ctr_res + 1) * (v->ctr_res + 1) * 3;
  std::vector<uint16_t> buf(npix);
  for (int f = 0; f < nframes; f++) {
    if (chunk[f] < 0 || chunk[f] >= v->ctr_num) return -1;
    std::memcpy(buf.data(), depth + npix * f, sizeof(uint16_t) * npix);
    oracle_reproject(v->o, buf.data(), Ti + 16 * (size_t)f, T0, A0, Ai + 16 * (size_t)f, v->ctr.data() + per * chunk[f], v->ctr_res,
                     v->ctr_length);
    oracle_tsdf_integrate(v->o, buf.data(), Ti + 16 * (size_t)f);
  }
  return 0;
}
int er_tsdf_sync(er_tsdf*) { return 0; }
int er_tsdf_extract(er_tsdf* v, float* xyzi, size_t cap, size_t* n) {
  const size_t all = oracle_tsdf_extract(v->o, nullptr, 0);
  std::vector<float> tmp(4 * (all ? all : 1));
  oracle_tsdf_extract(v->o, tmp.data(), all);
  size_t m = 0;
  for (size_t i = 0; i < all; i++) {
    // voxel index -> block coordinates (TSDFVolume.cpp:118-123 writes i + (xi - 256) * 64)
    int b[3];
    for (int a = 0; a < 3; a++) {
      const int g = (int)tmp[4 * i + a] + 256 * 64;
      b[a] = g / 64;
    }
    if (v->world > 1 && ((b[0] + b[1] + b[2]) % v->world) != v->rank) continue;
    if (xyzi && m < cap) std::memcpy(xyzi + 4 * m, &tmp[4 * i], 4 * sizeof(float));
    m++;
  }
  *n = m;
  return 0;
}

int er_tsdf_save(er_tsdf*, const char*) { return -1; }     // (checkpoints are exercised on the GPU only)
int er_tsdf_load(er_tsdf*, const char*) { return -1; }
void* er_host_alloc(size_t bytes) { return std::malloc(bytes ? bytes : 1); }
void er_host_free(void* p) { std::free(p); }

er_cloud* er_cloud_upload(int device, const float* xyz, const float* normals, size_t n, double cell_size) {
  er_cloud* c = new er_cloud();
  c->c = oracle_cloud_create(xyz, normals, (int)n, cell_size);
  c->device = device;
  return c;
}
void er_cloud_free(er_cloud* c) { if (c) { oracle_cloud_free(c->c); delete c; } }
size_t er_cloud_size(const er_cloud* c) { return c ? (size_t)oracle_cloud_size(c->c) : 0; }
int er_icp_count_inliers(er_cloud* tgt, er_cloud* src, const double T[16], double dist, int* cnt) {
  *cnt = oracle_icp_count_inliers(tgt->c, src->c, T, dist, 1);
  return 0;
}
int er_icp_count_inliers_batch(er_cloud* const* clouds, int nclouds, const int* tgt_idx, const int* src_idx, int npairs, const double* T,
                               double dist, int* cnt) {
  for (int p = 0; p < npairs; p++) {
    if (tgt_idx[p] < 0 || tgt_idx[p] >= nclouds || src_idx[p] < 0 || src_idx[p] >= nclouds) return -2;
    if (clouds[tgt_idx[p]]->device != clouds[0]->device || clouds[src_idx[p]]->device != clouds[0]->device) return -1;
    cnt[p] = oracle_icp_count_inliers(clouds[tgt_idx[p]]->c, clouds[src_idx[p]]->c, T + 16 * (size_t)p, dist, 1);
  }
  return 0;
}
int er_icp_correspondences_batch(er_cloud* const* clouds, int nclouds, const int* tgt_idx, const int* src_idx, int npairs,
                                 const double* T, double dist, double normal_cos, int32_t* pairs, size_t cap, size_t* row_start,
                                 double* info36) {
  size_t at = 0;
  row_start[0] = 0;
  for (int p = 0; p < npairs; p++) {
    if (tgt_idx[p] < 0 || tgt_idx[p] >= nclouds || src_idx[p] < 0 || src_idx[p] >= nclouds) return -2;
    if (clouds[tgt_idx[p]]->device != clouds[0]->device || clouds[src_idx[p]]->device != clouds[0]->device) return -1;
    const int n = oracle_icp_correspondences(clouds[tgt_idx[p]]->c, clouds[src_idx[p]]->c, T + 16 * (size_t)p, dist, normal_cos,
                                             pairs + 2 * at, (int)(cap - at), info36 ? info36 + 36 * (size_t)p : nullptr, 1);
    at += (size_t)n;
    row_start[p + 1] = at;
  }
  return 0;
}
int er_icp_release(int) { return 0; }
int er_icp_align_batch(er_cloud* const* clouds, int nclouds, const int* tgt_idx, const int* src_idx, int npairs, const double* Tinit,
                       double max_dist, int max_iter, double eps, int fixed_iters, double* Tout, int* iters, double* fitness) {
  std::vector<oracle_cloud*> oc((size_t)nclouds);
  for (int i = 0; i < nclouds; i++) oc[(size_t)i] = clouds[i]->c;
  for (int p = 0; p < npairs; p++)                      // one device per batch, like the real library
    if (clouds[tgt_idx[p]]->device != clouds[0]->device || clouds[src_idx[p]]->device != clouds[0]->device) return -1;
  oracle_icp_align_batch(oc.data(), tgt_idx, src_idx, npairs, Tinit, max_dist, max_iter, eps, fixed_iters, Tout, iters, fitness, 1);
  return 0;
}
int er_icp_correspondences(er_cloud* tgt, er_cloud* src, const double T[16], double dist, double normal_cos, int32_t* pairs, size_t cap,
                           size_t* n, double info36[36]) {
  *n = (size_t)oracle_icp_correspondences(tgt->c, src->c, T, dist, normal_cos, pairs, (int)cap, info36, 1);
  return 0;
}

}  // extern "C"
