// er_fragopt.cu -- device side of the FragmentOptimizer path: per-point lattice coordinates and
// their updates (class PointCloud, FragmentOptimizer/PointCloud.h) and the assembly of the SLAC / rigid normal equations
// over all correspondences (the pair loops of COptApp::OptimizeSLAC / OptimizeRigid, FragmentOptimizer/OptApp.cpp:312-385,
// 473-570).  All arithmetic lives in er_fragopt_math.h, which the CPU emulation compiles too; this file is the data
// layout, the launches and the C ABI (include/elastic_b200.h).
//
// Layout in HBM
//   points of all fragments, concatenated (cloud_off[l] = first point of fragment l):
//     statA int4   (base, r0, r1, r2)          lattice cell + trilinear residuals, fixed after loading
//     statB float4 (n0x, n0y, n0z, fragment)   the loaded normal / cell edge, fixed
//     pos   float4 (x, y, z, -)                current position     (Point::p_)
//     nrm   float4 (x, y, z, -)                current normal       (Point::n_)
//   correspondences of all pairs as int4 (global point index of pi, of pj, pair, -), sorted by the 64-bit key
//     (pair, base of pi, base of pj): every run of equal keys shares the 60 column indices of its Jacobian rows.
//   SLAC: JJ dense row-major M x M doubles (upper triangle used; M = 6 num + nper = 2487 for 50 fragments, 49 MB:
//     resident in the 126 MB L2, which is where the flush atomics land), Jb[M], score.
//
// slac_assemble_kernel: persistent 128-thread CTAs, each owning a contiguous range of the sorted correspondences.
//   Per batch of 128: every thread forms one row (gathers of two points, ~150 flops) into shared memory (64 doubles per
//   row: 60 entries, b, padding); then the CTA walks the batch, thread t < 120 adding the 4x4 tile (ty, tx) of
//   row^T row to 16 double registers, threads 120..127 adding b*row (and b*b).  The sums are added to JJ / Jb with
//   atomics only when the key changes.  Bound: FP64 FMA issue (1830 + 60 multiply-adds per correspondence against
//   ~100 bytes of gathers).
#include <cub/device/device_radix_sort.cuh>
#include <cusolverDn.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdint>
#include <mutex>
#include <vector>

#include "er_common.h"
#include "er_fragopt_math.h"

using namespace er_fo;

namespace {

constexpr int kThreads = 128;
// rows staged per batch (<= kThreads; 64 doubles of shared memory each) and resident CTAs per SM the kernel is compiled for
#ifndef ER_FO_BATCH
#define ER_FO_BATCH 96
#endif
#ifndef ER_FO_MINB
#define ER_FO_MINB 4
#endif
constexpr int kBatch = ER_FO_BATCH;
static_assert(kBatch <= kThreads && kBatch % 32 == 0, "ER_FO_BATCH");

struct FoView {
  Lattice L;
  int num, M;
  const int4* statA;
  const float4* statB;
  float4* pos;
  float4* nrm;
  const int* pair_i;
  const int* pair_j;
  const int4* corr;
  long long total;
};

struct DevAdd {
  __device__ __forceinline__ void operator()(double* p, double v) const { atomicAdd(p, v); }
};

__device__ __forceinline__ void load_point(const FoView& v, int g, int* base, float r[3], float p[3]) {
  const int4 a = __ldg(&v.statA[g]);
  const float4 q = v.pos[g];
  *base = a.x;
  r[0] = __int_as_float(a.y); r[1] = __int_as_float(a.z); r[2] = __int_as_float(a.w);
  p[0] = q.x; p[1] = q.y; p[2] = q.z;
}

__global__ void __launch_bounds__(kThreads, ER_FO_MINB)
slac_assemble_kernel(FoView v, const double* __restrict__ rot_t, double* JJ, double* Jb, double* score, unsigned long long* flushes) {
  extern __shared__ __align__(16) double rows[];   // [kBatch][kRowStride]
  __shared__ SlacKey keys[kBatch];
  __shared__ int brk[kBatch];
  const int tid = threadIdx.x;
  int ty, tx;
  fo_tile_of(tid < 120 ? tid : 0, &ty, &tx);
  const long long begin = v.total * (long long)blockIdx.x / gridDim.x;
  const long long end = v.total * (long long)(blockIdx.x + 1) / gridDim.x;
  double acc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.0;
  SlacKey cur;
  cur.pair = -1; cur.bi = -1; cur.bj = -1;
  for (long long base = begin; base < end; base += kBatch) {
    const int cnt = (int)min((long long)kBatch, end - base);
    if (tid < cnt) {
      const int4 c = __ldg(&v.corr[base + tid]);
      int bi, bj;
      float ri[3], rj[3], pi[3], pj[3];
      load_point(v, c.x, &bi, ri, pi);
      load_point(v, c.y, &bj, rj, pj);
      const float4 nn = v.nrm[c.x];
      const float ni[3] = {nn.x, nn.y, nn.z};
      const int i = __ldg(&v.pair_i[c.z]), j = __ldg(&v.pair_j[c.z]);
      fo_slac_row(pi, ni, ri, pj, rj, rot_t + 9 * i, rot_t + 9 * j, &rows[tid * kRowStride], tid & 7);
      SlacKey k;
      k.pair = c.z; k.bi = bi; k.bj = bj;
      keys[tid] = k;
    }
    __syncthreads();
    if (tid < cnt) brk[tid] = !(keys[tid] == (tid == 0 ? cur : keys[tid - 1]));
    __syncthreads();
    for (int c = 0; c < cnt; ++c) {
      if (brk[c]) {                                   // uniform over the CTA
        if (cur.pair >= 0) {
          fo_slac_flush(v.L, v.num, v.M, v.pair_i[cur.pair], v.pair_j[cur.pair], cur, tid, ty, tx, acc, JJ, Jb, score, DevAdd());
          if (tid == 0) atomicAdd(flushes, 1ull);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = 0.0;
        cur = keys[c];
      }
      fo_slac_mac(tid, ty, tx, &rows[c * kRowStride], c & 7, acc);
    }
    __syncthreads();
  }
  if (cur.pair >= 0) {
    fo_slac_flush(v.L, v.num, v.M, v.pair_i[cur.pair], v.pair_j[cur.pair], cur, tid, ty, tx, acc, JJ, Jb, score, DevAdd());
    if (tid == 0) atomicAdd(flushes, 1ull);
  }
}

// Nonrigid (OptApp.cpp:159-206): the same walk over the sorted list with 48-entry rows (24 unknowns of fragment i's
// lattice, 24 of fragment j's) and no right-hand side; 96-thread CTAs, 78 threads own the 4x4 tiles of the 12 x 12-tile upper
// triangle, rows are staged 96 at a time (36 KB of shared memory).  A = dense row-major N x N, N = num * nper.
__global__ void __launch_bounds__(kNrThreads, 5)
nonrigid_assemble_kernel(FoView v, double weight, long long N, double* A, unsigned long long* flushes) {
  __shared__ __align__(16) double rows[kNrThreads * kNrStride];
  __shared__ SlacKey keys[kNrThreads];
  __shared__ int brk[kNrThreads];
  const int tid = threadIdx.x;
  int ty, tx;
  fo_nr_tile_of(tid < kNrTiles ? tid : 0, &ty, &tx);
  const long long begin = v.total * (long long)blockIdx.x / gridDim.x;
  const long long end = v.total * (long long)(blockIdx.x + 1) / gridDim.x;
  double acc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.0;
  SlacKey cur;
  cur.pair = -1; cur.bi = -1; cur.bj = -1;
  for (long long base = begin; base < end; base += kNrThreads) {
    const int cnt = (int)min((long long)kNrThreads, end - base);
    if (tid < cnt) {
      const int4 c = __ldg(&v.corr[base + tid]);
      const int4 a = __ldg(&v.statA[c.x]), b = __ldg(&v.statA[c.y]);
      const float4 nn = v.nrm[c.x];
      const float ri[3] = {__int_as_float(a.y), __int_as_float(a.z), __int_as_float(a.w)};
      const float rj[3] = {__int_as_float(b.y), __int_as_float(b.z), __int_as_float(b.w)};
      const float ni[3] = {nn.x, nn.y, nn.z};
      fo_nonrigid_row(ri, ni, rj, weight, &rows[tid * kNrStride], tid & 7);
      SlacKey k;
      k.pair = c.z; k.bi = a.x; k.bj = b.x;
      keys[tid] = k;
    }
    __syncthreads();
    if (tid < cnt) brk[tid] = !(keys[tid] == (tid == 0 ? cur : keys[tid - 1]));
    __syncthreads();
    for (int c = 0; c < cnt; ++c) {
      if (brk[c]) {
        if (cur.pair >= 0) {
          fo_nr_flush(v.L, N, v.pair_i[cur.pair], v.pair_j[cur.pair], cur, tid, ty, tx, acc, A, DevAdd());
          if (tid == 0) atomicAdd(flushes, 1ull);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = 0.0;
        cur = keys[c];
      }
      if (tid < kNrTiles) fo_tile_mac<kNrStride>(ty, tx, &rows[c * kNrStride], c & 7, acc);
    }
    __syncthreads();
  }
  if (cur.pair >= 0) {
    fo_nr_flush(v.L, N, v.pair_i[cur.pair], v.pair_j[cur.pair], cur, tid, ty, tx, acc, A, DevAdd());
    if (tid == 0) atomicAdd(flushes, 1ull);
  }
}

__global__ void add_triplets_kernel(size_t nnz, const int* __restrict__ r, const int* __restrict__ c, const double* __restrict__ val, long long N,
                                    double* A) {
  const size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k < nnz) atomicAdd(&A[(long long)r[k] * N + c[k]], val[k]);
}

// Rigid: 12-entry rows; a CTA owns a range of one pair's correspondences, every thread sums its rows' products
// (78 upper entries of val^T val, 12 of b*val, b*b) and the CTA adds them once.  out = [npairs][91].
constexpr int kRigidSums = 91;
__global__ void __launch_bounds__(kThreads)
rigid_assemble_kernel(FoView v, const long long* __restrict__ pair_start, int chunks, double* out) {
  __shared__ double sums[kRigidSums];
  const int pair = blockIdx.x;
  const long long b0 = pair_start[pair], b1 = pair_start[pair + 1];
  const long long n = b1 - b0;
  const long long begin = b0 + n * (long long)blockIdx.y / chunks, end = b0 + n * (long long)(blockIdx.y + 1) / chunks;
  if (threadIdx.x < kRigidSums) sums[threadIdx.x] = 0.0;
  __syncthreads();
  double acc[kRigidSums];
#pragma unroll
  for (int e = 0; e < kRigidSums; ++e) acc[e] = 0.0;
  bool any = false;
  for (long long k = begin + threadIdx.x; k < end; k += kThreads) {
    const int4 c = __ldg(&v.corr[k]);
    const float4 pi = v.pos[c.x], ni = v.nrm[c.x], pj = v.pos[c.y];
    const float fpi[3] = {pi.x, pi.y, pi.z}, fni[3] = {ni.x, ni.y, ni.z}, fpj[3] = {pj.x, pj.y, pj.z};
    double val[12], b;
    fo_rigid_row(fpi, fni, fpj, val, &b);
    int e = 0;
#pragma unroll
    for (int r = 0; r < 12; ++r)
#pragma unroll
      for (int s = r; s < 12; ++s) acc[e++] += val[r] * val[s];
#pragma unroll
    for (int r = 0; r < 12; ++r) acc[78 + r] += b * val[r];
    acc[90] += b * b;
    any = true;
  }
  if (any) {
#pragma unroll
    for (int e = 0; e < kRigidSums; ++e) atomicAdd(&sums[e], acc[e]);
  }
  __syncthreads();
  if (threadIdx.x < kRigidSums && sums[threadIdx.x] != 0.0) atomicAdd(&out[(size_t)pair * kRigidSums + threadIdx.x], sums[threadIdx.x]);
}

__global__ void update_pose_kernel(FoView v, int npoints, const float* __restrict__ poses) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= npoints) return;
  const int l = __float_as_int(__ldg(&v.statB[g]).w);
  float4 p4 = v.pos[g], n4 = v.nrm[g];
  float p[3] = {p4.x, p4.y, p4.z}, n[3] = {n4.x, n4.y, n4.z};
  float m[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) m[e] = __ldg(&poses[16 * l + e]);
  fo_update_pose(m, p, n);
  v.pos[g] = make_float4(p[0], p[1], p[2], 0.0f);
  v.nrm[g] = make_float4(n[0], n[1], n[2], 0.0f);
}

// mode 0: normals only (UpdateAllNormal); mode 1: normals and positions (UpdateAllPointPN)
__global__ void update_lattice_kernel(FoView v, int npoints, const double* __restrict__ ctr, int mode) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= npoints) return;
  const int4 a = __ldg(&v.statA[g]);
  const float4 b = __ldg(&v.statB[g]);
  StaticPoint s;
  s.base = a.x; s.valid = 1;
  s.r[0] = __int_as_float(a.y); s.r[1] = __int_as_float(a.z); s.r[2] = __int_as_float(a.w);
  s.n0[0] = b.x; s.n0[1] = b.y; s.n0[2] = b.z;
  const int l = __float_as_int(b.w);
  const double* c = ctr + (size_t)l * v.L.nper;
  float n[3];
  fo_update_normal(v.L, s, c, n);
  v.nrm[g] = make_float4(n[0], n[1], n[2], 0.0f);
  if (mode == 1) {
    double pos[3];
    fo_update_point(v.L, s, c, pos);
    v.pos[g] = make_float4((float)pos[0], (float)pos[1], (float)pos[2], 0.0f);
  }
}

// xyzn rows of one fragment; with ctr: the warped position (UpdatePoint) instead of the stored one
__global__ void get_points_kernel(FoView v, int first, int n, const double* __restrict__ ctr, float* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int g = first + t;
  float4 p = v.pos[g];
  const float4 nn = v.nrm[g];
  if (ctr) {
    const int4 a = __ldg(&v.statA[g]);
    const float4 b = __ldg(&v.statB[g]);
    StaticPoint s;
    s.base = a.x; s.valid = 1;
    s.r[0] = __int_as_float(a.y); s.r[1] = __int_as_float(a.z); s.r[2] = __int_as_float(a.w);
    s.n0[0] = b.x; s.n0[1] = b.y; s.n0[2] = b.z;
    double pos[3];
    fo_update_point(v.L, s, ctr + (size_t)__float_as_int(b.w) * v.L.nper, pos);
    p.x = (float)pos[0]; p.y = (float)pos[1]; p.z = (float)pos[2];
  }
  out[6 * (size_t)t + 0] = p.x; out[6 * (size_t)t + 1] = p.y; out[6 * (size_t)t + 2] = p.z;
  out[6 * (size_t)t + 3] = nn.x; out[6 * (size_t)t + 4] = nn.y; out[6 * (size_t)t + 5] = nn.z;
}

__global__ void corr_keys_kernel(const int4* __restrict__ statA, const int4* __restrict__ corr, long long n, int base_bits,
                                 unsigned long long* keys, unsigned int* vals) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int4 c = corr[k];
  const unsigned long long bi = (unsigned)statA[c.x].x, bj = (unsigned)statA[c.y].x;
  keys[k] = ((unsigned long long)(unsigned)c.z << (2 * base_bits)) | (bi << base_bits) | bj;
  vals[k] = (unsigned int)k;
}

__global__ void corr_gather_kernel(const int4* __restrict__ in, const unsigned int* __restrict__ order, long long n, int4* out) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k < n) out[k] = in[order[k]];
}

__global__ void add_matrix_kernel(double* __restrict__ dst, const double* __restrict__ src, size_t n) {
  const size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k < n) dst[k] += src[k];
}

// A = base + data on the stored triangle (row-major upper = column-major lower for cuSOLVER)
__global__ void add_upper_kernel(const double* __restrict__ base, const double* __restrict__ data, double* A, int M) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c < M && c >= r) A[(size_t)r * M + c] = base[(size_t)r * M + c] + data[(size_t)r * M + c];
}

}  // namespace

// cuSOLVER (dense Cholesky of the SLAC system) is bound at the first er_fragopt_slac_solve, not at load time: the TSDF and
// ICP paths, and every executable's start-up, do not pay for mapping cuSOLVER + cuBLAS.
namespace {
struct Solver {
  void* dl = nullptr;
  decltype(&cusolverDnCreate) create = nullptr;
  decltype(&cusolverDnDestroy) destroy = nullptr;
  decltype(&cusolverDnSetStream) set_stream = nullptr;
  decltype(&cusolverDnDpotrf_bufferSize) potrf_buffer = nullptr;
  decltype(&cusolverDnDpotrf) potrf = nullptr;
  decltype(&cusolverDnDpotrs) potrs = nullptr;
  bool ok = false;
};
Solver& solver_api() {
  static Solver s;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libcusolver.so.11", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so"}) {
      s.dl = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (s.dl) break;
    }
    if (s.dl) {
      s.create = (decltype(s.create))dlsym(s.dl, "cusolverDnCreate");
      s.destroy = (decltype(s.destroy))dlsym(s.dl, "cusolverDnDestroy");
      s.set_stream = (decltype(s.set_stream))dlsym(s.dl, "cusolverDnSetStream");
      s.potrf_buffer = (decltype(s.potrf_buffer))dlsym(s.dl, "cusolverDnDpotrf_bufferSize");
      s.potrf = (decltype(s.potrf))dlsym(s.dl, "cusolverDnDpotrf");
      s.potrs = (decltype(s.potrs))dlsym(s.dl, "cusolverDnDpotrs");
      s.ok = s.create && s.destroy && s.set_stream && s.potrf_buffer && s.potrf && s.potrs;
    }
  });
  return s;
}
}  // namespace

struct er_fragopt {
  int device = 0;
  cudaStream_t stream = nullptr;
  int num = 0, res = 0;
  double length = 0;
  Lattice L;
  int M = 0;
  int sms = 148;
  std::vector<std::vector<int4>> h_statA;
  std::vector<std::vector<float4>> h_statB, h_pos, h_nrm;
  std::vector<int> cloud_off;
  bool uploaded = false;
  int npoints = 0;
  int4* d_statA = nullptr;
  float4 *d_statB = nullptr, *d_pos = nullptr, *d_nrm = nullptr;
  int npairs = 0;
  int *d_pair_i = nullptr, *d_pair_j = nullptr;
  long long* d_pair_start = nullptr;
  int4* d_corr = nullptr;
  long long total = 0;
  double *d_rot = nullptr, *d_JJ = nullptr, *d_Jb = nullptr, *d_score = nullptr, *d_base = nullptr, *d_A = nullptr, *d_rhs = nullptr;
  double* d_ctr = nullptr;
  float* d_poses = nullptr;
  double* d_rigid = nullptr;
  double* d_AA = nullptr;            // nonrigid: dense N x N
  double* d_peer = nullptr;          // er_fragopt_slac_accumulate: the other device's matrix
  double* d_nr_rhs = nullptr;
  long long nr_N = 0;
  bool nr_factored = false;
  unsigned long long* d_flushes = nullptr;
  bool have_base = false;
  cusolverDnHandle_t solver = nullptr;
  double* d_work = nullptr;
  int work_len = 0;
  int* d_info = nullptr;
  cudaEvent_t ev[8] = {};
  cudaEvent_t k0 = nullptr, k1 = nullptr;
  uint64_t launches = 0;
  uint64_t last_rows = 0;
};

namespace {
FoView view_of(const er_fragopt* h) {
  FoView v;
  v.L = h->L; v.num = h->num; v.M = h->M;
  v.statA = h->d_statA; v.statB = h->d_statB; v.pos = h->d_pos; v.nrm = h->d_nrm;
  v.pair_i = h->d_pair_i; v.pair_j = h->d_pair_j; v.corr = h->d_corr; v.total = h->total;
  return v;
}

int upload_points(er_fragopt* h) {
  if (h->uploaded) return ER_OK;
  h->cloud_off.assign(h->num + 1, 0);
  for (int l = 0; l < h->num; ++l) h->cloud_off[l + 1] = h->cloud_off[l] + (int)h->h_statA[l].size();
  h->npoints = h->cloud_off[h->num];
  const size_t n = (size_t)std::max(h->npoints, 1);
  ER_CUDA(cudaMalloc(&h->d_statA, n * sizeof(int4)));
  ER_CUDA(cudaMalloc(&h->d_statB, n * sizeof(float4)));
  ER_CUDA(cudaMalloc(&h->d_pos, n * sizeof(float4)));
  ER_CUDA(cudaMalloc(&h->d_nrm, n * sizeof(float4)));
  for (int l = 0; l < h->num; ++l) {
    const size_t m = h->h_statA[l].size(), o = (size_t)h->cloud_off[l];
    if (!m) continue;
    ER_CUDA(cudaMemcpyAsync(h->d_statA + o, h->h_statA[l].data(), m * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    ER_CUDA(cudaMemcpyAsync(h->d_statB + o, h->h_statB[l].data(), m * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
    ER_CUDA(cudaMemcpyAsync(h->d_pos + o, h->h_pos[l].data(), m * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
    ER_CUDA(cudaMemcpyAsync(h->d_nrm + o, h->h_nrm[l].data(), m * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
  }
  ER_CUDA(cudaStreamSynchronize(h->stream));
  for (int l = 0; l < h->num; ++l) {      // the host copies are no longer needed
    std::vector<int4>().swap(h->h_statA[l]);
    std::vector<float4>().swap(h->h_statB[l]);
    std::vector<float4>().swap(h->h_pos[l]);
    std::vector<float4>().swap(h->h_nrm[l]);
  }
  h->uploaded = true;
  return ER_OK;
}
}  // namespace

extern "C" {

er_fragopt* er_fragopt_create(int device, int num, int resolution, double length) {
  if (num <= 0 || resolution <= 0 || resolution > 64 || !(length > 0)) { er::set_error("er_fragopt_create: bad arguments"); return nullptr; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    er::set_error("er_fragopt_create: CUDA device %d not available (%d devices); this library has no CPU path", device, ndev);
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) { er::set_error("er_fragopt_create: cudaSetDevice(%d) failed", device); return nullptr; }
  er_fragopt* h = new er_fragopt();
  h->device = device; h->num = num; h->res = resolution; h->length = length;
  h->L = fo_lattice(resolution, (float)length);
  h->M = 6 * num + h->L.nper;
  h->h_statA.resize(num); h->h_statB.resize(num); h->h_pos.resize(num); h->h_nrm.resize(num);
  cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device);
  bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int e = 0; e < 8 && ok; ++e) ok = cudaEventCreate(&h->ev[e]) == cudaSuccess;
  ok = ok && cudaEventCreate(&h->k0) == cudaSuccess && cudaEventCreate(&h->k1) == cudaSuccess;
  ok = ok && cudaMalloc(&h->d_rot, sizeof(double) * 9 * num) == cudaSuccess;
  ok = ok && cudaMalloc(&h->d_poses, sizeof(float) * 16 * num) == cudaSuccess;
  ok = ok && cudaMalloc(&h->d_ctr, sizeof(double) * (size_t)num * h->L.nper) == cudaSuccess;
  ok = ok && cudaMalloc(&h->d_Jb, sizeof(double) * h->M) == cudaSuccess;
  ok = ok && cudaMalloc(&h->d_rhs, sizeof(double) * h->M) == cudaSuccess;
  ok = ok && cudaMalloc(&h->d_score, sizeof(double)) == cudaSuccess;
  ok = ok && cudaMalloc(&h->d_flushes, sizeof(unsigned long long)) == cudaSuccess;
  ok = ok && cudaMalloc(&h->d_info, sizeof(int)) == cudaSuccess;
  if (ok) ok = cudaFuncSetAttribute(slac_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(kBatch * kRowStride * sizeof(double))) == cudaSuccess;
  if (!ok) {
    er::set_error("er_fragopt_create: %s", cudaGetErrorString(cudaGetLastError()));
    er_fragopt_destroy(h);
    return nullptr;
  }
  return h;
}

void er_fragopt_destroy(er_fragopt* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  void* ptrs[] = {h->d_statA, h->d_statB, h->d_pos, h->d_nrm, h->d_pair_i, h->d_pair_j, h->d_pair_start, h->d_corr, h->d_rot,
                  h->d_JJ, h->d_Jb, h->d_score, h->d_base, h->d_A, h->d_rhs, h->d_ctr, h->d_poses, h->d_rigid, h->d_flushes,
                  h->d_work, h->d_info, h->d_AA, h->d_nr_rhs, h->d_peer};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (h->solver) solver_api().destroy(h->solver);
  for (int e = 0; e < 8; ++e) if (h->ev[e]) cudaEventDestroy(h->ev[e]);
  if (h->k0) cudaEventDestroy(h->k0);
  if (h->k1) cudaEventDestroy(h->k1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int er_fragopt_set_cloud(er_fragopt* h, int l, const float* xyzn, size_t n, size_t* n_kept) {
  ER_ARG(h && l >= 0 && l < h->num && (n == 0 || xyzn), "er_fragopt_set_cloud: bad arguments");
  ER_ARG(!h->uploaded, "er_fragopt_set_cloud: clouds are fixed once er_fragopt_set_pairs has been called");
  ER_ARG(n < (size_t)1 << 30, "er_fragopt_set_cloud: too many points");
  std::vector<int4>& A = h->h_statA[l];
  std::vector<float4>& B = h->h_statB[l];
  std::vector<float4>& P = h->h_pos[l];
  std::vector<float4>& N = h->h_nrm[l];
  A.resize(n); B.resize(n); P.resize(n); N.resize(n);
  long long first_bad = (long long)n;
#pragma omp parallel for schedule(static) reduction(min : first_bad)
  for (long long k = 0; k < (long long)n; ++k) {
    StaticPoint s;
    float p[3], nn[3];
    const bool ok = fo_get_coordinate(h->L, xyzn + 6 * k, &s, p, nn);
    int4 a;
    a.x = s.base;
    memcpy(&a.y, &s.r[0], 4); memcpy(&a.z, &s.r[1], 4); memcpy(&a.w, &s.r[2], 4);
    A[k] = a;
    float lw;
    memcpy(&lw, &l, 4);
    B[k] = make_float4(s.n0[0], s.n0[1], s.n0[2], lw);
    P[k] = make_float4(p[0], p[1], p[2], 0.0f);
    N[k] = make_float4(nn[0], nn[1], nn[2], 0.0f);
    if (!ok && k < first_bad) first_bad = k;
  }
  size_t kept = n;
  if (first_bad < (long long)n) {        // PointCloud.cpp:37-40: the failing point stays as a zero-filled record, loading stops
    kept = (size_t)first_bad + 1;
    A.resize(kept); B.resize(kept); P.resize(kept); N.resize(kept);
  }
  if (n_kept) *n_kept = kept;
  return ER_OK;
}

int er_fragopt_set_pairs(er_fragopt* h, int npairs, const int* idx0, const int* idx1, const size_t* row_start, const int32_t* rows) {
  ER_ARG(h && npairs >= 0 && (npairs == 0 || (idx0 && idx1 && row_start)), "er_fragopt_set_pairs: bad arguments");
  ER_CUDA(cudaSetDevice(h->device));
  int rc = upload_points(h);
  if (rc != ER_OK) return rc;
  const long long total = npairs ? (long long)row_start[npairs] : 0;
  ER_ARG(total < (1ll << 31) - 1024, "er_fragopt_set_pairs: more than 2^31 correspondences");
  ER_ARG(total == 0 || rows, "er_fragopt_set_pairs: rows missing");
  ER_ARG(npairs < (1 << 22), "er_fragopt_set_pairs: too many pairs");
  std::vector<long long> ps(npairs + 1, 0);
  for (int p = 0; p < npairs; ++p) {
    ER_ARG(idx0[p] >= 0 && idx0[p] < h->num && idx1[p] >= 0 && idx1[p] < h->num, "er_fragopt_set_pairs: fragment index out of range");
    ER_ARG(row_start[p + 1] >= row_start[p], "er_fragopt_set_pairs: row_start must ascend");
    ps[p] = (long long)row_start[p];
  }
  ps[npairs] = total;
  // global point indices; a row that points outside its cloud is an error (the reference would read out of bounds)
  std::vector<int4> hc((size_t)std::max(total, 1ll));
  int bad = 0;
  for (int p = 0; p < npairs; ++p) {
    const int o0 = h->cloud_off[idx0[p]], n0 = h->cloud_off[idx0[p] + 1] - o0;
    const int o1 = h->cloud_off[idx1[p]], n1 = h->cloud_off[idx1[p] + 1] - o1;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (long long k = ps[p]; k < ps[p + 1]; ++k) {
      const int a = rows[2 * k], b = rows[2 * k + 1];
      if (a < 0 || a >= n0 || b < 0 || b >= n1) { ++bad; hc[k] = make_int4(o0, o1, p, 0); continue; }
      hc[k] = make_int4(o0 + a, o1 + b, p, 0);
    }
  }
  if (bad) { er::set_error("er_fragopt_set_pairs: %d correspondence rows point outside their clouds", bad); return ER_ERR_ARG; }
  for (void* p : {(void*)h->d_pair_i, (void*)h->d_pair_j, (void*)h->d_pair_start, (void*)h->d_corr, (void*)h->d_rigid}) if (p) cudaFree(p);
  h->d_pair_i = h->d_pair_j = nullptr; h->d_pair_start = nullptr; h->d_corr = nullptr; h->d_rigid = nullptr;
  h->npairs = npairs; h->total = total;
  const size_t np = (size_t)std::max(npairs, 1);
  ER_CUDA(cudaMalloc(&h->d_pair_i, np * sizeof(int)));
  ER_CUDA(cudaMalloc(&h->d_pair_j, np * sizeof(int)));
  ER_CUDA(cudaMalloc(&h->d_pair_start, (np + 1) * sizeof(long long)));
  ER_CUDA(cudaMalloc(&h->d_rigid, np * kRigidSums * sizeof(double)));
  ER_CUDA(cudaMalloc(&h->d_corr, hc.size() * sizeof(int4)));
  if (npairs) {
    ER_CUDA(cudaMemcpyAsync(h->d_pair_i, idx0, npairs * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    ER_CUDA(cudaMemcpyAsync(h->d_pair_j, idx1, npairs * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  }
  ER_CUDA(cudaMemcpyAsync(h->d_pair_start, ps.data(), (npairs + 1) * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
  if (total) {
    // sort by (pair, base of pi, base of pj): LSD radix sort, stable, so equal keys keep their file order
    int4* d_in = nullptr;
    unsigned long long *d_k0 = nullptr, *d_k1 = nullptr;
    unsigned int *d_v0 = nullptr, *d_v1 = nullptr;
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    int base_bits = 1;
    while ((1 << base_bits) < h->L.nper / 3) ++base_bits;
    int pair_bits = 1;
    while ((1ll << pair_bits) < npairs) ++pair_bits;
    const int end_bit = 2 * base_bits + pair_bits;
    cudaError_t e = cudaMalloc(&d_in, total * sizeof(int4));
    if (e == cudaSuccess) e = cudaMalloc(&d_k0, total * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_k1, total * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_v0, total * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_v1, total * 4);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_k0, d_k1, d_v0, d_v1, (int)total, 0, end_bit, h->stream);
    if (e == cudaSuccess) e = cudaMalloc(&d_tmp, std::max(tmp_bytes, (size_t)16));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, hc.data(), total * sizeof(int4), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
      const unsigned blocks = (unsigned)((total + 255) / 256);
      corr_keys_kernel<<<blocks, 256, 0, h->stream>>>(h->d_statA, d_in, total, base_bits, d_k0, d_v0);
      e = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_k0, d_k1, d_v0, d_v1, (int)total, 0, end_bit, h->stream);
      if (e == cudaSuccess) {
        corr_gather_kernel<<<blocks, 256, 0, h->stream>>>(d_in, d_v1, total, h->d_corr);
        h->launches += 3;
        e = cudaStreamSynchronize(h->stream);
      }
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    for (void* p : {(void*)d_in, (void*)d_k0, (void*)d_k1, (void*)d_v0, (void*)d_v1, d_tmp}) if (p) cudaFree(p);
    if (e != cudaSuccess) { er::set_error("er_fragopt_set_pairs: %s", cudaGetErrorString(e)); return ER_ERR_CUDA; }
  } else {
    ER_CUDA(cudaStreamSynchronize(h->stream));
  }
  return ER_OK;
}

int er_fragopt_update_pose(er_fragopt* h, const double* T) {
  ER_ARG(h && T, "er_fragopt_update_pose: bad arguments");
  ER_CUDA(cudaSetDevice(h->device));
  int rc = upload_points(h);
  if (rc != ER_OK) return rc;
  std::vector<float> m((size_t)16 * h->num);
  for (size_t e = 0; e < m.size(); ++e) m[e] = (float)T[e];               // pose.cast<float>()
  ER_CUDA(cudaMemcpyAsync(h->d_poses, m.data(), m.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  if (h->npoints) {
    update_pose_kernel<<<(h->npoints + 255) / 256, 256, 0, h->stream>>>(view_of(h), h->npoints, h->d_poses);
    ++h->launches;
  }
  ER_CUDA(cudaStreamSynchronize(h->stream));      // m is a local
  ER_CUDA(cudaGetLastError());
  return ER_OK;
}

static int update_lattice(er_fragopt* h, const double* ctr, int mode) {
  ER_CUDA(cudaSetDevice(h->device));
  int rc = upload_points(h);
  if (rc != ER_OK) return rc;
  ER_CUDA(cudaMemcpyAsync(h->d_ctr, ctr, sizeof(double) * (size_t)h->num * h->L.nper, cudaMemcpyHostToDevice, h->stream));
  if (h->npoints) {
    update_lattice_kernel<<<(h->npoints + 255) / 256, 256, 0, h->stream>>>(view_of(h), h->npoints, h->d_ctr, mode);
    ++h->launches;
  }
  ER_CUDA(cudaStreamSynchronize(h->stream));
  ER_CUDA(cudaGetLastError());
  return ER_OK;
}

int er_fragopt_update_normals(er_fragopt* h, const double* ctr) {
  ER_ARG(h && ctr, "er_fragopt_update_normals: bad arguments");
  return update_lattice(h, ctr, 0);
}

int er_fragopt_update_points_pn(er_fragopt* h, const double* expand_ctr) {
  ER_ARG(h && expand_ctr, "er_fragopt_update_points_pn: bad arguments");
  return update_lattice(h, expand_ctr, 1);
}

int er_fragopt_get_points(er_fragopt* h, int l, float* xyzn, size_t cap_rows, const double* ctr) {
  ER_ARG(h && l >= 0 && l < h->num && xyzn, "er_fragopt_get_points: bad arguments");
  ER_CUDA(cudaSetDevice(h->device));
  int rc = upload_points(h);
  if (rc != ER_OK) return rc;
  const int first = h->cloud_off[l], n = h->cloud_off[l + 1] - first;
  ER_ARG((size_t)n <= cap_rows, "er_fragopt_get_points: buffer too small");
  if (!n) return ER_OK;
  float* d_out = nullptr;
  ER_CUDA(cudaMalloc(&d_out, sizeof(float) * 6 * (size_t)n));
  cudaError_t e = cudaSuccess;
  if (ctr) e = cudaMemcpyAsync(h->d_ctr, ctr, sizeof(double) * (size_t)h->num * h->L.nper, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    get_points_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(view_of(h), first, n, ctr ? h->d_ctr : nullptr, d_out);
    ++h->launches;
    e = cudaMemcpyAsync(xyzn, d_out, sizeof(float) * 6 * (size_t)n, cudaMemcpyDeviceToHost, h->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_out);
  if (e != cudaSuccess) { er::set_error("er_fragopt_get_points: %s", cudaGetErrorString(e)); return ER_ERR_CUDA; }
  return ER_OK;
}

int er_fragopt_slac_assemble(er_fragopt* h, const double* rot_t, double* Jb, double* score) {
  ER_ARG(h && rot_t && Jb && score, "er_fragopt_slac_assemble: bad arguments");
  ER_ARG(h->d_corr || h->total == 0, "er_fragopt_slac_assemble: call er_fragopt_set_pairs first");
  ER_CUDA(cudaSetDevice(h->device));
  const size_t mm = (size_t)h->M * h->M;
  if (!h->d_JJ) ER_CUDA(cudaMalloc(&h->d_JJ, mm * sizeof(double)));
  ER_CUDA(cudaMemcpyAsync(h->d_rot, rot_t, sizeof(double) * 9 * h->num, cudaMemcpyHostToDevice, h->stream));
  ER_CUDA(cudaMemsetAsync(h->d_JJ, 0, mm * sizeof(double), h->stream));
  ER_CUDA(cudaMemsetAsync(h->d_Jb, 0, sizeof(double) * h->M, h->stream));
  ER_CUDA(cudaMemsetAsync(h->d_score, 0, sizeof(double), h->stream));
  ER_CUDA(cudaMemsetAsync(h->d_flushes, 0, sizeof(unsigned long long), h->stream));
  ER_CUDA(cudaEventRecord(h->k0, h->stream));
  if (h->total) {
    const long long batches = (h->total + kBatch - 1) / kBatch;
    const int grid = (int)std::min<long long>(batches, (long long)ER_FO_MINB * h->sms);
    slac_assemble_kernel<<<grid, kThreads, kBatch * kRowStride * sizeof(double), h->stream>>>(view_of(h), h->d_rot, h->d_JJ, h->d_Jb,
                                                                                                  h->d_score, h->d_flushes);
    ++h->launches;
  }
  ER_CUDA(cudaEventRecord(h->k1, h->stream));
  h->last_rows = (uint64_t)h->total;
  ER_CUDA(cudaMemcpyAsync(Jb, h->d_Jb, sizeof(double) * h->M, cudaMemcpyDeviceToHost, h->stream));
  ER_CUDA(cudaMemcpyAsync(score, h->d_score, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  ER_CUDA(cudaStreamSynchronize(h->stream));
  ER_CUDA(cudaGetLastError());
  return ER_OK;
}

int er_fragopt_slac_accumulate(er_fragopt* dst, er_fragopt* src) {
  ER_ARG(dst && src && dst != src && dst->d_JJ && src->d_JJ && dst->M == src->M, "er_fragopt_slac_accumulate: both handles must have assembled the same system");
  const size_t n = (size_t)dst->M * dst->M;
  ER_CUDA(cudaSetDevice(dst->device));
  if (!dst->d_peer) ER_CUDA(cudaMalloc(&dst->d_peer, n * sizeof(double)));
  // src's er_fragopt_slac_assemble returned synchronised; cudaMemcpyPeer stages through the host where peer access is unavailable
  ER_CUDA(cudaMemcpyPeerAsync(dst->d_peer, dst->device, src->d_JJ, src->device, n * sizeof(double), dst->stream));
  add_matrix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, dst->stream>>>(dst->d_JJ, dst->d_peer, n);
  ++dst->launches;
  ER_CUDA(cudaStreamSynchronize(dst->stream));
  ER_CUDA(cudaGetLastError());
  return ER_OK;
}

int er_fragopt_slac_get_jj(er_fragopt* h, double* jj) {
  ER_ARG(h && jj && h->d_JJ, "er_fragopt_slac_get_jj: nothing assembled yet");
  ER_CUDA(cudaSetDevice(h->device));
  ER_CUDA(cudaMemcpyAsync(jj, h->d_JJ, sizeof(double) * (size_t)h->M * h->M, cudaMemcpyDeviceToHost, h->stream));
  ER_CUDA(cudaStreamSynchronize(h->stream));
  return ER_OK;
}

int er_fragopt_slac_solve(er_fragopt* h, const double* base, const double* rhs, double* x) {
  ER_ARG(h && rhs && x && h->d_JJ, "er_fragopt_slac_solve: assemble first");
  ER_ARG(base || h->have_base, "er_fragopt_slac_solve: no base matrix given yet");
  ER_CUDA(cudaSetDevice(h->device));
  const int M = h->M;
  const size_t mm = (size_t)M * M;
  if (!h->d_base) ER_CUDA(cudaMalloc(&h->d_base, mm * sizeof(double)));
  if (!h->d_A) ER_CUDA(cudaMalloc(&h->d_A, mm * sizeof(double)));
  if (base) {
    ER_CUDA(cudaMemcpyAsync(h->d_base, base, mm * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    h->have_base = true;
  }
#define ER_SOLVER(expr)                                                                          \
  do {                                                                                           \
    cusolverStatus_t s_ = (expr);                                                                \
    if (s_ != CUSOLVER_STATUS_SUCCESS) {                                                         \
      er::set_error("%s failed: cusolver status %d (%s:%d)", #expr, (int)s_, __FILE__, __LINE__); \
      return ER_ERR_CUDA;                                                                        \
    }                                                                                            \
  } while (0)
  Solver& S = solver_api();
  if (!S.ok) { er::set_error("er_fragopt_slac_solve: cannot load libcusolver"); return ER_ERR_CUDA; }
  if (!h->solver) {
    ER_SOLVER(S.create(&h->solver));
    ER_SOLVER(S.set_stream(h->solver, h->stream));
  }
  // row-major upper == column-major lower
  int need = 0;
  ER_SOLVER(S.potrf_buffer(h->solver, CUBLAS_FILL_MODE_LOWER, M, h->d_A, M, &need));
  if (need > h->work_len) {
    if (h->d_work) cudaFree(h->d_work);
    h->d_work = nullptr;
    ER_CUDA(cudaMalloc(&h->d_work, sizeof(double) * (size_t)need));
    h->work_len = need;
  }
  add_upper_kernel<<<dim3((M + 255) / 256, M), 256, 0, h->stream>>>(h->d_base, h->d_JJ, h->d_A, M);
  ++h->launches;
  ER_CUDA(cudaMemcpyAsync(h->d_rhs, rhs, sizeof(double) * M, cudaMemcpyHostToDevice, h->stream));
  ER_SOLVER(S.potrf(h->solver, CUBLAS_FILL_MODE_LOWER, M, h->d_A, M, h->d_work, h->work_len, h->d_info));
  int info = 0;
  ER_CUDA(cudaMemcpyAsync(&info, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  ER_CUDA(cudaStreamSynchronize(h->stream));
  if (info != 0) { er::set_error("er_fragopt_slac_solve: the matrix is not positive definite (potrf info = %d)", info); return ER_ERR_ARG; }
  ER_SOLVER(S.potrs(h->solver, CUBLAS_FILL_MODE_LOWER, M, 1, h->d_A, M, h->d_rhs, M, h->d_info));
  ER_CUDA(cudaMemcpyAsync(x, h->d_rhs, sizeof(double) * M, cudaMemcpyDeviceToHost, h->stream));
  ER_CUDA(cudaStreamSynchronize(h->stream));
  ER_CUDA(cudaGetLastError());
#undef ER_SOLVER
  return ER_OK;
}

int er_fragopt_rigid_assemble(er_fragopt* h, double* JJ, double* Jb, double* score) {
  ER_ARG(h && JJ && Jb && score, "er_fragopt_rigid_assemble: bad arguments");
  ER_ARG(h->d_corr || h->total == 0, "er_fragopt_rigid_assemble: call er_fragopt_set_pairs first");
  ER_CUDA(cudaSetDevice(h->device));
  const int n6 = 6 * h->num;
  std::fill(JJ, JJ + (size_t)n6 * n6, 0.0);
  std::fill(Jb, Jb + n6, 0.0);
  *score = 0.0;
  if (!h->npairs) return ER_OK;
  std::vector<double> sums((size_t)h->npairs * kRigidSums);
  ER_CUDA(cudaMemsetAsync(h->d_rigid, 0, sums.size() * sizeof(double), h->stream));
  if (h->total) {
    const int chunks = (int)std::max<long long>(1, std::min<long long>(64, (4ll * h->sms + h->npairs - 1) / h->npairs));
    rigid_assemble_kernel<<<dim3(h->npairs, chunks), kThreads, 0, h->stream>>>(view_of(h), h->d_pair_start, chunks, h->d_rigid);
    ++h->launches;
  }
  ER_CUDA(cudaMemcpyAsync(sums.data(), h->d_rigid, sums.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  std::vector<int> pi(h->npairs), pj(h->npairs);
  ER_CUDA(cudaMemcpyAsync(pi.data(), h->d_pair_i, h->npairs * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  ER_CUDA(cudaMemcpyAsync(pj.data(), h->d_pair_j, h->npairs * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  ER_CUDA(cudaStreamSynchronize(h->stream));
  ER_CUDA(cudaGetLastError());
  for (int p = 0; p < h->npairs; ++p) {      // pairs in list order: the sum order is fixed
    const double* s = &sums[(size_t)p * kRigidSums];
    int idx[12];
    for (int r = 0; r < 6; ++r) { idx[r] = pi[p] * 6 + r; idx[6 + r] = pj[p] * 6 + r; }
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
  return ER_OK;
}

// ------------------------------------------------------------------------------------------------ nonrigid
static int load_solver(er_fragopt* h, const char* who) {
  Solver& S = solver_api();
  if (!S.ok) { er::set_error("%s: cannot load libcusolver", who); return ER_ERR_CUDA; }
  if (!h->solver) {
    if (S.create(&h->solver) != CUSOLVER_STATUS_SUCCESS || S.set_stream(h->solver, h->stream) != CUSOLVER_STATUS_SUCCESS) {
      er::set_error("%s: cusolverDnCreate failed", who);
      return ER_ERR_CUDA;
    }
  }
  return ER_OK;
}

int er_fragopt_nonrigid_assemble(er_fragopt* h, double weight) {
  ER_ARG(h, "er_fragopt_nonrigid_assemble: null handle");
  ER_ARG(h->d_corr || h->total == 0, "er_fragopt_nonrigid_assemble: call er_fragopt_set_pairs first");
  ER_CUDA(cudaSetDevice(h->device));
  const long long N = (long long)h->num * h->L.nper;
  if (!h->d_AA) {
    size_t free_b = 0, total_b = 0;
    ER_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const double need = (double)N * (double)N * 8.0;
    if (need > 0.9 * (double)free_b) {
      er::set_error("er_fragopt_nonrigid_assemble: the dense %lld x %lld system needs %.1f GB, %.1f GB are free (use --slac for this many fragments)",
                    N, N, need / 1e9, free_b / 1e9);
      return ER_ERR_CAPACITY;
    }
    ER_CUDA(cudaMalloc(&h->d_AA, (size_t)N * (size_t)N * sizeof(double)));
    ER_CUDA(cudaMalloc(&h->d_nr_rhs, (size_t)N * sizeof(double)));
    h->nr_N = N;
  }
  ER_CUDA(cudaMemsetAsync(h->d_AA, 0, (size_t)N * (size_t)N * sizeof(double), h->stream));
  ER_CUDA(cudaMemsetAsync(h->d_flushes, 0, sizeof(unsigned long long), h->stream));
  ER_CUDA(cudaEventRecord(h->k0, h->stream));
  if (h->total) {
    const long long batches = (h->total + kNrThreads - 1) / kNrThreads;
    const int grid = (int)std::min<long long>(batches, 5ll * h->sms);
    nonrigid_assemble_kernel<<<grid, kNrThreads, 0, h->stream>>>(view_of(h), weight, N, h->d_AA, h->d_flushes);
    ++h->launches;
  }
  ER_CUDA(cudaEventRecord(h->k1, h->stream));
  h->last_rows = (uint64_t)h->total;
  h->nr_factored = false;
  ER_CUDA(cudaStreamSynchronize(h->stream));
  ER_CUDA(cudaGetLastError());
  return ER_OK;
}

int er_fragopt_nonrigid_get_aa(er_fragopt* h, double* aa) {
  ER_ARG(h && aa && h->d_AA && !h->nr_factored, "er_fragopt_nonrigid_get_aa: assemble first (and before factoring)");
  ER_CUDA(cudaSetDevice(h->device));
  ER_CUDA(cudaMemcpyAsync(aa, h->d_AA, (size_t)h->nr_N * (size_t)h->nr_N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  ER_CUDA(cudaStreamSynchronize(h->stream));
  return ER_OK;
}

int er_fragopt_nonrigid_factor(er_fragopt* h, size_t nnz, const int32_t* rows, const int32_t* cols, const double* vals) {
  ER_ARG(h && h->d_AA && !h->nr_factored, "er_fragopt_nonrigid_factor: assemble first");
  ER_ARG(nnz == 0 || (rows && cols && vals), "er_fragopt_nonrigid_factor: bad arguments");
  ER_CUDA(cudaSetDevice(h->device));
  int rc = load_solver(h, "er_fragopt_nonrigid_factor");
  if (rc != ER_OK) return rc;
  const long long N = h->nr_N;
  ER_ARG(N < (1ll << 31), "er_fragopt_nonrigid_factor: system too large for cuSOLVER's 32-bit interface");
  for (size_t k = 0; k < nnz; ++k)
    ER_ARG(rows[k] >= 0 && rows[k] <= cols[k] && cols[k] < N, "er_fragopt_nonrigid_factor: triplets must lie in the upper triangle");
  if (nnz) {
    int *d_r = nullptr, *d_c = nullptr;
    double* d_v = nullptr;
    cudaError_t e = cudaMalloc(&d_r, nnz * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&d_c, nnz * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&d_v, nnz * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_r, rows, nnz * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_c, cols, nnz * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_v, vals, nnz * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
      add_triplets_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, h->stream>>>(nnz, d_r, d_c, d_v, N, h->d_AA);
      ++h->launches;
      e = cudaStreamSynchronize(h->stream);
    }
    for (void* p : {(void*)d_r, (void*)d_c, (void*)d_v}) if (p) cudaFree(p);
    if (e != cudaSuccess) { er::set_error("er_fragopt_nonrigid_factor: %s", cudaGetErrorString(e)); return ER_ERR_CUDA; }
  }
  Solver& S = solver_api();
  int need = 0;
  if (S.potrf_buffer(h->solver, CUBLAS_FILL_MODE_LOWER, (int)N, h->d_AA, (int)N, &need) != CUSOLVER_STATUS_SUCCESS) {
    er::set_error("er_fragopt_nonrigid_factor: potrf_bufferSize failed");
    return ER_ERR_CUDA;
  }
  if (need > h->work_len) {
    if (h->d_work) cudaFree(h->d_work);
    h->d_work = nullptr;
    ER_CUDA(cudaMalloc(&h->d_work, sizeof(double) * (size_t)need));
    h->work_len = need;
  }
  if (S.potrf(h->solver, CUBLAS_FILL_MODE_LOWER, (int)N, h->d_AA, (int)N, h->d_work, h->work_len, h->d_info) != CUSOLVER_STATUS_SUCCESS) {
    er::set_error("er_fragopt_nonrigid_factor: potrf failed");
    return ER_ERR_CUDA;
  }
  int info = 0;
  ER_CUDA(cudaMemcpyAsync(&info, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  ER_CUDA(cudaStreamSynchronize(h->stream));
  if (info != 0) { er::set_error("er_fragopt_nonrigid_factor: the matrix is not positive definite (potrf info = %d)", info); return ER_ERR_ARG; }
  h->nr_factored = true;
  return ER_OK;
}

int er_fragopt_nonrigid_solve(er_fragopt* h, const double* rhs, double* x) {
  ER_ARG(h && rhs && x && h->nr_factored, "er_fragopt_nonrigid_solve: factor first");
  ER_CUDA(cudaSetDevice(h->device));
  const long long N = h->nr_N;
  Solver& S = solver_api();
  ER_CUDA(cudaMemcpyAsync(h->d_nr_rhs, rhs, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, h->stream));
  if (S.potrs(h->solver, CUBLAS_FILL_MODE_LOWER, (int)N, 1, h->d_AA, (int)N, h->d_nr_rhs, (int)N, h->d_info) != CUSOLVER_STATUS_SUCCESS) {
    er::set_error("er_fragopt_nonrigid_solve: potrs failed");
    return ER_ERR_CUDA;
  }
  ER_CUDA(cudaMemcpyAsync(x, h->d_nr_rhs, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, h->stream));
  ER_CUDA(cudaStreamSynchronize(h->stream));
  ER_CUDA(cudaGetLastError());
  return ER_OK;
}

int er_fragopt_event_record(er_fragopt* h, int slot) {
  ER_ARG(h && slot >= 0 && slot < 8, "er_fragopt_event_record: bad slot");
  ER_CUDA(cudaSetDevice(h->device));
  ER_CUDA(cudaEventRecord(h->ev[slot], h->stream));
  return ER_OK;
}

int er_fragopt_event_elapsed_ms(er_fragopt* h, int a, int b, float* ms) {
  ER_ARG(h && ms && a >= 0 && a < 8 && b >= 0 && b < 8, "er_fragopt_event_elapsed_ms: bad arguments");
  ER_CUDA(cudaSetDevice(h->device));
  ER_CUDA(cudaEventSynchronize(h->ev[b]));
  ER_CUDA(cudaEventElapsedTime(ms, h->ev[a], h->ev[b]));
  return ER_OK;
}

uint64_t er_fragopt_launch_count(const er_fragopt* h) { return h ? h->launches : 0; }

int er_fragopt_sync(er_fragopt* h) {
  ER_ARG(h, "er_fragopt_sync: null handle");
  ER_CUDA(cudaSetDevice(h->device));
  ER_CUDA(cudaStreamSynchronize(h->stream));
  return ER_OK;
}

int er_fragopt_last_assemble_stats(er_fragopt* h, float* kernel_ms, uint64_t* rows, uint64_t* flushes) {
  ER_ARG(h && (h->d_JJ || h->d_AA), "er_fragopt_last_assemble_stats: nothing assembled yet");
  ER_CUDA(cudaSetDevice(h->device));
  if (kernel_ms) {
    ER_CUDA(cudaEventSynchronize(h->k1));
    ER_CUDA(cudaEventElapsedTime(kernel_ms, h->k0, h->k1));
  }
  if (rows) *rows = h->last_rows;
  if (flushes) {
    unsigned long long f = 0;
    ER_CUDA(cudaMemcpy(&f, h->d_flushes, sizeof f, cudaMemcpyDeviceToHost));
    *flushes = f;
  }
  return ER_OK;
}

}  // extern "C"
