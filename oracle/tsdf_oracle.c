/*
 * oracle/tsdf_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's TSDF integration path, used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker
 * for the CUDA path.  Nothing under elasticreconstruction_b200/ may call into it.
 *
 * Parity status: PINNED.  tests/test_oracle_tsdf.py checks this file bit-for-bit
 * against the verbatim reference binary oracle/_ref/Integrate_ref (built from the
 * untouched sources under /root/reference/Integrate by oracle/Makefile) and against
 * the golden fixtures under tests/golden/ that were produced by that binary.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference).  Arithmetic types and evaluation order are the reference's:
 * build with -ffp-contract=off and without -mfma (see oracle/Makefile); a fused
 * multiply-add anywhere in the projection flips round() for a few voxels per
 * million.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ER_BLOCK_RES 64            /* Integrate/TSDFVolume.cpp:57  */
#define ER_GRID_HALF 256           /* Integrate/TSDFVolume.h:62-68 */
#define ER_GRID_DIM 512

typedef struct {
  int key, xi, yi, zi;
  float *sdf, *weight;             /* Integrate/TSDFVolumeUnit.h:98-113: two float[64^3], zeroed */
} oracle_block;

typedef struct {
  int cols, rows;
  float fx, fy, cx, cy, icp_trunc, integration_trunc;   /* CameraParam, TSDFVolumeUnit.h:65-70 */
  double unit_length;              /* TSDFVolume.cpp:10 */
  double tsdf_trunc;               /* TSDFVolume.cpp:11 */
  oracle_block *blocks;            /* insertion order (the reference uses an unordered_map) */
  int nblocks, cap;
  int *table;                      /* open-addressing key -> block index */
  int table_cap;
  float *scaled;
  int threads;
  /* statistics of the last integrate call (for the roofline's N_updated) */
  long long last_updated_voxels;
  int last_touched_blocks;
  int *last_touched_keys;          /* keys of the blocks the last frame touched, in first-touch order */
} oracle_tsdf;

/* ------------------------------------------------------------------ small double 4x4 helpers */

/* Eigen 3.1.2 fixed-size product, coefficient by coefficient:
 * ((l(r,0)*x0 + l(r,1)*x1) + l(r,2)*x2) + l(r,3)*x3, multiply then add, no fma
 * (FragmentOptimizer/external/Eigen/src/Core/products/CoeffBasedProduct.h:239-256, 381-408).
 * Matrices here are row-major double[16]. */
static void mat4_mul_vec(const double *m, const double *v, double *o) {
  for (int r = 0; r < 4; r++)
    o[r] = ((m[r * 4 + 0] * v[0] + m[r * 4 + 1] * v[1]) + m[r * 4 + 2] * v[2]) + m[r * 4 + 3] * v[3];
}
static void mat4_mul(const double *a, const double *b, double *o) {
  double t[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++)
      t[r * 4 + c] = ((a[r * 4 + 0] * b[0 * 4 + c] + a[r * 4 + 1] * b[1 * 4 + c]) + a[r * 4 + 2] * b[2 * 4 + c]) +
                     a[r * 4 + 3] * b[3 * 4 + c];
  memcpy(o, t, sizeof t);
}

/* Matrix4d::inverse() in Eigen 3.1.2 with SSE2 enabled (every x86-64 build) is the 2x2-block
 * adjugate scheme of FragmentOptimizer/external/Eigen/src/LU/arch/Inverse_SSE.h:163-325.
 * It is restated here lane by lane on a two-double pair so that every intermediate rounding
 * is the same as in the reference build.  `m` and `out` are row-major; Eigen stores
 * column-major, so element k of its storage is m[(k%4)*4 + k/4]. */
typedef struct { double lo, hi; } pd2;
static pd2 pd_mul(pd2 a, pd2 b) { pd2 r = {a.lo * b.lo, a.hi * b.hi}; return r; }
static pd2 pd_sub(pd2 a, pd2 b) { pd2 r = {a.lo - b.lo, a.hi - b.hi}; return r; }
static pd2 pd_add(pd2 a, pd2 b) { pd2 r = {a.lo + b.lo, a.hi + b.hi}; return r; }
static pd2 pd_shuf(pd2 a, pd2 b, int imm) { pd2 r = {(imm & 1) ? a.hi : a.lo, (imm & 2) ? b.hi : b.lo}; return r; }
static pd2 pd_sub_lo(pd2 a, pd2 b) { pd2 r = {a.lo - b.lo, a.hi}; return r; }
static pd2 pd_add_lo(pd2 a, pd2 b) { pd2 r = {a.lo + b.lo, a.hi}; return r; }
static pd2 pd_mul_lo(pd2 a, pd2 b) { pd2 r = {a.lo * b.lo, a.hi}; return r; }

void oracle_mat4_inverse(const double *m, double *out) {
  double s[16], o[16];
  for (int k = 0; k < 16; k++) s[k] = m[(k % 4) * 4 + k / 4];
#define LD(i) ((pd2){s[i], s[(i) + 1]})
  pd2 A1 = LD(0), B1 = LD(2), A2 = LD(4), B2 = LD(6), C1 = LD(8), D1 = LD(10), C2 = LD(12), D2 = LD(14);
#undef LD
  pd2 iA1, iA2, iB1, iB2, iC1, iC2, iD1, iD2, DC1, DC2, AB1, AB2, dA, dB, dC, dD, det, d1, d2, rd;
  dA = pd_mul(A1, pd_shuf(A2, A2, 1)); dA = pd_sub_lo(dA, pd_shuf(dA, dA, 3));
  dB = pd_mul(B1, pd_shuf(B2, B2, 1)); dB = pd_sub_lo(dB, pd_shuf(dB, dB, 3));
  AB1 = pd_mul(B1, pd_shuf(A2, A2, 3));
  AB2 = pd_mul(B2, pd_shuf(A1, A1, 0));
  AB1 = pd_sub(AB1, pd_mul(B2, pd_shuf(A1, A1, 3)));
  AB2 = pd_sub(AB2, pd_mul(B1, pd_shuf(A2, A2, 0)));
  dC = pd_mul(C1, pd_shuf(C2, C2, 1)); dC = pd_sub_lo(dC, pd_shuf(dC, dC, 3));
  dD = pd_mul(D1, pd_shuf(D2, D2, 1)); dD = pd_sub_lo(dD, pd_shuf(dD, dD, 3));
  DC1 = pd_mul(C1, pd_shuf(D2, D2, 3));
  DC2 = pd_mul(C2, pd_shuf(D1, D1, 0));
  DC1 = pd_sub(DC1, pd_mul(C2, pd_shuf(D1, D1, 3)));
  DC2 = pd_sub(DC2, pd_mul(C1, pd_shuf(D2, D2, 0)));
  d1 = pd_mul(AB1, pd_shuf(DC1, DC2, 0));
  d2 = pd_mul(AB2, pd_shuf(DC1, DC2, 3));
  rd = pd_add(d1, d2);
  rd = pd_add_lo(rd, pd_shuf(rd, rd, 3));
  iD1 = pd_mul(AB1, pd_shuf(C1, C1, 0));
  iD2 = pd_mul(AB1, pd_shuf(C2, C2, 0));
  iD1 = pd_add(iD1, pd_mul(AB2, pd_shuf(C1, C1, 3)));
  iD2 = pd_add(iD2, pd_mul(AB2, pd_shuf(C2, C2, 3)));
  iA1 = pd_mul(DC1, pd_shuf(B1, B1, 0));
  iA2 = pd_mul(DC1, pd_shuf(B2, B2, 0));
  iA1 = pd_add(iA1, pd_mul(DC2, pd_shuf(B1, B1, 3)));
  iA2 = pd_add(iA2, pd_mul(DC2, pd_shuf(B2, B2, 3)));
  dA = pd_shuf(dA, dA, 0);
  iD1 = pd_sub(pd_mul(D1, dA), iD1);
  iD2 = pd_sub(pd_mul(D2, dA), iD2);
  dD = pd_shuf(dD, dD, 0);
  iA1 = pd_sub(pd_mul(A1, dD), iA1);
  iA2 = pd_sub(pd_mul(A2, dD), iA2);
  d1 = pd_mul_lo(dA, dD);
  d2 = pd_mul_lo(dB, dC);
  iB1 = pd_mul(D1, pd_shuf(AB2, AB1, 1));
  iB2 = pd_mul(D2, pd_shuf(AB2, AB1, 1));
  iB1 = pd_sub(iB1, pd_mul(pd_shuf(D1, D1, 1), pd_shuf(AB2, AB1, 2)));
  iB2 = pd_sub(iB2, pd_mul(pd_shuf(D2, D2, 1), pd_shuf(AB2, AB1, 2)));
  det = pd_add_lo(d1, d2);
  det = pd_sub_lo(det, rd);
  iC1 = pd_mul(A1, pd_shuf(DC2, DC1, 1));
  iC2 = pd_mul(A2, pd_shuf(DC2, DC1, 1));
  iC1 = pd_sub(iC1, pd_mul(pd_shuf(A1, A1, 1), pd_shuf(DC2, DC1, 2)));
  iC2 = pd_sub(iC2, pd_mul(pd_shuf(A2, A2, 1), pd_shuf(DC2, DC1, 2)));
  rd.lo = 1.0 / det.lo;
  rd.hi = rd.lo;
  dB = pd_shuf(dB, dB, 0);
  iB1 = pd_sub(pd_mul(C1, dB), iB1);
  iB2 = pd_sub(pd_mul(C2, dB), iB2);
  d1.lo = rd.lo;  d1.hi = -rd.hi;   /* xor with the sign mask (+,-) */
  d2.lo = -rd.lo; d2.hi = rd.hi;    /* (-,+) */
  dC = pd_shuf(dC, dC, 0);
  iC1 = pd_sub(pd_mul(B1, dC), iC1);
  iC2 = pd_sub(pd_mul(B2, dC), iC2);
#define ST(i, v) do { pd2 t_ = (v); o[i] = t_.lo; o[(i) + 1] = t_.hi; } while (0)
  ST(0, pd_mul(pd_shuf(iA2, iA1, 3), d1));
  ST(4, pd_mul(pd_shuf(iA2, iA1, 0), d2));
  ST(2, pd_mul(pd_shuf(iB2, iB1, 3), d1));
  ST(6, pd_mul(pd_shuf(iB2, iB1, 0), d2));
  ST(8, pd_mul(pd_shuf(iC2, iC1, 3), d1));
  ST(12, pd_mul(pd_shuf(iC2, iC1, 0), d2));
  ST(10, pd_mul(pd_shuf(iD2, iD1, 3), d1));
  ST(14, pd_mul(pd_shuf(iD2, iD1, 0), d2));
#undef ST
  for (int k = 0; k < 16; k++) out[(k % 4) * 4 + k / 4] = o[k];
}

void oracle_mat4_mul(const double *a, const double *b, double *o) { mat4_mul(a, b, o); }

/* ------------------------------------------------------------------ volume container */

oracle_tsdf *oracle_tsdf_create(int cols, int rows) {
  oracle_tsdf *o = (oracle_tsdf *)calloc(1, sizeof *o);
  o->cols = cols; o->rows = rows;
  /* CameraParam defaults, Integrate/TSDFVolumeUnit.h:69 */
  o->fx = 525.0f; o->fy = 525.0f; o->cx = 319.5f; o->cy = 239.5f; o->icp_trunc = 2.5f; o->integration_trunc = 2.5f;
  o->unit_length = 3.0 / 512.0;   /* Integrate/TSDFVolume.cpp:10 */
  o->tsdf_trunc = 0.03;           /* Integrate/TSDFVolume.cpp:11 */
  o->cap = 64;
  o->blocks = (oracle_block *)calloc((size_t)o->cap, sizeof(oracle_block));
  o->table_cap = 1 << 12;
  o->table = (int *)malloc(sizeof(int) * (size_t)o->table_cap);
  for (int i = 0; i < o->table_cap; i++) o->table[i] = -1;
  o->scaled = (float *)calloc((size_t)cols * rows, sizeof(float));
  o->threads = 8;                 /* num_threads( 8 ), Integrate/TSDFVolume.cpp:21,71 */
  return o;
}

void oracle_tsdf_destroy(oracle_tsdf *o) {
  if (!o) return;
  for (int i = 0; i < o->nblocks; i++) { free(o->blocks[i].sdf); free(o->blocks[i].weight); }
  free(o->blocks); free(o->table); free(o->scaled); free(o->last_touched_keys); free(o);
}

void oracle_tsdf_set_camera(oracle_tsdf *o, float fx, float fy, float cx, float cy, float integration_trunc) {
  o->fx = fx; o->fy = fy; o->cx = cx; o->cy = cy; o->integration_trunc = integration_trunc;
}
void oracle_tsdf_set_threads(oracle_tsdf *o, int t) { o->threads = t > 0 ? t : 1; }
/* test hook: the reference hard-codes these (TSDFVolume.cpp:10-11); BASELINE config 1 asks for 12 mm */
void oracle_tsdf_set_voxel(oracle_tsdf *o, double unit_length, double tsdf_trunc) {
  o->unit_length = unit_length; o->tsdf_trunc = tsdf_trunc;
}

static unsigned hash_u32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

static int table_find(const oracle_tsdf *o, int key) {
  unsigned h = hash_u32((unsigned)key) & (unsigned)(o->table_cap - 1);
  for (;;) {
    int b = o->table[h];
    if (b < 0) return -1;
    if (o->blocks[b].key == key) return b;
    h = (h + 1) & (unsigned)(o->table_cap - 1);
  }
}
static void table_put(oracle_tsdf *o, int key, int b) {
  unsigned h = hash_u32((unsigned)key) & (unsigned)(o->table_cap - 1);
  while (o->table[h] >= 0) h = (h + 1) & (unsigned)(o->table_cap - 1);
  o->table[h] = b;
}
static int block_get_or_create(oracle_tsdf *o, int key, int xi, int yi, int zi) {
  int b = table_find(o, key);
  if (b >= 0) return b;
  if (o->nblocks == o->cap) {
    o->cap *= 2;
    o->blocks = (oracle_block *)realloc(o->blocks, sizeof(oracle_block) * (size_t)o->cap);
  }
  if ((o->nblocks + 1) * 2 > o->table_cap) {
    o->table_cap *= 2;
    o->table = (int *)realloc(o->table, sizeof(int) * (size_t)o->table_cap);
    for (int i = 0; i < o->table_cap; i++) o->table[i] = -1;
    for (int i = 0; i < o->nblocks; i++) table_put(o, o->blocks[i].key, i);
  }
  b = o->nblocks++;
  size_t vol = (size_t)ER_BLOCK_RES * ER_BLOCK_RES * ER_BLOCK_RES;
  o->blocks[b].key = key; o->blocks[b].xi = xi; o->blocks[b].yi = yi; o->blocks[b].zi = zi;
  o->blocks[b].sdf = (float *)calloc(vol, sizeof(float));      /* TSDFVolumeUnit.cpp:10-14 */
  o->blocks[b].weight = (float *)calloc(vol, sizeof(float));
  table_put(o, key, b);
  return b;
}

/* TSDFVolume::round, Integrate/TSDFVolume.h:36-38 */
static int ref_round(double x) { return (int)floor(x + 0.5); }

/* ------------------------------------------------------------------ a1: ScaleDepth */
/* Integrate/TSDFVolume.cpp:19-36 */
void oracle_scale_depth(const oracle_tsdf *o, const uint16_t *depth, float *scaled) {
#pragma omp parallel for num_threads(o->threads)
  for (int y = 0; y < o->rows; y++) {
    for (int x = 0; x < o->cols; x++) {
      unsigned short d = depth[y * o->cols + x];
      float xl = ((float)x - o->cx) / o->fx;
      float yl = ((float)y - o->cy) / o->fy;
      float lambda = sqrtf(xl * xl + yl * yl + 1.0f);
      float res = (float)(int)d * lambda / 1000.f;
      if (res > o->integration_trunc) scaled[y * o->cols + x] = 0.0f;
      else scaled[y * o->cols + x] = (float)(int)d * lambda / 1000.f;
    }
  }
}

/* ------------------------------------------------------------------ a3: IntegrateVolumeUnit */
/* Integrate/TSDFVolume.cpp:69-102.  trans / trans_inv are the float casts of the double pose and
 * of its double inverse (TSDFVolume.cpp:59), row-major here. */
static long long integrate_volume_unit(const oracle_tsdf *o, const float *scaled, const float *trans,
                                       const float *trans_inv, oracle_block *vol, float x_shift, float y_shift,
                                       float z_shift) {
  long long updated = 0;
  const int R = ER_BLOCK_RES;
#pragma omp parallel for num_threads(o->threads) reduction(+ : updated)
  for (int i = 0; i < R; i++) {
    for (int j = 0; j < R; j++) {
      for (int k = 0; k < R; k++) {
        /* int * double + float -> double, narrowed by the Vector4f constructor (TSDFVolume.cpp:75) */
        float gx = (float)(i * o->unit_length + x_shift);
        float gy = (float)(j * o->unit_length + y_shift);
        float gz = (float)(k * o->unit_length + z_shift);
        /* tgv = trans_inv * gridv, Eigen float product: ((m0*x + m1*y) + m2*z) + m3*1 */
        float t0 = ((trans_inv[0] * gx + trans_inv[1] * gy) + trans_inv[2] * gz) + trans_inv[3] * 1.0f;
        float t1 = ((trans_inv[4] * gx + trans_inv[5] * gy) + trans_inv[6] * gz) + trans_inv[7] * 1.0f;
        float t2 = ((trans_inv[8] * gx + trans_inv[9] * gy) + trans_inv[10] * gz) + trans_inv[11] * 1.0f;
        if (t2 > 0) {
          int coox = ref_round((double)(t0 * o->fx / t2 + o->cx));   /* float expr, widened by round(double) */
          int cooy = ref_round((double)(t1 * o->fy / t2 + o->cy));
          if (coox >= 0 && coox < o->cols && cooy >= 0 && cooy < o->rows) {
            float dp_scaled = scaled[cooy * o->cols + coox];
            if (dp_scaled > 0.001f) {
              float rx = gx - trans[3];
              float ry = gy - trans[7];
              float rz = gz - trans[11];
              float sdf = dp_scaled - sqrtf(rx * rx + ry * ry + rz * rz);
              if ((double)sdf >= -o->tsdf_trunc) {                  /* double compare, TSDFVolume.cpp:87 */
                float q = (float)((double)sdf / o->tsdf_trunc);     /* double divide, narrowed by min<float> */
                float tsdf = (q < 1.0f) ? q : 1.0f;                 /* std::min<float>(1.0f, q) */
                float w = 1.0f;
                int lll = (i * R + j) * R + k;
                vol->sdf[lll] = (vol->sdf[lll] * vol->weight[lll] + w * tsdf) / (vol->weight[lll] + w);
                vol->weight[lll] += w;
                updated++;
              }
            }
          }
        }
      }
    }
  }
  return updated;
}

/* ------------------------------------------------------------------ a2: Integrate */
/* Integrate/TSDFVolume.cpp:38-67.  T is row-major camera->world double[16].  Returns the number of
 * blocks touched (== IntegrateVolumeUnit calls) in this frame. */
int oracle_tsdf_integrate_scaled(oracle_tsdf *o, const uint16_t *depth, const float *scaled, const double *T) {
  double Tinv[16];
  oracle_mat4_inverse(T, Tinv);
  float Tf[16], Tinvf[16];
  for (int i = 0; i < 16; i++) { Tf[i] = (float)T[i]; Tinvf[i] = (float)Tinv[i]; }

  int touched_cap = 256, ntouched = 0;
  int *touched = (int *)malloc(sizeof(int) * (size_t)touched_cap);
  long long updated = 0;
  for (int v = 0; v < o->rows; v++) {
    for (int u = 0; u < o->cols; u++) {
      unsigned short d = depth[v * o->cols + u];
      if (d > 0) {
        /* UVD2XYZ, TSDFVolume.h:40-49: (u - cx_) is int - float = float; the rest is double */
        double z = d / 1000.0;
        double x = ((float)u - o->cx) * z / o->fx;
        double y = ((float)v - o->cy) * z / o->fy;
        double p[4] = {x, y, z, 1.0}, vv[4];
        mat4_mul_vec(T, p, vv);
        int xi = ((int)floor(vv[0] / o->unit_length + 0.5) + (ER_GRID_HALF * ER_BLOCK_RES)) / ER_BLOCK_RES;
        int yi = ((int)floor(vv[1] / o->unit_length + 0.5) + (ER_GRID_HALF * ER_BLOCK_RES)) / ER_BLOCK_RES;
        int zi = ((int)floor(vv[2] / o->unit_length + 0.5) + (ER_GRID_HALF * ER_BLOCK_RES)) / ER_BLOCK_RES;
        int key = xi * ER_GRID_DIM * ER_GRID_DIM + yi * ER_GRID_DIM + zi;   /* hash_key, TSDFVolume.h:62-64 */
        int seen = 0;
        for (int t = ntouched - 1; t >= 0; t--) if (touched[t] == key) { seen = 1; break; }
        if (!seen) {
          if (ntouched == touched_cap) { touched_cap *= 2; touched = (int *)realloc(touched, sizeof(int) * (size_t)touched_cap); }
          touched[ntouched++] = key;
          int b = block_get_or_create(o, key, xi, yi, zi);
          /* I2F, TSDFVolume.h:66-68: double product narrowed to float */
          float xs = (float)((xi - ER_GRID_HALF) * ER_BLOCK_RES * o->unit_length);
          float ys = (float)((yi - ER_GRID_HALF) * ER_BLOCK_RES * o->unit_length);
          float zs = (float)((zi - ER_GRID_HALF) * ER_BLOCK_RES * o->unit_length);
          updated += integrate_volume_unit(o, scaled, Tf, Tinvf, &o->blocks[b], xs, ys, zs);
        }
      }
    }
  }
  free(o->last_touched_keys);
  o->last_touched_keys = touched;
  o->last_updated_voxels = updated;
  o->last_touched_blocks = ntouched;
  return ntouched;
}

/* ScaleDepth + Integrate, the pair of calls at Integrate/IntegrateApp.cpp:224-225 */
int oracle_tsdf_integrate(oracle_tsdf *o, const uint16_t *depth, const double *T) {
  oracle_scale_depth(o, depth, o->scaled);
  return oracle_tsdf_integrate_scaled(o, depth, o->scaled, T);
}

long long oracle_tsdf_last_updated_voxels(const oracle_tsdf *o) { return o->last_updated_voxels; }
int oracle_tsdf_last_touched(const oracle_tsdf *o, int *keys) {
  if (keys) for (int i = 0; i < o->last_touched_blocks; i++) keys[i] = o->last_touched_keys[i];
  return o->last_touched_blocks;
}
int oracle_tsdf_num_blocks(const oracle_tsdf *o) { return o->nblocks; }
void oracle_tsdf_block_keys(const oracle_tsdf *o, int *keys) { for (int i = 0; i < o->nblocks; i++) keys[i] = o->blocks[i].key; }
int oracle_tsdf_read_block(const oracle_tsdf *o, int key, float *sdf, float *weight) {
  int b = table_find(o, key);
  if (b < 0) return -1;
  size_t vol = (size_t)ER_BLOCK_RES * ER_BLOCK_RES * ER_BLOCK_RES;
  memcpy(sdf, o->blocks[b].sdf, vol * sizeof(float));
  memcpy(weight, o->blocks[b].weight, vol * sizeof(float));
  return 0;
}

/* ------------------------------------------------------------------ a5: SaveWorld predicate */
/* Integrate/TSDFVolume.cpp:104-132.  Writes up to cap records (x,y,z,intensity) and returns the total
 * count (call with cap 0 to size the buffer).  Coordinates are global voxel indices as floats. */
size_t oracle_tsdf_extract(const oracle_tsdf *o, float *xyzi, size_t cap) {
  size_t n = 0;
  const int R = ER_BLOCK_RES;
  for (int b = 0; b < o->nblocks; b++) {
    const oracle_block *u = &o->blocks[b];
    const float *sdf = u->sdf, *w = u->weight;
    for (int i = 0; i < R; i++)
      for (int j = 0; j < R; j++)
        for (int k = 0; k < R; k++, sdf++, w++) {
          if (*w != 0.0f && *sdf < 0.98f && *sdf >= -0.98f) {
            if (n < cap) {
              xyzi[4 * n + 0] = (float)(i + (u->xi - ER_GRID_HALF) * R);
              xyzi[4 * n + 1] = (float)(j + (u->yi - ER_GRID_HALF) * R);
              xyzi[4 * n + 2] = (float)(k + (u->zi - ER_GRID_HALF) * R);
              xyzi[4 * n + 3] = *sdf;
            }
            n++;
          }
        }
  }
  return n;
}

/* ------------------------------------------------------------------ a6: ControlGrid */
/* Integrate/ControlGrid.h:44-87.  ctr = (res+1)^3 float xyz, index i + j(res+1) + k(res+1)^2. */
static int grid_get_position(const float *ctr, int res, float unit_length, const float *pt, float *pos) {
  int corner[3] = {(int)floorf(pt[0] / unit_length), (int)floorf(pt[1] / unit_length), (int)floorf(pt[2] / unit_length)};
  if (corner[0] < 0 || corner[0] >= res || corner[1] < 0 || corner[1] >= res || corner[2] < 0 || corner[2] >= res) return 0;
  float r[3] = {pt[0] / unit_length - (float)corner[0], pt[1] / unit_length - (float)corner[1],
                pt[2] / unit_length - (float)corner[2]};
  int n1 = res + 1, n2 = n1 * n1;
#define GI(i, j, k) ((i) + (j) * n1 + (k) * n2)
  int idx[8] = {GI(corner[0], corner[1], corner[2]),         GI(corner[0], corner[1], corner[2] + 1),
                GI(corner[0], corner[1] + 1, corner[2]),     GI(corner[0], corner[1] + 1, corner[2] + 1),
                GI(corner[0] + 1, corner[1], corner[2]),     GI(corner[0] + 1, corner[1], corner[2] + 1),
                GI(corner[0] + 1, corner[1] + 1, corner[2]), GI(corner[0] + 1, corner[1] + 1, corner[2] + 1)};
#undef GI
  float val[8];
  val[0] = (1 - r[0]) * (1 - r[1]) * (1 - r[2]);
  val[1] = (1 - r[0]) * (1 - r[1]) * (r[2]);
  val[2] = (1 - r[0]) * (r[1]) * (1 - r[2]);
  val[3] = (1 - r[0]) * (r[1]) * (r[2]);
  val[4] = (r[0]) * (1 - r[1]) * (1 - r[2]);
  val[5] = (r[0]) * (1 - r[1]) * (r[2]);
  val[6] = (r[0]) * (r[1]) * (1 - r[2]);
  val[7] = (r[0]) * (r[1]) * (r[2]);
  /* GetPosition: left-nested sum of scalar*Vector3f, per coefficient (ControlGrid.h:82-87) */
  for (int c = 0; c < 3; c++) {
    float acc = val[0] * ctr[3 * idx[0] + c] + val[1] * ctr[3 * idx[1] + c];
    for (int t = 2; t < 8; t++) acc = acc + val[t] * ctr[3 * idx[t] + c];
    pos[c] = acc;
  }
  return 1;
}

/* ------------------------------------------------------------------ a7: Reproject */
/* Integrate/IntegrateApp.cpp:228-269.  depth is rewritten in place.  Ti = traj_[frame-1],
 * T0 = traj_[0], A0 = seg_traj_[0], Ai = seg_traj_[frame-1] (row-major double[16]);
 * ctr = lattice of chunk (frame-1)/interval; unit_length = float(length)/res (ControlGrid.cpp:19). */
void oracle_reproject(const oracle_tsdf *o, uint16_t *depth, const double *Ti, const double *T0, const double *A0,
                      const double *Ai, const float *ctr, int res, float length) {
  const int n = o->cols * o->rows;
  uint16_t *buf = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)n);
  for (int i = 0; i < n; i++) { buf[i] = depth[i]; depth[i] = 0; }
  double Ti_inv[16], A0_inv[16], adj[16];
  oracle_mat4_inverse(Ti, Ti_inv);
  oracle_mat4_inverse(A0, A0_inv);
  mat4_mul(Ti_inv, T0, adj);       /* left to right: (Ti^-1 * T0) * A0^-1 */
  mat4_mul(adj, A0_inv, adj);
  float unit_length = length / (float)res;
  for (int v = 0; v < o->rows; v++) {
    for (int u = 0; u < o->cols; u++) {
      unsigned short d = buf[v * o->cols + u];
      if (d > 0) {
        double z = d / 1000.0;
        double x = ((float)u - o->cx) * z / o->fx;
        double y = ((float)v - o->cy) * z / o->fy;
        double p[4] = {x, y, z, 1.0}, dummy[4];
        mat4_mul_vec(Ai, p, dummy);
        float pt[3] = {(float)dummy[0], (float)dummy[1], (float)dummy[2]}, pos[3];
        if (grid_get_position(ctr, res, unit_length, pt, pos)) {
          double q[4] = {pos[0], pos[1], pos[2], 1.0}, rp[4];
          mat4_mul_vec(adj, q, rp);
          /* XYZ2UVD, TSDFVolume.h:51-60 (bounds are the literal 640 x 480) */
          if (rp[2] > 0) {
            int uu = ref_round(rp[0] * o->fx / rp[2] + o->cx);
            int vv = ref_round(rp[1] * o->fy / rp[2] + o->cy);
            unsigned short dd = (unsigned short)ref_round(rp[2] * 1000.0);
            if (uu >= 0 && uu < 640 && vv >= 0 && vv < 480) {
              unsigned short ddd = depth[vv * o->cols + uu];
              if (ddd == 0 || ddd > dd) depth[vv * o->cols + uu] = dd;
            }
          }
        }
      }
    }
  }
  free(buf);
}
