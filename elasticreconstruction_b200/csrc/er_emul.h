// er_emul.h -- lets the shared arithmetic headers (er_voxel_math.h, er_nn_math.h) compile both as device code under
// nvcc and as plain host code under g++ (CPU tests in tests/emul/).  Under nvcc this header only defines ER_VOXEL_FN
// and ER_RCP_APPROX; under a host compiler it maps the CUDA intrinsics those headers use onto IEEE float/double
// operators (build with -ffp-contract=off so that no multiply-add is fused), libm's correctly rounded fmaf/sqrtf, and
// small emulations of the conversion / rounding-mode instructions.  The stand-in for rcp.approx can be perturbed by a
// test (er_voxel_rcp_bias, in ulps): results must not depend on its last bits.
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
#define ER_VOXEL_FN __device__ __forceinline__
#define ER_RCP_APPROX(r, b) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b))
#else
#define __ldg(p) (*(p))
using std::max;
using std::min;
#include <cstdint>
#include <cstring>
#define ER_VOXEL_FN static inline
extern int er_voxel_rcp_bias;          // test knob: ulps added to the emulated approximate reciprocal
static inline float er_emul_rcp(float b) {
  float r = (float)(1.0 / (double)b);
  int32_t bits;
  std::memcpy(&bits, &r, 4);
  bits += er_voxel_rcp_bias;
  std::memcpy(&r, &bits, 4);
  return r;
}
static inline int er_emul_f2i_rz(float x) {           // cvt.rzi.s32.f32: saturating, NaN -> 0
  if (x != x) return 0;
  if (x >= 2147483648.0f) return 2147483647;
  if (x <= -2147483648.0f) return (-2147483647 - 1);
  return (int)x;
}
#define ER_RCP_APPROX(r, b) ((r) = er_emul_rcp(b))
#define __fmul_rn(a, b) ((float)((float)(a) * (float)(b)))
#define __fadd_rn(a, b) ((float)((float)(a) + (float)(b)))
#define __fsub_rn(a, b) ((float)((float)(a) - (float)(b)))
#define __fdiv_rn(a, b) ((float)((float)(a) / (float)(b)))
#define __fmaf_rn(a, b, c) fmaf((a), (b), (c))
#define __fsqrt_rn(a) sqrtf(a)
#define __ddiv_rn(a, b) ((double)(a) / (double)(b))
#define __float2int_rz(a) er_emul_f2i_rz(a)
static inline int er_emul_f2i_rd(float x) {           // cvt.rmi.s32.f32: floor, saturating, NaN -> 0
  if (x != x) return 0;
  const float f = floorf(x);
  if (f >= 2147483648.0f) return 2147483647;
  if (f <= -2147483648.0f) return (-2147483647 - 1);
  return (int)f;
}
static inline float er_emul_fadd_rz(float a, float b) {   // add.rz.f32 through the exact double sum when it is exact
  const double s = (double)a + (double)b;                 // exact unless the exponents differ by more than 29 bits
  float r = (float)s;                                     // round to nearest ...
  if ((double)r != s && fabs((double)r) > fabs(s)) r = nextafterf(r, 0.0f);   // ... stepped back towards zero
  return r;
}
#define __float2int_rd(a) er_emul_f2i_rd(a)
#define __fadd_rz(a, b) er_emul_fadd_rz((a), (b))
#define __dadd_rn(a, b) ((double)(a) + (double)(b))
#define __dmul_rn(a, b) ((double)(a) * (double)(b))
static inline int __double2int_rd(double x) {        // cvt.rmi.s32.f64 (callers range-check first)
  return (int)floor(x);
}
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
#endif
