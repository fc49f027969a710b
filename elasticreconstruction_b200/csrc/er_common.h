// er_common.h -- shared host-side plumbing of libelastic_b200 (error reporting, 4x4 double math).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "elastic_b200.h"

namespace er {

void set_error(const char* fmt, ...);

#define ER_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      er::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);   \
      return ER_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

#define ER_ARG(cond, msg)            \
  do {                               \
    if (!(cond)) {                   \
      er::set_error("%s", msg);      \
      return ER_ERR_ARG;             \
    }                                \
  } while (0)

// Row-major 4x4 double helpers that round exactly like the reference's Eigen 3.1.2 expressions
// (sequential multiply-then-add, no fma); see er_host_math.cpp.
void mat4_mul(const double* a, const double* b, double* o);
void mat4_inverse(const double* m, double* o);

}  // namespace er
