// er_host_math.cpp -- host-side double 4x4 arithmetic with the reference's rounding behaviour.
//
// The per-voxel projection is only bit-identical to the reference if the float matrices handed to the
// kernel are: the reference forms trans_inv = transformation.inverse() in double (Integrate/
// TSDFVolume.cpp:40) and narrows it to float (TSDFVolume.cpp:59).  Eigen 3.1.2 evaluates that inverse
// with the 2x2-block adjugate scheme (vendored copy: FragmentOptimizer/external/Eigen/src/LU/arch/
// Inverse_SSE.h:163-325):
//
//     M = | A B |      M^-1 = 1/det * | (A|D| - B D# C)#   (C|B| - D B# A)# |
//         | C D |                     | (B|C| - A C# D)#   (D|A| - C A# B)# |
//     det = |A||D| + |B||C| - tr(A# B D# C)          (# = 2x2 adjugate)
//
// applied to the column-major storage.  Below the same formula is written out as scalar expressions in
// that operation order (each product and sum rounded once; this TU is compiled with -ffp-contract=off).
#include <cmath>

#include "er_common.h"

namespace er {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

void mat4_mul(const double* a, const double* b, double* o) {
  double t[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      double acc = a[r * 4 + 0] * b[c];
      acc = acc + a[r * 4 + 1] * b[4 + c];
      acc = acc + a[r * 4 + 2] * b[8 + c];
      acc = acc + a[r * 4 + 3] * b[12 + c];
      t[r * 4 + c] = acc;
    }
  std::memcpy(o, t, sizeof t);
}

void mat4_inverse(const double* m, double* out) {
  // s = column-major storage of m:  s[c*4+r] = m(r,c)
  double s[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) s[c * 4 + r] = m[r * 4 + c];

  // 2x2 block determinants
  const double dA = s[0] * s[5] - s[1] * s[4];
  const double dB = s[2] * s[7] - s[3] * s[6];
  const double dC = s[8] * s[13] - s[9] * s[12];
  const double dD = s[10] * s[15] - s[11] * s[14];
  // AB = A# B, DC = D# C  (two rows each, stored as x=first lane, y=second lane)
  const double AB1x = s[2] * s[5] - s[6] * s[1], AB1y = s[3] * s[5] - s[7] * s[1];
  const double AB2x = s[6] * s[0] - s[2] * s[4], AB2y = s[7] * s[0] - s[3] * s[4];
  const double DC1x = s[8] * s[15] - s[12] * s[11], DC1y = s[9] * s[15] - s[13] * s[11];
  const double DC2x = s[12] * s[10] - s[8] * s[14], DC2y = s[13] * s[10] - s[9] * s[14];
  // tr(AB * DC)
  const double trx = AB1x * DC1x + AB2x * DC1y;
  const double try_ = AB1y * DC2x + AB2y * DC2y;
  const double tr = trx + try_;
  // C A# B and B D# C
  double iD1x = AB1x * s[8] + AB2x * s[9], iD1y = AB1y * s[8] + AB2y * s[9];
  double iD2x = AB1x * s[12] + AB2x * s[13], iD2y = AB1y * s[12] + AB2y * s[13];
  double iA1x = DC1x * s[2] + DC2x * s[3], iA1y = DC1y * s[2] + DC2y * s[3];
  double iA2x = DC1x * s[6] + DC2x * s[7], iA2y = DC1y * s[6] + DC2y * s[7];
  iD1x = s[10] * dA - iD1x; iD1y = s[11] * dA - iD1y;
  iD2x = s[14] * dA - iD2x; iD2y = s[15] * dA - iD2y;
  iA1x = s[0] * dD - iA1x;  iA1y = s[1] * dD - iA1y;
  iA2x = s[4] * dD - iA2x;  iA2y = s[5] * dD - iA2y;
  const double d1 = dA * dD;
  const double d2 = dB * dC;
  // D B# A and A C# D
  double iB1x = s[10] * AB2y - s[11] * AB2x, iB1y = s[11] * AB1x - s[10] * AB1y;
  double iB2x = s[14] * AB2y - s[15] * AB2x, iB2y = s[15] * AB1x - s[14] * AB1y;
  const double det = (d1 + d2) - tr;
  double iC1x = s[0] * DC2y - s[1] * DC2x, iC1y = s[1] * DC1x - s[0] * DC1y;
  double iC2x = s[4] * DC2y - s[5] * DC2x, iC2y = s[5] * DC1x - s[4] * DC1y;
  const double rd = 1.0 / det;
  iB1x = s[8] * dB - iB1x;  iB1y = s[9] * dB - iB1y;
  iB2x = s[12] * dB - iB2x; iB2y = s[13] * dB - iB2y;
  iC1x = s[2] * dC - iC1x;  iC1y = s[3] * dC - iC1y;
  iC2x = s[6] * dC - iC2x;  iC2y = s[7] * dC - iC2y;

  double o[16];
  o[0] = iA2y * rd;     o[1] = iA1y * -rd;
  o[4] = iA2x * -rd;    o[5] = iA1x * rd;
  o[2] = iB2y * rd;     o[3] = iB1y * -rd;
  o[6] = iB2x * -rd;    o[7] = iB1x * rd;
  o[8] = iC2y * rd;     o[9] = iC1y * -rd;
  o[12] = iC2x * -rd;   o[13] = iC1x * rd;
  o[10] = iD2y * rd;    o[11] = iD1y * -rd;
  o[14] = iD2x * -rd;   o[15] = iD1x * rd;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) out[r * 4 + c] = o[c * 4 + r];
}

}  // namespace er

extern "C" {
const char* er_last_error(void) { return er::g_err; }
int er_version(void) { return 100; }
const char* er_build_arch(void) { return "sm_100a"; }
void er_mat4_mul(const double a[16], const double b[16], double out[16]) { er::mat4_mul(a, b, out); }
void er_mat4_inverse(const double m[16], double out[16]) { er::mat4_inverse(m, out); }
}
