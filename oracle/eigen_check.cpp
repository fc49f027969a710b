// eigen_check.cpp -- TEST ONLY.  The three Eigen operations PCL's ICP performs around the restated arithmetic of
// icp_oracle.c, executed by the REAL Eigen 3.1.2 the reference vendors (FragmentOptimizer/external/Eigen, compiled
// where it lies; `make -C oracle eigen`), so that tests/test_oracle_icp.py can check the restatement against them:
//   * x = ATA.inverse() * ATb                       (TransformationEstimationPointToPlaneLLS::estimateRigidTransformation)
//   * final = T_inc * final,  Matrix4f * Matrix4f    (IterativeClosestPoint::computeTransformation)
//   * p' = T * (x, y, z, 1),  Matrix4f * Vector4f    (IterativeClosestPoint::transformCloud)
// PCL itself is not available (see oracle/icp_oracle.c, "parity unpinned"); this pins the Eigen half of that gap.
#include <Eigen/Dense>

extern "C" {

void eigen_inverse_solve6(const double* A_rowmajor, const double* b, double* x) {
  Eigen::Matrix<double, 6, 6> ATA;
  Eigen::Matrix<double, 6, 1> ATb;
  for (int r = 0; r < 6; r++) {
    for (int c = 0; c < 6; c++) ATA(r, c) = A_rowmajor[r * 6 + c];
    ATb(r) = b[r];
  }
  Eigen::Matrix<double, 6, 1> sol = static_cast<Eigen::Matrix<double, 6, 1> >(ATA.inverse() * ATb);
  for (int r = 0; r < 6; r++) x[r] = sol(r);
}

void eigen_mat4f_mul(const float* a_rowmajor, const float* b_rowmajor, float* o_rowmajor) {
  Eigen::Matrix4f A, B;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) { A(r, c) = a_rowmajor[r * 4 + c]; B(r, c) = b_rowmajor[r * 4 + c]; }
  const Eigen::Matrix4f O = A * B;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) o_rowmajor[r * 4 + c] = O(r, c);
}

void eigen_mat4f_point(const float* m_rowmajor, const float* p3, float* o3) {
  Eigen::Matrix4f M;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) M(r, c) = m_rowmajor[r * 4 + c];
  const Eigen::Vector4f q = M * Eigen::Vector4f(p3[0], p3[1], p3[2], 1.0f);
  o3[0] = q[0]; o3[1] = q[1]; o3[2] = q[2];
}

}  // extern "C"
