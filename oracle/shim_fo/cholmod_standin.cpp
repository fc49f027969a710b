// TEST SCAFFOLDING (oracle build only) -- not product code.
//
// The reference links SuiteSparse's CHOLMOD through Eigen's CholmodSupport
// (FragmentOptimizer/HashSparseMatrix.h:7; OptApp.cpp:209-211,389-391,575-577).  The tree vendors CHOLMOD's headers
// (FragmentOptimizer/external/SuiteSparse/Include) but only Windows import libraries, and no SuiteSparse is installed
// here.  This file implements the nine CHOLMOD entry points Eigen 3.1.2's CholmodSupport.h calls, against the
// vendored cholmod.h, on top of the SAME vendored Eigen's SimplicialLLT (sparse LL^T with AMD ordering): an exact
// Cholesky solve in double, i.e. the same mathematical operation with a different elimination order.  Everything
// else of FragmentOptimizer_ref (assembly of the normal equations, regulariser, pose/lattice updates, file formats)
// is the reference's own unmodified code.  Solutions differ from CHOLMOD's at rounding level only
// (cond(A) * 1e-16 relative); the parity tolerances of tests/test_fragopt_*.py are stated with that in mind.
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>
#include "cholmod.h"
#include "Eigen/Core"
#include "Eigen/SparseCore"
#include "Eigen/OrderingMethods"
#include "Eigen/SparseCholesky"

namespace {
typedef Eigen::SparseMatrix<double> SpMat;
typedef Eigen::SimplicialLLT<SpMat, Eigen::Upper> Solver;
std::map<cholmod_factor*, Solver*>& table() {
  static std::map<cholmod_factor*, Solver*> t;
  return t;
}
// copies the stored triangle of a cholmod_sparse view (CSC, int indices, double values)
SpMat to_eigen(const cholmod_sparse* A) {
  const int* Ap = (const int*)A->p;
  const int* Ai = (const int*)A->i;
  const int* Anz = (const int*)A->nz;
  const double* Ax = (const double*)A->x;
  std::vector<Eigen::Triplet<double> > tri;
  for (size_t j = 0; j < A->ncol; ++j) {
    int b = Ap[j], e = A->packed ? Ap[j + 1] : Ap[j] + Anz[j];
    for (int k = b; k < e; ++k) {
      int i = Ai[k];
      if (A->stype > 0 && i > (int)j) continue;   // upper triangle is the matrix
      if (A->stype < 0 && i < (int)j) continue;
      tri.push_back(Eigen::Triplet<double>(A->stype < 0 ? (int)j : i, A->stype < 0 ? i : (int)j, Ax[k]));
    }
  }
  SpMat M((int)A->nrow, (int)A->ncol);
  M.setFromTriplets(tri.begin(), tri.end());
  return M;
}
}  // namespace

extern "C" {
int cholmod_start(cholmod_common* c) { std::memset(c, 0, sizeof *c); return 1; }
int cholmod_finish(cholmod_common*) { return 1; }

cholmod_factor* cholmod_analyze(cholmod_sparse* A, cholmod_common*) {
  cholmod_factor* L = (cholmod_factor*)std::calloc(1, sizeof(cholmod_factor));
  L->n = A->nrow;
  Solver* s = new Solver();
  SpMat M = to_eigen(A);
  s->analyzePattern(M);
  table()[L] = s;
  return L;
}
int cholmod_factorize(cholmod_sparse* A, cholmod_factor* L, cholmod_common*) {
  SpMat M = to_eigen(A);
  Solver* s = table()[L];
  s->factorize(M);
  return s->info() == Eigen::Success;
}
cholmod_dense* cholmod_solve(int sys, cholmod_factor* L, cholmod_dense* B, cholmod_common*) {
  if (sys != CHOLMOD_A) return 0;
  Solver* s = table()[L];
  cholmod_dense* X = (cholmod_dense*)std::calloc(1, sizeof(cholmod_dense));
  X->nrow = B->nrow; X->ncol = B->ncol; X->nzmax = B->nrow * B->ncol; X->d = B->nrow;
  X->xtype = B->xtype; X->dtype = B->dtype;
  X->x = std::malloc(sizeof(double) * X->nzmax);
  for (size_t c = 0; c < B->ncol; ++c) {
    Eigen::Map<const Eigen::VectorXd> b((const double*)B->x + c * B->d, B->nrow);
    Eigen::VectorXd x = s->solve(b);
    std::memcpy((double*)X->x + c * X->d, x.data(), sizeof(double) * B->nrow);
  }
  return X;
}
cholmod_sparse* cholmod_spsolve(int, cholmod_factor*, cholmod_sparse*, cholmod_common*) { return 0; }  // not reached
int cholmod_free_factor(cholmod_factor** L, cholmod_common*) {
  if (L && *L) { delete table()[*L]; table().erase(*L); std::free(*L); *L = 0; }
  return 1;
}
int cholmod_free_dense(cholmod_dense** X, cholmod_common*) {
  if (X && *X) { std::free((*X)->x); std::free(*X); *X = 0; }
  return 1;
}
int cholmod_free_sparse(cholmod_sparse** X, cholmod_common*) {
  if (X && *X) { std::free(*X); *X = 0; }
  return 1;
}
}
