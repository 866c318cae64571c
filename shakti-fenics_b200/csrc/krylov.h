// Flexible right-preconditioned restarted GMRES (FGMRES) and BiCGStab.  Vectors, the Gram-Schmidt passes and
// their reductions live on the device; the (m+1) x m Hessenberg / Givens algebra and the convergence decision
// run on the host from one small read per iteration (written by a kernel into page-locked host memory).
// Replaces PETSc KSP(preonly)+PC(lu) behind DOLFINx NewtonSolver (reference solvers.py:52,179).
#pragma once
#include <functional>

#include "device.h"

namespace shakti {

// y(owned) = A x ; x is an n_local buffer whose ghost part the operator may overwrite (halo)
using ApplyFn = std::function<void(double* x_local, double* y_owned)>;
// z(owned part of an n_local buffer) = M^-1 r(owned)
using PrecFn = std::function<void(const double* r_owned, double* z_local)>;
// in-place sum over ranks of `count` doubles in device memory (no-op on one GPU)
using AllReduceFn = std::function<void(double* dev, int count)>;

struct KrylovResult {
  int iterations = 0;
  double relres = 0.0;   // estimated ||b - A x|| / ||b||
  bool converged = false;
};

class Gmres {
 public:
  void init(int64_t n_owned, int64_t n_local, int restart, int sm_count, cudaStream_t s);
  // Solve A x = b, x0 = 0.  x, b: owned-length device vectors.
  KrylovResult solve(const ApplyFn& A, const PrecFn& M, const AllReduceFn& allreduce, const double* b,
                     double* x, double rtol, double atol, int max_it);
  int restart() const { return m_; }
  Reducer& reducer() { return red_; }

 private:
  int64_t n_ = 0, nl_ = 0, ld_ = 0, ldz_ = 0;
  int m_ = 0;
  cudaStream_t s_ = 0;
  DevBuf<double> V_, Z_, z_, u_, r_, small_;
  Reducer red_;
  double* host_status_ = nullptr;  // pinned staging of the Gram-Schmidt coefficients / y
  double *h_, *h2_, *y_, *scal_;   // device scalars inside small_ (Hessenberg algebra runs on the host)
 public:
  ~Gmres();
};

class BiCgStab {
 public:
  void init(int64_t n_owned, int64_t n_local, int sm_count, cudaStream_t s);
  KrylovResult solve(const ApplyFn& A, const PrecFn& M, const AllReduceFn& allreduce, const double* b,
                     double* x, double rtol, double atol, int max_it);
 private:
  int64_t n_ = 0, nl_ = 0;
  cudaStream_t s_ = 0;
  DevBuf<double> r_, r0_, p_, v_, s_v_, t_, ph_, sh_, dots_;
  Reducer red_;
  double* host_ = nullptr;
 public:
  ~BiCgStab();
};

}  // namespace shakti
