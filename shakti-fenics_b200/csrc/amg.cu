// Smoothed-aggregation AMG (see amg.h).  Symbolic phase on the host (patterns are fixed by
// the mesh), numeric phase and V-cycle on the device.
#include "amg.h"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace shakti {

// ------------------------------------------------------------------ host: aggregation + patterns

// Greedy (Vanek) aggregation on the pattern graph of the square block of A.
// Returns agg[i] in [0,n_agg) or -1 for excluded rows.
static int aggregate(const HostCsr& A, int32_t n, const std::vector<uint8_t>& excl, std::vector<int32_t>& agg) {
  agg.assign(n, -1);
  std::vector<uint8_t> free_(n, 1);
  for (int32_t i = 0; i < n; ++i)
    if (!excl.empty() && excl[i]) free_[i] = 0;
  int32_t na = 0;
  auto nb = [&](int32_t i, auto&& f) {
    for (int32_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
      const int32_t j = A.col[k];
      if (j < n && j != i && (excl.empty() || !excl[j])) f(j);
    }
  };
  // pass 1: roots whose whole neighbourhood is free
  for (int32_t i = 0; i < n; ++i) {
    if (!free_[i]) continue;
    bool ok = true;
    nb(i, [&](int32_t j) { ok &= (free_[j] != 0); });
    if (!ok) continue;
    agg[i] = na;
    free_[i] = 0;
    nb(i, [&](int32_t j) { agg[j] = na; free_[j] = 0; });
    ++na;
  }
  // pass 2: attach leftovers to a neighbouring pass-1 aggregate (smallest so far)
  {
    std::vector<int32_t> snap(agg);
    std::vector<int32_t> size(na, 0);
    for (int32_t i = 0; i < n; ++i)
      if (snap[i] >= 0) size[snap[i]]++;
    for (int32_t i = 0; i < n; ++i) {
      if (!free_[i]) continue;
      int32_t best = -1;
      nb(i, [&](int32_t j) {
        if (snap[j] >= 0 && (best < 0 || size[snap[j]] < size[best])) best = snap[j];
      });
      if (best >= 0) { agg[i] = best; size[best]++; free_[i] = 0; }
    }
  }
  // pass 3: whatever is left forms aggregates with its free neighbours
  for (int32_t i = 0; i < n; ++i) {
    if (!free_[i]) continue;
    agg[i] = na;
    free_[i] = 0;
    nb(i, [&](int32_t j) { if (free_[j]) { agg[j] = na; free_[j] = 0; } });
    ++na;
  }
  return na;
}

// P pattern: row i -> sorted unique aggregates of the (non-excluded, in-block) neighbours of i
static HostCsr prolongator_pattern(const HostCsr& A, int32_t n, int32_t na, const std::vector<int32_t>& agg, bool smoothed) {
  HostCsr P;
  P.n_rows = n;
  P.n_cols = na;
  P.rowptr.assign(n + 1, 0);
  std::vector<int32_t> tmp;
  for (int32_t i = 0; i < n; ++i) {
    tmp.clear();
    if (agg[i] >= 0) {
      if (smoothed) {
        for (int32_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
          const int32_t j = A.col[k];
          if (j < n && agg[j] >= 0) tmp.push_back(agg[j]);
        }
        std::sort(tmp.begin(), tmp.end());
        tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      } else {
        tmp.push_back(agg[i]);
      }
    }
    P.col.insert(P.col.end(), tmp.begin(), tmp.end());
    P.rowptr[i + 1] = (int32_t)P.col.size();
  }
  return P;
}

// pattern of A(:, 0:B.n_rows) * B
static HostCsr product_pattern(const HostCsr& A, const HostCsr& B) {
  HostCsr C;
  C.n_rows = A.n_rows;
  C.n_cols = B.n_cols;
  C.rowptr.assign(A.n_rows + 1, 0);
  std::vector<int32_t> mark(B.n_cols, -1), tmp;
  for (int64_t i = 0; i < A.n_rows; ++i) {
    tmp.clear();
    for (int32_t k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
      const int32_t j = A.col[k];
      if (j >= B.n_rows) continue;
      for (int32_t l = B.rowptr[j]; l < B.rowptr[j + 1]; ++l) {
        const int32_t c = B.col[l];
        if (mark[c] != (int32_t)i) { mark[c] = (int32_t)i; tmp.push_back(c); }
      }
    }
    std::sort(tmp.begin(), tmp.end());
    C.col.insert(C.col.end(), tmp.begin(), tmp.end());
    if (C.col.size() > 2000000000ULL) throw Error(SHAKTI_ERR_INVALID, "AMG product pattern too large");
    C.rowptr[i + 1] = (int32_t)C.col.size();
  }
  return C;
}

// ------------------------------------------------------------------ device kernels

// dinv = 1/diag
__global__ void amg_dinv_kernel(int32_t n, const int32_t* __restrict__ diag_pos, const double* __restrict__ val,
                                double* __restrict__ dinv) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double d = val[diag_pos[i]];
  dinv[i] = d != 0.0 ? 1.0 / d : 0.0;
}

// P = (I - omega D^-1 A) T : every A entry (i,j) adds (delta_ij - omega a_ij / a_ii) to
// P[i, agg(j)]; pmap holds the position of agg(j) inside P's row i (255: dropped).
__global__ void __launch_bounds__(256)
amg_prolongator_kernel(SellView A, const uint8_t* __restrict__ pmap, const int32_t* __restrict__ diag_pos,
                       const double* __restrict__ dinv, double omega, const int32_t* __restrict__ Pslice,
                       double* __restrict__ Pval) {
  const int32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= A.n_rows) return;
  const int32_t slice = row >> 5, lane = row & 31;
  const int32_t base = A.slice_ptr[slice];
  const int32_t w = (A.slice_ptr[slice + 1] - base) >> 5;
  const int32_t pbase = Pslice[slice] + lane;
  const int32_t dpos = diag_pos[row];
  const double s = omega * dinv[row];
  for (int k = 0; k < w; ++k) {
    const int32_t p = base + 32 * k + lane;
    const uint8_t t = pmap[p];
    if (t == 255) continue;
    const double v = (p == dpos ? 1.0 : 0.0) - s * A.val[p];
    Pval[pbase + 32 * (int32_t)t] += v;
  }
}

// out[p] = src[map[p]] (map < 0 -> 0)
__global__ void amg_gather_vals_kernel(int64_t n, const int32_t* __restrict__ map, const double* __restrict__ src,
                                       double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t m = map[i];
    out[i] = m >= 0 ? src[m] : 0.0;
  }
}

// C = A(:, 0:B.n_rows) * B with a known pattern of C (values zeroed beforehand): one thread
// per row, positions found by binary search in C's sorted row.
__global__ void __launch_bounds__(128)
amg_spgemm_kernel(SellView A, const int32_t* __restrict__ Alen, SellView B, const int32_t* __restrict__ Blen,
                  int32_t Bn_rows, const int32_t* __restrict__ Cslice, const int32_t* __restrict__ Ccol,
                  const int32_t* __restrict__ Clen, double* __restrict__ Cval) {
  const int32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= A.n_rows) return;
  const int32_t lane = row & 31;
  const int32_t abase = A.slice_ptr[row >> 5] + lane;
  const int32_t cbase = Cslice[row >> 5] + lane;
  const int32_t clen = Clen[row];
  const int32_t alen = Alen[row];
  for (int ka = 0; ka < alen; ++ka) {
    const int32_t j = A.col[abase + 32 * ka];
    if (j >= Bn_rows) continue;
    const double a = A.val[abase + 32 * ka];
    if (a == 0.0) continue;
    const int32_t bbase = B.slice_ptr[j >> 5] + (j & 31);
    const int32_t blen = Blen[j];
    for (int kb = 0; kb < blen; ++kb) {
      const int32_t c = B.col[bbase + 32 * kb];
      const double b = B.val[bbase + 32 * kb];
      int lo = 0, hi = clen - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (Ccol[cbase + 32 * mid] < c) lo = mid + 1; else hi = mid;
      }
      Cval[cbase + 32 * lo] += a * b;
    }
  }
}

// dense coarse operator: D[i*ld + j] (ld = 2n), right half = identity
__global__ void amg_dense_fill_kernel(SellView A, const int32_t* __restrict__ Alen, int32_t n, double* __restrict__ D) {
  const int32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const int32_t base = A.slice_ptr[row >> 5] + (row & 31);
  for (int k = 0; k < Alen[row]; ++k) {
    const int32_t j = A.col[base + 32 * k];
    if (j < n) D[(size_t)row * 2 * n + j] += A.val[base + 32 * k];
  }
  D[(size_t)row * 2 * n + n + row] = 1.0;
}

// Gauss-Jordan with partial pivoting on [A | I] -> [I | A^-1], one thread block.
__global__ void __launch_bounds__(1024)
amg_dense_invert_kernel(int32_t n, double* __restrict__ D, int* __restrict__ info) {
  extern __shared__ double fcol[];   // n factors
  __shared__ int piv;
  __shared__ double pval;
  const int ld = 2 * n;
  for (int k = 0; k < n; ++k) {
    if (threadIdx.x < 32) {   // warp-parallel pivot search (first maximum wins)
      int p = k;
      double best = -1.0;
      for (int i = k + (int)threadIdx.x; i < n; i += 32) {
        const double v = fabs(D[(size_t)i * ld + k]);
        if (v > best) { best = v; p = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int op = __shfl_xor_sync(0xffffffffu, p, o);
        if (ob > best || (ob == best && op < p)) { best = ob; p = op; }
      }
      if (threadIdx.x == 0) {
        piv = p;
        pval = D[(size_t)p * ld + k];
        if (!(best > 0.0)) *info = k + 1;
      }
    }
    __syncthreads();
    const int p = piv;
    const double inv = pval != 0.0 ? 1.0 / pval : 0.0;
    // swap rows k,p and scale the pivot row
    for (int j = threadIdx.x; j < ld; j += blockDim.x) {
      const double a = D[(size_t)p * ld + j], b = D[(size_t)k * ld + j];
      D[(size_t)p * ld + j] = b;
      D[(size_t)k * ld + j] = a * inv;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) fcol[i] = (i == k) ? 0.0 : D[(size_t)i * ld + k];
    __syncthreads();
    for (int idx = threadIdx.x; idx < n * ld; idx += blockDim.x) {
      const int i = idx / ld, j = idx - i * ld;
      const double f = fcol[i];
      if (f != 0.0) D[(size_t)i * ld + j] -= f * D[(size_t)k * ld + j];
    }
    __syncthreads();
  }
}

// x = Ainv b, Ainv = right half of D; one warp per row
__global__ void amg_dense_apply_kernel(int32_t n, const double* __restrict__ D, const double* __restrict__ b,
                                       double* __restrict__ x) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const double* row = D + (size_t)warp * 2 * n + n;
  double acc = 0.0;
  for (int j = lane; j < n; j += 32) acc += row[j] * b[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) x[warp] = acc;
}

// ------------------------------------------------------------------ hierarchy
struct AmgLevel {
  int32_t n = 0, n_cols = 0, n_coarse = 0;
  int64_t nnz = 0;
  DevSell A;                  // levels >= 1 (level 0 uses the caller's matrix)
  DevBuf<int32_t> diag_pos;   // levels >= 1
  DevBuf<double> dinv;
  DevSell P, R, AP;
  DevBuf<uint8_t> pmap;
  DevBuf<int32_t> tmap;
  DevBuf<double> x, x2, b, r;
  bool last = false;
};

struct Amg::Impl {
  AmgOptions opt;
  cudaStream_t s = 0;
  std::vector<std::unique_ptr<AmgLevel>> lv;
  DevBuf<double> dense;
  DevBuf<int> info;
  bool dense_coarse = false;
  double op_complexity = 0.0;
};

Amg::Amg() : p_(new Impl()) {}
Amg::~Amg() = default;
int Amg::levels() const { return (int)p_->lv.size(); }
double Amg::operator_complexity() const { return p_->op_complexity; }

static std::vector<int32_t> diag_positions(const HostCsr& A, const HostSell& S, int32_t n) {
  std::vector<int32_t> d(n, 0);
  for (int32_t r = 0; r < n; ++r) {
    const int32_t* b = A.col.data() + A.rowptr[r];
    const int32_t* e = A.col.data() + A.rowptr[r + 1];
    const int32_t* p = std::lower_bound(b, e, r);
    if (p == e || *p != r) throw Error(SHAKTI_ERR_INVALID, "AMG: matrix row without diagonal entry");
    d[r] = (int32_t)S.pos(r, (int)(p - b));
  }
  return d;
}

void Amg::setup(const HostCsr& A0, const HostSell& S0, const std::vector<uint8_t>& exclude, const AmgOptions& opt,
                int sm_count, cudaStream_t s) {
  (void)sm_count;
  Impl& I = *p_;
  I.opt = opt;
  I.s = s;
  I.lv.clear();
  refreshes_ = 0;
  HostCsr Acur;            // levels >= 1 own their pattern
  HostSell Scur;
  const HostCsr* A = &A0;
  const HostSell* S = &S0;
  std::vector<uint8_t> excl = exclude;
  double nnz_sum = 0.0;
  const double nnz0 = (double)A0.nnz();
  for (int l = 0;; ++l) {
    std::unique_ptr<AmgLevel> L(new AmgLevel());
    L->n = (int32_t)A->n_rows;
    L->n_cols = (int32_t)A->n_cols;
    L->nnz = A->nnz();
    nnz_sum += (double)L->nnz;
    if (l > 0) {
      L->A.upload_pattern(*S, A->nnz());
      L->diag_pos.upload(diag_positions(*A, *S, L->n));
    }
    L->dinv.alloc_zero(std::max(L->n, 1), s);
    L->x.alloc_zero(std::max(L->n_cols, 1), s);
    L->x2.alloc_zero(std::max(L->n_cols, 1), s);
    L->r.alloc_zero(std::max(L->n, 1), s);
    if (l > 0) L->b.alloc_zero(std::max(L->n, 1), s);
    const bool stop = (L->n <= opt.coarse_size) || (l + 1 >= opt.max_levels);
    std::vector<int32_t> agg;
    int32_t na = 0;
    if (!stop) {
      na = aggregate(*A, L->n, excl, agg);
      if (na <= 0 || na > 0.85 * L->n) na = 0;  // coarsening stalled
    }
    if (stop || na == 0) {
      L->last = true;
      I.lv.push_back(std::move(L));
      break;
    }
    L->n_coarse = na;
    HostCsr P = prolongator_pattern(*A, L->n, na, agg, opt.prolong_omega != 0.0);
    HostSell PS = sell_from_csr(P);
    // pmap in A's SELL layout
    {
      std::vector<uint8_t> pm(S->padded(), 255);
      for (int32_t i = 0; i < L->n; ++i) {
        const int32_t* pb = P.col.data() + P.rowptr[i];
        const int32_t* pe = P.col.data() + P.rowptr[i + 1];
        if (pb == pe) continue;
        for (int32_t k = A->rowptr[i]; k < A->rowptr[i + 1]; ++k) {
          const int32_t j = A->col[k];
          if (j >= L->n || agg[j] < 0) continue;
          if (opt.prolong_omega == 0.0 && j != i) continue;
          const int32_t* q = std::lower_bound(pb, pe, agg[j]);
          const int64_t t = q - pb;
          if (t > 254) throw Error(SHAKTI_ERR_INVALID, "AMG: prolongator row too long");
          pm[S->pos(i, k - A->rowptr[i])] = (uint8_t)t;
        }
      }
      L->pmap.upload(pm);
    }
    L->P.upload_pattern(PS, P.nnz());
    std::vector<int32_t> tentry;
    HostCsr R = csr_transpose(P, &tentry);
    HostSell RS = sell_from_csr(R);
    {
      std::vector<int32_t> ppos = sell_positions(P, PS);
      std::vector<int32_t> tm(RS.padded(), -1);
      for (int64_t r = 0; r < R.n_rows; ++r)
        for (int32_t k = R.rowptr[r]; k < R.rowptr[r + 1]; ++k) tm[RS.pos(r, k - R.rowptr[r])] = ppos[tentry[k]];
      L->tmap.upload(tm);
    }
    L->R.upload_pattern(RS, R.nnz());
    HostCsr AP = product_pattern(*A, P);
    HostSell APS = sell_from_csr(AP);
    L->AP.upload_pattern(APS, AP.nnz());
    HostCsr Ac = product_pattern(R, AP);
    Ac.n_cols = na;
    I.lv.push_back(std::move(L));
    Acur = std::move(Ac);
    Scur = sell_from_csr(Acur);
    A = &Acur;
    S = &Scur;
    excl.clear();
  }
  AmgLevel& last = *I.lv.back();
  I.dense_coarse = last.n <= 512 && I.lv.size() > 1 ? true : (last.n <= 512);
  if (I.dense_coarse) {
    I.dense.alloc_zero((size_t)2 * last.n * last.n, s);
    I.info.alloc_zero(1, s);
  }
  I.op_complexity = nnz0 > 0 ? nnz_sum / nnz0 : 0.0;
  SHAKTI_CUDA(cudaStreamSynchronize(s));
}

static void spgemm(const DevSell& A, const DevSell& B, DevSell& C, cudaStream_t s) {
  SHAKTI_CUDA(cudaMemsetAsync(C.val.p, 0, sizeof(double) * C.padded, s));
  if (A.n_rows == 0) return;
  SHAKTI_LAUNCH(amg_spgemm_kernel, div_up(A.n_rows, 128), 128, 0, s, view(A), A.rowlen.p, view(B), B.rowlen.p,
                B.n_rows, C.slice_ptr.p, C.col.p, C.rowlen.p, C.val.p);
}

void Amg::refresh(const DevSell& Afine, const int32_t* fine_diag_pos) {
  Impl& I = *p_;
  cudaStream_t s = I.s;
  for (size_t l = 0; l < I.lv.size(); ++l) {
    AmgLevel& L = *I.lv[l];
    const DevSell& A = (l == 0) ? Afine : L.A;
    const int32_t* dpos = (l == 0) ? fine_diag_pos : L.diag_pos.p;
    if (L.n > 0) SHAKTI_LAUNCH(amg_dinv_kernel, div_up(L.n, 256), 256, 0, s, L.n, dpos, A.val.p, L.dinv.p);
    if (L.last) break;
    SHAKTI_CUDA(cudaMemsetAsync(L.P.val.p, 0, sizeof(double) * L.P.padded, s));
    SHAKTI_LAUNCH(amg_prolongator_kernel, div_up(L.n, 256), 256, 0, s, view(A), L.pmap.p, dpos, L.dinv.p,
                  I.opt.prolong_omega, L.P.slice_ptr.p, L.P.val.p);
    SHAKTI_LAUNCH(amg_gather_vals_kernel, (int)std::min<int64_t>(148 * 8, std::max<int64_t>(1, (L.R.padded + 255) / 256)), 256, 0, s,
                  L.R.padded, L.tmap.p, L.P.val.p, L.R.val.p);
    // A restricted to its square block times P, then R * (AP)
    {
      DevSell& AP = L.AP;
      SHAKTI_CUDA(cudaMemsetAsync(AP.val.p, 0, sizeof(double) * AP.padded, s));
      SHAKTI_LAUNCH(amg_spgemm_kernel, div_up(L.n, 128), 128, 0, s, view(A), A.rowlen.p, view(L.P), L.P.rowlen.p,
                    L.P.n_rows, AP.slice_ptr.p, AP.col.p, AP.rowlen.p, AP.val.p);
    }
    spgemm(L.R, L.AP, I.lv[l + 1]->A, s);
  }
  if (I.dense_coarse) {
    AmgLevel& L = *I.lv.back();
    const DevSell& A = (I.lv.size() == 1) ? Afine : L.A;
    const int n = L.n;
    if (n > 0) {
      SHAKTI_CUDA(cudaMemsetAsync(I.dense.p, 0, sizeof(double) * 2 * (size_t)n * n, s));
      SHAKTI_CUDA(cudaMemsetAsync(I.info.p, 0, sizeof(int), s));
      SHAKTI_LAUNCH(amg_dense_fill_kernel, div_up(n, 128), 128, 0, s, view(A), A.rowlen.p, n, I.dense.p);
      SHAKTI_LAUNCH(amg_dense_invert_kernel, 1, 1024, n * sizeof(double), s, n, I.dense.p, I.info.p);
    }
  }
  ++refreshes_;
}

void Amg::apply(const DevSell& Afine, const double* rin, double* z) {
  Impl& I = *p_;
  cudaStream_t s = I.s;
  const double om = I.opt.smoother_omega;
  const int nl = (int)I.lv.size();
  // downward sweep
  for (int l = 0; l < nl; ++l) {
    AmgLevel& L = *I.lv[l];
    const DevSell& A = (l == 0) ? Afine : L.A;
    const double* b = (l == 0) ? rin : L.b.p;
    if (L.n == 0) continue;
    if (L.last) {
      if (I.dense_coarse) {
        SHAKTI_LAUNCH(amg_dense_apply_kernel, div_up((int64_t)L.n * 32, 128), 128, 0, s, L.n, I.dense.p, b, L.x.p);
      } else {
        // no dense solve available: a few damped-Jacobi sweeps
        launch_pointwise_mul(L.n, L.dinv.p, b, om, L.x.p, s);
        for (int k = 0; k < 8; ++k) {
          launch_jacobi(view(A), L.dinv.p, b, L.x.p, L.x2.p, om, s);
          std::swap(L.x.p, L.x2.p);
        }
      }
      break;
    }
    // pre-smoothing from a zero guess: first sweep is a scaling
    launch_pointwise_mul(L.n, L.dinv.p, b, om, L.x.p, s);
    for (int k = 1; k < I.opt.presmooth; ++k) {
      launch_jacobi(view(A), L.dinv.p, b, L.x.p, L.x2.p, om, s);
      std::swap(L.x.p, L.x2.p);
    }
    if (I.opt.presmooth == 0) launch_fill(L.n, 0.0, L.x.p, s);
    launch_residual(view(A), L.x.p, b, L.r.p, s);
    launch_spmv(view(L.R), L.r.p, I.lv[l + 1]->b.p, s);
  }
  // upward sweep
  for (int l = nl - 2; l >= 0; --l) {
    AmgLevel& L = *I.lv[l];
    const DevSell& A = (l == 0) ? Afine : L.A;
    const double* b = (l == 0) ? rin : L.b.p;
    if (L.n == 0) continue;
    launch_spmv_add(view(L.P), I.lv[l + 1]->x.p, L.x.p, s);
    for (int k = 0; k < I.opt.postsmooth; ++k) {
      launch_jacobi(view(A), L.dinv.p, b, L.x.p, L.x2.p, om, s);
      std::swap(L.x.p, L.x2.p);
    }
  }
  AmgLevel& L0 = *I.lv[0];
  if (L0.n) SHAKTI_CUDA(cudaMemcpyAsync(z, L0.x.p, sizeof(double) * L0.n, cudaMemcpyDeviceToDevice, s));
}

}  // namespace shakti
