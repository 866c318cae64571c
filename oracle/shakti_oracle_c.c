/* C restatement of the element kernels of the SHAKTI transient path -- TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py: checker and timed CPU baseline, never part of the product path).
 * PARITY UNPINNED like the rest of oracle/: the reference (agstub/shakti-fenics) ships no golden vectors
 * and its arithmetic lives in FFCx-generated C that cannot be produced here.
 *
 * Why it exists: in the reference the element work is compiled C (FFCx `tabulate_tensor`, called from the
 * DOLFINx assembler once per cell), spread over the MPI ranks.  The numpy oracle evaluates the same
 * formulas ~40x slower than compiled code, which would make the CPU baseline of bench.py an unfairly slow
 * stand-in.  This file is the same per-quadrature-point algorithm as ShaktiOracle.element_FJ / kbar /
 * the nodal-update expressions (no closed-form splitting, one rule for every term), in plain C with an
 * OpenMP loop over cells, so that the baseline's assembly runs at compiled speed on all host threads.
 * tests/test_oracle.py holds it to the numpy oracle at 1e-13.
 *
 * Reference locations (paths relative to /root/reference):
 *   weak form F ........................ source/solvers.py:35-45
 *   Jacobian dF/dN ..................... source/solvers.py:51 (ufl.derivative inside NonlinearProblem)
 *   Head / WaterFlux / Reynolds ........ source/constitutive.py:6-20
 *   Melt / Closure ..................... source/constitutive.py:22-31
 *   constants .......................... source/params.py:4-11
 *   q / melt_n / b expressions ......... source/solvers.py:143,162,165
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared; no fast-math: the order of operations is the source's)
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* params[8] = g, rho_i, rho_w, nu, Lh, omega, n, A   (source/params.py:4-11, same order as oracle Params) */
enum { P_G = 0, P_RHOI, P_RHOW, P_NU, P_LH, P_OMEGA, P_N, P_A };

int shakti_oracle_c_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static inline double pow_abs(double x, double e) {
  /* abs(x)**e; integer exponents 0..3 as numpy evaluates them for the reference's n = 3 */
  const double a = fabs(x);
  if (e == 2.0) return a * a;
  if (e == 1.0) return a;
  if (e == 0.0) return 1.0;
  return pow(a, e);
}

typedef struct {
  double gx[3], gy[3], detabs;
} geo_t;

static inline geo_t geometry(const double* xy, const int32_t* c) {
  /* constant P1 gradients: rows of the inverse Jacobian (ShaktiOracle._geometry) */
  geo_t g;
  const double x0 = xy[2 * c[0]], y0 = xy[2 * c[0] + 1];
  const double d1x = xy[2 * c[1]] - x0, d1y = xy[2 * c[1] + 1] - y0;
  const double d2x = xy[2 * c[2]] - x0, d2y = xy[2 * c[2] + 1] - y0;
  const double det = d1x * d2y - d2x * d1y;
  g.gx[1] = d2y / det;
  g.gy[1] = -d2x / det;
  g.gx[2] = -d1y / det;
  g.gy[2] = d1x / det;
  g.gx[0] = -g.gx[1] - g.gx[2];
  g.gy[0] = -g.gy[1] - g.gy[2];
  g.detabs = fabs(det);
  return g;
}

/* Element residual Fe (ne,3) and Jacobian Je (ne,3,3; may be NULL) -- ShaktiOracle.element_FJ.
 * q is (nv,2) interleaved as in the reference's blocked vector space. */
void shakti_oracle_c_element_FJ(int64_t ne, const int32_t* cells, const double* xy, const double* N, const double* N_n,
                                const double* b, const double* q, const double* G, const double* melt_n,
                                const double* storage, const double* inputs, const double* z_b, const double* z_s,
                                const double* params, int nq, const double* qpts, const double* qwts, double dt,
                                double* Fe, double* Je) {
  const double g = params[P_G], rho_i = params[P_RHOI], rho_w = params[P_RHOW], nu = params[P_NU];
  const double Lh = params[P_LH], omega = params[P_OMEGA], n = params[P_N], A = params[P_A];
  const double cm = 1.0 / rho_i - 1.0 / rho_w;
  const double rwg = rho_w * g;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < ne; ++e) {
    const int32_t* c = cells + 3 * e;
    const geo_t ge = geometry(xy, c);
    double h[3], Nc[3], Nn[3], bc[3], qx[3], qy[3], Gc[3], mc[3], sc[3], ic[3];
    for (int a = 0; a < 3; ++a) {
      const int32_t v = c[a];
      Nc[a] = N[v];
      h[a] = z_b[v] + (rho_i / rho_w) * (z_s[v] - z_b[v]) - Nc[a] / rwg; /* Head, constitutive.py:6-9 */
      Nn[a] = N_n[v]; bc[a] = b[v]; qx[a] = q[2 * v]; qy[a] = q[2 * v + 1];
      Gc[a] = G[v]; mc[a] = melt_n[v]; sc[a] = storage[v]; ic[a] = inputs[v];
    }
    double ghx = 0, ghy = 0, gbx = 0, gby = 0, gmx = 0, gmy = 0;
    for (int a = 0; a < 3; ++a) {
      ghx += h[a] * ge.gx[a];  ghy += h[a] * ge.gy[a];
      gbx += bc[a] * ge.gx[a]; gby += bc[a] * ge.gy[a];
      gmx += mc[a] * ge.gx[a]; gmy += mc[a] * ge.gy[a];
    }
    const double gb2 = gbx * gbx + gby * gby, gmgb = gmx * gbx + gmy * gby;
    double F[3] = {0, 0, 0}, J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int k = 0; k < nq; ++k) {
      const double xi = qpts[2 * k], eta = qpts[2 * k + 1];
      const double lam[3] = {1.0 - xi - eta, xi, eta};
      const double wd = qwts[k] * ge.detabs;
      double bq = 0, Nq = 0, Nnq = 0, Gq = 0, mq = 0, sq = 0, iq = 0, qqx = 0, qqy = 0;
      for (int a = 0; a < 3; ++a) {
        bq += bc[a] * lam[a]; Nq += Nc[a] * lam[a]; Nnq += Nn[a] * lam[a]; Gq += Gc[a] * lam[a];
        mq += mc[a] * lam[a]; sq += sc[a] * lam[a]; iq += ic[a] * lam[a];
        qqx += qx[a] * lam[a]; qqy += qy[a] * lam[a];
      }
      const double Re = sqrt(qqx * qqx + qqy * qqy) / nu;                          /* constitutive.py:18-20 */
      const double ab = fabs(bq);
      const double K = ab * ab * ab * g / (12.0 * nu * (1.0 + omega * Re));         /* constitutive.py:11-16 */
      const double qgh = qqx * ghx + qqy * ghy;
      const double m0 = (Gq - rwg * qgh) / Lh;                                      /* constitutive.py:25 */
      const double mdiff = (gb2 * mq + bq * gmgb) / (1.0 + gb2);                    /* constitutive.py:26 */
      const double clos = A * bq * Nq * pow_abs(Nq, n - 1.0);                       /* constitutive.py:31 */
      const double lake = sq * (1.0 / (rwg * dt)) * (Nq - Nnq);                     /* solvers.py:42 */
      const double R = cm * (m0 + mdiff) - clos - lake - iq;
      for (int a = 0; a < 3; ++a)                                                   /* solvers.py:45 */
        F[a] += (wd * K) * (ghx * ge.gx[a] + ghy * ge.gy[a]) + (wd * R) * lam[a];
      if (Je) {
        const double sgn = Nq > 0 ? 1.0 : (Nq < 0 ? -1.0 : 0.0);
        const double dclos = A * bq * (pow_abs(Nq, n - 1.0) + Nq * (n - 1.0) * pow_abs(Nq, n - 2.0) * sgn);
        const double dreact = dclos + sq / (rwg * dt);
        const double kj = -(wd * K / rwg);
        for (int bb = 0; bb < 3; ++bb) {
          /* d m0 / dN [phi_b] = (q . grad phi_b) / Lh */
          const double dR_adv = (cm / Lh) * (qqx * ge.gx[bb] + qqy * ge.gy[bb]);
          const double dR = dR_adv - dreact * lam[bb];
          for (int a = 0; a < 3; ++a)
            J[a][bb] += kj * (ge.gx[a] * ge.gx[bb] + ge.gy[a] * ge.gy[bb]) + wd * lam[a] * dR;
        }
      }
    }
    for (int a = 0; a < 3; ++a) Fe[3 * e + a] = F[a];
    if (Je)
      for (int a = 0; a < 3; ++a)
        for (int bb = 0; bb < 3; ++bb) Je[9 * e + 3 * a + bb] = J[a][bb];
  }
}

/* |detJ| sum_k w_k K(b(xi_k), |q(xi_k)|) per cell -- ShaktiOracle.kbar (constitutive.py:11-20) */
void shakti_oracle_c_kbar(int64_t ne, const int32_t* cells, const double* xy, const double* b, const double* q,
                          const double* params, int nq, const double* qpts, const double* qwts, double* out) {
  const double g = params[P_G], nu = params[P_NU], omega = params[P_OMEGA];
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < ne; ++e) {
    const int32_t* c = cells + 3 * e;
    const geo_t ge = geometry(xy, c);
    double acc = 0;
    for (int k = 0; k < nq; ++k) {
      const double xi = qpts[2 * k], eta = qpts[2 * k + 1];
      const double lam[3] = {1.0 - xi - eta, xi, eta};
      double bq = 0, qqx = 0, qqy = 0;
      for (int a = 0; a < 3; ++a) { bq += b[c[a]] * lam[a]; qqx += q[2 * c[a]] * lam[a]; qqy += q[2 * c[a] + 1] * lam[a]; }
      const double Re = sqrt(qqx * qqx + qqy * qqy) / nu;
      const double ab = fabs(bq);
      acc += qwts[k] * (ab * ab * ab * g / (12.0 * nu * (1.0 + omega * Re)));
    }
    out[e] = acc * ge.detabs;
  }
}

/* The three Function.interpolate(Expression) updates, evaluated per vertex in its WINNING cell (highest cell
 * index containing the vertex; win_cell/win_loc as ShaktiOracle._winning_cells), in the reference's order
 * (solvers.py:186-197): q from the old q, melt_n from the new q and the old melt_n, b from the new q, the new
 * melt_n and the old b, then the clamp b >= b_min.  All three read N (the new one) through the head.
 * q_new (nv,2), melt_new (nv), b_new (nv) are outputs; vertices in no cell keep their values. */
void shakti_oracle_c_nodal_updates(int64_t nv, const int32_t* cells, const int64_t* win_cell, const int64_t* win_loc,
                                   const double* xy, const double* N, const double* b, const double* q, const double* G,
                                   const double* melt_n, const double* z_b, const double* z_s, const double* params,
                                   double dt, double b_min, double* q_new, double* melt_new, double* b_new) {
  const double g = params[P_G], rho_i = params[P_RHOI], rho_w = params[P_RHOW], nu = params[P_NU];
  const double Lh = params[P_LH], omega = params[P_OMEGA], n = params[P_N], A = params[P_A];
  const double rwg = rho_w * g;
  /* pass 1: q */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nv; ++i) {
    q_new[2 * i] = q[2 * i]; q_new[2 * i + 1] = q[2 * i + 1];
    if (win_cell[i] < 0) continue;
    const int32_t* c = cells + 3 * win_cell[i];
    const geo_t ge = geometry(xy, c);
    double ghx = 0, ghy = 0;
    for (int a = 0; a < 3; ++a) {
      const int32_t v = c[a];
      const double h = z_b[v] + (rho_i / rho_w) * (z_s[v] - z_b[v]) - N[v] / rwg;
      ghx += h * ge.gx[a]; ghy += h * ge.gy[a];
    }
    const double Re = sqrt(q[2 * i] * q[2 * i] + q[2 * i + 1] * q[2 * i + 1]) / nu;       /* constitutive.py:18-20 */
    const double ab = fabs(b[i]);
    const double K = ab * ab * ab * g / (12.0 * nu * (1.0 + omega * Re));                 /* constitutive.py:11-16 */
    q_new[2 * i] = -K * ghx; q_new[2 * i + 1] = -K * ghy;
  }
  /* pass 2: melt_n (new q at the vertex; old melt_n, b in the cell gradients) */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nv; ++i) {
    melt_new[i] = melt_n[i];
    if (win_cell[i] < 0) continue;
    const int32_t* c = cells + 3 * win_cell[i];
    const geo_t ge = geometry(xy, c);
    double ghx = 0, ghy = 0, gbx = 0, gby = 0, gmx = 0, gmy = 0;
    for (int a = 0; a < 3; ++a) {
      const int32_t v = c[a];
      const double h = z_b[v] + (rho_i / rho_w) * (z_s[v] - z_b[v]) - N[v] / rwg;
      ghx += h * ge.gx[a]; ghy += h * ge.gy[a];
      gbx += b[v] * ge.gx[a]; gby += b[v] * ge.gy[a];
      gmx += melt_n[v] * ge.gx[a]; gmy += melt_n[v] * ge.gy[a];
    }
    const double gb2 = gbx * gbx + gby * gby, gmgb = gmx * gbx + gmy * gby;
    const double m0 = (G[i] - rwg * (q_new[2 * i] * ghx + q_new[2 * i + 1] * ghy)) / Lh;  /* constitutive.py:25 */
    const double mdiff = (gb2 * melt_n[i] + b[i] * gmgb) / (1.0 + gb2);                   /* constitutive.py:26 */
    melt_new[i] = m0 + mdiff;
  }
  /* pass 3: b (new q, new melt_n; old b), then the clamp of solvers.py:196-197 */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nv; ++i) {
    b_new[i] = b[i];
    if (win_cell[i] >= 0) {
      const int32_t* c = cells + 3 * win_cell[i];
      const geo_t ge = geometry(xy, c);
      double ghx = 0, ghy = 0, gbx = 0, gby = 0, gmx = 0, gmy = 0;
      for (int a = 0; a < 3; ++a) {
        const int32_t v = c[a];
        const double h = z_b[v] + (rho_i / rho_w) * (z_s[v] - z_b[v]) - N[v] / rwg;
        ghx += h * ge.gx[a]; ghy += h * ge.gy[a];
        gbx += b[v] * ge.gx[a]; gby += b[v] * ge.gy[a];
        gmx += melt_new[v] * ge.gx[a]; gmy += melt_new[v] * ge.gy[a];
      }
      const double gb2 = gbx * gbx + gby * gby, gmgb = gmx * gbx + gmy * gby;
      const double m0 = (G[i] - rwg * (q_new[2 * i] * ghx + q_new[2 * i + 1] * ghy)) / Lh;
      const double mdiff = (gb2 * melt_new[i] + b[i] * gmgb) / (1.0 + gb2);
      const double melt = m0 + mdiff;
      const double clos = A * b[i] * N[i] * pow_abs(N[i], n - 1.0);                       /* constitutive.py:31 */
      b_new[i] = b[i] + dt * (melt / rho_i - clos);                                       /* solvers.py:162 */
    }
    if (b_new[i] < b_min) b_new[i] = b_min;                                               /* solvers.py:196-197 */
  }
  (void)win_loc;
}
