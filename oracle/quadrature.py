"""Quadrature tables on the reference triangle {x,y>=0, x+y<=1} (weights sum to 1/2).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED.

The reference never names a rule: FFCx picks one from UFL's estimated degree (7 for both the
residual ``solvers.py:45`` and its Gateaux derivative ``solvers.py:51``) and asks Basix for
its default scheme of that degree.  Basix' default for a triangle of degree <= 30 is a
Xiao-Gimbutas table (15 points for degree 7) whose digits are not obtainable offline, so the
table is a *data input* everywhere; the stand-in used here is the collapsed Gauss-Jacobi rule
of the same degree (what Basix itself uses above degree 30): (m+2)//2 points per direction.
Every polynomial part of the forms (true degree <= 5) is rule independent; only the
transmissivity integral  int_T K(b,|q|) dx  depends on the table.
"""
import numpy as np
from scipy.special import roots_jacobi, roots_legendre


def gauss_jacobi_triangle(degree):
    """Collapsed (Duffy) Gauss-Jacobi rule exact for total degree ``degree``.

    Returns (pts (n,2), wts (n,)) with sum(wts) == 1/2.
    """
    m = (degree + 2) // 2
    xu, wu = roots_jacobi(m, 1.0, 0.0)       # weight (1-x) on [-1,1]
    xv, wv = roots_legendre(m)
    u = 0.5 * (1.0 + xu)
    v = 0.5 * (1.0 + xv)
    pts = np.empty((m * m, 2))
    wts = np.empty(m * m)
    k = 0
    for i in range(m):
        for j in range(m):
            pts[k, 0] = u[i]
            pts[k, 1] = v[j] * (1.0 - u[i])
            wts[k] = 0.25 * wu[i] * 0.5 * wv[j]
            k += 1
    return pts, wts


def radon7():
    """Radon's 7-point degree-5 rule (closed form), weights normalised to area 1/2."""
    s15 = np.sqrt(15.0)
    a1 = (6.0 - s15) / 21.0
    a2 = (6.0 + s15) / 21.0
    w0 = 9.0 / 40.0
    w1 = (155.0 - s15) / 1200.0
    w2 = (155.0 + s15) / 1200.0
    bary = [(1 / 3, 1 / 3, 1 / 3)]
    wts = [w0]
    for a, w in ((a1, w1), (a2, w2)):
        c = 1.0 - 2.0 * a
        bary += [(c, a, a), (a, c, a), (a, a, c)]
        wts += [w, w, w]
    bary = np.array(bary)
    # reference coordinates (x,y) = (lambda1, lambda2)
    return bary[:, 1:3].copy(), 0.5 * np.array(wts)


def check_degree(pts, wts, degree=7, tol=1e-14):
    """Validate a user-supplied table (e.g. the one a FEniCSx golden dump records from Basix): weights sum to
    the area 1/2, all points inside the reference triangle, every monomial x^a y^b with a + b <= degree
    integrated exactly (exact value a! b! / (a + b + 2)!).  Returns the largest absolute moment error;
    raises ValueError when the table is not a rule of that degree."""
    from math import factorial
    pts, wts = np.asarray(pts, dtype=np.float64), np.asarray(wts, dtype=np.float64)
    if pts.ndim != 2 or pts.shape[1] != 2 or wts.shape != (pts.shape[0],):
        raise ValueError("quadrature table must be pts (n,2), wts (n,)")
    if np.any(pts < -1e-14) or np.any(pts.sum(axis=1) > 1 + 1e-14):
        raise ValueError("quadrature point outside the reference triangle")
    worst = 0.0
    for a in range(degree + 1):
        for b in range(degree + 1 - a):
            exact = factorial(a) * factorial(b) / factorial(a + b + 2)
            err = abs(float((wts * pts[:, 0] ** a * pts[:, 1] ** b).sum()) - exact)
            worst = max(worst, err)
            if err > tol + 1e-13 * exact:
                raise ValueError(f"table does not integrate x^{a} y^{b} exactly (error {err:.2e}): not of degree {degree}")
    return worst


def default_table():
    """Degree-7 table used when the caller supplies none (stand-in for Basix XG-7)."""
    return gauss_jacobi_triangle(7)
