#!/usr/bin/env python3
"""CPU prototype (scipy) of the design choices of csrc/amg.cu — the evidence behind DESIGN.md §3/§4.

It builds a P1 diffusion + small reaction operator with a strongly varying coefficient on a jittered
triangle mesh (the structure of the SHAKTI Jacobian: -(div K grad) - c), orders it along a Z-curve like
the library, and measures right-preconditioned GMRES iteration counts for one V-cycle of

  * smoothed aggregation with strength-of-connection aggregation (theta 0.5^level), Jacobi or
    Chebyshev smoothing;
  * the DISTRIBUTED variant of the library: contiguous row ranges as "ranks", aggregates never cross
    ranks, restriction = full P^T ("galerkin"), P^T with the cross-rank entries dropped ("trunc", what
    the library does: no reverse communication), or the tentative T^T ("tent").

Typical output (n = 40 000):  1 part: 19 / 19 / 250+   8 parts: 19 / 20 / 300   (galerkin / trunc / tent)

Self-contained: numpy + scipy only (it does not use oracle/ nor the CUDA library).
"""
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def mesh(n, jitter=0.25, seed=1):
    rng = np.random.default_rng(seed)
    xs = np.linspace(0, 1, n + 1)
    X, Y = np.meshgrid(xs, xs)
    xy = np.stack([X.ravel(), Y.ravel()], 1)
    inner = np.zeros((n + 1, n + 1), bool)
    inner[1:-1, 1:-1] = True
    xy[inner.ravel()] += rng.uniform(-jitter, jitter, (int(inner.sum()), 2)) / n
    i, j = np.meshgrid(np.arange(n), np.arange(n))
    v00 = (j * (n + 1) + i).ravel()
    cells = np.concatenate([np.stack([v00, v00 + 1, v00 + n + 2], 1), np.stack([v00, v00 + n + 2, v00 + n + 1], 1)])
    return xy, cells


def operator(xy, cells):
    """-(stiffness with K = b^3, b varying over 2 decades) - small mass term, Dirichlet on x = 0, Z-curve order."""
    X = xy[cells]
    d1, d2 = X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]
    det = d1[:, 0] * d2[:, 1] - d2[:, 0] * d1[:, 1]
    g = np.empty((len(cells), 3, 2))
    g[:, 1, 0], g[:, 1, 1] = d2[:, 1] / det, -d2[:, 0] / det
    g[:, 2, 0], g[:, 2, 1] = -d1[:, 1] / det, d1[:, 0] / det
    g[:, 0] = -g[:, 1] - g[:, 2]
    c = X.mean(axis=1)
    K = (10.0 ** (np.sin(6 * c[:, 0]) * np.cos(5 * c[:, 1]))) ** 3
    Ke = -(K * np.abs(det) / 2)[:, None, None] * np.einsum("eak,ebk->eab", g, g)
    Me = -1e-3 * (np.abs(det) / 24)[:, None, None] * (np.ones((3, 3)) + np.eye(3))[None]
    rows, cols = np.repeat(cells, 3, axis=1).ravel(), np.tile(cells, (1, 3)).ravel()
    n = xy.shape[0]
    A = sp.csr_matrix(((Ke + Me).ravel(), (rows, cols)), shape=(n, n)).tolil()
    bc = np.nonzero(np.isclose(xy[:, 0], 0.0))[0]
    A = A.tocsr()
    keep = np.ones(n, bool)
    keep[bc] = False
    D = sp.diags(keep.astype(float))
    A = (D @ A @ D + sp.diags((~keep).astype(float))).tocsr()
    # Z-curve order
    q = np.minimum((xy * 1023).astype(np.int64), 1023)
    key = np.zeros(n, np.int64)
    for bit in range(10):
        key |= ((q[:, 0] >> bit) & 1) << (2 * bit) | ((q[:, 1] >> bit) & 1) << (2 * bit + 1)
    order = np.argsort(key, kind="stable")
    return A[order][:, order].tocsr(), (~keep)[order]


def strength(A, theta):
    if theta <= 0:
        return A
    C = A.tocoo()
    d = np.abs(A.diagonal())
    keep = (np.abs(C.data) >= theta * np.sqrt(d[C.row] * d[C.col])) | (C.row == C.col)
    return sp.csr_matrix((C.data[keep], (C.row[keep], C.col[keep])), shape=A.shape)


def aggregate(S, excl):
    """The greedy three-pass aggregation of csrc/amg.cu (aggregate())."""
    n, ip, ix = S.shape[0], S.indptr, S.indices
    agg = -np.ones(n, int)
    free = ~excl.copy()
    nbrs = lambda i: [j for j in ix[ip[i]:ip[i + 1]] if j != i and not excl[j]]
    na = 0
    for i in range(n):
        if free[i] and all(free[j] for j in nbrs(i)):
            for j in [i] + nbrs(i):
                agg[j], free[j] = na, False
            na += 1
    snap = agg.copy()
    size = np.bincount(snap[snap >= 0], minlength=na)
    for i in range(n):
        if free[i]:
            cand = [snap[j] for j in nbrs(i) if snap[j] >= 0]
            if cand:
                best = min(cand, key=lambda a: size[a])
                agg[i], free[i] = best, False
                size[best] += 1
    for i in range(n):
        if free[i]:
            agg[i], free[i] = na, False
            for j in nbrs(i):
                if free[j]:
                    agg[j], free[j] = na, False
            na += 1
    return agg, na


def hierarchy(A, excl, nparts=1, mode="trunc", theta=0.08, omega=0.67, coarse=128):
    levels, l = [], 0
    part = (np.arange(A.shape[0]) * nparts) // A.shape[0]
    while True:
        n, D = A.shape[0], A.diagonal()
        if n <= coarse or l >= 10:
            levels.append(dict(A=A, D=D, last=True))
            return levels
        S = strength(A, theta * 0.5 ** l).tocoo()
        same = part[S.row] == part[S.col]                       # aggregates never cross ranks
        S = sp.csr_matrix((S.data[same], (S.row[same], S.col[same])), shape=A.shape)
        agg, na = aggregate(S, excl)
        m = agg >= 0
        T = sp.csr_matrix((np.ones(m.sum()), (np.nonzero(m)[0], agg[m])), shape=(n, na))
        cpart = np.zeros(na, int)
        cpart[agg[m]] = part[m]
        P = (sp.diags((~excl).astype(float)) @ ((sp.eye(n) - omega * sp.diags(1.0 / D) @ A) @ T)).tocsr()
        if mode == "galerkin":
            R = P.T.tocsr()
        elif mode == "tent":
            R = T.T.tocsr()
        else:
            C = P.tocoo()
            k = part[C.row] == cpart[C.col]
            R = sp.csr_matrix((C.data[k], (C.col[k], C.row[k])), shape=(na, n))
        levels.append(dict(A=A, D=D, P=P, R=R, last=False))
        A, excl, part, l = (R @ A @ P).tocsr(), np.zeros(na, bool), cpart, l + 1


def smooth(L, b, x, kind, sweeps, zero):
    A, D = L["A"], L["D"]
    if kind == "jacobi":
        for k in range(sweeps):
            x = 0.67 * b / D if (zero and k == 0) else x + 0.67 * (b - A @ x) / D
        return x
    hi = L.setdefault("lmax", float(np.max(np.abs(A).sum(axis=1).A1 / np.abs(D))))      # Gershgorin bound
    lo = hi / 5.0
    theta, delta = (hi + lo) / 2, (hi - lo) / 2
    sigma = theta / delta
    rho = 1 / sigma
    r = b / D if zero else (b - A @ x) / D
    d = r / theta
    x = x + d if not zero else d
    for _ in range(1, sweeps):
        rho_n = 1 / (2 * sigma - rho)
        d = rho_n * rho * d + 2 * rho_n / delta * ((b - A @ x) / D)
        x, rho = x + d, rho_n
    return x


def vcycle(levels, l, b, kind="cheb", sweeps=2):
    L = levels[l]
    if L["last"]:
        return spla.spsolve(L["A"].tocsc(), b)
    x = smooth(L, b, None, kind, sweeps, True)
    x = x + L["P"] @ vcycle(levels, l + 1, L["R"] @ (b - L["A"] @ x), kind, sweeps)
    return smooth(L, b, x, kind, sweeps, False)


def gmres_iterations(A, b, M, rtol=1e-12, maxit=300):
    beta = np.linalg.norm(b)
    V, H, g, cs, sn = [b / beta], np.zeros((maxit + 1, maxit)), np.zeros(maxit + 1), [], []
    g[0] = beta
    for j in range(maxit):
        w = A @ M(V[j])
        for i in range(j + 1):
            H[i, j] = V[i] @ w
            w = w - H[i, j] * V[i]
        H[j + 1, j] = np.linalg.norm(w)
        V.append(w / H[j + 1, j])
        for i in range(j):
            H[i, j], H[i + 1, j] = cs[i] * H[i, j] + sn[i] * H[i + 1, j], -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
        d = np.hypot(H[j, j], H[j + 1, j])
        cs.append(H[j, j] / d)
        sn.append(H[j + 1, j] / d)
        g[j + 1], g[j] = -sn[j] * g[j], cs[j] * g[j]
        if abs(g[j + 1]) <= rtol * beta:
            return j + 1
    return maxit


def experiment(n=100, parts=(1, 8), modes=("galerkin", "trunc", "tent"), kind="cheb", maxit=300):
    xy, cells = mesh(n)
    A, excl = operator(xy, cells)
    rhs = np.where(excl, 0.0, np.random.default_rng(0).standard_normal(A.shape[0]))
    out = {}
    for p in parts:
        for mode in modes:
            lev = hierarchy(A, excl, p, mode)
            out[(p, mode)] = gmres_iterations(A, rhs, lambda r: vcycle(lev, 0, r, kind), maxit=maxit)
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    for (p, mode), its in experiment(n).items():
        print(f"ranks {p:2d}  restriction {mode:9s}  GMRES iterations {its}")
