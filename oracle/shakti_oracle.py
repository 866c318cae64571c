"""CPU restatement of the SHAKTI transient path of agstub/shakti-fenics.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED (no runnable reference, no
golden vectors in the reference tree).

What is restated, with the reference location each piece follows
(paths relative to /root/reference):

* constants .................. source/params.py:4-11
* Head/WaterFlux/Reynolds/Melt/Closure ... source/constitutive.py:6-31
* Dirichlet dofs + value ..... source/solvers.py:17-26
* weak form F ................ source/solvers.py:35-45
* Jacobian = dF/dN ........... source/solvers.py:51 (NonlinearProblem builds ufl.derivative)
* Newton ..................... source/solvers.py:52,179 (DOLFINx NewtonSolver defaults)
* nodal updates q, melt_n, b . source/solvers.py:143,162,165,186-197
* dt schedule, N_n <- N ...... source/solvers.py:81,174-176,228-229

Third-party semantics reproduced (marked [EXT] in SURVEY.md): one quadrature rule for every
term of the form (the table is an input); Dirichlet rows/cols zeroed with unit diagonal,
lifting with scale -1, F[bc] = N[bc] - g; sparse LU as the linear solve; ``Function.interpolate
(Expression)`` evaluates the expression in every cell at its three vertices from the OLD
coefficient values and then writes cell by cell, so the highest-index cell containing a vertex
wins.

Everything is written per quadrature point / per cell exactly as a generated form kernel would
do it (no closed-form splitting); the CUDA path splits the integrals differently and must
agree with this to 1e-12.
"""
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import quadrature


@dataclass
class Params:
    """source/params.py:4-11 (rho_i, rho_w and n are Python ints there)."""
    g: float = 9.81
    rho_i: float = 917
    rho_w: float = 1000
    nu: float = 1.787e-6
    Lh: float = 3.34e5
    omega: float = 1e-3
    n: float = 3
    A: float = 2.24e-24


def csr_pattern(n_vert, cells):
    """Sorted-unique-column CSR pattern of the P1 Jacobian (incl. diagonal), int32."""
    cells = np.asarray(cells, dtype=np.int64)
    rows = np.repeat(cells, 3, axis=1).ravel()
    cols = np.tile(cells, (1, 3)).ravel()
    key = np.unique(rows * n_vert + cols)
    r = (key // n_vert).astype(np.int64)
    c = (key % n_vert).astype(np.int32)
    rowptr = np.zeros(n_vert + 1, dtype=np.int64)
    np.add.at(rowptr, r + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    return rowptr, c


def boundary_facets(cells):
    """Edges that belong to exactly one cell, as an (m,2) vertex array."""
    cells = np.asarray(cells)
    e = np.concatenate([cells[:, [0, 1]], cells[:, [1, 2]], cells[:, [2, 0]]])
    es = np.sort(e, axis=1)
    key = es[:, 0].astype(np.int64) * (cells.max() + 1) + es[:, 1]
    uniq, idx, cnt = np.unique(key, return_index=True, return_counts=True)
    return es[idx[cnt == 1]]


def dirichlet_dofs(xy, cells, marker):
    """solvers.py:22-23: boundary facets whose vertices ALL satisfy ``marker(x)`` (x is (3,n),
    as DOLFINx passes it), then the dofs of those facets; sorted unique."""
    f = boundary_facets(cells)
    x3 = np.zeros((3, xy.shape[0]))
    x3[0], x3[1] = xy[:, 0], xy[:, 1]
    m = np.asarray(marker(x3), dtype=bool)
    keep = m[f[:, 0]] & m[f[:, 1]]
    return np.unique(f[keep].ravel()).astype(np.int32)


class ShaktiOracle:
    def __init__(self, xy, cells, params=None, quad=None, newton_r0="initial_residual", backend="numpy",
                 permc_spec="COLAMD"):
        """backend: "numpy" (the restatement every parity test uses) or "c" (oracle/shakti_oracle_c.c: the same
        per-quadrature-point formulas compiled, OpenMP over cells -- what bench.py times as the CPU baseline, since
        the reference's element kernels are compiled C too).  permc_spec: SuperLU column ordering of the LU solve
        ("MMD_AT_PLUS_A" halves the fill on these symmetric-pattern matrices; the reference's PETSc LU packages
        order with nested dissection / AMD as well)."""
        if backend not in ("numpy", "c"):
            raise ValueError(backend)
        self.backend, self.permc_spec = backend, permc_spec
        self.xy = np.ascontiguousarray(xy, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.nv = self.xy.shape[0]
        self.ne = self.cells.shape[0]
        self.p = params or Params()
        self.qpts, self.qwts = quad if quad is not None else quadrature.default_table()
        nv = self.nv
        z = lambda: np.zeros(nv)
        # model_setup.py:44-53 fields; solvers.py:129-156 state
        self.z_b, self.z_s, self.G, self.inputs, self.storage = z(), z(), z(), z(), z()
        self.b, self.N, self.N_n, self.melt_n = z(), z(), z(), z()
        self.q = np.zeros((nv, 2))
        self.bc_dofs = np.zeros(0, dtype=np.int32)
        self.N_bdry = 0.0
        self.b_min = 1.0e-5                       # model_setup.py:53
        # Newton (DOLFINx NewtonSolver defaults)
        self.rtol, self.atol, self.max_it = 1e-9, 1e-10, 50
        self.newton_r0 = newton_r0
        self._residual0 = 0.0
        self.relaxation = 1.0                     # NewtonSolver.relaxation_parameter default
        self.line_search, self.backtracks = 0, 0  # extension (off = reference behaviour), see newton()
        self.rowptr, self.col = csr_pattern(nv, self.cells)
        self._geometry()
        self._slots()
        self._winning_cells()

    # ---------------------------------------------------------------- set-up
    def _geometry(self):
        X = self.xy[self.cells]                    # (ne,3,2)
        d1 = X[:, 1] - X[:, 0]
        d2 = X[:, 2] - X[:, 0]
        det = d1[:, 0] * d2[:, 1] - d2[:, 0] * d1[:, 1]
        g = np.empty((self.ne, 3, 2))
        g[:, 1, 0] = d2[:, 1] / det
        g[:, 1, 1] = -d2[:, 0] / det
        g[:, 2, 0] = -d1[:, 1] / det
        g[:, 2, 1] = d1[:, 0] / det
        g[:, 0] = -g[:, 1] - g[:, 2]
        self.gradphi = g
        self.detabs = np.abs(det)

    def _slots(self):
        """CSR position of every (cell, a, b) entry."""
        c = self.cells.astype(np.int64)
        rows = np.repeat(c, 3, axis=1)             # (ne,9) row-major a,b
        cols = np.tile(c, (1, 3))
        # the global key array (row-major, columns sorted within a row) is sorted
        key = np.repeat(np.arange(self.nv, dtype=np.int64),
                        np.diff(self.rowptr)) * self.nv + self.col
        want = rows * self.nv + cols
        slot = np.searchsorted(key, want)
        assert np.array_equal(key[slot], want)
        self.slot = slot

    def _winning_cells(self):
        """Highest-index cell containing each vertex and the vertex' local index in it."""
        ce = np.repeat(np.arange(self.ne), 3)
        win = np.full(self.nv, -1, dtype=np.int64)
        np.maximum.at(win, self.cells.ravel(), ce)
        self.win_cell = win
        loc = np.full(self.nv, -1, dtype=np.int64)
        has = win >= 0
        wc = self.cells[win[has]]
        vid = np.nonzero(has)[0]
        loc[has] = np.argmax(wc == vid[:, None], axis=1)
        self.win_loc = loc

    def set_dirichlet(self, dofs, value):
        self.bc_dofs = np.unique(np.asarray(dofs, dtype=np.int32))
        self.N_bdry = float(value)

    # ---------------------------------------------------------------- physics
    def head(self, N):
        """constitutive.py:6-9"""
        p = self.p
        return self.z_b + (p.rho_i / p.rho_w) * (self.z_s - self.z_b) - N / (p.rho_w * p.g)

    def _cell_grad(self, f):
        return np.einsum("ea,eak->ek", f[self.cells], self.gradphi)

    def kbar(self):
        """|detJ| sum_k w_k K(b(xi_k), |q(xi_k)|)   (constitutive.py:11-20)."""
        p = self.p
        if self.backend == "c":
            from . import cbackend
            return cbackend.kbar(self)
        bc = self.b[self.cells]
        qc = self.q[self.cells]
        out = np.zeros(self.ne)
        for (xi, eta), w in zip(self.qpts, self.qwts):
            lam = np.array([1.0 - xi - eta, xi, eta])
            bq = bc @ lam
            qq = np.einsum("eak,a->ek", qc, lam)
            Re = np.power(qq[:, 0] * qq[:, 0] + qq[:, 1] * qq[:, 1], 0.5) / p.nu
            K = np.abs(bq) ** 3 * p.g / (12 * p.nu * (1 + p.omega * Re))
            out += w * K
        return out * self.detabs

    def element_FJ(self, dt, N=None, want_J=True):
        """Element residual (ne,3) and Jacobian (ne,3,3), every term at every quadrature point."""
        p = self.p
        N = self.N if N is None else N
        if self.backend == "c":
            from . import cbackend
            return cbackend.element_FJ(self, dt, N, want_J)
        c = self.cells
        gp = self.gradphi
        h = self.head(N)
        gh = self._cell_grad(h)
        gb = self._cell_grad(self.b)
        gm = self._cell_grad(self.melt_n)
        gb2 = np.einsum("ek,ek->e", gb, gb)
        gmgb = np.einsum("ek,ek->e", gm, gb)
        Nc, Nnc, bc = N[c], self.N_n[c], self.b[c]
        qc, Gc, mc = self.q[c], self.G[c], self.melt_n[c]
        sc, ic = self.storage[c], self.inputs[c]
        cm = 1 / p.rho_i - 1 / p.rho_w
        Fe = np.zeros((self.ne, 3))
        Je = np.zeros((self.ne, 3, 3)) if want_J else None
        ghgp = np.einsum("ek,eak->ea", gh, gp)               # grad h . grad phi_a
        gpgp = np.einsum("eak,ebk->eab", gp, gp)
        for (xi, eta), w in zip(self.qpts, self.qwts):
            lam = np.array([1.0 - xi - eta, xi, eta])
            wd = w * self.detabs
            bq, Nq, Nnq = bc @ lam, Nc @ lam, Nnc @ lam
            Gq, mq, sq, iq = Gc @ lam, mc @ lam, sc @ lam, ic @ lam
            qq = np.einsum("eak,a->ek", qc, lam)
            Re = np.power(qq[:, 0] * qq[:, 0] + qq[:, 1] * qq[:, 1], 0.5) / p.nu
            K = np.abs(bq) ** 3 * p.g / (12 * p.nu * (1 + p.omega * Re))
            # -dot(water_flux, grad v) = K grad h . grad v          (solvers.py:45)
            Fe += (wd * K)[:, None] * ghgp
            qgh = np.einsum("ek,ek->e", qq, gh)
            m0 = (Gq - p.rho_w * p.g * qgh) / p.Lh               # constitutive.py:25
            mdiff = (gb2 * mq + bq * gmgb) / (1 + gb2)           # constitutive.py:26, P1 cellwise
            clos = p.A * bq * Nq * np.abs(Nq) ** (p.n - 1)       # constitutive.py:31
            lake = sq * (1 / (p.rho_w * p.g * dt)) * (Nq - Nnq)  # solvers.py:42
            R = cm * (m0 + mdiff) - clos - lake - iq
            Fe += (wd * R)[:, None] * lam[None, :]
            if want_J:
                Je += -(wd * K / (p.rho_w * p.g))[:, None, None] * gpgp
                # d m0/dN[phi_b] = (q . grad phi_b)/Lh
                qgp = np.einsum("ek,ebk->eb", qq, gp)
                dclos = p.A * bq * (np.abs(Nq) ** (p.n - 1)
                                    + Nq * (p.n - 1) * np.abs(Nq) ** (p.n - 2) * np.sign(Nq))
                dR = (cm / p.Lh) * qgp - (dclos + sq / (p.rho_w * p.g * dt))[:, None] * lam[None, :]
                Je += wd[:, None, None] * lam[None, :, None] * dR[:, None, :]
        return Fe, Je

    def assemble(self, dt, N=None, want_J=True):
        """Global residual (nv,) and CSR Jacobian values (nnz,) with Dirichlet handling:
        lifting scale -1, F[bc] = N[bc] - g, bc rows/cols zeroed, unit bc diagonal."""
        N = self.N if N is None else N
        Fe, Je = self.element_FJ(dt, N, want_J=True)
        c = self.cells
        isbc = np.zeros(self.nv, dtype=bool)
        isbc[self.bc_dofs] = True
        bce = isbc[c]                                           # (ne,3)
        if self.bc_dofs.size:
            gx = np.where(bce, self.N_bdry - N[c], 0.0)         # (g - x) on bc columns
            Fe = Fe + np.einsum("eab,eb->ea", Je, gx)
        F = np.zeros(self.nv)
        np.add.at(F, c.ravel(), Fe.ravel())
        F[self.bc_dofs] = N[self.bc_dofs] - self.N_bdry
        if not want_J:
            return F, None
        keep = (~bce)[:, :, None] & (~bce)[:, None, :]
        Jz = np.where(keep, Je, 0.0)
        vals = np.zeros(self.col.size)
        np.add.at(vals, self.slot.ravel(), Jz.ravel())
        if self.bc_dofs.size:
            key = np.repeat(np.arange(self.nv, dtype=np.int64), np.diff(self.rowptr)) * self.nv + self.col
            d = np.searchsorted(key, self.bc_dofs.astype(np.int64) * (self.nv + 1))
            vals[d] += 1.0
        return F, vals

    def jacobian_matrix(self, vals):
        return sp.csr_matrix((vals, self.col, self.rowptr), shape=(self.nv, self.nv))

    # ---------------------------------------------------------------- Newton
    def newton(self, dt):
        """DOLFINx NewtonSolver.solve with defaults: criterion 'residual', relaxation 1,
        rtol 1e-9, atol 1e-10, max_it 50, LU linear solve, raise on non-convergence.

        newton_r0 == 'dolfinx': the relative test divides by ``residual0`` which the C++ class
        sets to ||dx||_2 of the first iteration and never resets between solves (0 before the
        very first solve).  newton_r0 == 'initial_residual': r0 = ||F||_2 at iteration 0
        (legacy-DOLFIN definition)."""
        def check(r):
            if self.newton_r0 == "dolfinx":
                rel = r / self._residual0 if self._residual0 > 0 else np.inf
            else:
                rel = r / self._r_init if self._r_init > 0 else 0.0
            return rel < self.rtol or r < self.atol

        F, _ = self.assemble(dt, want_J=False)
        r = np.linalg.norm(F)
        self._r_init = r
        conv = check(r)
        it = 0
        self.residual_history = [r]
        self.step_lengths = []
        while not conv and it < self.max_it:
            F, vals = self.assemble(dt)
            J = self.jacobian_matrix(vals).tocsc()
            dx = spla.splu(J, permc_spec=self.permc_spec).solve(F)
            lam = self.relaxation                     # NewtonSolver.relaxation_parameter (1 in the reference)
            self.N = self.N - lam * dx
            it += 1
            F, _ = self.assemble(dt, want_J=False)
            if it == 1 and self.newton_r0 == "dolfinx":
                self._residual0 = np.linalg.norm(dx)
            r_old, r = r, np.linalg.norm(F)
            # backtracking line search -- NOT in the reference (DOLFINx has none); restated here only as the
            # checker of the library's opt-in newton_line_search (same rule, same update order)
            nb = 0
            while nb < self.line_search and not (r <= (1.0 - 1e-4 * lam) * r_old):
                self.N = self.N + 0.5 * lam * dx
                lam *= 0.5
                nb += 1
                F, _ = self.assemble(dt, want_J=False)
                r = np.linalg.norm(F)
            self.backtracks += nb
            self.step_lengths.append(lam)
            self.residual_history.append(r)
            conv = check(r)
        if not conv:
            raise RuntimeError("Newton solver did not converge")
        return it, conv

    # ---------------------------------------------------------------- nodal updates
    def _expr_at_vertices(self, vals_cell_vertex):
        """Write a (ne,3[,k]) per-cell-vertex array cell by cell: last (highest) cell wins."""
        w, l = self.win_cell, self.win_loc
        return vals_cell_vertex[w, l]

    def update_q(self):
        """solvers.py:143,186: q <- WaterFlux(b, Head(N), Reynolds(q_old)) at cell vertices."""
        p = self.p
        c = self.cells
        gh = self._cell_grad(self.head(self.N))                  # (ne,2)
        bv = self.b[c]                                           # (ne,3)
        qv = self.q[c]                                           # (ne,3,2)
        Re = np.power(qv[..., 0] * qv[..., 0] + qv[..., 1] * qv[..., 1], 0.5) / p.nu
        p1 = -(np.abs(bv) ** 3)[..., None] * p.g * gh[:, None, :]
        p2 = 12 * p.nu * (1 + p.omega * Re)
        self.q = self._expr_at_vertices(p1 / p2[..., None])

    def _melt_cell_vertex(self):
        """Melt(q, Head(N), G, b, melt_n) evaluated at the three vertices of every cell."""
        p = self.p
        c = self.cells
        gh = self._cell_grad(self.head(self.N))
        gb = self._cell_grad(self.b)
        gm = self._cell_grad(self.melt_n)
        gb2 = np.einsum("ek,ek->e", gb, gb)[:, None]
        gmgb = np.einsum("ek,ek->e", gm, gb)[:, None]
        qgh = np.einsum("eak,ek->ea", self.q[c], gh)
        m0 = (self.G[c] - p.rho_w * p.g * qgh) / p.Lh
        mdiff = (gb2 * self.melt_n[c] + self.b[c] * gmgb) / (1 + gb2)
        return m0 + mdiff

    def update_melt(self):
        """solvers.py:165,189 (new q, new N, old b, old melt_n)."""
        self.melt_n = self._expr_at_vertices(self._melt_cell_vertex())

    def update_b(self, dt):
        """solvers.py:162,192,196 (new q, new N, old b, NEW melt_n), then clamp at b_min."""
        p = self.p
        c = self.cells
        melt = self._melt_cell_vertex()
        Nv = self.N[c]
        clos = p.A * self.b[c] * Nv * np.abs(Nv) ** (p.n - 1)
        bnew = self._expr_at_vertices(self.b[c] + dt * (melt / p.rho_i - clos))
        bnew[bnew < self.b_min] = self.b_min
        self.b = bnew

    # ---------------------------------------------------------------- time loop
    def step(self, dt):
        """One pass of solvers.py:179-229 (without output)."""
        it, conv = self.newton(dt)
        if self.backend == "c":
            from . import cbackend
            self.q, self.melt_n, self.b = cbackend.nodal_updates(self, dt)
        else:
            self.update_q()
            self.update_melt()
            self.update_b(dt)
        self.N_n = self.N.copy()
        return it, conv

    def start(self):
        """solvers.py:48: N.interpolate(N_n), done once when the solver is built."""
        self.N = self.N_n.copy()

    @staticmethod
    def dt_schedule(timesteps):
        """solvers.py:81,174-176."""
        t = np.asarray(timesteps, dtype=np.float64)
        dts = np.empty(t.size)
        dts[0] = 0.1 * np.abs(t[1] - t[0])
        dts[1:] = np.abs(t[1:] - t[:-1])
        return dts

    def run(self, timesteps, nsteps=None):
        dts = self.dt_schedule(timesteps)
        if nsteps is not None:
            dts = dts[:nsteps]
        self.start()
        its = []
        for dt in dts:
            it, _ = self.step(dt)
            its.append(it)
        return its
