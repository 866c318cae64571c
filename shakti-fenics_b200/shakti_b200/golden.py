"""Golden dumps: a directory of arrays written by a REAL FEniCSx run of the reference (see
``tools/dump_fenicsx_golden.py``) and the checks that pin this repo against it.

No such dump exists yet: FEniCSx is not installed here and cannot be (DESIGN.md
§1 "parity unpinned").  The loader and the checks are exercised with dumps of the same format
written by the CPU oracle (``write_dump``; the oracle-side adapter lives in tests/common.py), so the moment a maintainer produces a real dump the
open questions (quadrature table, Newton r0, cell order, last-cell-wins) are settled by
running ``check_oracle`` / ``check_model`` on it.

Directory layout
  meta.json                 {"N_bdry", "dts": [...], "params": {...}, "producer": "..."}
  geometry_x.npy (nv,2)     cells.npy (ne,3) int32      bc_dofs.npy int32
  quadrature_points.npy (nq,2)   quadrature_weights.npy (nq,)      # rule of the forms (reference triangle, sum = 1/2)
  initial.npz               z_b z_s G inputs storage b N_n q(nv,2) melt_n
  F0.npy (nv,)              J0_indptr.npy J0_indices.npy J0_data.npy     # residual/Jacobian at the initial state, dt = dts[0]
  step_0000.npz ...         N b q melt_n niter                            # state after each time step
"""
import json
from pathlib import Path

import numpy as np

FIELDS = ("z_b", "z_s", "G", "inputs", "storage", "b", "N_n", "melt_n")


class Dump:
    def __init__(self, path):
        p = Path(path)
        self.path = p
        self.meta = json.loads((p / "meta.json").read_text())
        self.xy = np.load(p / "geometry_x.npy")[:, :2].copy()
        self.cells = np.load(p / "cells.npy").astype(np.int32)
        self.bc_dofs = np.load(p / "bc_dofs.npy").astype(np.int32)
        self.quad = (np.load(p / "quadrature_points.npy")[:, :2], np.load(p / "quadrature_weights.npy"))
        self.initial = dict(np.load(p / "initial.npz"))
        self.F0 = np.load(p / "F0.npy")
        self.J0 = (np.load(p / "J0_indptr.npy"), np.load(p / "J0_indices.npy"), np.load(p / "J0_data.npy"))
        self.steps = [dict(np.load(f)) for f in sorted(p.glob("step_*.npz"))]
        self.dts = list(self.meta["dts"])
        self.N_bdry = float(self.meta["N_bdry"])


def write_dump(path, oracle, dts, producer="oracle"):
    """Write a dump in the golden format from an oracle-like object at its initial state (test helper;
    the object is passed in: this module never imports ``oracle/``)."""
    p = Path(path)
    p.mkdir(parents=True, exist_ok=True)
    o = oracle
    np.save(p / "geometry_x.npy", o.xy)
    np.save(p / "cells.npy", o.cells)
    np.save(p / "bc_dofs.npy", o.bc_dofs)
    np.save(p / "quadrature_points.npy", o.qpts)
    np.save(p / "quadrature_weights.npy", o.qwts)
    np.savez(p / "initial.npz", q=o.q, **{k: getattr(o, k) for k in FIELDS})
    F, vals = o.assemble(dts[0])
    np.save(p / "F0.npy", F)
    np.save(p / "J0_indptr.npy", o.rowptr)
    np.save(p / "J0_indices.npy", o.col)
    np.save(p / "J0_data.npy", vals)
    for i, dt in enumerate(dts):
        it, _ = o.step(dt)
        np.savez(p / f"step_{i:04d}.npz", N=o.N, b=o.b, q=o.q, melt_n=o.melt_n, niter=it)
    prm = {k: getattr(o.p, k) for k in ("g", "rho_i", "rho_w", "nu", "Lh", "omega", "n", "A")}
    (p / "meta.json").write_text(json.dumps(dict(N_bdry=o.N_bdry, dts=list(map(float, dts)), params=prm, producer=producer)))


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


def _csr_dense_compare(ind_a, col_a, val_a, ind_b, col_b, val_b):
    """Jacobians may carry different explicit-zero patterns (PETSc keeps the full P1 pattern): compare as matrices."""
    import scipy.sparse as sp
    n = len(ind_a) - 1
    A = sp.csr_matrix((val_a, col_a, ind_a), shape=(n, n))
    B = sp.csr_matrix((val_b, col_b, ind_b), shape=(n, n))
    return float(abs(A - B).max() / max(abs(B).max(), 1e-300))


def check(dump, stepper, tol_FJ=1e-12, tol_fields=1e-8, pattern=None):
    """``stepper`` adapts the implementation under test:
        stepper.assemble(dt) -> (F, (indptr, indices, data))
        stepper.step(dt) -> niter ; stepper.state() -> dict(N, b, q, melt_n)
    Returns a report dict and raises AssertionError on the first violated tolerance."""
    rep = {}
    F, (ip, ix, dv) = stepper.assemble(dump.dts[0])
    rep["F"] = _rel(F, dump.F0)
    rep["J"] = _csr_dense_compare(ip, ix, dv, *dump.J0)
    assert rep["F"] < tol_FJ and rep["J"] < tol_FJ, rep
    if pattern is not None:                       # bit-exact pattern when the dump carries the full P1 pattern
        rep["pattern_equal"] = bool(np.array_equal(pattern[0], dump.J0[0]) and np.array_equal(pattern[1], dump.J0[1]))
    for i, (dt, ref) in enumerate(zip(dump.dts, dump.steps)):
        it = stepper.step(dt)
        st = stepper.state()
        errs = {k: _rel(st[k], ref[k]) for k in ("N", "b", "q", "melt_n")}
        rep[f"step{i}"] = dict(niter=(int(it), int(ref["niter"])), **errs)
        assert all(v < tol_fields for v in errs.values()), rep
    return rep


class ModelStepper:
    """The CUDA path through the C ABI."""

    def __init__(self, dump, **opt):
        from . import capi
        p = capi.Params()
        for k, v in dump.meta["params"].items():
            setattr(p, k, float(v))
        m = capi.Model(dump.xy, dump.cells, params=p, **opt)
        m.set_quadrature(*dump.quad)
        for k in FIELDS:
            m.set_field(k, dump.initial[k])
        m.set_flux(dump.initial["q"])
        m.set_dirichlet(dump.bc_dofs, dump.N_bdry)
        m.start()
        self.m = m

    def assemble(self, dt):
        F, J = self.m.assemble(dt)
        rp, col = self.m.csr()
        return F, (rp, col, J)

    def step(self, dt):
        return self.m.step(dt)[0]

    def state(self):
        m = self.m
        return dict(N=m.get_field("N"), b=m.get_field("b"), q=m.get_flux(), melt_n=m.get_field("melt_n"))
