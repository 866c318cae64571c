#!/usr/bin/env python3
"""Write a golden dump (format: shakti_b200/golden.py) from a REAL FEniCSx run of the reference.

Run where dolfinx (0.8 / 0.9), petsc4py and the reference checkout are available, SERIALLY:

    cd <reference>/source && python3 <this repo>/tools/dump_fenicsx_golden.py <setup_module> <out_dir> [nsteps]

UNTESTED here (FEniCSx is not installable in the build container); it only uses public DOLFINx
API calls the reference itself uses plus `basix.make_quadrature`, `V.dofmap.list`,
`assemble_vector/assemble_matrix`.  What it pins (SURVEY.md §8c i-vi): the quadrature table the forms
get, the dof map, F and J at the initial state, Newton iteration counts and the fields after each step.
"""
import json
import sys
from pathlib import Path

import numpy as np


def main():
    setup_name, out = sys.argv[1], Path(sys.argv[2])
    nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    sys.path.insert(0, "../setups")
    import importlib
    import basix
    from mpi4py import MPI
    from dolfinx.fem import Constant, Function, Expression, form
    from dolfinx.fem.petsc import NonlinearProblem, assemble_vector, assemble_matrix, apply_lifting, set_bc
    from dolfinx.nls.petsc import NewtonSolver
    import ufl
    import params
    from constitutive import Melt, Closure, Head, WaterFlux, Reynolds
    from solvers import get_bcs

    md = importlib.import_module(setup_name).initialize(MPI.COMM_WORLD)
    assert md.size == 1, "dump serially: the parallel dof order is partition dependent"
    out.mkdir(parents=True, exist_ok=True)
    V, Vq = md.V, md.V_flux
    np.save(out / "geometry_x.npy", md.domain.geometry.x[:, :2])
    cells = np.asarray(V.dofmap.list).reshape(-1, 3).astype(np.int32)
    assert np.array_equal(cells, np.asarray(md.domain.geometry.dofmap).reshape(-1, 3)), "geometry node != dof index"
    np.save(out / "cells.npy", cells)

    # state exactly as solvers.solve builds it
    N, q, b, N_n, melt_n = Function(V), Function(Vq), Function(V), Function(V), Function(V)
    b.interpolate(md.b_init); N_n.interpolate(md.N_init)
    q.sub(0).interpolate(md.q_init.sub(0)); q.sub(1).interpolate(md.q_init.sub(1))
    storage = md.lake_bdry if md.storage_on else Function(V)
    dts = [0.1 * abs(md.timesteps[1] - md.timesteps[0])] + [abs(md.timesteps[i] - md.timesteps[i - 1]) for i in range(1, nsteps)]
    dt = Constant(md.domain, dts[0])
    np.savez(out / "initial.npz", z_b=md.z_b.x.array, z_s=md.z_s.x.array, G=md.G.x.array, inputs=md.inputs.x.array,
             storage=storage.x.array, b=b.x.array, N_n=N_n.x.array, q=q.x.array.reshape(-1, 2), melt_n=melt_n.x.array)

    bcs = get_bcs(md)
    np.save(out / "bc_dofs.npy", np.sort(bcs[0]._cpp_object.dof_indices()[0]).astype(np.int32) if bcs else np.zeros(0, np.int32))
    v = ufl.TestFunction(V)
    head = Head(N, md.z_b, md.z_s)
    F = -ufl.dot(WaterFlux(b, head, Reynolds(q)), ufl.grad(v)) * ufl.dx + (
        (1 / params.rho_i - 1 / params.rho_w) * Melt(q, head, md.G, b, melt_n) - Closure(b, N)
        - storage * (1 / (params.rho_w * params.g * dt)) * (N - N_n) - md.inputs) * v * ufl.dx
    N.interpolate(N_n)
    problem = NonlinearProblem(F, N, bcs=bcs)
    # the quadrature rule FFCx gives these forms: estimated degree -> basix default scheme
    from ufl.algorithms import estimate_total_polynomial_degree
    deg = estimate_total_polynomial_degree(F)
    pts, wts = basix.make_quadrature(basix.CellType.triangle, deg)
    np.save(out / "quadrature_points.npy", pts); np.save(out / "quadrature_weights.npy", wts)

    # F and J at the initial state, assembled the way NonlinearProblem.F / .J do
    L, a = problem.L, problem.a
    vec = assemble_vector(L); apply_lifting(vec, [a], bcs=[bcs], x0=[N.x.petsc_vec], alpha=-1.0)
    vec.ghostUpdate(addv=1, mode=1); set_bc(vec, bcs, N.x.petsc_vec, -1.0)
    np.save(out / "F0.npy", vec.array.copy())
    A = assemble_matrix(a, bcs=bcs); A.assemble()
    ip, ix, dv = A.getValuesCSR()
    np.save(out / "J0_indptr.npy", ip); np.save(out / "J0_indices.npy", ix); np.save(out / "J0_data.npy", dv)

    solver = NewtonSolver(md.comm, problem)
    q_expr = Expression(WaterFlux(b, Head(N, md.z_b, md.z_s), Reynolds(q)), Vq.element.interpolation_points())
    melt_expr = Expression(Melt(q, Head(N, md.z_b, md.z_s), md.G, b, melt_n), V.element.interpolation_points())
    b_expr = Expression(b + dt * (Melt(q, Head(N, md.z_b, md.z_s), md.G, b, melt_n) / params.rho_i - Closure(b, N)),
                        V.element.interpolation_points())
    for i in range(nsteps):
        dt.value = dts[i]
        niter, converged = solver.solve(N)
        q.interpolate(q_expr); melt_n.interpolate(melt_expr); b.interpolate(b_expr)
        b.x.array[b.x.array < md.b_min] = md.b_min
        np.savez(out / f"step_{i:04d}.npz", N=N.x.array, b=b.x.array, q=q.x.array.reshape(-1, 2), melt_n=melt_n.x.array, niter=niter)
        N_n.x.array[:] = N.x.array
    prm = {k: float(getattr(params, k)) for k in ("g", "rho_i", "rho_w", "nu", "Lh", "omega", "n", "A")}
    import dolfinx
    (out / "meta.json").write_text(json.dumps(dict(N_bdry=float(md.N_bdry), dts=[float(d) for d in dts], params=prm,
                                                    producer=f"dolfinx {dolfinx.__version__}", quadrature_degree=int(deg))))
    print("golden dump written to", out)


if __name__ == "__main__":
    main()
