"""The CPU oracle against independent mathematics (the reference ships no tests or golden
vectors and cannot run here: SURVEY.md §8c, "parity unpinned")."""
import json
from math import factorial
from pathlib import Path

import numpy as np
import pytest

from common import make_case, make_oracle, relinf
from oracle import quadrature
from oracle.shakti_oracle import Params, ShaktiOracle, csr_pattern, dirichlet_dofs

GOLDEN = json.loads((Path(__file__).parent / "golden" / "element_golden.json").read_text())


@pytest.mark.parametrize("rule,deg", [(quadrature.gauss_jacobi_triangle(7), 7), (quadrature.radon7(), 5),
                                      (quadrature.gauss_jacobi_triangle(10), 10)])
def test_quadrature_exact_for_monomials(rule, deg):
    p, w = rule
    assert abs(w.sum() - 0.5) < 1e-15
    for a in range(deg + 1):
        for b in range(deg + 1 - a):
            exact = factorial(a) * factorial(b) / factorial(a + b + 2)
            assert abs((w * p[:, 0] ** a * p[:, 1] ** b).sum() - exact) < 2e-16 + 1e-14 * exact


def test_check_degree_accepts_rules_of_the_degree_and_rejects_others():
    """Any table handed to shakti_set_quadrature / ShaktiOracle(quad=) -- in particular the one a FEniCSx golden
    dump records from Basix -- can be validated before use."""
    assert quadrature.check_degree(*quadrature.gauss_jacobi_triangle(7), degree=7) < 1e-15
    assert quadrature.check_degree(*quadrature.radon7(), degree=5) < 1e-15
    with pytest.raises(ValueError):
        quadrature.check_degree(*quadrature.radon7(), degree=7)            # a degree-5 rule is not a degree-7 rule
    p, w = quadrature.gauss_jacobi_triangle(7)
    with pytest.raises(ValueError):
        quadrature.check_degree(p, 1.001 * w, degree=7)                    # wrong area
    with pytest.raises(ValueError):
        quadrature.check_degree(p + np.array([0.9, 0.0]), w, degree=7)     # points outside the triangle


def _single_element_oracle(case):
    xy = np.array(case["xy"])
    o = ShaktiOracle(xy, np.array([[0, 1, 2]], dtype=np.int32))
    f = case["fields"]
    for k in ("z_b", "z_s", "G", "inputs", "storage", "b", "N_n", "melt_n"):
        getattr(o, k)[:] = f[k]
    o.N = np.array(f["N"])
    o.q[:, 0], o.q[:, 1] = f["qx"], f["qy"]
    return o


@pytest.mark.parametrize("idx", range(len(GOLDEN["cases"])))
def test_element_integrals_match_exact_sympy_fixture(idx):
    """tests/golden/element_golden.json holds EXACT (sympy rational) integrals of the weak form
    solvers.py:45 and its derivative on single triangles, see make_element_golden.py."""
    case = GOLDEN["cases"][idx]
    o = _single_element_oracle(case)
    Fe, Je = o.element_FJ(case["dt"])
    Fe, Je = Fe[0], Je[0]
    if case["kind"] == "poly":          # fixture excludes the (non-polynomial) K term: remove it
        kb = o.kbar()[0]
        gh = o._cell_grad(o.head(o.N))[0]
        gp = o.gradphi[0]
        Fe = Fe - kb * gp @ gh
        Je = Je + kb / (o.p.rho_w * o.p.g) * gp @ gp.T
    Fg, Jg = np.array(case["F"]), np.array(case["J"])
    assert np.max(np.abs(Fe - Fg)) <= 1e-12 * np.max(np.abs(Fg))
    assert np.max(np.abs(Je - Jg)) <= 1e-12 * np.max(np.abs(Jg))


def test_jacobian_matches_finite_differences():
    c = make_case(nx=10, ny=8, seed=3)
    o = make_oracle(*c)
    F, vals = o.assemble(3600.0)
    J = o.jacobian_matrix(vals)
    rng = np.random.default_rng(0)
    v = rng.standard_normal(o.nv)
    v[o.bc_dofs] = 0
    eps = 0.1
    Fp, _ = o.assemble(3600.0, o.N + eps * v, want_J=False)
    Fm, _ = o.assemble(3600.0, o.N - eps * v, want_J=False)
    fd, an = (Fp - Fm) / (2 * eps), J @ v
    interior = np.ones(o.nv, bool)
    interior[o.bc_dofs] = False
    assert np.max(np.abs(fd - an)[interior]) < 1e-7 * np.max(np.abs(an)[interior])


def test_patch_linear_head_gives_zero_flux_divergence():
    """With b constant, q = 0 and every reaction term off, a linear head field has zero interior residual."""
    from shakti_b200 import meshgen
    xy, cells = meshgen.rectangle(8, 6, 8e3, 6e3, jitter=0.2, diagonal="random")
    p = Params(A=0.0)
    o = ShaktiOracle(xy, cells, params=p)
    o.b[:] = 2e-3
    o.z_b[:] = 0.0
    o.z_s[:] = 0.0
    o.N = -(p.rho_w * p.g) * (0.01 * xy[:, 0] + 0.02 * xy[:, 1])     # h = -N/(rho_w g) linear
    o.N_n = o.N.copy()
    F, _ = o.assemble(3600.0, want_J=False)
    from oracle.shakti_oracle import boundary_facets
    bnd = np.unique(boundary_facets(cells))
    interior = np.setdiff1d(np.arange(o.nv), bnd)
    assert np.max(np.abs(F[interior])) < 1e-12 * np.max(np.abs(F[bnd]))


def test_stiffness_rows_sum_to_zero():
    c = make_case(nx=8, ny=6, seed=1, storage=False)
    o = make_oracle(*c)
    o.p.A = 0.0
    o.set_dirichlet([], 0.0)
    o.q[:] = 0.0        # removes the advection part
    _, vals = o.assemble(3600.0)
    J = o.jacobian_matrix(vals)
    assert np.max(np.abs(J @ np.ones(o.nv))) < 1e-12 * np.max(np.abs(vals))


def test_dirichlet_handling():
    c = make_case(nx=8, ny=6, seed=2)
    o = make_oracle(*c)
    o.N[o.bc_dofs] = o.N_bdry + 7.0   # violate the BC: lifting must act
    F, vals = o.assemble(3600.0)
    J = o.jacobian_matrix(vals).toarray()
    bc = o.bc_dofs
    assert np.allclose(F[bc], 7.0)
    assert np.allclose(J[bc][:, bc], np.eye(bc.size))
    off = np.ones(o.nv, bool)
    off[bc] = False
    assert np.all(J[bc][:, off] == 0) and np.all(J[off][:, bc] == 0)
    # one Newton update restores the boundary value exactly
    dx = np.linalg.solve(J, F)
    assert np.allclose((o.N - dx)[bc], o.N_bdry)


def test_last_cell_wins_interpolation():
    c = make_case(nx=6, ny=5, seed=5)
    o = make_oracle(*c)
    vals = np.arange(o.ne * 3, dtype=float).reshape(o.ne, 3)
    out = o._expr_at_vertices(vals)
    ref = np.zeros(o.nv)
    for cidx in range(o.ne):            # the literal DOLFINx loop
        for i in range(3):
            ref[o.cells[cidx, i]] = vals[cidx, i]
    assert np.array_equal(out, ref)


def test_dt_schedule_and_step_order():
    t = np.array([0.0, 10.0, 25.0, 27.0])
    assert np.allclose(ShaktiOracle.dt_schedule(t), [1.0, 10.0, 15.0, 2.0])
    c = make_case(nx=8, ny=6, seed=6)
    o = make_oracle(*c)
    its = o.run(np.linspace(0, 3 * 3600.0, 4), nsteps=3)
    assert len(its) == 3 and np.array_equal(o.N, o.N_n)
    assert o.b.min() >= o.b_min


@pytest.mark.parametrize("mode", ["initial_residual", "dolfinx"])
def test_newton_modes_converge_to_same_root(mode):
    c = make_case(nx=8, ny=6, seed=7)
    o = make_oracle(*c, newton_r0=mode)
    it, conv = o.newton(3600.0)
    assert conv and it >= 1
    F, _ = o.assemble(3600.0, want_J=False)
    denom = o.residual_history[0] if mode == "initial_residual" else o._residual0   # ||F_0|| or ||dx_1||
    assert np.linalg.norm(F) < 1e-9 * denom


def test_newton_raises_when_not_converged():
    c = make_case(nx=8, ny=6, seed=8)
    o = make_oracle(*c)
    o.max_it = 0
    o.rtol = o.atol = 0.0
    with pytest.raises(RuntimeError):
        o.newton(3600.0)


def test_csr_pattern_and_dirichlet_dofs_small():
    xy = np.array([[0, 0], [1, 0], [1, 1], [0, 1.0]])
    cells = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int32)
    rp, col = csr_pattern(4, cells)
    assert rp.tolist() == [0, 4, 7, 11, 14]
    assert col.tolist() == [0, 1, 2, 3, 0, 1, 2, 0, 1, 2, 3, 0, 2, 3]
    d = dirichlet_dofs(xy, cells, lambda x: np.isclose(x[0], 0.0))
    assert d.tolist() == [0, 3]


def test_newton_relaxation_and_line_search_restatement():
    """The oracle's relaxation parameter (DOLFINx NewtonSolver.relaxation_parameter) and the backtracking rule the
    library's opt-in newton_line_search is checked against: defaults reproduce the plain iteration exactly; an
    overshooting step (2.5 dx) is halved once per iteration; every accepted step satisfies the decrease test."""
    c = make_case(seed=1)
    o0 = make_oracle(*c)
    it0, _ = o0.newton(3600.0)
    o1 = make_oracle(*c)
    o1.line_search = 4
    it1, _ = o1.newton(3600.0)
    assert it1 == it0 and o1.backtracks == 0 and np.array_equal(o1.N, o0.N)
    o2 = make_oracle(*c)
    o2.relaxation, o2.line_search = 2.5, 3
    it2, conv = o2.newton(3600.0)
    assert conv and o2.backtracks == it2 and set(o2.step_lengths) == {1.25}
    r = np.array(o2.residual_history)
    assert np.all(r[1:] <= (1 - 1e-4 * 1.25) * r[:-1])
    # same root whatever the path (a linearly convergent iteration stops ~1e-7 away from it at rtol 1e-9)
    assert relinf(o2.N, o0.N) < 1e-6
    o3 = make_oracle(*c)
    o3.relaxation = 0.7
    it3, _ = o3.newton(3600.0)
    assert it3 > it0 and relinf(o3.N, o0.N) < 1e-6


@pytest.fixture(scope="module")
def c_backend():
    """oracle/shakti_oracle_c.c, built on demand (gcc is part of the image; __graft_entry__.build() does the same)."""
    import subprocess
    from oracle import cbackend
    subprocess.run(["make", "-C", str(Path(__file__).resolve().parent.parent / "oracle")], check=True, capture_output=True)
    assert cbackend.available() and cbackend.threads() >= 1
    return cbackend


@pytest.mark.parametrize("kw", [dict(seed=3), dict(seed=4, neg_b=True, turbulent=False), dict(seed=5, storage=False)],
                         ids=["rough-fields", "negative-gap", "no-storage"])
def test_c_backend_is_the_same_restatement(c_backend, kw):
    """The compiled element kernels (the CPU baseline bench.py times) against the numpy oracle: element residual and
    Jacobian, Kbar and three full time steps incl. the last-cell-wins nodal updates and the clamp."""
    c = make_case(nx=30, ny=20, **kw)
    o1, o2 = make_oracle(*c), make_oracle(*c, backend="c")
    for dt in (3600.0, 360.0):
        (F1, J1), (F2, J2) = o1.element_FJ(dt), o2.element_FJ(dt)
        assert relinf(F2, F1) < 1e-13 and relinf(J2, J1) < 1e-13
        assert o2.element_FJ(dt, want_J=False)[1] is None
    assert relinf(o2.kbar(), o1.kbar()) < 1e-13
    # the three nodal updates from an identical state (N moved away from N_n so that every term is active)
    o1, o2 = make_oracle(*c), make_oracle(*c, backend="c")
    o1.N[:] = o2.N[:] = c[2]["N_n"] * 1.01
    o1.update_q(); o1.update_melt(); o1.update_b(3600.0)
    q, melt, b = c_backend.nodal_updates(o2, 3600.0)
    assert relinf(q, o1.q) < 1e-13 and relinf(melt, o1.melt_n) < 1e-12 and relinf(b, o1.b) < 1e-13
    assert (b == o1.b_min).sum() == (o1.b == o1.b_min).sum()
    if kw.get("neg_b"):
        return                               # the first Newton solve of this state is chaotic (LU-conditioning bound)
    o1, o2 = make_oracle(*c), make_oracle(*c, backend="c", permc_spec="MMD_AT_PLUS_A")
    for dt in (360.0, 3600.0, 3600.0):
        assert o1.step(dt)[0] == o2.step(dt)[0]
    assert relinf(o2.N, o1.N) < 1e-10 and relinf(o2.b, o1.b) < 1e-10 and relinf(o2.melt_n, o1.melt_n) < 1e-9


def test_line_search_is_not_a_cure_on_the_negative_gap_start_state():
    """Why newton_line_search is off by default ("Step length"): on the reference's hard start state
    (setup_cooke2.py:66: gap height negative at ~40 % of the nodes) the plain Newton iteration of the reference wanders
    -- the residual goes up and down -- and converges; forcing ||F||_2 to decrease monotonically stalls instead."""
    c = make_case(seed=4, neg_b=True, turbulent=False)
    o = make_oracle(*c)
    it, conv = o.newton(360.0)
    r = np.array(o.residual_history)
    assert conv and it > 10 and np.any(r[2:] > r[1:-1])       # non-monotone, yet it converges
    o = make_oracle(*c)
    o.line_search = 6
    with pytest.raises(RuntimeError):
        o.newton(360.0)
    r = np.array(o.residual_history)
    assert o.backtracks > 0 and r[-1] > 1e-3 * r[1]            # stalled far from the root


def test_oracle_sensitivity_on_the_negative_gap_start_state():
    """How well defined is the reference answer on its own hard start state?  The same oracle with two LU column
    orderings (different rounding, same mathematics) agrees to ~3e-11 in N after the 22 wandering Newton iterations:
    that is the floor any other solver of this step can be held to, and the context of the 1e-6 bound of
    tests/test_gpu_parity.py::test_negative_gap_height_first_step (Krylov solves stop on the residual)."""
    c = make_case(seed=4, neg_b=True, turbulent=False)
    o1, o2 = make_oracle(*c), make_oracle(*c, permc_spec="MMD_AT_PLUS_A")
    it1, _ = o1.newton(360.0)
    it2, _ = o2.newton(360.0)
    assert it1 == it2 > 10
    d = relinf(o2.N, o1.N)
    assert 1e-14 < d < 1e-8, d
