"""Parity at BASELINE.json's sizes.  configs[1] (C2, 250k triangles) is small enough for the oracle,
so it is compared directly; at 1M dofs (C4's fields) the CUDA path is checked through
size-independent properties the reference's algorithm guarantees."""
import numpy as np
import pytest

from common import relinf

pytestmark = pytest.mark.gpu


def _oracle_for(case):
    from oracle.shakti_oracle import ShaktiOracle
    o = ShaktiOracle(case.xy, case.cells)
    for k in ("z_b", "z_s", "G", "inputs", "storage", "b", "N_n", "melt_n"):
        getattr(o, k)[:] = case.fields[k]
    o.q[:] = case.fields["q"]
    o.set_dirichlet(case.bc_dofs, case.N_bdry)
    o.start()
    return o


def test_c2_rect_250k_triangles_against_oracle():
    from shakti_b200 import capi, configs
    case = configs.rect_steady()                      # 500 x 250 x 2 = 250 000 triangles, 125 751 dofs
    assert case.cells.shape[0] == 250000
    o = _oracle_for(case)
    m = capi.Model(case.xy, case.cells)
    try:
        configs.apply_case(m, case)
        rp, col = m.csr()
        assert np.array_equal(rp, o.rowptr) and np.array_equal(col, o.col)          # bit-exact pattern
        F, J = m.assemble(360.0)
        Fo, Jo = o.assemble(360.0)
        assert relinf(F, Fo) < 1e-12 and relinf(J, Jo) < 1e-12
        dts = case.dts(3)
        its_o = [o.step(dt)[0] for dt in dts]
        its_m = list(m.run(dts))
        assert its_m == its_o
        for name, ref in (("N", o.N), ("b", o.b), ("melt_n", o.melt_n)):
            assert relinf(m.get_field(name), ref) < 1e-8, name
        assert relinf(m.get_flux(), o.q) < 1e-8
    finally:
        m.close()


@pytest.fixture(scope="module")
def big():
    from shakti_b200 import capi, configs
    case = configs.dofs16m(nside=1000, nsteps=8)      # C4's fields on 1M dofs
    m = capi.Model(case.xy, case.cells)
    configs.apply_case(m, case)
    yield case, m
    m.close()


def test_large_jacobian_is_the_derivative_of_the_residual(big):
    """J v == d/d eps F(N + eps v), all on the device (central differences)."""
    case, m = big
    rng = np.random.default_rng(0)
    N0 = m.get_field("N")
    v = rng.standard_normal(case.n_vert)
    v[case.bc_dofs] = 0.0
    m.assemble(3600.0)
    Jv = m.spmv(v)
    eps = 0.5
    m.set_field("N", N0 + eps * v); Fp, _ = m.assemble(3600.0, want_J=False)
    m.set_field("N", N0 - eps * v); Fm, _ = m.assemble(3600.0, want_J=False)
    m.set_field("N", N0)
    fd = (Fp - Fm) / (2 * eps)
    interior = np.ones(case.n_vert, bool)
    interior[case.bc_dofs] = False
    assert np.max(np.abs(fd - Jv)[interior]) < 1e-6 * np.max(np.abs(Jv)[interior])


def test_large_step_properties(big):
    case, m = big
    from shakti_b200 import capi
    dts = case.dts(3)
    for dt in dts:
        it, conv = m.step(dt)
        assert conv and it >= 1
        st = m.stats()
        assert st["last_residual"] <= 1e-9 * st["last_residual0"] or st["last_residual"] < 1e-10   # Newton exit test
    b = m.get_field("b")
    assert b.min() >= 1e-5 and np.isfinite(b).all()
    N, Nn = m.get_field("N"), m.get_field("N_n")
    assert np.array_equal(N, Nn)                                            # solvers.py:228
    assert np.allclose(N[case.bc_dofs], case.N_bdry, rtol=0, atol=1e-6)     # Dirichlet rows
    # linear solve: residual of J dx = F reaches the requested tolerance
    F, _ = m.assemble(3600.0, want_J=False)
    rhs = F.copy()
    dx, its, relres = m.linear_solve(rhs)
    r = rhs - m.spmv(dx)
    r[case.bc_dofs] = 0.0
    assert np.linalg.norm(r) <= 5e-12 * np.linalg.norm(rhs) and its < 60


def test_large_step_is_reproducible():
    """Atomics-free assembly, ordered reductions: two runs give identical bits."""
    from shakti_b200 import capi, configs
    case = configs.dofs16m(nside=400, nsteps=6)
    out = []
    for _ in range(2):
        m = capi.Model(case.xy, case.cells)
        configs.apply_case(m, case)
        m.run(case.dts(2))
        out.append((m.get_field("N"), m.get_field("b")))
        m.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def test_model_against_golden_dump(tmp_path):
    """The checker a real FEniCSx dump would be run through (here: a dump written by the oracle)."""
    from common import make_case, make_oracle
    from shakti_b200 import golden
    c = make_case(nx=16, ny=12, seed=13)
    golden.write_dump(tmp_path / "dump", make_oracle(*c), [360.0, 3600.0, 3600.0])
    d = golden.Dump(tmp_path / "dump")
    s = golden.ModelStepper(d)
    try:
        rep = golden.check(d, s, pattern=s.m.csr())
        assert rep["pattern_equal"] and all(rep[f"step{i}"]["niter"][0] == rep[f"step{i}"]["niter"][1] for i in range(3))
    finally:
        s.m.close()


def test_model_against_committed_golden_dump():
    """The CUDA path against the committed fixture tests/golden/small_dump."""
    from pathlib import Path
    from shakti_b200 import golden
    d = golden.Dump(Path(__file__).parent / "golden" / "small_dump")
    s = golden.ModelStepper(d)
    try:
        rep = golden.check(d, s, pattern=s.m.csr())
        assert rep["pattern_equal"]
        assert all(rep[f"step{i}"]["niter"][0] == rep[f"step{i}"]["niter"][1] for i in range(len(d.steps)))
    finally:
        s.m.close()
