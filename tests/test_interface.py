"""Host mirror of the reference interface: params / constitutive / model_setup / solvers /
main keep the reference's names, argument order and attributes."""
import importlib
import inspect
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "shakti-fenics_b200" / "source"
SETUPS = ROOT / "shakti-fenics_b200" / "setups"
for p in (str(SRC), str(SETUPS)):
    if p not in sys.path:
        sys.path.insert(0, p)

from common import make_case, make_oracle, relinf  # noqa: E402
from shakti_b200 import fem  # noqa: E402
from shakti_b200.ufl_lite import interpolate_expression  # noqa: E402


def test_params_match_reference_values():
    import params
    assert (params.g, params.rho_i, params.rho_w, params.nu, params.Lh, params.omega, params.n, params.A) == \
        (9.81, 917, 1000, 1.787e-6, 3.34e5, 1e-3, 3, 2.24e-24)
    assert isinstance(params.rho_i, int) and isinstance(params.rho_w, int) and isinstance(params.n, int)


def test_constitutive_signatures():
    import constitutive as c
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(c.Head) == ["N", "z_b", "z_s"]
    assert sig(c.WaterFlux) == ["b", "h", "Re"]
    assert sig(c.Reynolds) == ["q"]
    assert sig(c.Melt) == ["q", "h", "G", "b_n", "melt_n"]
    assert sig(c.Closure) == ["b", "N"]
    assert sig(c.BackgroundGradient) == ["z_b", "z_s"] and sig(c.BackgroundPotential) == ["z_b", "z_s"]


def test_solver_signatures():
    import solvers
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(solvers.get_bcs) == ["md"]
    assert sig(solvers.pde_solver) == ["md", "N", "N_n", "b", "q", "melt_n", "storage", "dt"]
    assert sig(solvers.solve) == ["md"]


def test_constitutive_expressions_reproduce_oracle_nodal_updates():
    """The reference's own expression text (solvers.py:143,162,165 with constitutive.py),
    evaluated by the mini-UFL with last-cell-wins interpolation, equals the oracle."""
    import constitutive as cst
    c = make_case()
    o = make_oracle(*c)
    xy, cells, f, bc, Nb = c
    mesh = fem.Mesh(xy, cells)
    V = fem.functionspace(mesh, ("CG", 1))
    Vq = fem.functionspace(mesh, fem.element('P', 'triangle', 1, shape=(2,)))
    F = {k: fem.Function(V) for k in ("z_b", "z_s", "G", "b", "N", "melt_n")}
    for k in ("z_b", "z_s", "G", "b", "melt_n"):
        F[k].x.array[:] = f[k]
    F["N"].x.array[:] = f["N_n"]
    o.N = f["N_n"].copy()
    q = fem.Function(Vq)
    q.x.array[:] = f["q"].reshape(-1)
    head = lambda: cst.Head(F["N"], F["z_b"], F["z_s"])
    interpolate_expression(q, cst.WaterFlux(F["b"], head(), cst.Reynolds(q)))
    o.update_q()
    assert relinf(q.x.array.reshape(-1, 2), o.q) < 1e-13
    interpolate_expression(F["melt_n"], cst.Melt(q, head(), F["G"], F["b"], F["melt_n"]))
    o.update_melt()
    assert relinf(F["melt_n"].x.array, o.melt_n) < 1e-13
    dt = 3600.0
    interpolate_expression(F["b"], F["b"] + dt * (cst.Melt(q, head(), F["G"], F["b"], F["melt_n"]) / cst.rho_i
                                                   - cst.Closure(F["b"], F["N"])))
    o.update_b(dt)
    assert relinf(np.maximum(F["b"].x.array, 1e-5), o.b) < 1e-13


def test_background_potential():
    import constitutive as cst
    import params
    xy, cells, f, *_ = make_case(nx=6, ny=4)
    mesh = fem.Mesh(xy, cells)
    V = fem.functionspace(mesh, ("CG", 1))
    zb, zs = fem.Function(V), fem.Function(V)
    zb.x.array[:], zs.x.array[:] = f["z_b"], f["z_s"]
    pot = fem.Function(V)
    interpolate_expression(pot, cst.BackgroundPotential(zb, zs))
    ref = params.rho_i * params.g * f["z_s"] + (params.rho_w - params.rho_i) * params.g * f["z_b"]
    assert relinf(pot.x.array, ref) < 1e-13


def test_model_setup_attributes_and_helpers():
    from model_setup import model_setup, points_in_polygon
    from shakti_b200 import meshgen
    xy, cells = meshgen.rectangle(10, 8, 10e3, 8e3)
    md = model_setup(fem.comm_world(), fem.Mesh(xy, cells))
    for a in ("comm", "rank", "size", "domain", "x", "y", "V", "V_flux", "mask", "OutflowBoundary", "bounds", "outflow_on",
              "storage_on", "z_b", "z_s", "G", "inputs", "b_init", "N_init", "q_init", "lake_bdry", "N_bdry", "b_min",
              "outline", "lake_name", "results_name", "setup_name", "timesteps", "nt_save", "nt_check"):
        assert hasattr(md, a), a
    assert md.b_min == 1e-5 and md.N_bdry == 0.0 and md.outflow_on and md.storage_on
    assert md.mask.all() and md.q_init.x.array.size == 2 * xy.shape[0]
    assert md.bounds[0] == pytest.approx(-10e3) and md.bounds[1] == pytest.approx(20e3)   # buffer = 10 x max gap
    # gridded data interpolation (bilinear, extrapolating)
    xd, yd = np.linspace(-20e3, 30e3, 26), np.linspace(-20e3, 30e3, 26)
    fgrid = 2.0 * xd[None, :] + 3.0 * yd[:, None]
    md.interp_data("z_b", xd, yd, fgrid)
    assert np.allclose(md.z_b.x.array, 2 * md.x + 3 * md.y)
    md.set_lake_bdry(np.array([[2e3, 2e3], [6e3, 2e3], [6e3, 6e3], [2e3, 6e3]]))
    inside = (md.x > 2e3) & (md.x < 6e3) & (md.y > 2e3) & (md.y < 6e3)
    assert np.array_equal(md.lake_bdry.x.array[inside], np.ones(inside.sum()))
    assert md.lake_bdry.x.array[(md.x < 1e3)].sum() == 0
    md.N_init.interpolate(lambda x: 5.0 + 0 * x[0])
    assert np.all(md.N_init.x.array == 5.0)
    md.q_init.sub(1).interpolate(lambda x: 0 * x[0] + 2.0)
    assert np.all(md.q_init.x.array[1::2] == 2.0) and np.all(md.q_init.x.array[0::2] == 0.0)


def test_get_bcs_matches_oracle():
    import solvers
    from oracle.shakti_oracle import dirichlet_dofs
    import setup_cooke2_like
    md = setup_cooke2_like.initialize(fem.comm_world())
    bcs = solvers.get_bcs(md)
    assert len(bcs) == 1 and bcs[0].value == 3.7e5
    ref = dirichlet_dofs(md.domain.xy, md.domain.cells, md.OutflowBoundary)
    assert np.array_equal(bcs[0].dofs, ref) and ref.size > 0
    md.outflow_on = False
    assert solvers.get_bcs(md) == []


@pytest.mark.parametrize("name,nv", [("setup_rect250k", 125751), ("setup_cooke2_like", 12321)])
def test_setups_initialize(name, nv):
    mod = importlib.import_module(name)
    md = mod.initialize(fem.comm_world())
    assert md.V.dofmap.index_map.size_global == nv
    assert md.setup_name == name and md.results_name and md.nt_save > 0
    assert md.timesteps.size % md.nt_save == 0


def test_dof_helpers():
    from dof_helpers import dofs_to_serial
    rng = np.random.default_rng(0)
    nodes = rng.random((50, 2)) * 1e5
    perm = rng.permutation(50)
    m = dofs_to_serial(nodes[perm], nodes)
    assert np.allclose(nodes[perm][m], nodes)


def test_synthetic_configs_sizes():
    from shakti_b200 import configs
    c = configs.rect_steady()
    assert c.n_vert == 125751 and c.cells.shape[0] == 250000
    c = configs.dofs16m(nside=40)
    assert c.n_vert == 1600 and np.allclose(c.dts(3), [360.0, 3600.0, 3600.0])
    c5 = configs.lakes_fill_drain(nside=80)
    assert c5.storage_on and c5.fields["storage"].sum() > 0
    lake = c5.meta["lake"] > 0
    pulse = configs.lake_pulse_inputs(c5, c5.meta["t_pulse"])
    assert np.allclose(pulse[lake], 11 * c5.meta["inputs0"][lake]) and np.allclose(pulse[~lake], c5.meta["inputs0"][~lake])


def test_background_gradient_and_flux_of_linear_head():
    """grad/dot of the mini-UFL on a linear field: exact constant gradient in every cell."""
    import constitutive as cst
    from shakti_b200.ufl_lite import as_expr
    xy, cells, *_ = make_case(nx=6, ny=4)
    mesh = fem.Mesh(xy, cells)
    V = fem.functionspace(mesh, ("CG", 1))
    zb, zs = fem.Function(V), fem.Function(V)
    zb.x.array[:] = 3.0 + 0.01 * xy[:, 0]
    zs.x.array[:] = 3.0 + 0.01 * xy[:, 0] + 1000.0 - 0.02 * xy[:, 1]
    g = cst.BackgroundGradient(zb, zs)
    assert np.allclose(g.val[..., 0], 0.01) and np.allclose(g.val[..., 1], -0.02 * 0.917)
    b = fem.Function(V)
    b.x.array[:] = 2e-3
    N = fem.Function(V)
    qv = fem.Function(fem.functionspace(mesh, fem.element('P', 'triangle', 1, shape=(2,))))
    q = cst.WaterFlux(b, cst.Head(N, zb, zs), cst.Reynolds(qv))
    K = cst.Transmissivity(2e-3, 0.0)
    assert np.allclose(q.val[..., 0], -K * 0.01) and np.allclose(q.val[..., 1], K * 0.02 * 0.917)
    # Closure is pointwise: A b N |N|^(n-1), signed in b and N
    N.x.array[:] = -4.0e5
    c = as_expr(cst.Closure(b, N))
    assert np.allclose(c.val, cst.A * 2e-3 * (-4.0e5) ** 3)


def test_main_usage_error():
    import subprocess
    r = subprocess.run([sys.executable, "main.py"], cwd=str(SRC), capture_output=True, text=True)
    assert r.returncode != 0 and "usage" in (r.stdout + r.stderr)


def test_async_saver_places_owned_slices_and_double_buffers(monkeypatch):
    """Host logic of the save path (solvers._AsyncSaver, reference solvers.py:199-225): rows are placed through the
    owned-index map, two buffer sets alternate, a row only reaches the arrays when flushed -- with a fake device
    model (no GPU): the 'device' state is a numpy array that keeps changing after each enqueue."""
    import solvers
    from shakti_b200 import capi

    class FakePinned:
        def __init__(self, n):
            self.array = np.zeros(int(n))

    class FakeComm:
        def gather(self, obj, root=0):
            return [obj]

    class FakeMd:
        size, rank, comm = 1, 0, FakeComm()

    class FakeModel:
        def __init__(self, perm):
            self.perm, self.state, self.inflight, self.waits = perm, None, None, 0

        def owned(self):
            return self.perm

        def save_outputs_async(self, b, N, qx, qy, owned_only=False):
            assert owned_only
            assert self.inflight is None, "a second enqueue before the previous copies were awaited"
            self.inflight = ([b, N, qx, qy], [f[self.perm].copy() for f in self.state])   # snapshot at enqueue time

        def wait_outputs(self):
            self.waits += 1
            if self.inflight is not None:
                bufs, vals = self.inflight
                for dst, v in zip(bufs, vals):
                    dst[:] = v
                self.inflight = None

    monkeypatch.setattr(capi, "PinnedArray", FakePinned)
    nd, nrows = 37, 4
    rng = np.random.default_rng(3)
    model = FakeModel(rng.permutation(nd).astype(np.int32))
    saver = solvers._AsyncSaver(FakeMd(), model)
    arrays = tuple(np.full((nrows, nd), np.nan) for _ in range(4))
    truth = []
    for j in range(nrows):
        model.state = [rng.standard_normal(nd) for _ in range(4)]
        truth.append([f.copy() for f in model.state])
        saver.flush(arrays)                      # previous row comes home
        if j > 0:
            assert all(np.array_equal(arrays[k][j - 1], truth[j - 1][k]) for k in range(4))
        saver.enqueue(j)
        assert np.isnan(arrays[0][j]).all()      # not there before the flush
        model.state = [f + 100.0 for f in model.state]   # the run goes on: the snapshot must not see this
    saver.flush(arrays)
    saver.flush(arrays)                          # idempotent
    for j in range(nrows):
        for k in range(4):
            assert np.array_equal(arrays[k][j], truth[j][k])
    assert saver.sets[0][0] is not saver.sets[1][0]


def test_newton_solver_forwards_dolfinx_attributes():
    """B200NewtonSolver mirrors the DOLFINx NewtonSolver attributes a user may set after pde_solver() returned
    (reference solvers.py:52 leaves them at their defaults): rtol, atol, max_it and relaxation_parameter reach the
    library's options before the next solve, and only when they changed.  Fake model, no GPU."""
    import solvers
    from types import SimpleNamespace

    class FakeModel:
        def __init__(self):
            self.options = SimpleNamespace(newton_rtol=1e-9, newton_atol=1e-10, newton_max_it=50, newton_relaxation=1.0)
            self.pushed = []

        def set_options(self, **kw):
            self.pushed.append(kw)
            for k, v in kw.items():
                setattr(self.options, k, v)

        def newton_solve(self, dt):
            return 3, True

    s = solvers.B200NewtonSolver.__new__(solvers.B200NewtonSolver)
    s.model, s.dt, s.sync_host = FakeModel(), SimpleNamespace(value=3600.0), False
    assert s.solve(None) == (3, True) and s.model.pushed == []          # defaults: nothing to push
    s.relaxation_parameter = 0.8
    s.rtol = 1e-8
    s.solve(None)
    assert s.model.pushed == [dict(newton_rtol=1e-8, newton_atol=1e-10, newton_max_it=50, newton_relaxation=0.8)]
    s.solve(None)
    assert len(s.model.pushed) == 1                                      # unchanged: not pushed again
