"""End to end through the reference-shaped entry points on the GPU: setups -> model_setup ->
solvers.solve(md) -> .npy files, compared with the oracle stepping the same inputs."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "shakti-fenics_b200" / "source", ROOT / "shakti-fenics_b200" / "setups"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

from common import make_oracle, relinf  # noqa: E402

pytestmark = pytest.mark.gpu


def small_case():
    from shakti_b200 import configs
    c = configs.cooke2_like(nsteps=12)
    # shrink: keep the construction (random b_init incl. negative values, lake storage, potential-based outflow)
    return c


def build_md(tmp_path, nt_save=3):
    from _synthetic import md_from_case
    from shakti_b200 import fem
    case = small_case()
    case.timesteps = np.linspace(0, 11 * 3600.0, 12)
    md = md_from_case(fem.comm_world(), case, str(ROOT / "shakti-fenics_b200" / "setups" / "setup_cooke2_like.py"),
                      nt_save=nt_save, nt_check=2 * nt_save, results_name=str(tmp_path / "results" / "run"),
                      solver_options=dict(linear_rtol=1e-13))
    return md, case


def test_solve_md_outputs_match_oracle(tmp_path):
    import solvers
    md, case = build_md(tmp_path)
    solvers.solve(md)
    out = Path(md.results_name)
    for f in ("t.npy", "nodes_x.npy", "nodes_y.npy", "b.npy", "N.npy", "qx.npy", "qy.npy", "setup_cooke2_like.py"):
        assert (out / f).exists(), f
    nt, nd = 12, case.n_vert
    b, N, qx, qy = (np.load(out / f"{k}.npy") for k in ("b", "N", "qx", "qy"))
    assert b.shape == N.shape == qx.shape == qy.shape == (nt // md.nt_save, nd)
    assert np.allclose(np.load(out / "t.npy"), np.linspace(0, md.timesteps.max(), nt // md.nt_save))
    assert np.array_equal(np.load(out / "nodes_x.npy"), md.x)
    # oracle with the same Dirichlet dofs, saving after steps 0, nt_save, ...
    bc = solvers.get_bcs(md)[0].dofs
    o = make_oracle(case.xy, case.cells, case.fields, bc, case.N_bdry)
    dts = o.dt_schedule(md.timesteps)
    j = 0
    for i, dt in enumerate(dts):
        o.step(dt)
        if i % md.nt_save == 0:
            # the step-0 state (random b_init with negative values) is very badly conditioned:
            # LU and Krylov agree to ~1e-6 there, and to 1e-8 once b has been clamped
            tol = 1e-5 if i == 0 else 1e-8
            assert relinf(N[j], o.N) < tol, (i, relinf(N[j], o.N))
            assert relinf(b[j], o.b) < tol
            assert relinf(np.stack([qx[j], qy[j]], 1), o.q) < 10 * tol
            j += 1
    assert j == nt // md.nt_save
    assert relinf(md.final["N"].x.array, o.N) < 1e-8 and relinf(md.final["b"].x.array, o.b) < 1e-8


def test_existing_results_directory_exits_with_code_1(tmp_path, capsys):
    import solvers
    md, _ = build_md(tmp_path)
    Path(md.results_name).mkdir(parents=True)
    with pytest.raises(SystemExit) as e:
        solvers.solve(md)
    assert e.value.code == 1
    assert "already exists" in capsys.readouterr().out


def test_pde_solver_object(tmp_path):
    """pde_solver(...).solve(N) -> (niter, converged) with N synchronised back to the host."""
    import solvers
    from shakti_b200.fem import Function
    md, case = build_md(tmp_path)
    N, N_n, b, melt_n, storage = (Function(md.V) for _ in range(5))
    q = Function(md.V_flux)
    b.interpolate(md.b_init)
    N_n.interpolate(md.N_init)
    dt = solvers.Constant(md.domain, 360.0)
    solver = solvers.pde_solver(md, N, N_n, b, q, melt_n, md.lake_bdry, dt)
    assert np.array_equal(N.x.array, N_n.x.array)          # solvers.py:48
    niter, converged = solver.solve(N)
    assert converged and niter >= 1 and isinstance(niter, int)
    bc = solvers.get_bcs(md)[0].dofs
    assert np.allclose(N.x.array[bc], md.N_bdry)
    o = make_oracle(case.xy, case.cells, case.fields, bc, case.N_bdry)
    it_o, _ = o.newton(360.0)
    assert niter == it_o and relinf(N.x.array, o.N) < 1e-5
    solver.max_it = 0
    solver.rtol = solver.atol = 0.0
    with pytest.raises(RuntimeError):
        solver.solve(N)


def test_main_cli(tmp_path):
    """python3 main.py <setup_module> run from source/ (reference source/main.py, notebooks/example.ipynb:61)."""
    import subprocess
    src = ROOT / "shakti-fenics_b200" / "source"
    setup = tmp_path / "setup_cli_case.py"
    setup.write_text(
        "import sys\n"
        f"sys.path.insert(0, {str(ROOT / 'shakti-fenics_b200' / 'setups')!r})\n"
        "import numpy as np\n"
        "from _synthetic import md_from_case\n"
        "from shakti_b200 import configs\n\n"
        "def initialize(comm):\n"
        "    case = configs.rect_steady(nx=40, ny=20, nsteps=8)\n"
        f"    return md_from_case(comm, case, __file__, nt_save=4, results_name={str(tmp_path / 'out')!r})\n")
    import os
    # extend, never replace: CI / driver hooks may already export PYTHONPATH
    env = {**os.environ, "PYTHONPATH": os.pathsep.join(filter(None, [str(tmp_path), os.environ.get("PYTHONPATH")]))}
    r = subprocess.run([sys.executable, "main.py", "setup_cli_case"], cwd=str(src), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    N = np.load(tmp_path / "out" / "N.npy")
    assert N.shape == (2, 41 * 21) and np.isfinite(N).all() and np.abs(N[1] - 0.37e6).max() > 1.0
    # a second run into the same directory must refuse, with exit code 1 (solvers.py:91-102)
    r2 = subprocess.run([sys.executable, "main.py", "setup_cli_case"], cwd=str(src), env=env, capture_output=True, text=True, timeout=600)
    assert r2.returncode == 1 and "already exists" in r2.stdout


def test_time_dependent_inputs_and_resume(tmp_path):
    """Extensions of SURVEY §8(f)4: md.inputs_of_t and checkpoint/resume give the same fields as one straight run."""
    import solvers
    from _synthetic import md_from_case
    from shakti_b200 import configs, fem

    def make(results, nsteps):
        case = configs.lakes_fill_drain(nside=40, nsteps=13)
        case.meta["t_pulse"], case.meta["tau"] = 5 * 3600.0, 3 * 3600.0
        md = md_from_case(fem.comm_world(), case, str(ROOT / "shakti-fenics_b200" / "setups" / "setup_lakes64m.py"), nt_save=2,
                          nt_check=2, results_name=str(results))
        md.timesteps = case.timesteps[:nsteps]
        md.inputs_of_t = lambda t: configs.lake_pulse_inputs(case, t)
        return md, case

    md, case = make(tmp_path / "straight", 12)
    solvers.solve(md)
    N_ref, b_ref = np.load(tmp_path / "straight" / "N.npy"), np.load(tmp_path / "straight" / "b.npy")
    assert np.abs(N_ref[-1] - N_ref[0]).max() > 0
    # forcing really is time dependent: a run with static inputs differs
    md0, _ = make(tmp_path / "static", 12)
    md0.inputs_of_t = None
    solvers.solve(md0)
    assert np.abs(np.load(tmp_path / "static" / "N.npy")[-1] - N_ref[-1]).max() > 1e-8 * np.abs(N_ref[-1]).max()
    # stop after 8 steps (last checkpoint after step index 6), resume to 12
    md1, _ = make(tmp_path / "resumed", 8)
    md1.resume = True
    solvers.solve(md1)
    md2, _ = make(tmp_path / "resumed", 12)
    md2.resume = True
    solvers.solve(md2)
    N2, b2 = np.load(tmp_path / "resumed" / "N.npy"), np.load(tmp_path / "resumed" / "b.npy")
    assert N2.shape == N_ref.shape
    assert relinf(N2[-1], N_ref[-1]) < 1e-8 and relinf(b2[-1], b_ref[-1]) < 1e-8


def test_step_counter_and_amg_refresh_policy_on_the_split_path(tmp_path):
    """solvers.solve(md) drives newton_solve / update_* separately: the library's step counter (and with it
    amg_refresh_every) must advance exactly as it does through shakti_run."""
    import solvers
    from shakti_b200 import capi, configs
    md, case = build_md(tmp_path)
    md.solver_options = dict(amg_refresh_every=2)
    solvers.solve(md)
    st = md.solver.model.stats()
    nt = md.timesteps.size
    assert st["steps"] == nt
    m = capi.Model(case.xy, case.cells, amg_refresh_every=2, b_min=float(md.b_min))
    try:
        configs.apply_case(m, case)
        bc = solvers.get_bcs(md)[0].dofs
        m.set_dirichlet(bc, md.N_bdry)
        m.start()
        dts = np.concatenate([[0.1 * 3600.0], np.full(nt - 1, 3600.0)])
        m.run(dts)
        st2 = m.stats()
        assert st2["steps"] == nt
        assert abs(st["amg_refreshes"] - st2["amg_refreshes"]) <= 1, (st["amg_refreshes"], st2["amg_refreshes"])
        assert nt // 2 <= st2["amg_refreshes"] < nt
    finally:
        m.close()
