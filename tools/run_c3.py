"""C3 (BASELINE.json configs[2]): synthetic ice-sheet margin mesh, 4M P1 triangles (2 003 001 dofs), transient
1000 steps of 3600 s in the turbulent K(b,Re) regime, on one B200.

  python tools/run_c3.py [nsteps] [rebuild_every] [out_prefix]

Logs Newton / Krylov iteration counts, AMG refreshes and rebuilds, min/max of b and N and the step time every
`report` steps, checks every step that Newton converged, b >= b_min and all fields are finite, and saves the
fields after step 10 (compared on the CPU with the LU oracle by tools/check_c3_against_oracle.py)."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "shakti-fenics_b200"))
import numpy as np
from shakti_b200 import capi, configs

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rebuild_every = int(sys.argv[2]) if len(sys.argv) > 2 else 0
prefix = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/c3"
report = 50
t0 = time.perf_counter()
case = configs.margin_turbulent(nsteps=nsteps + 1)
m = capi.Model(case.xy, case.cells)
configs.apply_case(m, case)
dts = case.dts()
setup_s = time.perf_counter() - t0
print(json.dumps(dict(event="setup", dofs=case.n_vert, cells=int(case.cells.shape[0]), seconds=round(setup_s, 1))), flush=True)
log, t_run = [], 0.0
st_prev = m.stats()
theta = 0.08
for lo in range(0, nsteps, report):
    hi = min(nsteps, lo + report)
    if rebuild_every and lo > 0 and lo % rebuild_every == 0:
        theta = 0.08 if theta != 0.08 else 0.0800001      # any change of an AMG option drops the hierarchy: next solve re-aggregates
        tb = time.perf_counter()
        m.set_options(amg_strength_theta=theta)
    its, ms = m.run_timed(dts[lo:hi])
    t_run += ms / 1e3
    st = m.stats()
    b, N = m.get_field("b"), m.get_field("N")
    q = m.get_flux()
    ok = bool(np.isfinite(b).all() and np.isfinite(N).all() and np.isfinite(q).all() and b.min() >= 1e-5)
    rec = dict(event="steps", first=lo, last=hi - 1, ms_per_step=round(ms / (hi - lo), 2), newton=[int(i) for i in np.bincount(its)],
               newton_per_step=float(np.mean(its)), krylov_per_step=(st["linear_its"] - st_prev["linear_its"]) / (hi - lo),
               krylov_per_solve=(st["linear_its"] - st_prev["linear_its"]) / max(1, st["newton_its"] - st_prev["newton_its"]),
               amg_refreshes=st["amg_refreshes"] - st_prev["amg_refreshes"], amg_levels=st["amg_levels"],
               b_min=float(b.min()), b_max=float(b.max()), N_min=float(N.min()), N_max=float(N.max()),
               q_max=float(np.abs(q).max()), omega_Re_max=float(1e-3 * np.hypot(q[:, 0], q[:, 1]).max() / 1.787e-6), finite_and_clamped=ok)
    print(json.dumps(rec), flush=True)
    log.append(rec)
    st_prev = st
    assert ok, rec
    if lo == 0:
        # fields after step 10 for the CPU-side comparison with the LU oracle: rerun 10 steps on a second model
        m2 = capi.Model(case.xy, case.cells)
        configs.apply_case(m2, case)
        its10 = m2.run(dts[:10])
        np.savez_compressed(prefix + "_step10.npz", N=m2.get_field("N"), b=m2.get_field("b"), newton=its10)
        m2.close()
summary = dict(event="summary", steps=nsteps, rebuild_every=rebuild_every, seconds=round(t_run, 2), steps_per_s=round(nsteps / t_run, 2),
               newton_total=m.stats()["newton_its"], krylov_total=m.stats()["linear_its"], amg_refreshes=m.stats()["amg_refreshes"])
print(json.dumps(summary), flush=True)
Path(prefix + f"_log_rebuild{rebuild_every}.json").write_text(json.dumps(dict(setup_seconds=setup_s, log=log, summary=summary), indent=1))
