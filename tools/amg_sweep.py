"""Sweep linear-solver options on the GPU: iterations and wall time of one Jacobian solve."""
import sys, time, json, itertools
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "shakti-fenics_b200"))
import numpy as np
from shakti_b200 import capi, configs

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
case = configs.dofs16m(nside=nside, nsteps=8)
m = capi.Model(case.xy, case.cells, precond="amg")
configs.apply_case(m, case)
its = m.run(case.dts(2))
print("warm steps newton its", its, m.stats()["linear_its"], flush=True)
F, _ = m.assemble(3600.0, want_J=False)
rhs = F.copy()
import os
cfgs = []
for pre, post in ((2, 2), (1, 2), (2, 1), (1, 3), (2, 3), (3, 3), (0, 3), (0, 2)):
    cfgs.append(dict(linear_solver="gmres", amg_smoother=1, amg_presmooth=pre, amg_postsmooth=post))
cfgs.append(dict(linear_solver="bicgstab", amg_smoother=1, amg_presmooth=2, amg_postsmooth=2))
cfgs.append(dict(linear_solver="gmres", amg_smoother=1, amg_presmooth=2, amg_postsmooth=2, amg_prolong_omega=0.55))
cfgs.append(dict(linear_solver="gmres", amg_smoother=1, amg_presmooth=2, amg_postsmooth=2, amg_prolong_omega=0.8))
cfgs.append(dict(linear_solver="gmres", amg_smoother=1, amg_presmooth=2, amg_postsmooth=2, amg_strength_theta=0.04))
cfgs.append(dict(linear_solver="gmres", amg_smoother=1, amg_presmooth=2, amg_postsmooth=2, amg_strength_theta=0.12))
for cfg in cfgs:
    full = dict(amg_cheby_ratio=5.0, amg_strength_theta=0.08, amg_prolong_omega=0.67); full.update(cfg)
    m.set_options(linear_rtol=1e-12, linear_max_it=400, **full)
    try:
        m.assemble(3600.0)
        t0 = time.perf_counter(); dx, it, rr = m.linear_solve(rhs); t1 = time.perf_counter() - t0   # includes AMG (re)build
        t2 = 1e9
        for _ in range(2):
            t0 = time.perf_counter(); dx, it, rr = m.linear_solve(rhs); t2 = min(t2, time.perf_counter() - t0)
        st = m.stats()
        print(json.dumps(dict(cfg=full, its=it, relres=rr, first_s=round(t1, 3), solve_ms=round(1e3 * t2, 2), levels=st["amg_levels"],
                              opc=round(st["amg_operator_complexity"], 3))), flush=True)
    except capi.ShaktiError as e:
        print("FAILED", full, e, flush=True)
