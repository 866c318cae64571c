import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "shakti-fenics_b200")): sys.path.insert(0, p)
import numpy as np
from shakti_b200 import capi, configs
case = configs.rect_steady(nx=40, ny=20, nsteps=8)
for kw in (dict(), dict(amg_cuda_graph=0), dict(amg_fp32_cycle=0), dict(amg_smoother=0), dict(amg_fp32_cycle=0, amg_cuda_graph=0), dict(precond="jacobi")):
    m = capi.Model(case.xy, case.cells, linear_max_it=300, **kw)
    configs.apply_case(m, case)
    out = []
    for dt in case.dts(6):
        k0 = m.stats()["linear_its"]
        try:
            it, cv = m.step(dt); out.append((it, m.stats()["linear_its"] - k0))
        except capi.ShaktiError as e:
            out.append(("FAIL", m.stats()["linear_its"] - k0, str(e)[-50:])); break
    print(kw, out, "levels", m.stats()["amg_levels"], flush=True)
    m.close()
