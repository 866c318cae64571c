import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "shakti-fenics_b200"), str(ROOT / "tests")): sys.path.insert(0, p)
import numpy as np
from common import make_case, make_model, make_oracle, relinf
from shakti_b200 import capi
c = make_case(seed=2)
o = make_oracle(*c)
dts = o.dt_schedule(np.linspace(0, 8 * 3600.0, 9))[:6]
for kw in (dict(amg_smoother=0, amg_presmooth=1, amg_postsmooth=1, amg_strength_theta=0.0), dict(amg_smoother=0, amg_presmooth=1, amg_postsmooth=1),
           dict(amg_smoother=1), dict(amg_smoother=1, amg_strength_theta=0.0), dict(amg_smoother=1, amg_cheby_ratio=30.0), dict(amg_smoother=1, amg_max_levels=1)):
    m = make_model(*c, precond="amg", linear_max_it=300, **kw)
    out = []
    for dt in dts:
        k0 = m.stats()["linear_its"]
        try:
            it, cv = m.step(dt)
            out.append((it, m.stats()["linear_its"] - k0))
        except capi.ShaktiError as e:
            out.append(("FAIL", m.stats()["linear_its"] - k0, str(e)[-60:])); break
    print(kw, out, "refreshes", m.stats()["amg_refreshes"], flush=True)
    m.close()
