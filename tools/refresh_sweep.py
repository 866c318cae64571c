"""How often must the AMG numbers be renewed?  C4 on one GPU: the same 12 time steps (state rolled back) with
amg_refresh_every = 2, 3, 4, 6; prints ms/step, Newton and Krylov counts, refreshes.  No torch import (start-up time)."""
import sys, json, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "shakti-fenics_b200"))
from shakti_b200 import capi, configs
t0 = time.perf_counter()
nside = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
nst = int(sys.argv[2]) if len(sys.argv) > 2 else 12
case = configs.dofs16m(nside=nside, nsteps=nst + 12)
m = capi.Model(case.xy, case.cells)
configs.apply_case(m, case)
dts = case.dts()
m.run(dts[:5])
m.snapshot()
print(json.dumps(dict(setup_and_warmup_s=round(time.perf_counter() - t0, 1))), flush=True)
for every in (2, 4, 3, 6, 2):
    m.rollback()
    m.set_options(amg_refresh_every=every)
    s0 = m.stats()
    its, ms = m.run_timed(dts[5:5 + nst])
    s1 = m.stats()
    print(json.dumps(dict(every=every, ms_per_step=round(ms / nst, 3), newton=int(sum(its)), krylov=int(s1["linear_its"] - s0["linear_its"]),
                          refreshes=int(s1["amg_refreshes"] - s0["amg_refreshes"]))), flush=True)
