"""CPU side of the C3 evidence: the LU oracle takes the first 10 steps of C3 (2 003 001 dofs; ~15-30 s per sparse
LU, run in the build container) and is compared with the fields the B200 run saved (tools/run_c3.py)."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "shakti-fenics_b200"))
import numpy as np
from oracle.shakti_oracle import ShaktiOracle
from shakti_b200 import configs

gpu = np.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/c3_step10.npz")
nst = int(sys.argv[2]) if len(sys.argv) > 2 else 10
case = configs.margin_turbulent(nsteps=1001)
o = ShaktiOracle(case.xy, case.cells)
for k in ("z_b", "z_s", "G", "inputs", "storage", "b", "N_n", "melt_n"):
    getattr(o, k)[:] = case.fields[k]
o.q[:] = case.fields["q"]
o.set_dirichlet(case.bc_dofs, case.N_bdry)
o.start()
its = []
t0 = time.perf_counter()
for i, dt in enumerate(case.dts(nst)):
    its.append(o.step(dt)[0])
    print(f"oracle step {i}: {its[-1]} Newton its, {time.perf_counter() - t0:.0f} s", flush=True)
rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
out = dict(steps=nst, dofs=case.n_vert, newton_oracle=[int(i) for i in its], newton_gpu=[int(i) for i in gpu["newton"][:nst]],
           N_rel_err=rel(gpu["N"], o.N), b_rel_err=rel(gpu["b"], o.b), oracle_seconds=round(time.perf_counter() - t0, 1))
print(json.dumps(out))
Path("profiles/r2_c3_vs_oracle.json").write_text(json.dumps(out, indent=1))
