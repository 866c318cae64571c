"""One time step of the C4 config between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "shakti-fenics_b200"))
import torch
from shakti_b200 import capi, configs

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 3
case = configs.dofs16m(nside=nside, nsteps=warm + 4)
m = capi.Model(case.xy, case.cells)
configs.apply_case(m, case)
dts = case.dts()
m.run(dts[:warm])
torch.cuda.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
t0 = time.perf_counter()
its, ms = m.run_timed(dts[warm:warm + 1])
rt.cudaProfilerStop()
st = m.stats()
print(json.dumps(dict(nside=nside, newton=int(its[0]), ms=ms, krylov_total=st["linear_its"], launches=st["kernel_launches"])))
