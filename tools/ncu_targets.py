"""ncu target for the kernels without a round-2 capture: between cudaProfilerStart/Stop one nodal pass
(update_q_melt_kernel + update_b_kernel, twice) and then ONE time step that begins with an AMG refresh
(amg_spgemm_table_kernel) and runs the V-cycle's fp32 SpMV kernels.  Use with
  ncu --profile-from-start off --kernel-name-base demangled -k regex:"update_q_melt_kernel|update_b_kernel|amg_spgemm_table_kernel|spmv_sell_kernel<.int.[0-2], float, .int.1>" -c 16 ...
(the nodal kernels come first in the window, so a small -c reaches every family)."""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "shakti-fenics_b200"))
import torch
from shakti_b200 import capi, configs

nside = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
warm = 4                                   # amg_refresh_every = 2: step index 4 starts with a refresh
case = configs.dofs16m(nside=nside, nsteps=warm + 4)
m = capi.Model(case.xy, case.cells)
configs.apply_case(m, case)
dts = case.dts()
m.run(dts[:warm])
st0 = m.stats()
torch.cuda.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
nodal_ms = m.time_kernel("nodal", reps=1, dt=3600.0)
its, ms = m.run_timed(dts[warm:warm + 1])
rt.cudaProfilerStop()
st = m.stats()
print(json.dumps(dict(nside=nside, dofs=case.n_vert, nodal_ms=nodal_ms, step_ms=ms, newton=int(its[0]),
                      amg_refreshes_in_step=st["amg_refreshes"] - st0["amg_refreshes"], nodal_GB=m.kernel_bytes("nodal") / 1e9)))
