"""Micro-benchmark the library's kernels on the C4 fields (CUDA events on the library stream)."""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "shakti-fenics_b200"))
from shakti_b200 import capi, configs
nside = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
kernels = sys.argv[2].split(",") if len(sys.argv) > 2 else ["assemble", "kbar", "nodal", "spmv"]
case = configs.dofs16m(nside=nside, nsteps=6)
m = capi.Model(case.xy, case.cells, **({"assembly_kernel": int(sys.argv[3])} if len(sys.argv) > 3 else {}))
configs.apply_case(m, case)
m.run(case.dts(2))
out = {}
for k in kernels:
    ms = m.time_kernel(k, reps=20); by = m.kernel_bytes(k)
    out[k] = dict(ms=round(ms, 4), GBps=round(by / 1e9 / (ms / 1e3), 1))
print(json.dumps(dict(nside=nside, **out)))
