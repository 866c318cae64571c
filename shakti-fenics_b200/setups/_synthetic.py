"""Shared code of the synthetic setup modules: turn a shakti_b200.configs.Case into a
``model_setup`` object the way setups/setup_cooke2.py (reference) builds one."""
import os
import sys
from pathlib import Path

_HERE = Path(__file__).resolve().parent
for _p in (str(_HERE.parent / "source"), str(_HERE.parent)):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
from model_setup import model_setup  # noqa: E402
from shakti_b200.fem import Mesh  # noqa: E402


def md_from_case(comm, case, setup_file, nt_save=24, nt_check=None, results_name=None, solver_options=None):
    domain = Mesh(case.xy, case.cells, comm)
    md = model_setup(comm, domain)
    md.setup_name = os.path.splitext(os.path.basename(setup_file))[0]
    md.lake_name = case.name
    md.N_bdry = case.N_bdry
    parent_dir = Path(setup_file).resolve().parent.parent
    md.results_name = results_name or f'{parent_dir}/results/{md.lake_name}_{int(md.N_bdry/1e3):d}kpa'

    f = case.fields
    md.z_b.x.array[:] = f["z_b"]
    md.z_s.x.array[:] = f["z_s"]
    md.G.x.array[:] = f["G"]
    md.inputs.x.array[:] = f["inputs"]
    md.lake_bdry.x.array[:] = f["storage"]
    md.b_init.x.array[:] = f["b"]
    md.N_init.x.array[:] = f["N_n"]
    md.q_init.x.array[:] = f["q"].reshape(-1)

    # outflow boundary marker, x has shape (3, npoints) as in DOLFINx
    if "outflow_marker" in case.meta:
        marker = case.meta["outflow_marker"]
        lookup = {(round(float(a), 3), round(float(b), 3)): bool(m) for a, b, m in zip(case.xy[:, 0], case.xy[:, 1], marker)}
        md.OutflowBoundary = lambda x: np.array([lookup.get((round(float(a), 3), round(float(b), 3)), False)
                                                 for a, b in zip(x[0], x[1])])
    else:
        x_out = case.xy[:, 0].min()
        md.OutflowBoundary = lambda x: np.isclose(x[0], x_out)
    md.outflow_on = True
    md.storage_on = case.storage_on

    md.timesteps = case.timesteps
    md.nt_save = nt_save
    md.nt_check = nt_check if nt_check is not None else 50 * nt_save
    md.solver_options = dict(solver_options or {})
    return md
