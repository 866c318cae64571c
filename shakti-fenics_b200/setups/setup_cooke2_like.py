# C1 of BASELINE.json: surrogate of setups/setup_cooke2.py (reference).  The Cook E2 mesh,
# BedMachine / ATL14 / AQ1 grids and the lake outline are not shipped with the reference, so a
# seeded synthetic stand-in of the same size and field ranges is used.
from _synthetic import md_from_case
from shakti_b200 import configs


def initialize(comm):
    days = 30                      # the reference runs 10*365 days; shortened default
    nt_per_day = 24
    case = configs.cooke2_like(nsteps=int(days * nt_per_day))
    t_final = (days / 365) * 3.154e7
    import numpy as np
    case.timesteps = np.linspace(0, t_final, int(days * nt_per_day))
    return md_from_case(comm, case, __file__, nt_save=nt_per_day, nt_check=50 * nt_per_day)
