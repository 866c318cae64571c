# C3 of BASELINE.json: synthetic ice-sheet margin mesh, 4M triangles, 1000 steps, turbulent K(b,Re)
from _synthetic import md_from_case
from shakti_b200 import configs


def initialize(comm):
    case = configs.margin_turbulent(nx=2000, ny=1000, nsteps=1000)
    return md_from_case(comm, case, __file__, nt_save=100)
