# C5 of BASELINE.json: 64M dofs (8000 x 8000 vertices) with 8 lake discs and storage, for 8 B200
from _synthetic import md_from_case
from shakti_b200 import configs


def initialize(comm):
    case = configs.lakes_fill_drain(nside=8000, nsteps=120)
    return md_from_case(comm, case, __file__, nt_save=60)
