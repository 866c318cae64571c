# C4 of BASELINE.json: 16M-dof synthetic mesh (4000 x 4000 vertices); run on 1/2/4/8 B200 with
#   torchrun --nproc-per-node N main.py setup_dofs16m
from _synthetic import md_from_case
from shakti_b200 import configs


def initialize(comm):
    case = configs.dofs16m(nside=4000, nsteps=120)
    return md_from_case(comm, case, __file__, nt_save=60)
