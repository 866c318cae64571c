# C2 of BASELINE.json: synthetic rectangular glacier bed, 250k P1 triangles, steady melt forcing
from _synthetic import md_from_case
from shakti_b200 import configs


def initialize(comm):
    case = configs.rect_steady(nx=500, ny=250, nsteps=240)
    return md_from_case(comm, case, __file__, nt_save=24)
