"""Model input container (parameter interface of the reference, source/model_setup.py:18-119):
same class name, attributes, defaults and helper methods, re-hosted on the DOLFINx-free shim
(shakti_b200.fem) so that setup modules written for the reference keep their shape.
"""
import os
import sys

import numpy as np
from scipy.interpolate import RegularGridInterpolator

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from shakti_b200.fem import Function, functionspace, element  # noqa: E402
from solvers import solve  # noqa: E402


def get_nested_attr(obj, attr_path):
    for attr in attr_path.split('.'):
        obj = getattr(obj, attr)
    return obj


def set_array_slice(obj, attr_path, values):
    arr = get_nested_attr(obj, attr_path)
    arr[:] = values


def points_in_polygon(px, py, poly):
    """Even-odd rule, vectorised over points; ``poly`` is an (m,2) vertex array."""
    poly = np.asarray(poly, dtype=np.float64)
    x0, y0 = poly[:, 0], poly[:, 1]
    x1, y1 = np.roll(x0, -1), np.roll(y0, -1)
    inside = np.zeros(px.shape, dtype=bool)
    for a, b, c, d in zip(x0, y0, x1, y1):
        crosses = ((b > py) != (d > py))
        with np.errstate(divide="ignore", invalid="ignore"):
            xint = (c - a) * (py - b) / (d - b) + a
        inside ^= crosses & (px < xint)
    return inside


class model_setup:
    def __init__(self, comm, domain):
        # process group (one process per GPU; reference: MPI communicator)
        self.comm = comm
        self.rank = comm.Get_rank()
        self.size = comm.Get_size()

        # domain, coordinates, function spaces
        self.domain = domain
        self.x = domain.geometry.x[:, 0]
        self.y = domain.geometry.x[:, 1]
        self.V = functionspace(domain, ("CG", 1))
        self.V_flux = functionspace(domain, element('P', domain.basix_cell(), 1, shape=(domain.geometry.dim,)))
        self.mask = self.ghost_mask(self.V)
        self.OutflowBoundary = None

        # bounding box (with buffer) used when interpolating gridded data
        buffer = self.get_buffer()
        self.bounds = [self.x.min() - buffer, self.x.max() + buffer,
                       self.y.min() - buffer, self.y.max() + buffer]

        # boundary-condition options
        self.outflow_on = True                  # Dirichlet N = N_bdry on the outflow boundary
        self.storage_on = True                  # lake storage term

        # physical input functions
        self.z_b = Function(self.V)             # bed elevation [m]
        self.z_s = Function(self.V)             # surface elevation [m]
        self.G = Function(self.V)               # geothermal heat flux [W/m^2]
        self.inputs = Function(self.V)          # water input to the bed [m/s]
        self.b_init = Function(self.V)          # initial gap height [m]
        self.N_init = Function(self.V)          # initial effective pressure [Pa]
        self.q_init = Function(self.V_flux)     # initial water flux [m^2/s]
        self.lake_bdry = Function(self.V)       # 1 inside the lake, 0 outside
        self.N_bdry = 0.0                       # effective pressure on the outflow boundary [Pa]
        self.b_min = 1.0e-5                     # lower bound of the gap height [m]

        # lake outline (GeoDataFrame in the reference; an (m,2) polygon array also works here)
        self.outline = None

        # names
        self.lake_name = None
        self.results_name = None
        self.setup_name = None

        # time stepping and output cadence
        self.timesteps = None
        self.nt_save = None
        self.nt_check = None

        # B200 solver options (extension; see include/shakti_b200.h shakti_options)
        self.solver_options = {}

    def set_lake_bdry(self, outline):
        if hasattr(outline, "geometry"):        # geopandas path of the reference (needs shapely)
            from shapely import Point
            for j in range(self.lake_bdry.x.array.size):
                point = Point(self.domain.geometry.x[j, 0], self.domain.geometry.x[j, 1])
                self.lake_bdry.x.array[j] = outline.geometry.contains(point).iloc[0]
        else:
            self.lake_bdry.x.array[:] = points_in_polygon(self.x, self.y, outline)
        self.lake_bdry.x.scatter_forward()

    def interp_data(self, var_name, x_d, y_d, f):
        # subset of the grid covering the (buffered) domain
        in_x = (x_d >= self.bounds[0]) & (x_d <= self.bounds[1])
        in_y = (y_d >= self.bounds[2]) & (y_d <= self.bounds[3])
        x_sub, y_sub = x_d[in_x], y_d[in_y]
        f_sub = f[np.ix_(in_y, in_x)]

        # bilinear interpolation (extrapolating outside the grid), evaluated at the mesh nodes
        f_interp = RegularGridInterpolator((x_sub, y_sub), f_sub.T, bounds_error=False, fill_value=None)
        values = f_interp(np.column_stack((self.x, self.y)))

        set_array_slice(self, f"{var_name}.x.array", values)
        get_nested_attr(self, f"{var_name}.x").scatter_forward()
        return f_interp

    def get_buffer(self):
        # ten times the largest gap between distinct node coordinates, per axis
        x_bfr, y_bfr = 0, 0
        x__ = self.comm.gather(self.x[self.mask], root=0)
        y__ = self.comm.gather(self.y[self.mask], root=0)
        if self.rank == 0:
            x__ = np.unique(np.concatenate(x__))
            y__ = np.unique(np.concatenate(y__))
            x_bfr = 10 * np.max(np.diff(x__))
            y_bfr = 10 * np.max(np.diff(y__))
        self.comm.barrier()
        x_bfr, y_bfr = self.comm.bcast(x_bfr, root=0), self.comm.bcast(y_bfr, root=0)
        return np.max([x_bfr, y_bfr])

    def ghost_mask(self, V):
        ghosts = V.dofmap.index_map.ghosts
        ghosts_local = V.dofmap.index_map.global_to_local(ghosts)
        size_local = V.dofmap.index_map.size_local
        num_ghosts = V.dofmap.index_map.num_ghosts
        mask = np.ones(size_local + num_ghosts, dtype=bool)
        mask[ghosts_local] = False
        return mask

    def solve(self):
        solve(self)
