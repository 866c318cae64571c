"""Model input container with the reference's interface (source/model_setup.py:18-119): class
``model_setup`` with the same attribute names, defaults and helper methods, so setup modules
written for the reference keep working.  It sits on the DOLFINx-free shim (shakti_b200.fem): the
Functions are host-side nodal arrays; the solver uploads them to the GPU when ``solve`` starts.

Attribute map (reference line in parentheses)
  comm, rank, size (21-23)       domain, x, y (26-28)        V, V_flux (29-30)      mask (31)
  OutflowBoundary (32)           bounds (36-37)               outflow_on, storage_on (40-41)
  z_b z_s G inputs b_init N_init q_init lake_bdry (44-51)     N_bdry (52)            b_min (53)
  outline (56)                   lake_name results_name setup_name (59-61)
  timesteps nt_save nt_check (64-66)
"""
import os
import sys

import numpy as np
from scipy.interpolate import RegularGridInterpolator

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from shakti_b200 import capi  # noqa: E402
from shakti_b200.fem import Function, functionspace, element  # noqa: E402


def _gpu_available():
    """Setup modules may be prepared on a machine without a GPU (the solve itself cannot run there)."""
    try:
        return capi.device_count() > 0
    except Exception:
        return False

from solvers import solve  # noqa: E402

# scalar P1 input fields: attribute name -> meaning [unit]
_SCALAR_FIELDS = {
    "z_b": "bed elevation [m]",
    "z_s": "surface elevation [m]",
    "G": "geothermal heat flux [W/m^2]",
    "inputs": "water input to the bed, e.g. moulins [m/s]",
    "b_init": "initial gap height [m]",
    "N_init": "initial effective pressure [Pa]",
    "lake_bdry": "lake indicator: 1 inside the lake, 0 outside",
}


def get_nested_attr(obj, attr_path):
    """``get_nested_attr(md, "z_b.x.array")`` -> md.z_b.x.array"""
    for name in attr_path.split('.'):
        obj = getattr(obj, name)
    return obj


def set_array_slice(obj, attr_path, values):
    get_nested_attr(obj, attr_path)[:] = values


def points_in_polygon(px, py, poly):
    """Even-odd rule, vectorised over the points; ``poly`` is an (m,2) array of vertices."""
    poly = np.asarray(poly, dtype=np.float64)
    xa, ya = poly[:, 0], poly[:, 1]
    xb, yb = np.roll(xa, -1), np.roll(ya, -1)
    inside = np.zeros(np.shape(px), dtype=bool)
    for x0, y0, x1, y1 in zip(xa, ya, xb, yb):
        straddles = (y0 > py) != (y1 > py)
        with np.errstate(divide="ignore", invalid="ignore"):
            x_cross = x0 + (py - y0) * (x1 - x0) / (y1 - y0)
        inside ^= straddles & (px < x_cross)
    return inside


class model_setup:
    def __init__(self, comm, domain):
        # process group: one process per GPU (an MPI communicator in the reference)
        self.comm, self.rank, self.size = comm, comm.Get_rank(), comm.Get_size()

        # mesh, node coordinates, P1 spaces for scalars and for the flux vector
        self.domain = domain
        self.x, self.y = domain.geometry.x[:, 0], domain.geometry.x[:, 1]
        self.V = functionspace(domain, ("CG", 1))
        self.V_flux = functionspace(domain, element('P', domain.basix_cell(), 1, shape=(domain.geometry.dim,)))
        self.mask = self.ghost_mask(self.V)
        self.OutflowBoundary = None                 # callable x(3,n) -> bool[n], set by the setup module

        # box used to crop gridded data sets before interpolation, padded by get_buffer()
        pad = self.get_buffer()
        self.bounds = [self.x.min() - pad, self.x.max() + pad, self.y.min() - pad, self.y.max() + pad]

        self.outflow_on = True                      # Dirichlet N = N_bdry on the outflow boundary
        self.storage_on = True                      # lake represented by a storage term

        for name in _SCALAR_FIELDS:                 # z_b, z_s, G, inputs, b_init, N_init, lake_bdry
            setattr(self, name, Function(self.V))
        self.q_init = Function(self.V_flux)         # initial water flux [m^2/s]
        self.N_bdry = 0.0                           # effective pressure on the outflow boundary [Pa]
        self.b_min = 1.0e-5                         # lower bound of the gap height [m]

        self.outline = None                         # lake outline (GeoDataFrame, or an (m,2) polygon here)
        self.lake_name = self.results_name = self.setup_name = None
        self.timesteps = self.nt_save = self.nt_check = None

        # extension: options of include/shakti_b200.h (shakti_options) for this run
        self.solver_options = {}

    # ------------------------------------------------------------------ input helpers
    def set_lake_bdry(self, outline):
        """Lake indicator at the mesh nodes from an outline (reference model_setup.py:68-72)."""
        if hasattr(outline, "geometry"):            # geopandas object, as in the reference (needs shapely)
            from shapely import Point
            xyz = self.domain.geometry.x
            for j in range(self.lake_bdry.x.array.size):
                self.lake_bdry.x.array[j] = outline.geometry.contains(Point(xyz[j, 0], xyz[j, 1])).iloc[0]
        else:                                       # plain polygon vertices
            if _gpu_available():                    # even-odd test on the device (csrc: points_in_polygon_kernel)
                self.lake_bdry.x.array[:] = capi.points_in_polygon(self.x, self.y, outline)
            else:
                self.lake_bdry.x.array[:] = points_in_polygon(self.x, self.y, outline)
        self.lake_bdry.x.scatter_forward()

    def interp_data(self, var_name, x_d, y_d, f):
        """Bilinear interpolation of a gridded field f[y, x] onto the nodes of Function ``var_name``
        (extrapolating outside the grid); returns the interpolant (reference model_setup.py:74-91)."""
        xmin, xmax, ymin, ymax = self.bounds
        keep_x = (x_d >= xmin) & (x_d <= xmax)
        keep_y = (y_d >= ymin) & (y_d <= ymax)
        x_sub, y_sub, f_sub = x_d[keep_x], y_d[keep_y], f[np.ix_(keep_y, keep_x)]
        interpolant = RegularGridInterpolator((x_sub, y_sub), f_sub.T, bounds_error=False, fill_value=None)
        if _gpu_available():                        # same bilinear / extrapolating rule on the device (interp_grid_kernel)
            values = capi.interp_grid(self.x, self.y, x_sub, y_sub, f_sub)
        else:
            values = interpolant(np.column_stack((self.x, self.y)))
        set_array_slice(self, f"{var_name}.x.array", values)
        get_nested_attr(self, f"{var_name}.x").scatter_forward()
        return interpolant

    def get_buffer(self):
        """Ten times the largest gap between distinct node coordinates, larger of the two axes
        (reference model_setup.py:93-106)."""
        gathered = [self.comm.gather(c[self.mask], root=0) for c in (self.x, self.y)]
        pads = (0, 0)
        if self.rank == 0:
            pads = tuple(10 * np.max(np.diff(np.unique(np.concatenate(g)))) for g in gathered)
        self.comm.barrier()
        pads = [self.comm.bcast(p, root=0) for p in pads]
        return np.max(pads)

    def ghost_mask(self, V):
        """True for owned dofs, False for ghosts (reference model_setup.py:108-116).  Every process holds
        the whole mesh here, so all entries are True; the CUDA library partitions internally."""
        imap = V.dofmap.index_map
        mask = np.ones(imap.size_local + imap.num_ghosts, dtype=bool)
        mask[imap.global_to_local(imap.ghosts)] = False
        return mask

    def solve(self):
        solve(self)
