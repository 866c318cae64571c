"""DOLFINx-free stand-ins for the handful of FEniCSx/mpi4py objects the reference's
``model_setup.py`` / ``solvers.py`` / setup modules touch.

Host-side containers only: meshes and P1 nodal arrays in numpy.  Nothing here computes the
hot path; that is the CUDA library behind ``shakti_b200.capi``.

Mirrors (reference file:line):
  * ``MPI.COMM_WORLD`` ................ source/main.py:11  -> ``Comm``
  * ``domain.geometry.x`` / topology .. source/model_setup.py:26-28 -> ``Mesh``
  * ``functionspace(domain, ("CG",1))`` source/model_setup.py:29 -> ``functionspace``
  * ``basix.ufl.element('P', cell, 1, shape=(2,))`` model_setup.py:30 -> ``element``
  * ``Function`` (.x.array, .x.scatter_forward, .interpolate, .sub) model_setup.py:44-51
  * ``V.dofmap.index_map`` .............. source/model_setup.py:108-116
"""
import numpy as np


class Comm:
    """One process per GPU.  With a single process this is the trivial communicator; under
    torchrun it forwards to torch.distributed (object collectives, host side only)."""

    def __init__(self):
        self._dist = None
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self._dist = dist
        except Exception:                                           # pragma: no cover
            self._dist = None

    def Get_rank(self):
        return self._dist.get_rank() if self._dist else 0

    def Get_size(self):
        return self._dist.get_world_size() if self._dist else 1

    def barrier(self):
        if self._dist:
            self._dist.barrier()

    Barrier = barrier

    def gather(self, obj, root=0):
        if not self._dist:
            return [obj]
        out = [None] * self.Get_size() if self.Get_rank() == root else None
        self._dist.gather_object(obj, out, dst=root)
        return out

    def bcast(self, obj, root=0):
        if not self._dist:
            return obj
        box = [obj]
        self._dist.broadcast_object_list(box, src=root)
        return box[0]


COMM_WORLD = None


def comm_world():
    global COMM_WORLD
    if COMM_WORLD is None:
        COMM_WORLD = Comm()
    return COMM_WORLD


class _Geometry:
    def __init__(self, xy):
        self.x = np.zeros((xy.shape[0], 3))
        self.x[:, :2] = xy
        self.dim = 2


class _Topology:
    dim = 2


class Mesh:
    """P1 triangle mesh: ``geometry.x`` (Nv,3) and ``cells`` (Ne,3) int32.  Every process holds
    the whole mesh (the CUDA library partitions it internally), so there are no ghosts here."""

    def __init__(self, xy, cells, comm=None):
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        self.geometry = _Geometry(xy[:, :2])
        self.topology = _Topology()
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.comm = comm or comm_world()

    def basix_cell(self):
        return "triangle"

    @property
    def xy(self):
        return self.geometry.x[:, :2]


class _IndexMap:
    def __init__(self, n):
        self.size_local = n
        self.size_global = n
        self.num_ghosts = 0
        self.ghosts = np.zeros(0, dtype=np.int64)

    def global_to_local(self, g):
        return np.asarray(g, dtype=np.int32)


class _DofMap:
    def __init__(self, mesh, bs):
        self.index_map = _IndexMap(mesh.geometry.x.shape[0])
        self.index_map_bs = bs
        self.list = mesh.cells


class _Element:
    def __init__(self, bs):
        self.bs = bs

    def interpolation_points(self):
        return np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])


class FunctionSpace:
    def __init__(self, mesh, bs=1):
        self.mesh = mesh
        self.bs = bs
        self.dofmap = _DofMap(mesh, bs)
        self.element = _Element(bs)

    def tabulate_dof_coordinates(self):
        return self.mesh.geometry.x


def element(family, cell, degree, shape=None):
    assert degree == 1 and family in ("P", "CG", "Lagrange"), "only P1 is supported"
    return ("P", 1, shape)


def functionspace(mesh, el):
    """``functionspace(domain, ("CG", 1))`` or with a vector ``element(..., shape=(2,))``."""
    shape = el[2] if len(el) > 2 else None
    bs = int(shape[0]) if shape else 1
    assert el[1] == 1, "only P1 is supported"
    return FunctionSpace(mesh, bs)


class _Vector:
    def __init__(self, n):
        self.array = np.zeros(n)

    def scatter_forward(self):      # no ghosts on the host side
        pass


class _SubFunction:
    """View of one component of a blocked (vector-P1) Function."""

    def __init__(self, parent, i):
        self.parent, self.i = parent, i
        self.function_space = FunctionSpace(parent.function_space.mesh, 1)

    @property
    def values(self):
        return self.parent.x.array[self.i::self.parent.function_space.bs]

    def interpolate(self, u):
        bs = self.parent.function_space.bs
        self.parent.x.array[self.i::bs] = _evaluate(u, self.function_space)


def _evaluate(u, V):
    X = V.mesh.geometry.x
    if isinstance(u, Function):
        assert u.function_space.bs == 1
        return u.x.array.copy()
    if isinstance(u, _SubFunction):
        return np.array(u.values)
    if callable(u):
        vals = np.asarray(u(X.T), dtype=np.float64)      # DOLFINx passes x with shape (3, npoints)
        return np.broadcast_to(vals, (X.shape[0],)).copy() if vals.ndim <= 1 else vals
    return np.full(X.shape[0], float(u))


class Function:
    """P1 nodal function; vector functions are stored blocked [x0,y0,x1,y1,...] like DOLFINx."""

    def __init__(self, V, name=None):
        self.function_space = V
        self.x = _Vector(V.mesh.geometry.x.shape[0] * V.bs)
        self.name = name

    def interpolate(self, u):
        V = self.function_space
        if V.bs == 1:
            self.x.array[:] = _evaluate(u, V)
        else:
            if isinstance(u, Function):
                self.x.array[:] = u.x.array
            else:
                vals = np.asarray(u(V.mesh.geometry.x.T), dtype=np.float64)   # (bs, npoints)
                self.x.array[:] = vals.T.reshape(-1)

    def sub(self, i):
        return _SubFunction(self, i)

    def copy(self):
        f = Function(self.function_space, self.name)
        f.x.array[:] = self.x.array
        return f

    # arithmetic builds cell-vertex expressions (see ufl_lite), as UFL does for dolfinx Functions
    def _e(self):
        from .ufl_lite import as_expr
        return as_expr(self)

    def __add__(self, o): return self._e() + o
    def __radd__(self, o): return o + self._e()
    def __sub__(self, o): return self._e() - o
    def __rsub__(self, o): return o - self._e()
    def __mul__(self, o): return self._e() * o
    def __rmul__(self, o): return o * self._e()
    def __truediv__(self, o): return self._e() / o
    def __rtruediv__(self, o): return o / self._e()
    def __neg__(self): return -self._e()
    def __abs__(self): return abs(self._e())
    def __pow__(self, p): return self._e() ** p


def boundary_facets(cells):
    """(m,2) vertex pairs of the facets that belong to exactly one cell."""
    e = np.concatenate([cells[:, [0, 1]], cells[:, [1, 2]], cells[:, [2, 0]]]).astype(np.int64)
    e.sort(axis=1)
    key = e[:, 0] * (int(cells.max()) + 1) + e[:, 1]
    _, idx, cnt = np.unique(key, return_index=True, return_counts=True)
    return e[idx[cnt == 1]]


def locate_entities_boundary(mesh, dim, marker):
    """Boundary facets (as vertex pairs) whose vertices all satisfy ``marker(x)``, x (3,n)."""
    assert dim == 1
    f = boundary_facets(mesh.cells)
    m = np.asarray(marker(mesh.geometry.x.T), dtype=bool)
    return f[m[f[:, 0]] & m[f[:, 1]]]


def locate_dofs_topological(V, dim, facets):
    return np.unique(np.asarray(facets).ravel()).astype(np.int32)


class DirichletBC:
    def __init__(self, value, dofs, V):
        self.value, self.dofs, self.function_space = float(value), np.asarray(dofs, dtype=np.int32), V


def dirichletbc(value, dofs, V):
    return DirichletBC(value, dofs, V)
