"""A very small stand-in for the UFL operators used by the reference's ``constitutive.py``
(``grad``, ``dot``, ``div``, ``abs``, ``**``): expressions of P1 ``Function``s are evaluated at the
three vertices of every cell, carrying exact first derivatives (forward mode), with the UFL
rule that second derivatives of P1 functions on affine cells vanish.

This is what ``Expression(expr, V.element.interpolation_points())`` produces in DOLFINx
(reference source/solvers.py:143-145,162,165); ``interpolate_expression`` then writes the
values cell by cell so the highest-index cell containing a vertex wins.

Host-side convenience for setups, plots and tests of the parameter interface; the solver's
hot path does NOT go through here (it runs the CUDA kernels behind shakti_b200.capi).
"""
import numpy as np

from .fem import Function, _SubFunction


def _cell_geometry(mesh):
    if getattr(mesh, "_gradphi", None) is None:
        X = mesh.geometry.x[:, :2][mesh.cells]
        d1, d2 = X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]
        det = d1[:, 0] * d2[:, 1] - d2[:, 0] * d1[:, 1]
        g = np.empty((mesh.cells.shape[0], 3, 2))
        g[:, 1, 0], g[:, 1, 1] = d2[:, 1] / det, -d2[:, 0] / det
        g[:, 2, 0], g[:, 2, 1] = -d1[:, 1] / det, d1[:, 0] / det
        g[:, 0] = -g[:, 1] - g[:, 2]
        mesh._gradphi = g
    return mesh._gradphi


class Scalar:
    """val (ne,3): value at each cell vertex; g (ne,3,2): gradient there."""

    def __init__(self, val, g, mesh):
        self.val, self.g, self.mesh = val, g, mesh

    @staticmethod
    def lift(u, like):
        if isinstance(u, Scalar):
            return u
        if isinstance(u, (Function, _SubFunction)):
            return as_expr(u)
        return Scalar(np.full(like.val.shape, float(u)), np.zeros(like.g.shape), like.mesh)

    def __add__(self, o):
        o = Scalar.lift(o, self)
        return Scalar(self.val + o.val, self.g + o.g, self.mesh)

    __radd__ = __add__

    def __neg__(self):
        return Scalar(-self.val, -self.g, self.mesh)

    def __sub__(self, o):
        return self + (-Scalar.lift(o, self))

    def __rsub__(self, o):
        return Scalar.lift(o, self) - self

    def __mul__(self, o):
        if isinstance(o, Vector):
            return o * self
        o = Scalar.lift(o, self)
        return Scalar(self.val * o.val, self.g * o.val[..., None] + self.val[..., None] * o.g, self.mesh)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = Scalar.lift(o, self)
        return Scalar(self.val / o.val, (self.g * o.val[..., None] - self.val[..., None] * o.g) / (o.val ** 2)[..., None],
                      self.mesh)

    def __rtruediv__(self, o):
        return Scalar.lift(o, self) / self

    def __abs__(self):
        return Scalar(np.abs(self.val), np.sign(self.val)[..., None] * self.g, self.mesh)

    def __pow__(self, p):
        p = float(p)
        with np.errstate(divide="ignore", invalid="ignore"):
            d = np.where(self.val != 0, p * np.power(self.val, p - 1), 0.0)
        return Scalar(np.power(self.val, p), d[..., None] * self.g, self.mesh)


class Vector:
    """val (ne,3,2), jac (ne,3,2,2) with jac[...,i,j] = d v_i / d x_j."""

    def __init__(self, val, jac, mesh):
        self.val, self.jac, self.mesh = val, jac, mesh

    def __mul__(self, s):
        if not isinstance(s, Scalar):
            return Vector(self.val * float(s), self.jac * float(s), self.mesh)
        return Vector(self.val * s.val[..., None],
                      self.jac * s.val[..., None, None] + self.val[..., :, None] * s.g[..., None, :], self.mesh)

    __rmul__ = __mul__

    def __neg__(self):
        return Vector(-self.val, -self.jac, self.mesh)

    def __truediv__(self, s):
        if not isinstance(s, Scalar):
            return self * (1.0 / float(s))
        return self * (1.0 / s)

    def __add__(self, o):
        return Vector(self.val + o.val, self.jac + o.jac, self.mesh)

    def __sub__(self, o):
        return Vector(self.val - o.val, self.jac - o.jac, self.mesh)

    def __getitem__(self, i):
        return Scalar(self.val[..., i], self.jac[..., i, :], self.mesh)


def as_expr(u):
    """P1 Function (scalar or blocked vector) -> Scalar / Vector of cell-vertex values."""
    if isinstance(u, (Scalar, Vector)):
        return u
    if isinstance(u, _SubFunction):
        mesh = u.parent.function_space.mesh
        nodal = np.asarray(u.values)
    else:
        mesh = u.function_space.mesh
        bs = u.function_space.bs
        if bs > 1:
            comps = [as_expr(u.sub(i)) for i in range(bs)]
            return Vector(np.stack([c.val for c in comps], -1), np.stack([c.g for c in comps], -2), mesh)
        nodal = u.x.array
    gp = _cell_geometry(mesh)
    v = nodal[mesh.cells]
    g = np.einsum("ea,eak->ek", v, gp)
    return Scalar(v, np.broadcast_to(g[:, None, :], (v.shape[0], 3, 2)).copy(), mesh)


def grad(f):
    f = as_expr(f)
    assert isinstance(f, Scalar)
    # second derivatives of P1 functions vanish cell-wise; for general expressions they are not tracked
    return Vector(f.g, np.zeros(f.g.shape + (2,)), f.mesh)


def dot(a, b):
    a, b = as_expr(a), as_expr(b)
    val = np.einsum("evi,evi->ev", a.val, b.val)
    g = np.einsum("evij,evi->evj", a.jac, b.val) + np.einsum("evi,evij->evj", a.val, b.jac)
    return Scalar(val, g, a.mesh)


def div(v):
    return Scalar(v.jac[..., 0, 0] + v.jac[..., 1, 1], np.zeros(v.val.shape), v.mesh)


def interpolate_expression(target, expr):
    """``target.interpolate(Expression(expr, interpolation_points))``: cell-by-cell write, the
    highest-index cell containing a vertex wins."""
    expr = as_expr(expr)
    mesh = expr.mesh
    bs = target.function_space.bs
    cells = mesh.cells
    flat = cells.ravel()
    if isinstance(expr, Scalar):
        assert bs == 1
        out = target.x.array
        out[flat] = expr.val.ravel()          # numpy assigns in order: the last (highest cell) value is kept
        # make the winner explicit (do not rely on assignment order)
        win = np.full(out.shape[0], -1, dtype=np.int64)
        np.maximum.at(win, flat, np.repeat(np.arange(cells.shape[0]), 3))
        ok = win >= 0
        loc = np.argmax(cells[win[ok]] == np.nonzero(ok)[0][:, None], axis=1)
        out[ok] = expr.val[win[ok], loc]
    else:
        for i in range(bs):
            comp = Function(type(target.function_space)(mesh, 1))
            interpolate_expression(comp, expr[i])
            target.x.array[i::bs] = comp.x.array
