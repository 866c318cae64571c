"""ctypes binding of oracle/shakti_oracle_c.c (TEST INFRASTRUCTURE ONLY, like everything under oracle/).

The C file is the per-quadrature-point algorithm of ``ShaktiOracle.element_FJ`` / ``kbar`` / the nodal
updates at compiled speed with an OpenMP loop over cells: the stand-in for the reference's FFCx-generated
element kernels when bench.py times the CPU baseline.  ``available()`` is False until ``make -C oracle`` (run by
``__graft_entry__.build()``) has produced the library; nothing here falls back silently -- callers ask."""
import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "_build" / "libshakti_oracle_c.so"
_lib = None


def available():
    return LIB_PATH.exists()


def load():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} not built: run `make -C oracle`")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.shakti_oracle_c_threads.restype = C.c_int
    return _lib


def threads():
    return int(load().shakti_oracle_c_threads())


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _params(p):
    return _d([p.g, p.rho_i, p.rho_w, p.nu, p.Lh, p.omega, p.n, p.A])


def element_FJ(o, dt, N, want_J):
    """-> Fe (ne,3), Je (ne,3,3) or None; `o` is a ShaktiOracle."""
    lib = load()
    Fe = np.empty((o.ne, 3))
    Je = np.empty((o.ne, 3, 3)) if want_J else None
    a = [_d(x) for x in (o.xy, N, o.N_n, o.b, o.q, o.G, o.melt_n, o.storage, o.inputs, o.z_b, o.z_s)]
    pr, qp, qw = _params(o.p), _d(o.qpts), _d(o.qwts)
    lib.shakti_oracle_c_element_FJ(C.c_int64(o.ne), _p(o.cells), *[_p(x) for x in a], _p(pr), C.c_int(len(qw)), _p(qp), _p(qw),
                                   C.c_double(dt), _p(Fe), _p(Je) if want_J else None)
    return Fe, Je


def kbar(o):
    lib = load()
    out = np.empty(o.ne)
    xy, b, q, pr, qp, qw = _d(o.xy), _d(o.b), _d(o.q), _params(o.p), _d(o.qpts), _d(o.qwts)
    lib.shakti_oracle_c_kbar(C.c_int64(o.ne), _p(o.cells), _p(xy), _p(b), _p(q), _p(pr), C.c_int(len(qw)), _p(qp), _p(qw), _p(out))
    return out


def nodal_updates(o, dt):
    """q, melt_n, b after solvers.py:186-197 (clamped) from the oracle's current N and old q, melt_n, b."""
    lib = load()
    wc = np.ascontiguousarray(o.win_cell, dtype=np.int64)
    wl = np.ascontiguousarray(o.win_loc, dtype=np.int64)
    a = [_d(x) for x in (o.xy, o.N, o.b, o.q, o.G, o.melt_n, o.z_b, o.z_s)]
    pr = _params(o.p)
    qn, mn, bn = np.empty((o.nv, 2)), np.empty(o.nv), np.empty(o.nv)
    lib.shakti_oracle_c_nodal_updates(C.c_int64(o.nv), _p(o.cells), _p(wc), _p(wl), *[_p(x) for x in a], _p(pr),
                                      C.c_double(dt), C.c_double(o.b_min), _p(qn), _p(mn), _p(bn))
    return qn, mn, bn
