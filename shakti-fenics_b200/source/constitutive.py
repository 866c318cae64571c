"""Constitutive relations of the SHAKTI model (parameter interface of the reference,
source/constitutive.py:6-41: same function names and argument order).

In the reference these return UFL expressions; here they return cell-vertex expressions of the
DOLFINx-free shim (shakti_b200.ufl_lite), which evaluate exactly what
``Expression(..., interpolation_points)`` evaluates.  They serve setups, diagnostics and tests
of the parameter interface.  The transient solver does not call them per step: the same laws
are implemented in the CUDA kernels (csrc/kernels.cu) with the constants of ``params.py``.
"""
import params as _p
from params import rho_i, rho_w, g, nu, omega, Lh, A, n
from shakti_b200.ufl_lite import grad, dot, div


def Head(N, z_b, z_s):
    """Hydraulic head [m] from effective pressure N, bed z_b and surface z_s:
    h = z_b + (rho_i/rho_w)(z_s - z_b) - N/(rho_w g)."""
    flotation = z_b + (rho_i / rho_w) * (z_s - z_b)
    return flotation - N / (rho_w * g)


def WaterFlux(b, h, Re):
    """Water discharge [m^2/s]: q = -|b|^3 g grad(h) / (12 nu (1 + omega Re))."""
    numerator = -(abs(b) ** 3) * g * grad(h)
    denominator = 12 * nu * (1 + omega * Re)
    return numerator / denominator


def Reynolds(q):
    """Local Reynolds number |q|/nu [-]."""
    return dot(q, q) ** 0.5 / nu


def Melt(q, h, G, b_n, melt_n):
    """Melt rate [kg m^-2 s^-1]: geothermal + dissipation, plus the lateral-melt diffusion term
    of Warburton et al. (2024), div(b m grad b / (1 + |grad b|^2))."""
    dissipation = rho_w * g * dot(q, grad(h))
    m0 = (G - dissipation) / Lh
    slope2 = dot(grad(b_n), grad(b_n))
    m_diff = div(b_n * melt_n * grad(b_n) / (1 + slope2))
    return m0 + m_diff


def Closure(b, N):
    """Viscous creep closure [m/s]: A b N |N|^(n-1)."""
    return A * b * N * abs(N) ** (n - 1)


def BackgroundGradient(z_b, z_s):
    """Hydraulic gradient of the zero-effective-pressure state [-]."""
    return grad(Head(0 * z_b, z_b, z_s))


def BackgroundPotential(z_b, z_s):
    """rho_w g Head(N = 0) [Pa]."""
    return rho_w * g * Head(0 * z_b, z_b, z_s)


def Transmissivity(b, q_norm):
    """K(b, |q|) = |b|^3 g / (12 nu (1 + omega |q|/nu)) for numpy arrays (diagnostics)."""
    import numpy as np
    return np.abs(b) ** 3 * _p.g / (12 * _p.nu * (1 + _p.omega * np.asarray(q_norm) / _p.nu))
