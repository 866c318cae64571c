"""Physical constants of the SHAKTI model (parameter interface of the reference,
source/params.py:4-11: same names, same values, same Python types).

They are NOT hard-coded in the CUDA kernels: ``solvers.pde_solver`` packs this module into the
``shakti_params`` struct of the C ABI (include/shakti_b200.h), so editing a value here changes
the device computation exactly as it changes the UFL forms in the reference.
"""

g = 9.81          # gravitational acceleration            [m s^-2]
rho_i = 917       # density of ice (int, as in the reference) [kg m^-3]
rho_w = 1000      # density of water (int)                 [kg m^-3]
nu = 1.787e-6     # kinematic viscosity of water           [m^2 s^-1]
Lh = 3.34e5       # latent heat of fusion                  [J kg^-1]
omega = 1e-3      # laminar/turbulent transition parameter of the flux law [-]
n = 3             # Glen's flow-law exponent (int: abs(N)**(n-1) is then a polynomial)
A = 2.24e-24      # Glen's flow-law rate factor            [Pa^-n s^-1]

NAMES = ("g", "rho_i", "rho_w", "nu", "Lh", "omega", "n", "A")


def as_dict():
    """The eight constants in the order of the C struct."""
    import sys
    mod = sys.modules[__name__]
    return {k: getattr(mod, k) for k in NAMES}
