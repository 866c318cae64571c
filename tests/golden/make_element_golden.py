"""Generate tests/golden/element_golden.json: EXACT element integrals of the SHAKTI weak form.

The reference cannot run here (no FEniCSx), so these fixtures are the independent pin of the
oracle: the residual (reference source/solvers.py:45) and its N-derivative (solvers.py:51) are
written symbolically with sympy exactly as the UFL form reads, on single P1 triangles with
dyadic-rational data, and integrated EXACTLY over the reference triangle.  Only the
polynomial parts are exact; the transmissivity integral int K(b,|q|) dx is polynomial only for
q == 0 (case "q0"), otherwise the fixture stores everything except the K term ("poly" part).

Run:  python tests/golden/make_element_golden.py   (needs sympy; a few seconds)
"""
import json
from pathlib import Path

import numpy as np
import sympy as sp

P = dict(g=9.81, rho_i=917, rho_w=1000, nu=1.787e-6, Lh=3.34e5, omega=1e-3, n=3, A=2.24e-24)


def R(x):
    return sp.Rational(float(x))          # exact value of the double


def exact_element(xy, fld, dt, with_K):
    xi, eta = sp.symbols("xi eta")
    lam = [1 - xi - eta, xi, eta]
    g, rho_i, rho_w, nu, Lh, A = [R(P[k]) for k in ("g", "rho_i", "rho_w", "nu", "Lh", "A")]
    X = [[R(v) for v in p] for p in xy]
    d1 = [X[1][0] - X[0][0], X[1][1] - X[0][1]]
    d2 = [X[2][0] - X[0][0], X[2][1] - X[0][1]]
    det = d1[0] * d2[1] - d2[0] * d1[1]
    gp = [None, (d2[1] / det, -d2[0] / det), (-d1[1] / det, d1[0] / det)]
    gp[0] = (-gp[1][0] - gp[2][0], -gp[1][1] - gp[2][1])
    detabs = abs(det)

    def P1(name):
        return sum(R(fld[name][i]) * lam[i] for i in range(3))

    def grad(name_or_vals):
        vals = fld[name_or_vals] if isinstance(name_or_vals, str) else name_or_vals
        return (sum(R(vals[i]) * gp[i][0] for i in range(3)), sum(R(vals[i]) * gp[i][1] for i in range(3)))

    N, N_n, b, G, melt, sto, inp = [P1(k) for k in ("N", "N_n", "b", "G", "melt_n", "storage", "inputs")]
    qx, qy = P1("qx"), P1("qy")
    # Head (constitutive.py:6-9) nodal -> gradient
    hn = [R(fld["z_b"][i]) + (rho_i / rho_w) * (R(fld["z_s"][i]) - R(fld["z_b"][i])) - R(fld["N"][i]) / (rho_w * g)
          for i in range(3)]
    gh = (sum(hn[i] * gp[i][0] for i in range(3)), sum(hn[i] * gp[i][1] for i in range(3)))
    gb, gm = grad("b"), grad("melt_n")
    gb2 = gb[0] ** 2 + gb[1] ** 2
    m0 = (G - rho_w * g * (qx * gh[0] + qy * gh[1])) / Lh                 # constitutive.py:25
    mdiff = (gb2 * melt + b * (gm[0] * gb[0] + gm[1] * gb[1])) / (1 + gb2)  # div(b m grad b/(1+|grad b|^2)), P1
    clos = A * b * N * N ** 2                                             # constitutive.py:31, n = 3
    lake = sto * (1 / (rho_w * g * R(dt))) * (N - N_n)                    # solvers.py:42
    cm = 1 / rho_i - 1 / rho_w
    Rint = cm * (m0 + mdiff) - clos - lake - inp                          # solvers.py:45

    def integrate(expr):
        return sp.integrate(sp.integrate(sp.expand(expr), (eta, 0, 1 - xi)), (xi, 0, 1)) * detabs

    K = (sp.Abs(b) ** 3) * g / (12 * nu)                                  # q == 0 => Re = 0
    F, J = [], []
    for a in range(3):
        Fa = integrate(Rint * lam[a])
        if with_K:
            # -dot(water_flux, grad v) = K grad h . grad phi_a ; b > 0 on the element in case q0
            Fa += integrate((b ** 3) * g / (12 * nu)) * (gh[0] * gp[a][0] + gh[1] * gp[a][1])
        F.append(Fa)
        row = []
        for bb in range(3):
            dR = cm * (qx * gp[bb][0] + qy * gp[bb][1]) / Lh - (3 * A * b * N ** 2 + sto / (rho_w * g * R(dt))) * lam[bb]
            Jab = integrate(dR * lam[a])
            if with_K:
                Jab += -integrate((b ** 3) * g / (12 * nu)) / (rho_w * g) * (gp[a][0] * gp[bb][0] + gp[a][1] * gp[bb][1])
            row.append(Jab)
        J.append(row)
    return [float(sp.N(v, 30)) for v in F], [[float(sp.N(v, 30)) for v in r] for r in J]


def main():
    rng = np.random.default_rng(20261018)
    cases = []
    tris = [[(0.0, 0.0), (2000.0, 100.0), (300.0, 1800.0)],      # counter-clockwise
            [(5000.0, 1000.0), (4000.0, 3000.0), (7000.0, 2500.0)],  # clockwise
            [(-100.0, 50.0), (400.0, -300.0), (900.0, 700.0)]]
    for k, xy in enumerate(tris):
        for name in ("q0", "poly"):
            fld = dict(
                z_b=(50 * rng.standard_normal(3)).tolist(), z_s=(1000 + 100 * rng.random(3)).tolist(),
                N=(3.7e5 * (1 + 0.1 * rng.standard_normal(3))).tolist(), N_n=(3.7e5 * (1 + 0.1 * rng.standard_normal(3))).tolist(),
                b=(1e-3 * (1 + rng.random(3))).tolist(), G=(0.05 + 0.01 * rng.random(3)).tolist(),
                melt_n=(1e-6 * rng.random(3)).tolist(), storage=rng.random(3).tolist(), inputs=(1e-8 * rng.random(3)).tolist(),
                qx=(0 * rng.random(3)).tolist() if name == "q0" else (2e-3 * rng.standard_normal(3)).tolist(),
                qy=(0 * rng.random(3)).tolist() if name == "q0" else (2e-3 * rng.standard_normal(3)).tolist())
            if k == 2:
                fld["N"][0] = -fld["N"][0]      # sign change of N inside the element
            dt = 3600.0 if k != 1 else 360.0
            F, J = exact_element(xy, fld, dt, with_K=(name == "q0"))
            cases.append(dict(kind=name, xy=xy, fields=fld, dt=dt, F=F, J=J))
    out = Path(__file__).with_name("element_golden.json")
    out.write_text(json.dumps(dict(params=P, cases=cases), indent=1))
    print("wrote", out, len(cases), "cases")


if __name__ == "__main__":
    main()
