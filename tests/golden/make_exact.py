#!/usr/bin/env python
"""Exact-rational element matrices of the pressure and velocity operators, derived DIRECTLY from the weak forms of the
reference and from nothing in this repository (no oracle/ code, no 1-D tables, no quadrature, no tensor-product tricks).

  python tests/golden/make_exact.py          -> tests/golden/exact_matrices.npz

What is computed: on a 2 x 1 Cartesian mesh (cells h_x x h_y, rational) every entry A[i, j] = "i-th test function against the
j-th trial function" of

  pressure (NS_stage 2):  NS.cc:1245-1248 (cell), 1277-1287 (interior face), 1312-1321 (boundary id 1)
  velocity (NS_stage 1):  NS.cc:951-958 (cell, TR_BDF2_stage 1), 1017-1036 (interior face), constant extrapolated velocity

as integrals of 2-D polynomials in exact rational arithmetic (fractions.Fraction).  On Cartesian cells the reference's
Gauss quadrature (n_q = degree + 1 points) integrates these integrands exactly (degree <= 2 p per variable), so the exact
integrals ARE what deal.II computes, up to round-off.  FE_DGQ(p) = Lagrange basis on the Gauss-Lobatto points, which are
rational for p = 1 ({0, 1}) and p = 2 ({0, 1/2, 1}).  deal.II conventions used (documented behaviour of the library, not code
of this repository): dofs numbered cell by cell, vector components blocked inside a cell, lexicographic inside a
component; phi_p = interior (first) cell of a face, phi_m = exterior, both evaluated with the normal n_plus of the
interior cell; on faces `inverse_jacobian(0)[dim - 1]` is the row of the face-normal direction, i.e. 1 / h_normal.
The irrational scalars (gamma = 2 - sqrt 2) only multiply whole rational matrices; those products are formed in float64.
"""
import math
import os
from fractions import Fraction as Fr

import numpy as np


# ---- 2-D polynomials with rational coefficients: {(ex, ey): coeff} -----------------------------------------------------
class Poly:
    def __init__(self, c=None):
        self.c = {k: v for k, v in (c or {}).items() if v != 0}

    @staticmethod
    def const(v):
        return Poly({(0, 0): Fr(v)})

    def __add__(self, o):
        o = o if isinstance(o, Poly) else Poly.const(o)
        r = dict(self.c)
        for k, v in o.c.items():
            r[k] = r.get(k, 0) + v
        return Poly(r)

    __radd__ = __add__

    def __neg__(self):
        return Poly({k: -v for k, v in self.c.items()})

    def __sub__(self, o):
        return self + (-(o if isinstance(o, Poly) else Poly.const(o)))

    def __mul__(self, o):
        if not isinstance(o, Poly):
            return Poly({k: v * Fr(o) for k, v in self.c.items()})
        r = {}
        for (a, b), v in self.c.items():
            for (c, d), w in o.c.items():
                r[(a + c, b + d)] = r.get((a + c, b + d), 0) + v * w
        return Poly(r)

    __rmul__ = __mul__

    def d(self, var):
        r = {}
        for (a, b), v in self.c.items():
            e = (a, b)[var]
            if e:
                k = (a - 1, b) if var == 0 else (a, b - 1)
                r[k] = r.get(k, 0) + v * e
        return Poly(r)

    def integrate_box(self, x0, x1, y0, y1):
        s = Fr(0)
        for (a, b), v in self.c.items():
            s += v * (x1 ** (a + 1) - x0 ** (a + 1)) / (a + 1) * (y1 ** (b + 1) - y0 ** (b + 1)) / (b + 1)
        return s

    def restrict(self, var, value):
        """polynomial in the OTHER variable (kept in its slot) on the line var = value"""
        r = {}
        for (a, b), v in self.c.items():
            k = (0, b) if var == 0 else (a, 0)
            r[k] = r.get(k, 0) + v * value ** ((a, b)[var])
        return Poly(r)

    def integrate_line(self, var, lo, hi):
        """integral over the remaining variable `var` of a restricted polynomial"""
        s = Fr(0)
        for (a, b), v in self.c.items():
            e = (a, b)[var]
            s += v * (hi ** (e + 1) - lo ** (e + 1)) / (e + 1)
        return s


X, Y = Poly({(1, 0): Fr(1)}), Poly({(0, 1): Fr(1)})


def lagrange_1d(nodes, i, t):
    """l_i(t) for the polynomial argument t (a Poly in one variable)"""
    r = Poly.const(1)
    for m, xm in enumerate(nodes):
        if m != i:
            r = r * (t - xm) * (Fr(1) / (nodes[i] - xm))
    return r


GL = {1: [Fr(0), Fr(1)], 2: [Fr(0), Fr(1, 2), Fr(1)]}


class Cell:
    def __init__(self, x0, y0, hx, hy, p):
        self.x0, self.y0, self.hx, self.hy = x0, y0, hx, hy
        nodes = GL[p]
        xi, eta = (X - x0) * (Fr(1) / hx), (Y - y0) * (Fr(1) / hy)
        # lexicographic: index = i + n * j, x fastest
        self.phi = [lagrange_1d(nodes, i, xi) * lagrange_1d(nodes, j, eta) for j in range(p + 1) for i in range(p + 1)]

    def box(self):
        return self.x0, self.x0 + self.hx, self.y0, self.y0 + self.hy


def face_integral(f, d, pos, lo, hi):
    """integral of the polynomial f over the face {x_d = pos}, other variable in [lo, hi]"""
    return f.restrict(d, pos).integrate_line(1 - d, lo, hi)


def mesh_2x1(hx, hy, p, periodic):
    cells = [Cell(Fr(0), Fr(0), hx, hy, p), Cell(hx, Fr(0), hx, hy, p)]
    # interior faces: (interior cell, exterior cell, direction d, position on the interior cell, position on the exterior cell)
    faces = [(0, 1, 0, hx, hx)]
    if periodic:
        faces.append((1, 0, 0, 2 * hx, Fr(0)))                     # wrap-around in x: upper face of cell 1 = lower face of cell 0
        faces += [(c, c, 1, hy, Fr(0)) for c in (0, 1)]            # one cell in y: its upper face meets its own lower face
    return cells, faces


def pressure_matrix(hx, hy, p, periodic, inv_coeff):
    """NS.cc:1230-1326.  Non-periodic: colorized boundary ids 0 (x = 0), 1 (x = 2 h_x), 2, 3; only id 1 has a term."""
    cells, faces = mesh_2x1(hx, hy, p, periodic)
    n = (p + 1) ** 2
    C_p = Fr((p + 1) ** 2)                                          # NS.cc:292
    A = [[Fr(0)] * (2 * n) for _ in range(2 * n)]
    h = (hx, hy)
    for c, cell in enumerate(cells):                               # NS.cc:1245-1248
        for i, pi in enumerate(cell.phi):
            for j, pj in enumerate(cell.phi):
                f = pi.d(0) * pj.d(0) + pi.d(1) * pj.d(1) + pi * pj * inv_coeff
                A[c * n + i][c * n + j] += f.integrate_box(*cell.box())
    for (ci, ce, d, pos_i, pos_e) in faces:                        # NS.cc:1277-1287
        sigma = C_p * Fr(1, 2) * (1 / h[d] + 1 / h[d])
        o_lo, o_hi = (cells[ci].y0, cells[ci].y0 + hy) if d == 0 else (cells[ci].x0, cells[ci].x0 + hx)
        for side_j, cj, pos_j in ((+1, ci, pos_i), (-1, ce, pos_e)):          # trial function lives on the interior / exterior cell
            for j, pj in enumerate(cells[cj].phi):
                # traces of the trial function on the face, as polynomials of the tangential variable
                val = pj.restrict(d, pos_j)
                dn = pj.d(d).restrict(d, pos_j)                                # d/dn with n = n_plus = +e_d
                avg_grad_n = Fr(1, 2) * dn
                jump = val * side_j
                for side_i, cti, pos_t in ((+1, ci, pos_i), (-1, ce, pos_e)):
                    for i, pi in enumerate(cells[cti].phi):
                        tv = pi.restrict(d, pos_t)
                        tdn = pi.d(d).restrict(d, pos_t)
                        # submit_value: +(-avg.n + sigma jump) on phi_p, the negative on phi_m; submit_gradient(-theta/2 jump n) on both
                        f = tv * (avg_grad_n * (-1) + jump * sigma) * side_i + tdn * (jump * Fr(-1, 2))
                        A[cti * n + i][cj * n + j] += f.integrate_line(1 - d, o_lo, o_hi)
    if not periodic:                                               # NS.cc:1306-1321, boundary id 1 = face x = 2 h_x of cell 1
        cell, d, pos = cells[1], 0, 2 * hx
        sigma = C_p * (1 / h[d])
        for j, pj in enumerate(cell.phi):
            val, dn = pj.restrict(d, pos), pj.d(d).restrict(d, pos)
            for i, pi in enumerate(cell.phi):
                tv, tdn = pi.restrict(d, pos), pi.d(d).restrict(d, pos)
                f = tv * (dn * (-1) + val * sigma) + tdn * (val * (-1))
                A[n + i][n + j] += f.integrate_line(1, cell.y0, cell.y0 + hy)
    return A


def velocity_parts(hx, hy, p, ue):
    """NS.cc:951-958 and 1017-1036 on the fully periodic 2 x 1 mesh with the constant extrapolated velocity ue = (ue_x, ue_y):
    A = 1/(gamma dt) M + a22/Re L + a22 Cv   (M, L, Cv rational)"""
    cells, faces = mesh_2x1(hx, hy, p, True)
    n = (p + 1) ** 2
    nd = 2 * n                                                      # dofs per cell: component-blocked
    C_u = Fr((p + 1) ** 2)                                          # NS.cc:293 with fe_degree_v = p
    h = (hx, hy)
    Z = lambda: [[Fr(0)] * (2 * nd) for _ in range(2 * nd)]
    M, L, Cv = Z(), Z(), Z()
    for c, cell in enumerate(cells):
        for comp in range(2):
            for i, pi in enumerate(cell.phi):
                for j, pj in enumerate(cell.phi):
                    r, s = c * nd + comp * n + i, c * nd + comp * n + j
                    M[r][s] += (pi * pj).integrate_box(*cell.box())                                    # submit_value(u / (gamma dt))
                    L[r][s] += (pi.d(0) * pj.d(0) + pi.d(1) * pj.d(1)).integrate_box(*cell.box())     # grad phi : grad u
                    # grad phi : (-(u (x) ue)):  sum_d d_d phi_comp * u_comp * ue_d
                    Cv[r][s] += (-(pi.d(0) * ue[0] + pi.d(1) * ue[1]) * pj).integrate_box(*cell.box())
    for (ci, ce, d, pos_i, pos_e) in faces:
        sigma = C_u * Fr(1, 2) * (1 / h[d] + 1 / h[d])
        lam = abs(ue[d])                                            # max(|ue_p . n|, |ue_m . n|), constant field
        o_lo, o_hi = (cells[ci].y0, cells[ci].y0 + hy) if d == 0 else (cells[ci].x0, cells[ci].x0 + hx)
        for comp in range(2):
            for side_j, cj, pos_j in ((+1, ci, pos_i), (-1, ce, pos_e)):
                for j, pj in enumerate(cells[cj].phi):
                    val, dn = pj.restrict(d, pos_j), pj.d(d).restrict(d, pos_j)
                    jump = val * side_j
                    for side_i, cti, pos_t in ((+1, ci, pos_i), (-1, ce, pos_e)):
                        for i, pi in enumerate(cells[cti].phi):
                            tv, tdn = pi.restrict(d, pos_t), pi.d(d).restrict(d, pos_t)
                            r, s = cti * nd + comp * n + i, cj * nd + comp * n + j
                            # a22/Re (-avg_grad n + sigma jump) on the value test (sign by side), -theta a22/Re 1/2 jump on d/dn phi
                            fL = tv * (dn * Fr(-1, 2) + jump * sigma) * side_i + tdn * (jump * Fr(-1, 2))
                            L[r][s] += fL.integrate_line(1 - d, o_lo, o_hi)
                            # a22 (avg (u (x) ue) n + 1/2 lambda jump):  (u (x) ue) n = u (ue . n)
                            fC = tv * (val * (Fr(1, 2) * ue[d]) + jump * (Fr(1, 2) * lam)) * side_i
                            Cv[r][s] += fC.integrate_line(1 - d, o_lo, o_hi)
    return M, L, Cv


def to_np(A):
    return np.array([[float(v) for v in row] for row in A], dtype=np.float64)


def main():
    hx, hy = Fr(1, 2), Fr(1, 3)
    dt, Re = 5e-3, 100.0
    gamma = 2.0 - math.sqrt(2.0)
    out = {"hx": float(hx), "hy": float(hy), "dt": dt, "Re": Re}
    for p in (1, 2):
        for periodic in (True, False):
            # 1/coeff enters linearly: A = K + (1/coeff) Mass, both rational
            K = pressure_matrix(hx, hy, p, periodic, Fr(0))
            KM = pressure_matrix(hx, hy, p, periodic, Fr(1))
            Kn, Mn = to_np(K), to_np(KM) - to_np(K)
            tag = f"pressure_q{p}_{'periodic' if periodic else 'box'}"
            out[tag + "_K"], out[tag + "_M"] = Kn, Mn
            for stage in (1, 2):
                coeff = 1.0e6 * gamma * dt * gamma * dt if stage == 1 else 1.0e6 * (1.0 - gamma) * dt * (1.0 - gamma) * dt   # NS.cc:1237
                out[f"{tag}_stage{stage}"] = Kn + Mn / coeff
    ue = (Fr(3, 4), Fr(-1, 3))
    out["ue"] = np.array([float(ue[0]), float(ue[1])])
    for p in (1, 2):
        M, L, Cv = (to_np(a) for a in velocity_parts(hx, hy, p, ue))
        a22 = 0.5                                                   # NS.cc:287
        out[f"velocity_q{p}_stage1"] = M / (gamma * dt) + a22 / Re * L + a22 * Cv
        out[f"velocity_q{p}_M"], out[f"velocity_q{p}_L"], out[f"velocity_q{p}_C"] = M, L, Cv
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "exact_matrices.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", v) for k, v in out.items() if "stage" in k})


if __name__ == "__main__":
    main()
