#!/usr/bin/env python
"""Exact-rational element matrices of the pressure operator on NON-ORTHOGONAL cells, derived directly from the weak forms
(NS.cc:1245-1248 cell, 1274-1287 interior face, 1311-1321 boundary id 1) with deal.II's general-mapping semantics and from nothing
in this repository (no oracle/ code, no quadrature, no metric tables).

  python tests/golden/make_exact_sheared.py          -> tests/golden/exact_sheared.npz

Mesh: 2 x 1 parallelograms, x = (X + s Y, Y) applied to the 2 x 1 Cartesian mesh with cells h_x x h_y; s = 3/4 makes every metric
quantity rational: J = [[h_x, s h_y], [0, h_y]], grad_x xi = (1, -s) / h_x with norm (5/4) / h_x, grad_x eta = (0, 1) / h_y, the
face xi = const has length h_y sqrt(1 + s^2) = (5/4) h_y and unit normal (4/5, -3/5), the face eta = const length h_x, normal (0, 1).
Everything is integrated in reference coordinates (xi, eta) in [0, 1]^2 per cell, where the integrands are polynomials (affine map).
Conventions as in make_exact.py; in addition: get_gradient = J^-T grad_xi, cell JxW = det J, face JxW = length of the
face image, penalty = C_p / 2 (|grad_x xi_d|_minus + |grad_x xi_d|_plus) on interior faces and C_p |grad_x xi_d| on the boundary.
Rotations / scalings keep cells orthogonal; this fixture pins what only shows on skewed cells: the off-diagonal metric terms, the
normal that is not a coordinate direction, the penalty scale |grad xi_d| != 1 / h_d and the face measure.
"""
import math
import os
import sys
from fractions import Fraction as Fr

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_exact import GL, Poly, X, Y, lagrange_1d, to_np   # exact 2-D polynomial arithmetic only (X = xi, Y = eta here)


def pressure_matrix_sheared(hx, hy, s, p, inv_coeff):
    n1 = p + 1
    n = n1 * n1
    nodes = GL[p]
    phi = [lagrange_1d(nodes, i, X) * lagrange_1d(nodes, j, Y) for j in range(n1) for i in range(n1)]   # on the reference cell
    # J^-T grad_xi: d/dx = (1/hx) d/dxi,  d/dy = (-s/hx) d/dxi + (1/hy) d/deta
    gx = lambda f: f.d(0) * (Fr(1) / hx)
    gy = lambda f: f.d(0) * (-s / hx) + f.d(1) * (Fr(1) / hy)
    det = hx * hy
    C_p = Fr(n1 * n1)                                                # NS.cc:292
    root = Fr(5, 4)
    assert root * root == 1 + s * s                                  # sqrt(1 + s^2) is rational
    A = [[Fr(0)] * (2 * n) for _ in range(2 * n)]
    for c in range(2):                                               # NS.cc:1245-1248
        for i, pi in enumerate(phi):
            for j, pj in enumerate(phi):
                f = gx(pi) * gx(pj) + gy(pi) * gy(pj) + pi * pj * inv_coeff
                A[c * n + i][c * n + j] += f.integrate_box(Fr(0), Fr(1), Fr(0), Fr(1)) * det
    # interior face: cell 0 (xi = 1, the interior side phi_p) | cell 1 (xi = 0, phi_m); n_plus = grad xi / |grad xi| = (1, -s) / root
    nx, ny = Fr(1) / root, -s / root
    length = hy * root                                               # face JxW (per unit of eta)
    sigma = C_p * Fr(1, 2) * (root / hx + root / hx)                 # NS.cc:1274-1275
    sides = ((+1, 0, Fr(1)), (-1, 1, Fr(0)))                         # (sign in the jump, cell, xi of the face in that cell)
    for side_j, cj, pos_j in sides:
        for j, pj in enumerate(phi):
            val = pj.restrict(0, pos_j)
            dn = (gx(pj) * nx + gy(pj) * ny).restrict(0, pos_j)      # grad u . n_plus
            jump = val * side_j
            for side_i, ci, pos_i in sides:
                for i, pi in enumerate(phi):
                    tv = pi.restrict(0, pos_i)
                    tdn = (gx(pi) * nx + gy(pi) * ny).restrict(0, pos_i)
                    # NS.cc:1284-1287: value test +-(-avg_grad . n + sigma jump); gradient test -theta/2 jump n . grad phi on both sides
                    f = tv * (dn * Fr(-1, 2) + jump * sigma) * side_i + tdn * (jump * Fr(-1, 2))
                    A[ci * n + i][cj * n + j] += f.integrate_line(1, Fr(0), Fr(1)) * length
    # boundary id 1: face xi = 1 of cell 1 (colorized box: the upper x face), NS.cc:1311-1321
    sigma_b = C_p * root / hx
    for j, pj in enumerate(phi):
        val, dn = pj.restrict(0, Fr(1)), (gx(pj) * nx + gy(pj) * ny).restrict(0, Fr(1))
        for i, pi in enumerate(phi):
            tv, tdn = pi.restrict(0, Fr(1)), (gx(pi) * nx + gy(pi) * ny).restrict(0, Fr(1))
            f = tv * (dn * (-1) + val * sigma_b) + tdn * (val * (-1))
            A[n + i][n + j] += f.integrate_line(1, Fr(0), Fr(1)) * length
    return A


def main():
    hx, hy, s = Fr(1, 2), Fr(1, 3), Fr(3, 4)
    dt = 5e-3
    gamma = 2.0 - math.sqrt(2.0)
    out = {"hx": float(hx), "hy": float(hy), "shear": float(s), "dt": dt}
    for p in (1, 2):
        K = to_np(pressure_matrix_sheared(hx, hy, s, p, Fr(0)))
        M = to_np(pressure_matrix_sheared(hx, hy, s, p, Fr(1))) - K
        out[f"pressure_q{p}_K"], out[f"pressure_q{p}_M"] = K, M
        for stage in (1, 2):
            coeff = 1.0e6 * gamma * dt * gamma * dt if stage == 1 else 1.0e6 * (1.0 - gamma) * dt * (1.0 - gamma) * dt   # NS.cc:1237
            out[f"pressure_q{p}_stage{stage}"] = K + M / coeff
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "exact_sheared.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", v) for k, v in out.items() if "stage" in k})


if __name__ == "__main__":
    main()
