"""1-D building blocks of FE_DGQ / QGauss (oracle; test infrastructure only).

Follows deal.II's definitions (not in /root/reference):

* ``FE_DGQ<dim>(p)``: Lagrange polynomials on the p+1 Gauss-Lobatto points of
  [0,1] (p = 0: the midpoint), tensor product, lexicographic numbering.
* ``QGauss<1>(n_q)``: Gauss-Legendre on [0,1], weights sum to 1.  The reference
  uses QGauss(p+2) for velocity forms and QGauss(p+1) for pressure forms
  (navier_stokes_TRBDF2_DG.cc:2320-2325).

All 1-D tables are computed in extended precision (np.longdouble) and rounded
to double once, so that they are correct to the last bit or two.
"""
from __future__ import annotations

import functools

import numpy as np

LD = np.longdouble


def _legendre_and_derivs(n: int, t):
    """P_n(t), P_n'(t), P_n''(t) by the three-term recurrence (t in [-1,1])."""
    t = np.asarray(t, dtype=LD)
    p0 = np.ones_like(t)
    if n == 0:
        return p0, np.zeros_like(t), np.zeros_like(t)
    p1 = t.copy()
    for k in range(1, n):
        p0, p1 = p1, ((2 * k + 1) * t * p1 - k * p0) / LD(k + 1)
    # p1 = P_n, p0 = P_{n-1}
    d1 = n * (t * p1 - p0) / (t * t - 1)                       # P_n'
    d2 = (2 * t * d1 - n * (n + 1) * p1) / (1 - t * t)          # P_n'' from Legendre ODE
    return p1, d1, d2


@functools.lru_cache(maxsize=None)
def gauss_legendre(nq: int):
    """Gauss-Legendre points/weights on [0,1] (QGauss<1>(nq))."""
    k = np.arange(1, nq + 1, dtype=LD)
    t = -np.cos((4 * k - 1) * LD(np.pi) / (4 * nq + 2)).astype(LD)   # ascending
    for _ in range(100):
        p, dp, _ = _legendre_and_derivs(nq, t)
        dt = p / dp
        t = t - dt
        if np.max(np.abs(dt)) < 1e-19:
            break
    _, dp, _ = _legendre_and_derivs(nq, t)
    w = 2 / ((1 - t * t) * dp * dp)
    x = (t + 1) / 2
    w = w / 2
    # enforce exact symmetry (deal.II computes half and mirrors)
    x = (x + (1 - x[::-1])) / 2
    w = (w + w[::-1]) / 2
    return x, w


@functools.lru_cache(maxsize=None)
def gauss_lobatto_nodes(n: int):
    """The n = p+1 support points of FE_DGQ(p) on [0,1] (QGaussLobatto<1>(n); midpoint for n=1)."""
    if n == 1:
        return np.array([0.5], dtype=LD)
    if n == 2:
        return np.array([0.0, 1.0], dtype=LD)
    p = n - 1
    m = n - 2                                     # interior points: roots of P_p'
    k = np.arange(1, m + 1, dtype=LD)
    t = -np.cos(k * LD(np.pi) / p).astype(LD)     # Chebyshev-Lobatto initial guess, ascending
    for _ in range(100):
        _, d1, d2 = _legendre_and_derivs(p, t)
        dt = d1 / d2
        t = t - dt
        if np.max(np.abs(dt)) < 1e-19:
            break
    t = np.concatenate(([LD(-1)], t, [LD(1)]))
    x = (t + 1) / 2
    x = (x + (1 - x[::-1])) / 2
    x[0] = 0
    x[-1] = 1
    return x


def lagrange_eval(nodes, x):
    """L[q][i] = l_i(x_q), D[q][i] = l_i'(x_q) for the Lagrange basis on ``nodes`` (longdouble)."""
    nodes = np.asarray(nodes, dtype=LD)
    x = np.atleast_1d(np.asarray(x, dtype=LD))
    n = len(nodes)
    L = np.ones((len(x), n), dtype=LD)
    D = np.zeros((len(x), n), dtype=LD)
    for i in range(n):
        denom = LD(1)
        for l in range(n):
            if l != i:
                denom *= nodes[i] - nodes[l]
        val = np.ones(len(x), dtype=LD)
        for l in range(n):
            if l != i:
                val = val * (x - nodes[l])
        L[:, i] = val / denom
        der = np.zeros(len(x), dtype=LD)
        for m in range(n):
            if m == i:
                continue
            prod = np.ones(len(x), dtype=LD)
            for l in range(n):
                if l != i and l != m:
                    prod = prod * (x - nodes[l])
            der = der + prod
        D[:, i] = der / denom
    return L, D


class Shape1D:
    """All 1-D tables for one (FE_DGQ(p), QGauss(nq)) pair.

    S[q][i] = l_i(xi_q), D[q][i] = l_i'(xi_q) (reference cell [0,1]);
    f0/f1 = l_i(0), l_i(1);  g0/g1 = l_i'(0), l_i'(1);  w = quadrature weights;
    P[c][j][i] = l_i((x_j + c)/2): embedding of the parent basis into child c
    (MGTransferMatrixFree for DG).
    """

    def __init__(self, degree: int, nq: int, dtype=np.float64):
        self.degree = degree
        self.n = degree + 1
        self.nq = nq
        nodes = gauss_lobatto_nodes(self.n)
        xq, wq = gauss_legendre(nq)
        S, D = lagrange_eval(nodes, xq)
        F, G = lagrange_eval(nodes, np.array([0, 1], dtype=LD))
        self.nodes_ld, self.xq_ld, self.wq_ld = nodes, xq, wq
        self.S_ld, self.D_ld, self.F_ld, self.G_ld = S, D, F, G
        P0, _ = lagrange_eval(nodes, nodes / 2)
        P1, _ = lagrange_eval(nodes, (nodes + 1) / 2)
        self.P_ld = np.stack([P0, P1])
        self.dtype = dtype
        cast = lambda a: np.asarray(a, dtype=np.float64).astype(dtype)
        self.nodes, self.xq, self.w = cast(nodes), cast(xq), cast(wq)
        self.S, self.D = cast(S), cast(D)
        self.f = cast(F)          # f[s][i] = l_i(s)
        self.g = cast(G)          # g[s][i] = l_i'(s)
        self.P = cast(self.P_ld)

    # 1-D reference matrices (longdouble), used by analytic tests and diagonals
    def mass_ld(self):
        return np.einsum("q,qi,qj->ij", self.wq_ld, self.S_ld, self.S_ld)

    def stiff_ld(self):
        return np.einsum("q,qi,qj->ij", self.wq_ld, self.D_ld, self.D_ld)


@functools.lru_cache(maxsize=None)
def shape(degree: int, nq: int, dtype_name: str = "float64") -> Shape1D:
    return Shape1D(degree, nq, np.dtype(dtype_name).type)
