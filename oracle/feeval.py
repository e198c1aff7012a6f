"""numpy stand-ins for deal.II's FEEvaluation / FEFaceEvaluation on Cartesian cells (oracle).

Semantics restated from SURVEY.md 9.3 (deal.II, not in /root/reference):
``get_value = sum_i phi_i(xi_q) u_i``; ``get_gradient = J^-T grad_xi u``;
``submit_value(v)`` stores ``v*JxW``; ``submit_gradient(g)`` stores ``J^-1 g JxW``;
``integrate`` applies the transposed shape matrices.  On a Cartesian cell
``J^-1 = diag(1/h_d)``, ``JxW = prod(h) * prod(w)``; face ``JxW = prod_{e != d} h_e * prod(w)``.

Arrays are [n_items, n_comp, (dim spatial axes, slowest..fastest = z,y,x)], matching the
DoF layout [cell][comp][k][j][i].  Everything is dense einsum-style: this is the slow,
obviously-correct path.  Test infrastructure only.
"""
from __future__ import annotations

import numpy as np


def _ax(dim, d, base=2):
    """array axis of spatial direction d (0 = x = fastest = last axis)."""
    return base + (dim - 1 - d)


def apply1d(T, M, axis):
    """contract M[q][i] with axis ``axis`` of T (size i) -> axis has size q."""
    out = np.tensordot(M, T, axes=([1], [axis]))
    return np.moveaxis(out, 0, axis)


def contract_vec(T, v, axis):
    return np.tensordot(T, v, axes=([axis], [0]))


def expand_vec(T, v, axis):
    """insert a new axis at ``axis`` filled with the outer product by v."""
    shape = [1] * (T.ndim + 1)
    shape[axis] = len(v)
    return np.expand_dims(T, axis) * v.reshape(shape)


class CellEval:
    """FEEvaluation<dim, p, nq, ncomp> on all given cells at once."""

    def __init__(self, dim, h, sh):
        self.dim, self.sh = dim, sh
        dt = sh.dtype
        self.h = np.asarray(h, dtype=np.float64)
        self.inv_h = (1.0 / self.h).astype(dt)
        w = sh.w
        W = np.ones((), dtype=dt)
        for _ in range(dim):
            W = np.multiply.outer(W, w)
        self.JxW = (W * dt(np.prod(self.h))).astype(dt)          # [q..]

    def values(self, u):
        v = u
        for d in range(self.dim):
            v = apply1d(v, self.sh.S, _ax(self.dim, d))
        return v

    def gradients(self, u):
        """[N, C, dim, q...]; component e = d/dx_e."""
        out = []
        for e in range(self.dim):
            g = u
            for d in range(self.dim):
                g = apply1d(g, self.sh.D if d == e else self.sh.S, _ax(self.dim, d))
            out.append(g * self.inv_h[e])
        return np.stack(out, axis=2)

    def integrate(self, val=None, grad=None):
        """val [N,C,q..] tested with phi; grad [N,C,dim,q..] tested with grad phi."""
        res = 0
        if val is not None:
            t = val * self.JxW
            for d in range(self.dim):
                t = apply1d(t, self.sh.S.T, _ax(self.dim, d))
            res = res + t
        if grad is not None:
            for e in range(self.dim):
                t = grad[:, :, e] * (self.JxW * self.inv_h[e])
                for d in range(self.dim):
                    t = apply1d(t, (self.sh.D if d == e else self.sh.S).T, _ax(self.dim, d))
                res = res + t
        return res


class FaceEval:
    """FEFaceEvaluation on face (direction d, side s) of all given cells at once.

    Face-point arrays are [N, C, (dim-1 remaining spatial axes, slowest..fastest)].
    """

    def __init__(self, dim, h, sh, d, side):
        self.dim, self.sh, self.d, self.side = dim, sh, d, side
        dt = sh.dtype
        self.h = np.asarray(h, dtype=np.float64)
        self.inv_h = (1.0 / self.h).astype(dt)
        self.rem = [e for e in range(dim) if e != d]              # tangential directions
        w = sh.w
        W = np.ones((), dtype=dt)
        for _ in range(dim - 1):
            W = np.multiply.outer(W, w)
        self.JxW = (W * dt(np.prod(self.h[self.rem]))).astype(dt)
        self.fv = sh.f[side]
        self.gv = sh.g[side]

    def _fax(self, e):
        """axis (in a face array) of tangential direction e."""
        order = sorted(self.rem, reverse=True)                    # slowest..fastest
        return 2 + order.index(e)

    def values(self, u):
        v = contract_vec(np.moveaxis(u, _ax(self.dim, self.d), -1), self.fv, u.ndim - 1)
        for e in self.rem:
            v = apply1d(v, self.sh.S, self._fax(e))
        return v

    def gradients(self, u):
        um = np.moveaxis(u, _ax(self.dim, self.d), -1)
        out = [None] * self.dim
        # normal component
        g = contract_vec(um, self.gv, u.ndim - 1)
        for e in self.rem:
            g = apply1d(g, self.sh.S, self._fax(e))
        out[self.d] = g * self.inv_h[self.d]
        # tangential components
        t0 = contract_vec(um, self.fv, u.ndim - 1)
        for e in self.rem:
            g = t0
            for e2 in self.rem:
                g = apply1d(g, self.sh.D if e2 == e else self.sh.S, self._fax(e2))
            out[e] = g * self.inv_h[e]
        return np.stack(out, axis=2)

    def integrate(self, val=None, grad=None, normal_derivative=None, normal_sign=None):
        """val tested with phi, grad [N,C,dim,...] with grad phi, normal_derivative with
        d(phi)/dn where n = normal_sign * e_d (the *interior* side's outward normal)."""
        res = 0
        ax_d = _ax(self.dim, self.d)
        if val is not None:
            t = val * self.JxW
            for e in self.rem:
                t = apply1d(t, self.sh.S.T, self._fax(e))
            res = res + expand_vec(t, self.fv, ax_d)
        if normal_derivative is not None:
            t = normal_derivative * (self.JxW * (self.inv_h[self.d] * normal_sign))
            for e in self.rem:
                t = apply1d(t, self.sh.S.T, self._fax(e))
            res = res + expand_vec(t, self.gv, ax_d)
        if grad is not None:
            t = grad[:, :, self.d] * (self.JxW * self.inv_h[self.d])
            for e in self.rem:
                t = apply1d(t, self.sh.S.T, self._fax(e))
            res = res + expand_vec(t, self.gv, ax_d)
            for e in self.rem:
                t = grad[:, :, e] * (self.JxW * self.inv_h[e])
                for e2 in self.rem:
                    t = apply1d(t, (self.sh.D if e2 == e else self.sh.S).T, self._fax(e2))
                res = res + expand_vec(t, self.fv, ax_d)
        return res
