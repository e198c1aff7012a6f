"""Right-hand-side linear forms of NavierStokesProjectionOperator (oracle; test infrastructure only).

Restates, face-centric and at quadrature points like the reference lambdas:
  vmult_rhs_velocity      NS.cc:788-798   = assemble_rhs_{cell,face,boundary}_term_velocity NS.cc:462-781
  vmult_rhs_pressure      NS.cc:915-925   = assemble_rhs_{cell,face,boundary}_term_pressure NS.cc:806-908
  vmult_grad_p_projection NS.cc:1363-1366 = assemble_rhs_cell_term_projection_grad_p        NS.cc:1336-1356
Mixed evaluators as in the reference: the pressure FE is evaluated on the velocity quadrature QGauss(p+2) in the
velocity forms (NS.cc:474) and the velocity FE on the pressure quadrature QGauss(p+1) in the pressure form (NS.cc:814).
Inflow data on boundary id 0: EquationData::Velocity (equation_data.h:47-57).
"""
from __future__ import annotations

import numpy as np

from .dg1d import shape
from .feeval import CellEval, FaceEval
from .mesh import CartesianMesh
from .operators import A21, A22, A31, A32, A33, GAMMA, THETA_V


def inflow_velocity(points, dim, Um=1.5, H=4.1):
    """EquationData::Velocity::value (equation_data.h:47-57): parabolic profile in component 0."""
    out = np.zeros(points.shape[:-1] + (dim,), dtype=np.float64)
    y = points[..., 1]
    out[..., 0] = 4.0 * Um * y * (H - y) / (H * H)
    return out


class RhsForms:
    def __init__(self, mesh: CartesianMesh, degree_p: int, Re: float, dt: float):
        self.mesh, self.dim = mesh, mesh.dim
        self.dp, self.dv = degree_p, degree_p + 1
        self.np_, self.nv = degree_p + 1, degree_p + 2
        self.Re, self.dt = float(Re), float(dt)
        self.C_u = 1.0 * (self.dv + 1) ** 2
        d, h = self.dim, mesh.h
        # (FE, quadrature) pairs: quad 0 = QGauss(nv), quad 1 = QGauss(np)
        self.vel_q0, self.pre_q0 = shape(self.dv, self.nv), shape(self.dp, self.nv)
        self.vel_q1, self.pre_q1 = shape(self.dv, self.np_), shape(self.dp, self.np_)
        self.cv0, self.cp0 = CellEval(d, h, self.vel_q0), CellEval(d, h, self.pre_q0)
        self.cv1, self.cp1 = CellEval(d, h, self.vel_q1), CellEval(d, h, self.pre_q1)
        F = lambda sh: {(e, s): FaceEval(d, h, sh, e, s) for e in range(d) for s in (0, 1)}
        self.fv0, self.fp0, self.fv1, self.fp1 = F(self.vel_q0), F(self.pre_q0), F(self.vel_q1), F(self.pre_q1)
        self.int_faces, self.bdry_faces = [], []
        for e in range(d):
            nb = mesh.nbr[:, 2 * e + 1]
            minus = np.nonzero(nb >= 0)[0]
            self.int_faces.append((minus, nb[minus]))
        for f in range(2 * d):
            nb = mesh.nbr[:, f]
            cells = np.nonzero(nb < 0)[0]
            self.bdry_faces.append((cells, -1 - nb[cells]))

    def _v(self, u):
        return np.asarray(u, dtype=np.float64).reshape((self.mesh.n_cells, self.dim) + (self.nv,) * self.dim)

    def _p(self, p):
        return np.asarray(p, dtype=np.float64).reshape((self.mesh.n_cells, 1) + (self.np_,) * self.dim)

    def _face_points(self, cells, e, s, xq):
        """physical coordinates of the face quadrature points: [N, (dim-1 face axes slowest..fastest), dim]"""
        m, d = self.mesh, self.dim
        rem = sorted([t for t in range(d) if t != e], reverse=True)     # slowest..fastest
        grids = np.meshgrid(*([np.asarray(xq, dtype=np.float64)] * (d - 1)), indexing="ij") if d > 1 else []
        pts = np.zeros((len(cells),) + (len(xq),) * (d - 1) + (d,))
        org = m.cell_origin()[cells]
        for k, t in enumerate(rem):
            pts[..., t] = org[:, t].reshape((-1,) + (1,) * (d - 1)) + grids[k] * m.h[t]
        pts[..., e] = (org[:, e] + s * m.h[e]).reshape((-1,) + (1,) * (d - 1))
        return pts

    # ------------------------------------------------------------------ velocity rhs
    def rhs_velocity(self, stage, srcs):
        d, Re, dt = self.dim, self.Re, self.dt
        N = self.mesh.n_cells
        out = np.zeros((N, d) + (self.nv,) * d)
        eye = np.eye(d).reshape((1, d, d) + (1,) * d)
        if stage == 1:                                           # NS.cc:466-513
            u_n, u_e, p_n = self._v(srcs[0]), self._v(srcs[1]), self._p(srcs[2])
            un, gn, ue, pn = self.cv0.values(u_n), self.cv0.gradients(u_n), self.cv0.values(u_e), self.cp0.values(p_n)
            tens = un[:, :, None] * ue[:, None, :]
            out += self.cv0.integrate(val=un / (GAMMA * dt), grad=-A21 / Re * gn + A21 * tens + pn[:, :, None] * eye)
        else:                                                    # NS.cc:514-551
            u_n, u_g, p_n = self._v(srcs[0]), self._v(srcs[1]), self._p(srcs[2])
            u_e = self._v(srcs[3])
            un, gn, ug, gg, pn = (self.cv0.values(u_n), self.cv0.gradients(u_n), self.cv0.values(u_g), self.cv0.gradients(u_g),
                                  self.cp0.values(p_n))
            out += self.cv0.integrate(val=ug / ((1.0 - GAMMA) * dt),
                                      grad=A32 * ug[:, :, None] * ug[:, None, :] + A31 * un[:, :, None] * un[:, None, :]
                                      - A32 / Re * gg - A31 / Re * gn + pn[:, :, None] * eye)
        # interior faces (NS.cc:563-610 / 611-659): n_plus = +e_d
        for e in range(d):
            minus, plus = self.int_faces[e]
            if len(minus) == 0:
                continue
            fp, fm, pp, pm = self.fv0[(e, 1)], self.fv0[(e, 0)], self.fp0[(e, 1)], self.fp0[(e, 0)]
            avg_p = 0.5 * (pp.values(p_n[minus]) + pm.values(p_n[plus]))             # [F,1,...]
            if stage == 1:
                gp, gm = fp.gradients(u_n[minus]), fm.gradients(u_n[plus])
                vp, vm = fp.values(u_n[minus]), fm.values(u_n[plus])
                ep, em = fp.values(u_e[minus]), fm.values(u_e[plus])
                flux = A21 / Re * 0.5 * (gp[:, :, e] + gm[:, :, e]) - A21 * 0.5 * (vp * ep[:, e:e + 1] + vm * em[:, e:e + 1])
            else:
                gp, gm = fp.gradients(u_n[minus]), fm.gradients(u_n[plus])
                hp, hm = fp.gradients(u_g[minus]), fm.gradients(u_g[plus])
                vp, vm = fp.values(u_n[minus]), fm.values(u_n[plus])
                wp, wm = fp.values(u_g[minus]), fm.values(u_g[plus])
                flux = (A31 / Re * 0.5 * (gp[:, :, e] + gm[:, :, e]) + A32 / Re * 0.5 * (hp[:, :, e] + hm[:, :, e])
                        - A31 * 0.5 * (vp * vp[:, e:e + 1] + vm * vm[:, e:e + 1]) - A32 * 0.5 * (wp * wp[:, e:e + 1] + wm * wm[:, e:e + 1]))
            nvec = np.zeros((1, d) + (1,) * (d - 1))
            nvec[0, e] = 1.0
            val = flux - avg_p * nvec
            out[minus] += fp.integrate(val=val)
            out[plus] += fm.integrate(val=-val)
        # boundary faces (NS.cc:671-724 / 725-780)
        a_b = A22 if stage == 1 else A33
        for f in range(2 * d):
            cells_all, ids_all = self.bdry_faces[f]
            e, s = f // 2, f % 2
            sgn = 2 * s - 1
            for bid in np.unique(ids_all):
                cells = cells_all[ids_all == bid]
                fe, pe = self.fv0[(e, s)], self.fp0[(e, s)]
                coef_jump = 0.0 if bid == 1 else self.C_u / self.mesh.h[e]
                aux = 0.0 if bid == 1 else 1.0
                p_old = pe.values(p_n[cells])
                ext = fe.values(u_e[cells])
                if bid == 0:
                    pts = self._face_points(cells, e, s, self.vel_q0.xq)
                    uD = np.moveaxis(inflow_velocity(pts, d), -1, 1)                # [N, dim, ...]
                else:
                    uD = np.zeros_like(ext)
                lam = 0.0 if bid == 1 else np.abs(sgn * ext[:, e:e + 1])
                if stage == 1:
                    g, v = fe.gradients(u_n[cells]), fe.values(u_n[cells])
                    base = (A21 / Re * g[:, :, e] - A21 * v * ext[:, e:e + 1]) * sgn
                else:
                    g, v = fe.gradients(u_n[cells]), fe.values(u_n[cells])
                    hg, w = fe.gradients(u_g[cells]), fe.values(u_g[cells])
                    base = (A31 / Re * g[:, :, e] + A32 / Re * hg[:, :, e] - A31 * v * v[:, e:e + 1] - A32 * w * w[:, e:e + 1]) * sgn
                nvec = np.zeros((1, d) + (1,) * (d - 1))
                nvec[0, e] = sgn
                val = (base - p_old * nvec + a_b / Re * 2.0 * coef_jump * uD - aux * a_b * uD * (sgn * ext[:, e:e + 1]) + a_b * lam * uD)
                out[cells] += fe.integrate(val=val, normal_derivative=-aux * THETA_V * a_b / Re * uD, normal_sign=sgn)
        return out.reshape(-1)

    # ------------------------------------------------------------------ pressure rhs (NS.cc:806-908)
    def rhs_pressure(self, stage, srcs):
        d, dt = self.dim, self.dt
        u_s, p_o = self._v(srcs[0]), self._p(srcs[1])
        g = GAMMA
        coeff = 1.0e6 * g * dt * g * dt if stage == 1 else 1.0e6 * (1.0 - g) * dt * (1.0 - g) * dt
        coeff_2 = g * dt if stage == 1 else (1.0 - g) * dt
        out = np.zeros((self.mesh.n_cells, 1) + (self.np_,) * d)
        us = self.cv1.values(u_s)                                                   # velocity FE on the pressure quadrature
        out += self.cp1.integrate(val=self.cp1.values(p_o) / coeff, grad=(us / coeff_2)[:, None])
        c = 1.0 / coeff_2
        for e in range(d):
            minus, plus = self.int_faces[e]
            if len(minus) == 0:
                continue
            avg = 0.5 * (self.fv1[(e, 1)].values(u_s[minus]) + self.fv1[(e, 0)].values(u_s[plus]))
            val = -c * avg[:, e:e + 1]
            out[minus] += self.fp1[(e, 1)].integrate(val=val)
            out[plus] += self.fp1[(e, 0)].integrate(val=-val)
        for f in range(2 * d):
            cells, _ = self.bdry_faces[f]
            if len(cells) == 0:
                continue
            e, s = f // 2, f % 2
            v = self.fv1[(e, s)].values(u_s[cells])
            out[cells] += self.fp1[(e, s)].integrate(val=-c * (2 * s - 1) * v[:, e:e + 1])
        return out.reshape(-1)

    # ------------------------------------------------------------------ grad p projection rhs (NS.cc:1336-1356)
    def rhs_grad_p(self, p):
        gp = self.cp0.gradients(self._p(p))[:, 0]                                   # [N, dim, q..]
        return self.cv0.integrate(val=gp).reshape(-1)
