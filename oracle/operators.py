"""Quadrature-point restatement of NavierStokesProjectionOperator (oracle; test infrastructure).

Every function cites the lines of
``NavierStokes_TRBDF2_DG/navier_stokes_TRBDF2_DG.cc`` (NS.cc) it follows.  The loops are
face-centric like ``MatrixFree::loop``: every interior face is visited once with an interior
("p", minus) and an exterior ("m", plus) evaluator that share the interior side's outward
normal, exactly the structure of the reference lambdas.  The CUDA kernels use a different
(cell-centric, Kronecker-factored) evaluation order, which is the point of the comparison.

Vectors are flat numpy arrays in the DG layout [cell][comp][k][j][i].
"""
from __future__ import annotations

import math

import numpy as np

from .dg1d import shape
from .feeval import CellEval, FaceEval
from .mesh import CartesianMesh

GAMMA = 2.0 - math.sqrt(2.0)                         # NS.cc:396, 2210
A21 = A22 = 0.5                                      # NS.cc:286-287
A31 = A32 = (1.0 - GAMMA) / (2.0 * (2.0 - GAMMA))    # NS.cc:396-397
A33 = 1.0 / (2.0 - GAMMA)
THETA_V = THETA_P = 1.0                              # NS.cc:290-291


class _OpBase:
    def __init__(self, mesh: CartesianMesh, degree: int, nq: int, ncomp: int, dtype=np.float64):
        self.mesh, self.degree, self.nq, self.ncomp = mesh, degree, nq, ncomp
        self.dim = mesh.dim
        self.dtype = np.dtype(dtype).type
        self.sh = shape(degree, nq, np.dtype(dtype).name)
        self.n = degree + 1
        self.dofs_per_cell = ncomp * self.n ** self.dim
        self.n_dofs = mesh.n_cells * self.dofs_per_cell
        self.cell = CellEval(self.dim, mesh.h, self.sh)
        self.face = {(d, s): FaceEval(self.dim, mesh.h, self.sh, d, s)
                     for d in range(self.dim) for s in (0, 1)}
        self.dt = 0.0
        self.TR_BDF2_stage = 1
        # interior face lists per direction: (minus cells, plus cells); boundary lists per face no.
        self.int_faces, self.bdry_faces = [], []
        for d in range(self.dim):
            nb = mesh.nbr[:, 2 * d + 1]
            minus = np.nonzero(nb >= 0)[0]
            self.int_faces.append((minus, nb[minus]))
        for f in range(2 * self.dim):
            nb = mesh.nbr[:, f]
            cells = np.nonzero(nb < 0)[0]
            self.bdry_faces.append((cells, -1 - nb[cells]))

    def m(self):
        return self.n_dofs

    def set_dt(self, dt):                            # NS.cc:415-418
        self.dt = float(dt)

    def set_TR_BDF2_stage(self, stage):              # NS.cc:424-430
        assert stage in (1, 2)
        self.TR_BDF2_stage = stage

    def _cells(self, v, ncomp=None):
        ncomp = self.ncomp if ncomp is None else ncomp
        return np.asarray(v).reshape((self.mesh.n_cells, ncomp) + (self.n,) * self.dim)

    def vmult(self, src):                            # Base::vmult: dst = 0; apply_add
        dst = np.zeros(self.n_dofs, dtype=self.dtype)
        self.apply_add(dst, np.asarray(src, dtype=self.dtype))
        return dst

    def sigma(self, C, d):
        """C * |(n J^-1)[dim-1]| = C / h_normal on Cartesian cells (NS.cc:1274-1275)."""
        return self.dtype(C / self.mesh.h[d])


class PressureOperator(_OpBase):
    """NS_stage == 2 branch of apply_add (NS.cc:1407-1414): rows B, C, D of SURVEY.md 8a."""

    def __init__(self, mesh, degree, dt, dtype=np.float64, nq=None):
        super().__init__(mesh, degree, degree + 1 if nq is None else nq, 1, dtype)
        self.C_p = 1.0 * (degree + 1) * (degree + 1)         # NS.cc:292
        self.set_dt(dt)

    def coeff(self):                                  # NS.cc:1237
        g, dt = GAMMA, self.dt
        return 1.0e6 * g * dt * g * dt if self.TR_BDF2_stage == 1 else 1.0e6 * (1.0 - g) * dt * (1.0 - g) * dt

    # --- NS.cc:1230-1252 -------------------------------------------------------------------
    def _cell_term(self, u):
        val = self.cell.values(u)
        grad = self.cell.gradients(u)
        return self.cell.integrate(val=self.dtype(1.0 / self.coeff()) * val, grad=grad)

    # --- NS.cc:1259-1292 -------------------------------------------------------------------
    def _face_term(self, d, u_p, u_m):
        fp, fm = self.face[(d, 1)], self.face[(d, 0)]         # interior side upper face, exterior lower
        coef_jump = self.dtype(self.C_p * 0.5 * (abs(1.0 / self.mesh.h[d]) + abs(1.0 / self.mesh.h[d])))
        n_plus = np.zeros(self.dim, dtype=self.dtype)
        n_plus[d] = 1
        gp, gm = fp.gradients(u_p), fm.gradients(u_m)
        avg_grad = self.dtype(0.5) * (gp + gm)
        jump = fp.values(u_p) - fm.values(u_m)
        avg_grad_n = np.tensordot(avg_grad, n_plus, axes=([2], [0]))
        sub_p = -avg_grad_n + coef_jump * jump
        sub_m = avg_grad_n - coef_jump * jump
        gsub = np.stack([self.dtype(-THETA_P * 0.5) * jump * n_plus[e] for e in range(self.dim)], axis=2)
        return fp.integrate(val=sub_p, grad=gsub), fm.integrate(val=sub_m, grad=gsub)

    # --- NS.cc:1299-1326 (operator: sigma; diagonal: 2 sigma, NS.cc:1997) ------------------
    def _boundary_term(self, f, u, penalty_factor=1.0):
        d, s = f // 2, f % 2
        fe = self.face[(d, s)]
        coef_jump = self.dtype(self.C_p * abs(1.0 / self.mesh.h[d]))
        n_plus = np.zeros(self.dim, dtype=self.dtype)
        n_plus[d] = 2 * s - 1
        grad = fe.gradients(u)
        pres = fe.values(u)
        val = -np.tensordot(grad, n_plus, axes=([2], [0])) + self.dtype(penalty_factor) * coef_jump * pres
        return fe.integrate(val=val, normal_derivative=self.dtype(-THETA_P) * pres, normal_sign=2 * s - 1)

    def apply_add(self, dst, src):
        u = self._cells(src)
        out = self._cells(dst)
        out += self._cell_term(u)
        for d in range(self.dim):
            minus, plus = self.int_faces[d]
            if len(minus) == 0:
                continue
            rp, rm = self._face_term(d, u[minus], u[plus])
            out[minus] += rp
            out[plus] += rm
        for f in range(2 * self.dim):
            cells, ids = self.bdry_faces[f]
            sel = cells[ids == 1]                     # only boundary_id == 1 (NS.cc:1308)
            if len(sel):
                out[sel] += self._boundary_term(f, u[sel])

    # --- compute_diagonal, NS_stage == 2 (NS.cc:1857-2008, 2018-2060) ----------------------
    def compute_diagonal(self):
        """Literal unit-vector probing, including the reference's quirks: the face probe sets
        dof i on BOTH sides at once (NS.cc:1921-1922) and the boundary probe uses 2*sigma
        (NS.cc:1997)."""
        N = self.mesh.n_cells
        npc = self.dofs_per_cell
        diag = np.zeros((N, npc), dtype=self.dtype)
        shape_c = (1,) + (self.n,) * self.dim
        for i in range(npc):
            e = np.zeros(npc, dtype=self.dtype)
            e[i] = 1
            unit = np.broadcast_to(e.reshape((1,) + shape_c), (N,) + shape_c)
            diag[:, i] += self._cell_term(unit).reshape(N, npc)[:, i]
            for d in range(self.dim):
                minus, plus = self.int_faces[d]
                if len(minus) == 0:
                    continue
                rp, rm = self._face_term(d, unit[minus], unit[plus])
                np.add.at(diag[:, i], minus, rp.reshape(len(minus), npc)[:, i])
                np.add.at(diag[:, i], plus, rm.reshape(len(minus), npc)[:, i])
            for f in range(2 * self.dim):
                cells, ids = self.bdry_faces[f]
                sel = cells[ids == 1]
                if len(sel):
                    r = self._boundary_term(f, unit[sel], penalty_factor=2.0)
                    diag[sel, i] += r.reshape(len(sel), npc)[:, i]
        self.inverse_diagonal = (self.dtype(1.0) / diag).reshape(-1)      # NS.cc:2055-2059
        return self.inverse_diagonal


class MassOperator(_OpBase):
    """NS_stage == 3: assemble_cell_term_projection_grad_p (NS.cc:1373-1390), velocity space,
    velocity quadrature."""

    def __init__(self, mesh, degree_v, dtype=np.float64, nq=None, ncomp=None):
        super().__init__(mesh, degree_v, degree_v + 1 if nq is None else nq,
                         mesh.dim if ncomp is None else ncomp, dtype)

    def apply_add(self, dst, src):
        u = self._cells(src)
        self._cells(dst)[...] += self.cell.integrate(val=self.cell.values(u))


class VelocityOperator(_OpBase):
    """NS_stage == 1 branch of apply_add (NS.cc:1399-1406): rows E, F, F' of SURVEY.md 8a."""

    def __init__(self, mesh, degree_v, Re, dt, dtype=np.float64, nq=None):
        super().__init__(mesh, degree_v, degree_v + 1 if nq is None else nq, mesh.dim, dtype)
        self.Re = float(Re)
        self.C_u = 1.0 * (degree_v + 1) * (degree_v + 1)      # NS.cc:293
        self.set_dt(dt)
        self.u_extr = np.zeros(self.n_dofs, dtype=self.dtype)

    def set_u_extr(self, src):                        # NS.cc:448-453
        self.u_extr = np.array(src, dtype=self.dtype, copy=True)

    def _stage(self):
        """(1/(gamma dt), a22) for stage 1 (NS.cc:957-958), (1/((1-gamma) dt), a33) for stage 2 (:982-983)."""
        if self.TR_BDF2_stage == 1:
            return 1.0 / (GAMMA * self.dt), A22
        return 1.0 / ((1.0 - GAMMA) * self.dt), A33

    # --- NS.cc:933-988 ---------------------------------------------------------------------
    def _cell_term(self, u, ue):
        alpha, a = self._stage()
        T = self.dtype
        u_q = self.cell.values(u)                         # [N, dim, q..]
        grad_u = self.cell.gradients(u)                   # [N, dim(comp), dim(deriv), q..]
        ue_q = self.cell.values(ue)
        tensor = u_q[:, :, None] * ue_q[:, None, :]       # outer_product(u, u_extr)[i][j] = u_i ue_j
        return self.cell.integrate(val=T(alpha) * u_q, grad=T(-a) * tensor + T(a / self.Re) * grad_u)

    # --- NS.cc:995-1085 --------------------------------------------------------------------
    def _face_term(self, d, u_p, u_m, ue_p, ue_m):
        _, a = self._stage()
        T = self.dtype
        fp, fm = self.face[(d, 1)], self.face[(d, 0)]
        coef_jump = T(self.C_u * 0.5 * (abs(1.0 / self.mesh.h[d]) + abs(1.0 / self.mesh.h[d])))
        vp, vm = fp.values(u_p), fm.values(u_m)
        ep, em = fp.values(ue_p), fm.values(ue_m)
        gp, gm = fp.gradients(u_p), fm.gradients(u_m)
        avg_grad_n = T(0.5) * (gp[:, :, d] + gm[:, :, d])           # (avg_grad * n_plus), n_plus = +e_d
        jump = vp - vm
        avg_tensor_n = T(0.5) * (vp * ep[:, d:d + 1] + vm * em[:, d:d + 1])   # (u (x) ue) n = u (ue . n)
        lam = np.maximum(np.abs(ep[:, d:d + 1]), np.abs(em[:, d:d + 1]))
        flux = T(a / self.Re) * (-avg_grad_n + coef_jump * jump) + T(a) * avg_tensor_n + T(0.5 * a) * lam * jump
        nd = T(-THETA_V * a / self.Re * 0.5) * jump
        return (fp.integrate(val=flux, normal_derivative=nd, normal_sign=1),
                fm.integrate(val=-flux, normal_derivative=nd, normal_sign=1))

    # --- NS.cc:1092-1223 -------------------------------------------------------------------
    def _boundary_term(self, f, bid, u, ue):
        _, a = self._stage()
        T = self.dtype
        d, s = f // 2, f % 2
        sgn = 2 * s - 1
        fe = self.face[(d, s)]
        coef_jump = T(self.C_u * abs(1.0 / self.mesh.h[d]))
        v = fe.values(u)
        e = fe.values(ue)
        g = fe.gradients(u)
        lam = np.abs(T(sgn) * e[:, d:d + 1])
        grad_n = T(sgn) * g[:, :, d]
        if bid != 1:
            coef_trasp = 0.0                                            # NS.cc:1114
            tens_n = v * (T(sgn) * e[:, d:d + 1])
            val = T(a / self.Re) * (-grad_n + T(2.0) * coef_jump * v) + T(a * coef_trasp) * tens_n + T(a) * lam * v
            nd = T(-THETA_V * a / self.Re) * v
        else:
            v_m = v.copy()
            g_m = g.copy()
            v_m[:, 1] = -v_m[:, 1]                                      # NS.cc:1146
            g_m[:, 0, 0] = -g_m[:, 0, 0]                                # NS.cc:1147
            g_m[:, 0, 1] = -g_m[:, 0, 1]                                # NS.cc:1148
            gavg_n = T(sgn) * (T(0.5) * (g + g_m))[:, :, d]
            val = (T(a / self.Re) * (-gavg_n + coef_jump * (v - v_m))
                   + T(a) * (T(0.5) * (v + v_m)) * (T(sgn) * e[:, d:d + 1])
                   + T(a * 0.5) * lam * (v - v_m))
            nd = T(-THETA_V * a / self.Re) * (v - v_m)
        return fe.integrate(val=val, normal_derivative=nd, normal_sign=sgn)

    def apply_add(self, dst, src):
        u = self._cells(src)
        ue = self._cells(self.u_extr)
        out = self._cells(dst)
        out += self._cell_term(u, ue)
        for d in range(self.dim):
            minus, plus = self.int_faces[d]
            if len(minus) == 0:
                continue
            rp, rm = self._face_term(d, u[minus], u[plus], ue[minus], ue[plus])
            out[minus] += rp
            out[plus] += rm
        for f in range(2 * self.dim):
            cells, ids = self.bdry_faces[f]
            for bid in np.unique(ids):
                sel = cells[ids == bid]
                out[sel] += self._boundary_term(f, int(bid), u[sel], ue[sel])

    # --- compute_diagonal, NS_stage == 1 (NS.cc:1431-1850) ---------------------------------
    def compute_diagonal(self):
        """Unit-vector probing as the reference does it: for every scalar dof i the probe sets ALL
        dim components of dof i to one (``tmp``, NS.cc:1442-1444, 1455) on BOTH sides of a face
        (NS.cc:1527-1530) and keeps the dim entries of dof i."""
        N = self.mesh.n_cells
        C = self.ncomp
        nsc = self.n ** self.dim
        diag = np.zeros((N, C, nsc), dtype=self.dtype)
        ue = self._cells(self.u_extr)
        shape_c = (C,) + (self.n,) * self.dim
        for i in range(nsc):
            e = np.zeros((C, nsc), dtype=self.dtype)
            e[:, i] = 1
            unit = np.broadcast_to(e.reshape((1,) + shape_c), (N,) + shape_c)
            diag[:, :, i] += self._cell_term(unit, ue).reshape(N, C, nsc)[:, :, i]
            for d in range(self.dim):
                minus, plus = self.int_faces[d]
                if len(minus) == 0:
                    continue
                rp, rm = self._face_term(d, unit[minus], unit[plus], ue[minus], ue[plus])
                np.add.at(diag[:, :, i], minus, rp.reshape(len(minus), C, nsc)[:, :, i])
                np.add.at(diag[:, :, i], plus, rm.reshape(len(minus), C, nsc)[:, :, i])
            for f in range(2 * self.dim):
                cells, ids = self.bdry_faces[f]
                for bid in np.unique(ids):
                    sel = cells[ids == bid]
                    r = self._boundary_term(f, int(bid), unit[sel], ue[sel])
                    diag[sel, :, i] += r.reshape(len(sel), C, nsc)[:, :, i]
        self.inverse_diagonal = (self.dtype(1.0) / diag).reshape(-1)      # NS.cc:2055-2059
        return self.inverse_diagonal
