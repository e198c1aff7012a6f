"""Pressure operator on a MappingQ1 (non-Cartesian) mesh -- dense restatement (oracle; test infrastructure only).

Follows assemble_cell_term_pressure (navier_stokes_TRBDF2_DG.cc:1230-1252), assemble_face_term_pressure (:1259-1292) and
assemble_boundary_term_pressure (:1299-1326) with deal.II's general-mapping semantics:
``get_gradient = J^-T grad_xi u``, ``submit_gradient(g)`` tests with ``J^-T grad_xi phi``, ``JxW = det J w`` in cells and
``|det J| |J^-T e_d| w`` on faces, both face evaluators share the interior side's outward normal, and the penalty uses
``|(n * inverse_jacobian(0))[dim-1]|`` = ``|n . grad_x xi_d|`` of both sides at the FIRST face quadrature point.

Everything is dense: every basis function of a cell is evaluated (value and PHYSICAL gradient) at every cell / face quadrature
point and the bilinear forms are summed face by face (face-centric, like MatrixFree::loop).  The CUDA kernel
(csrc/pressure_general.cu) is cell-centric, works in reference coordinates with folded metric tables and a collocation basis.

The topology is the refine_global hierarchy of ``CartesianMesh``; the geometry is a vertex map ``x = fn(X)`` applied to the
vertex lattice of every level (what GridTools::transform / a manifold do to the reference's triangulation).
"""
from __future__ import annotations

import numpy as np

from .dg1d import gauss_legendre, gauss_lobatto_nodes, lagrange_eval
from .mesh import CartesianMesh
from .operators import GAMMA, THETA_P


class GeneralMesh:
    def __init__(self, cart: CartesianMesh, fn):
        self.cart, self.fn, self.dim = cart, fn, cart.dim
        self.n_cells, self.nbr, self.level = cart.n_cells, cart.nbr, cart.level

    def vertex_lattice(self):
        """[nz+1][ny+1][nx+1][dim] (2-D: [ny+1][nx+1][dim]) physical vertex positions; flattened = dgx_mesh_set_vertices order"""
        m = self.cart
        axes = [m.lo[d] + m.h[d] * np.arange(m.ncells_dir[d] + 1) for d in range(m.dim)]
        grids = np.meshgrid(*axes[::-1], indexing="ij")                 # slowest (z) first
        X = np.stack(grids[::-1], axis=-1)                              # [..., dim] with components (x, y, z)
        return np.asarray(self.fn(X), dtype=np.float64)

    def cell_vertices(self):
        """[n_cells, 2^dim, dim] in deal.II vertex order (x fastest)"""
        lat, c, dim = self.vertex_lattice(), self.cart.coords, self.dim
        out = np.zeros((self.n_cells, 1 << dim, dim))
        for v in range(1 << dim):
            idx = tuple(c[:, d] + ((v >> d) & 1) for d in range(dim - 1, -1, -1))
            out[:, v] = lat[idx]
        return out

    def coarser(self):
        return GeneralMesh(self.cart.coarser(), self.fn)


def _tensor_tables(tabs):
    """tabs[d] = (L[q_d, i_d], D[q_d, i_d]) per direction -> B[q, i], dB[q, i, e] with x fastest in q and i"""
    dim = len(tabs)

    def kron(mats):                       # mats[d] for d = 0 (x) .. dim-1; result index = i_x + n i_y + ...
        out = mats[0]
        for d in range(1, dim):
            out = np.kron(mats[d], out)
        return out
    B = kron([tabs[d][0] for d in range(dim)])
    dB = np.stack([kron([tabs[d][1] if d == e else tabs[d][0] for d in range(dim)]) for e in range(dim)], axis=-1)
    return B, dB


def _points(tabs_x):
    """point coordinates [q, dim] for per-direction 1-D point arrays, x fastest"""
    dim = len(tabs_x)
    grids = np.meshgrid(*tabs_x[::-1], indexing="ij")
    return np.stack([g.reshape(-1) for g in grids[::-1]], axis=-1)


def q1_jacobian(X, pts):
    """inverse Jacobian invJ[c, q, e, k] = d xi_e / d x_k and det J [c, q] of the multilinear cells with vertices X[c, 2^dim, dim]
    at the reference points pts[q, dim]"""
    dim = X.shape[-1]
    J = np.zeros((X.shape[0], len(pts), dim, dim))                       # J[c, q, k, e] = d x_k / d xi_e
    for v in range(1 << dim):
        for e in range(dim):
            dN = np.ones(len(pts))
            for d in range(dim):
                b = (v >> d) & 1
                dN = dN * ((1.0 if b else -1.0) if d == e else (pts[:, d] if b else 1.0 - pts[:, d]))
            J[:, :, :, e] += X[:, v, None, :] * dN[None, :, None]
    return np.linalg.inv(J), np.linalg.det(J)


class PressureOperatorGeneral:
    def __init__(self, gmesh: GeneralMesh, degree: int, dt: float, dtype=np.float64):
        self.gm, self.degree, self.dim = gmesh, degree, gmesh.dim
        self.n = n = degree + 1
        self.dtype = np.dtype(dtype).type
        self.ND = n ** self.dim
        self.n_dofs = gmesh.n_cells * self.ND
        self.C_p = float(n * n)                                          # NS.cc:292
        self.dt, self.TR_BDF2_stage = float(dt), 1
        nodes = gauss_lobatto_nodes(n)
        xq, wq = gauss_legendre(n)                                       # QGauss(p + 1), NS.cc:1235, 2325
        self.xq, self.wq = np.asarray(xq, np.float64), np.asarray(wq, np.float64)
        ev = lambda pts: tuple(np.asarray(a, np.float64) for a in lagrange_eval(nodes, np.asarray(pts, dtype=np.longdouble)))
        self._tq, self._t01 = ev(xq), [ev([0.0]), ev([1.0])]
        self.X = gmesh.cell_vertices()
        dim = self.dim
        # cell quadrature
        self.B, self.dBref = _tensor_tables([self._tq] * dim)
        pts = _points([self.xq] * dim)
        w = np.ones(1)
        for d in range(dim):
            w = np.kron(self.wq, w)
        self.invJ, det = self._jac(pts)
        assert np.all(det > 0)
        self.JxW = det * w
        # faces
        self.face = {}
        for d in range(dim):
            for s in (0, 1):
                tabs = [self._t01[s] if e == d else self._tq for e in range(dim)]
                Bf, dBf = _tensor_tables(tabs)
                fp = _points([np.array([float(s)]) if e == d else self.xq for e in range(dim)])
                wf = np.ones(1)
                for e in range(dim):
                    if e != d:
                        wf = np.kron(self.wq, wf)
                invJ, det = self._jac(fp)
                gxi = invJ[:, :, d, :]                                   # grad_x xi_d
                norm = np.linalg.norm(gxi, axis=-1)
                normal = (2 * s - 1) * gxi / norm[..., None]
                self.face[(d, s)] = dict(B=Bf, dBref=dBf, invJ=invJ, normal=normal, JxW=det * norm * wf, norm=norm)

    # geometry ------------------------------------------------------------------------------------
    def _jac(self, pts):
        return q1_jacobian(self.X, pts)

    # operator ------------------------------------------------------------------------------------
    def set_dt(self, dt):
        self.dt = float(dt)

    def set_TR_BDF2_stage(self, stage):
        self.TR_BDF2_stage = stage

    def coeff(self):                                                     # NS.cc:1237
        g, dt = GAMMA, self.dt
        return 1.0e6 * g * dt * g * dt if self.TR_BDF2_stage == 1 else 1.0e6 * (1.0 - g) * dt * (1.0 - g) * dt

    def m(self):
        return self.n_dofs

    def _phys(self, dBref, invJ):
        """dB[c, q, i, k] = sum_e dBref[q, i, e] invJ[c, q, e, k]"""
        return np.einsum("qie,cqek->cqik", dBref, invJ)

    def vmult(self, src, boundary_penalty_factor=1.0):
        u = np.asarray(src, dtype=np.float64).reshape(self.gm.n_cells, self.ND)
        out = np.zeros_like(u)
        dim = self.dim
        # cells, NS.cc:1241-1250
        dB = self._phys(self.dBref, self.invJ)
        val = u @ self.B.T                                               # [c, q]
        grad = np.einsum("cqik,ci->cqk", dB, u)
        out += np.einsum("cqk,cqik,cq->ci", grad, dB, self.JxW) + np.einsum("cq,qi,cq->ci", val / self.coeff(), self.B, self.JxW)
        # interior faces, NS.cc:1266-1290: minus cell = the one below in direction d
        for d in range(dim):
            nb = self.gm.nbr[:, 2 * d + 1]
            cm = np.nonzero(nb >= 0)[0]
            cp_ = nb[cm]
            fp, fm = self.face[(d, 1)], self.face[(d, 0)]
            dBp = self._phys(fp["dBref"], fp["invJ"][cm])
            dBm = self._phys(fm["dBref"], fm["invJ"][cp_])
            n_plus, JxW = fp["normal"][cm], fp["JxW"][cm]
            coef = self.C_p * 0.5 * (np.abs(np.einsum("ck,ck->c", n_plus[:, 0], fp["invJ"][cm][:, 0, d, :])) +
                                     np.abs(np.einsum("ck,ck->c", n_plus[:, 0], fm["invJ"][cp_][:, 0, d, :])))        # NS.cc:1274-1275
            vp, vm = u[cm] @ fp["B"].T, u[cp_] @ fm["B"].T
            gp, gm_ = np.einsum("cqik,ci->cqk", dBp, u[cm]), np.einsum("cqik,ci->cqk", dBm, u[cp_])
            avg_n = np.einsum("cqk,cqk->cq", 0.5 * (gp + gm_), n_plus)
            jump = vp - vm
            sub_p = -avg_n + coef[:, None] * jump                        # NS.cc:1284
            sub_m = avg_n - coef[:, None] * jump                         # NS.cc:1285
            gsub = -THETA_P * 0.5 * jump                                 # NS.cc:1286-1287 (times n_plus)
            np.add.at(out, cm, np.einsum("cq,qi,cq->ci", sub_p, fp["B"], JxW) + np.einsum("cq,cqk,cqik,cq->ci", gsub, n_plus, dBp, JxW))
            np.add.at(out, cp_, np.einsum("cq,qi,cq->ci", sub_m, fm["B"], JxW) + np.einsum("cq,cqk,cqik,cq->ci", gsub, n_plus, dBm, JxW))
        # boundary faces with id 1, NS.cc:1306-1324
        for f in range(2 * dim):
            d, s = f // 2, f % 2
            cells = np.nonzero(self.gm.nbr[:, f] == -2)[0]
            if len(cells) == 0:
                continue
            fe = self.face[(d, s)]
            dBf = self._phys(fe["dBref"], fe["invJ"][cells])
            n_plus, JxW = fe["normal"][cells], fe["JxW"][cells]
            coef = self.C_p * np.abs(np.einsum("ck,ck->c", n_plus[:, 0], fe["invJ"][cells][:, 0, d, :]))                 # NS.cc:1311
            v = u[cells] @ fe["B"].T
            g_n = np.einsum("cqik,ci,cqk->cq", dBf, u[cells], n_plus)
            sub = -g_n + boundary_penalty_factor * coef[:, None] * v     # NS.cc:1320 (diagonal: 2 coef, NS.cc:1997)
            nd = -THETA_P * v                                            # NS.cc:1321: submit_normal_derivative
            np.add.at(out, cells, np.einsum("cq,qi,cq->ci", sub, fe["B"], JxW) + np.einsum("cq,cqk,cqik,cq->ci", nd, n_plus, dBf, JxW))
        return out.reshape(-1).astype(self.dtype)

    def compute_diagonal(self):
        """compute_diagonal (NS.cc:2018-2060) as the reference probes it: dof i set on both sides of every face (NS.cc:1921-1922),
        boundary probe with 2 sigma (NS.cc:1997), entry i kept, inverted (NS.cc:2055-2059)"""
        diag = np.zeros((self.gm.n_cells, self.ND))
        for i in range(self.ND):
            probe = np.zeros((self.gm.n_cells, self.ND))
            probe[:, i] = 1.0
            diag[:, i] = self.vmult(probe.reshape(-1), 2.0).reshape(self.gm.n_cells, self.ND)[:, i]
        self.inverse_diagonal = (1.0 / diag).reshape(-1).astype(self.dtype)
        return self.inverse_diagonal


class MassOperatorGeneral:
    """NS_stage 3 (assemble_cell_term_projection_grad_p, NS.cc:1373-1390) on MappingQ1 cells: per component
    M_ij = sum_q phi_i(x_q) phi_j(x_q) JxW_q with the velocity quadrature QGauss(degree_v + 1); dense"""

    def __init__(self, gmesh: GeneralMesh, degree_v: int, ncomp=None, dtype=np.float64):
        self.gm, self.degree, self.dim = gmesh, degree_v, gmesh.dim
        self.ncomp = gmesh.dim if ncomp is None else ncomp
        self.dtype = np.dtype(dtype).type
        n = degree_v + 1
        self.ND = n ** self.dim
        self.n_dofs = gmesh.n_cells * self.ncomp * self.ND
        nodes = gauss_lobatto_nodes(n)
        xq, wq = gauss_legendre(n)
        tq = tuple(np.asarray(a, np.float64) for a in lagrange_eval(nodes, xq))
        self.B, _ = _tensor_tables([tq] * self.dim)
        w = np.ones(1)
        for _d in range(self.dim):
            w = np.kron(np.asarray(wq, np.float64), w)
        _inv, det = q1_jacobian(gmesh.cell_vertices(), _points([np.asarray(xq, np.float64)] * self.dim))
        self.JxW = det * w
        self.M = np.einsum("qi,qj,cq->cij", self.B, self.B, self.JxW)

    def vmult(self, src):
        u = np.asarray(src, dtype=np.float64).reshape(self.gm.n_cells, self.ncomp, self.ND)
        return np.einsum("cij,cmj->cmi", self.M, u).reshape(-1).astype(self.dtype)

    def inverse(self, src):
        u = np.asarray(src, dtype=np.float64).reshape(self.gm.n_cells, self.ncomp, self.ND)
        return np.stack([np.linalg.solve(self.M[c], u[c].T).T for c in range(self.gm.n_cells)]).reshape(-1).astype(self.dtype)
