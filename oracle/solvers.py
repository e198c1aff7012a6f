"""Restatement of the deal.II solver stack the reference wires up (oracle; test infrastructure).

Configuration follows navier_stokes_TRBDF2_DG.cc (NS.cc):
  diffusion_step  NS.cc:2424-2464  Jacobi-preconditioned GMRES, abs tol eps*||rhs||
  projection_step NS.cc:2470-2548  CG preconditioned by a V-cycle (fp32 levels):
                                   Chebyshev(degree 3, range 15, 10 eig CG its) smoother with the
                                   level inverse diagonal, plain CG coarse solver sharing the outer
                                   SolverControl, MGTransferMatrixFree
  project_grad    NS.cc:2554-2577  unpreconditioned CG on the DG mass matrix, tol 1e-12*||rhs||
Algorithms are deal.II's (deal.II is not in /root/reference).
"""
from __future__ import annotations

import numpy as np

from .dg1d import shape
from .feeval import apply1d, _ax


class NoConvergence(RuntimeError):
    def __init__(self, last_step, last_residual):
        super().__init__(f"no convergence after {last_step} steps, residual {last_residual}")
        self.last_step, self.last_residual = last_step, last_residual


class SolverControl:
    """deal.II SolverControl: success when value <= tol, failure when step >= max_steps."""

    def __init__(self, max_steps, tol):
        self.max_steps, self.tol = int(max_steps), float(tol)
        self.last_step, self.last_value = 0, 0.0
        self.history = []

    def check(self, step, value):
        self.last_step, self.last_value = step, float(value)
        self.history.append(float(value))
        if value <= self.tol:
            return "success"
        if step >= self.max_steps or not np.isfinite(value):
            return "failure"
        return "iterate"


def solver_cg(A, x, b, precond=None, control=None, lanczos=None):
    """deal.II SolverCG.  ``A``/``precond`` are callables v -> A v.  x is updated
    in place.  ``lanczos``: optional dict receiving the tridiagonal (diagonal/offdiagonal) built
    the way SolverCG feeds its eigenvalue slot: entries are pushed from the second iteration on,
    using the previous alpha and the current beta."""
    dt = x.dtype.type
    r = b - A(x)
    res = float(np.sqrt(np.dot(r, r)))
    state = control.check(0, res)
    it = 0
    diag, offdiag = [], []
    eigen_beta_alpha = 0.0
    alpha = beta = prev_alpha = 0.0
    p = None
    r_dot_z = 0.0
    while state == "iterate":
        it += 1
        prev_alpha = alpha
        z = precond(r) if precond is not None else r
        if it > 1:
            prev_rz = r_dot_z
            r_dot_z = float(np.dot(r, z))
            beta = r_dot_z / prev_rz
            p = dt(beta) * p + z
        else:
            p = z.copy()
            r_dot_z = float(np.dot(r, z))
        v = A(p)
        alpha = r_dot_z / float(np.dot(p, v))
        x += dt(alpha) * p
        r = r - dt(alpha) * v
        res = float(np.sqrt(np.dot(r, r)))
        if it > 1:
            diag.append(1.0 / prev_alpha + eigen_beta_alpha)
            eigen_beta_alpha = beta / prev_alpha
            offdiag.append(np.sqrt(beta) / prev_alpha)
        state = control.check(it, res)
    if lanczos is not None:
        lanczos["diagonal"], lanczos["offdiagonal"] = diag, offdiag
    if state == "failure":
        raise NoConvergence(control.last_step, control.last_value)
    return it


def lanczos_eigenvalues(diag, offdiag):
    n = len(diag)
    if n == 0:
        return np.array([])
    T = np.diag(np.asarray(diag, dtype=np.float64))
    for i in range(n - 1):
        T[i, i + 1] = T[i + 1, i] = offdiag[i]
    return np.linalg.eigvalsh(T)


def solver_gmres(A, x, b, precond, control, max_basis=30):
    """Left-preconditioned restarted GMRES monitoring the preconditioned residual (deal.II
    SolverGMRES defaults of deal.II >= 9.6, which the reference requires: AdditionalData::max_basis_size = 30 is the Krylov
    dimension itself; before 9.6 max_n_tmp_vectors = 30 meant 28)."""
    dt = x.dtype.type
    kdim = max_basis
    it_total = 0
    while True:
        r = precond(b - A(x))
        rho = float(np.sqrt(np.dot(r, r)))
        state = control.check(it_total, rho)
        if state != "iterate":
            break
        V = [r / dt(rho)]
        H = np.zeros((kdim + 1, kdim))
        g = np.zeros(kdim + 1)
        g[0] = rho
        cs, sn = np.zeros(kdim), np.zeros(kdim)
        k_used = 0
        for k in range(kdim):
            w = precond(A(V[k]))
            # classical Gram-Schmidt with one re-orthogonalisation pass
            for _ in range(2):
                hcol = np.array([float(np.dot(w, vj)) for vj in V])
                for hj, vj in zip(hcol, V):
                    w = w - dt(hj) * vj
                H[: k + 1, k] += hcol
            hn = float(np.sqrt(np.dot(w, w)))
            H[k + 1, k] = hn
            for j in range(k):
                t = cs[j] * H[j, k] + sn[j] * H[j + 1, k]
                H[j + 1, k] = -sn[j] * H[j, k] + cs[j] * H[j + 1, k]
                H[j, k] = t
            den = np.hypot(H[k, k], H[k + 1, k])
            cs[k], sn[k] = H[k, k] / den, H[k + 1, k] / den
            H[k, k], H[k + 1, k] = den, 0.0
            g[k + 1] = -sn[k] * g[k]
            g[k] = cs[k] * g[k]
            it_total += 1
            k_used = k + 1
            rho = abs(g[k + 1])
            state = control.check(it_total, rho)
            if state != "iterate" or hn == 0.0:
                break
            V.append(w / dt(hn))
        y = np.linalg.solve(np.triu(H[:k_used, :k_used]), g[:k_used])
        for j in range(k_used):
            x += dt(y[j]) * V[j]
        if state != "iterate":
            break
    if state == "failure":
        raise NoConvergence(control.last_step, control.last_value)
    return it_total


class Chebyshev:
    """PreconditionChebyshev with a diagonal inner preconditioner (NS.cc:2493-2516)."""

    def __init__(self, A, inv_diag, degree=3, smoothing_range=15.0, eig_cg_n_iterations=10):
        self.A, self.inv_diag = A, inv_diag
        self.degree, self.smoothing_range, self.eig_its = degree, smoothing_range, eig_cg_n_iterations
        self.initialized = False

    def estimate_eigenvalues(self, n, dtype):
        # set_initial_guess: v[i] = (global index) % 11, minus its mean
        v = (np.arange(n, dtype=np.int64) % 11).astype(dtype)
        v = v - dtype(v.mean(dtype=np.float64))
        x = np.zeros(n, dtype=dtype)
        ctl = SolverControl(self.eig_its, 1e-10)
        lz = {}
        try:
            solver_cg(self.A, x, v, precond=lambda r: self.inv_diag * r, control=ctl, lanczos=lz)
        except NoConvergence:
            pass
        ev = lanczos_eigenvalues(lz["diagonal"], lz["offdiagonal"])
        self.min_eig_est, self.max_eig_est = (float(ev[0]), float(ev[-1])) if len(ev) else (1.0, 1.0)
        max_eig = 1.2 * self.max_eig_est
        alpha = max_eig / self.smoothing_range if self.smoothing_range > 1.0 else min(0.9 * max_eig, self.min_eig_est)
        self.delta = (max_eig - alpha) * 0.5
        self.theta = (max_eig + alpha) * 0.5
        self.initialized = True

    def _iterate(self, x, x_old, b, start_k):
        T = x.dtype.type
        rhok, sigma = self.delta / self.theta, self.theta / self.delta
        for k in range(self.degree - 1):
            rhokp = 1.0 / (2.0 * sigma - rhok)
            f1, f2 = rhokp * rhok, 2.0 * rhokp / self.delta
            rhok = rhokp
            t = self.A(x)
            if x_old is None:       # iteration_index == 1: x_old == 0
                x_new = T(1.0 + f1) * x + T(f2) * (self.inv_diag * (b - t))
            else:
                x_new = T(1.0 + f1) * x - T(f1) * x_old + T(f2) * (self.inv_diag * (b - t))
            x_old, x = x, x_new
        return x

    def vmult(self, b):
        """zero initial guess: degree-1 matrix-vector products."""
        if not self.initialized:
            self.estimate_eigenvalues(len(b), b.dtype.type)
        T = b.dtype.type
        x = T(1.0 / self.theta) * (self.inv_diag * b)
        if self.degree < 2 or abs(self.delta) < 1e-40:
            return x
        return self._iterate(x, None, b, 1)

    def step(self, x, b):
        """non-zero initial guess: degree matrix-vector products; returns the new x."""
        if not self.initialized:
            self.estimate_eigenvalues(len(b), b.dtype.type)
        T = b.dtype.type
        x1 = x + T(1.0 / self.theta) * (self.inv_diag * (b - self.A(x)))
        if self.degree < 2 or abs(self.delta) < 1e-40:
            return x1
        return self._iterate(x1, x, b, 2)


class Transfer:
    """MGTransferMatrixFree for FE_DGQ on a uniformly refined Cartesian hierarchy:
    child values = (P_cx (x) P_cy (x) P_cz) parent values; restriction is the transpose."""

    def __init__(self, mesh_fine, degree, dtype=np.float32):
        self.dim, self.n = mesh_fine.dim, degree + 1
        self.fine, self.coarse = mesh_fine, mesh_fine.coarser()
        self.P = shape(degree, degree + 1, np.dtype(dtype).name).P
        self.dtype = np.dtype(dtype).type
        # child c of parent j is fine cell 2^dim * j + c in the hierarchical order
        self.nchild = 1 << self.dim

    def prolongate(self, src_coarse):
        n, dim = self.n, self.dim
        uc = src_coarse.reshape((self.coarse.n_cells, 1) + (n,) * dim)
        out = np.empty((self.coarse.n_cells, self.nchild) + (n,) * dim, dtype=self.dtype)
        for c in range(self.nchild):
            t = uc
            for d in range(dim):
                t = apply1d(t, self.P[(c >> d) & 1], _ax(dim, d))
            out[:, c] = t[:, 0]
        return out.reshape(-1)

    def restrict_and_add(self, dst_coarse, src_fine):
        n, dim = self.n, self.dim
        uf = src_fine.reshape((self.coarse.n_cells, self.nchild, 1) + (n,) * dim)
        acc = dst_coarse.reshape((self.coarse.n_cells, 1) + (n,) * dim)
        for c in range(self.nchild):
            t = uf[:, c]
            for d in range(dim):
                t = apply1d(t, self.P[(c >> d) & 1].T, _ax(dim, d))
            acc += t
        return dst_coarse


class Multigrid:
    """Multigrid V-cycle + PreconditionMG (NS.cc:2531-2537).

    level_ops[l]: callable v -> A_l v (fp32), smoothers[l] (l >= 1): Chebyshev, transfers[l]:
    Transfer between level l and l-1, coarse_solve: callable b -> x on level 0."""

    def __init__(self, level_ops, smoothers, transfers, coarse_solve, dtype=np.float32):
        self.ops, self.smoothers, self.transfers, self.coarse = level_ops, smoothers, transfers, coarse_solve
        self.dtype = np.dtype(dtype).type
        self.nlevels = len(level_ops)

    def _level(self, l, defect):
        if l == 0:
            return self.coarse(defect)
        sol = self.smoothers[l].vmult(defect)                    # pre-smooth from zero
        t = defect - self.ops[l](sol)
        dc = np.zeros(self.transfers[l].coarse.n_cells * self.transfers[l].n ** self.transfers[l].dim,
                      dtype=self.dtype)
        self.transfers[l].restrict_and_add(dc, t)
        sc = self._level(l - 1, dc)
        sol = sol + self.transfers[l].prolongate(sc)             # prolongate_and_add
        return self.smoothers[l].step(sol, defect)               # post-smooth

    def vmult(self, src):
        """PreconditionMG::vmult: copy_to_mg (double -> float), V-cycle, copy_from_mg."""
        d = src.astype(self.dtype)
        return self._level(self.nlevels - 1, d).astype(src.dtype)
