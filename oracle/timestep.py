"""NavierStokesProjection::run's loop body (NS.cc:2934-2986) on the oracle classes (test infrastructure only)."""
from __future__ import annotations

import numpy as np

from .operators import GAMMA, MassOperator, VelocityOperator
from .projection import PressureMGSolver
from .rhs import RhsForms
from .solvers import SolverControl, solver_cg, solver_gmres


class OracleProjection:
    def __init__(self, mesh, degree_p=1, Re=100.0, dt=5e-3, max_its=10000, eps=1e-8):
        self.mesh, self.dp, self.Re, self.dt, self.max_its, self.eps = mesh, degree_p, Re, dt, max_its, eps
        self.vel = VelocityOperator(mesh, degree_p + 1, Re, dt)
        self.mass = MassOperator(mesh, degree_p + 1)
        self.rhs = RhsForms(mesh, degree_p, Re, dt)
        nv, npd = self.vel.n_dofs, mesh.n_cells * (degree_p + 1) ** mesh.dim
        z = lambda n: np.zeros(n)
        self.u_n, self.u_n_minus_1, self.u_extr, self.u_star, self.u_n_gamma, self.u_tmp, self.grad_pres_int = (z(nv) for _ in range(7))
        self.pres_n, self.pres_int = z(npd), z(npd)
        self.stage = 1
        self.iterations = []

    def interpolate_velocity(self):
        g = GAMMA
        # equ / equ / -= as NS.cc:2408-2416: u_tmp keeps the scaled second operand and is the initial guess of the next
        # project_grad solve (NS.cc:2576)
        if self.stage == 1:
            self.u_tmp = (g / (2.0 * (1.0 - g))) * self.u_n_minus_1
            self.u_extr = (1.0 + g / (2.0 * (1.0 - g))) * self.u_n - self.u_tmp
        else:
            self.u_tmp = ((1.0 - g) / g) * self.u_n
            self.u_extr = (1.0 + (1.0 - g) / g) * self.u_n_gamma - self.u_tmp

    def diffusion_step(self):
        self.vel.set_TR_BDF2_stage(self.stage)
        if self.stage == 1:
            rhs = self.rhs.rhs_velocity(1, [self.u_n, self.u_extr, self.pres_n])
        else:
            rhs = self.rhs.rhs_velocity(2, [self.u_n, self.u_n_gamma, self.pres_int, self.u_extr])
        self.vel.set_u_extr(self.u_extr)
        self.u_star = self.u_extr.copy()
        dinv = self.vel.compute_diagonal()
        ctl = SolverControl(self.max_its, self.eps * np.linalg.norm(rhs))
        its = solver_gmres(self.vel.vmult, self.u_star, rhs, lambda r: dinv * r, ctl)
        self.iterations.append((f"gmres_velocity_stage{self.stage}", its))

    def projection_step(self):
        old = self.pres_n if self.stage == 1 else self.pres_int
        rhs = self.rhs.rhs_pressure(self.stage, [self.u_star, old])
        solver = PressureMGSolver(self.mesh, self.dp, self.dt, self.stage)
        x = old.copy()
        its = solver.solve(x, rhs, self.eps, self.max_its)
        if self.stage == 1:
            self.pres_int = x
        else:
            self.pres_n = x
        self.iterations.append((f"cg_pressure_stage{self.stage}", its))

    def project_grad(self, flag):
        rhs = self.rhs.rhs_grad_p(self.pres_n if flag == 1 else self.pres_int)
        ctl = SolverControl(self.max_its, 1e-12 * np.linalg.norm(rhs))
        its = solver_cg(self.mass.vmult, self.u_tmp, rhs, None, ctl)
        self.iterations.append((f"cg_mass_flag{flag}", its))

    def advance(self):
        g, dt = GAMMA, self.dt
        self.stage = 1
        self.interpolate_velocity()
        self.diffusion_step()
        self.project_grad(1)
        self.u_star = self.u_star + g * dt * self.u_tmp
        self.u_tmp = g * dt * self.u_tmp
        self.projection_step()
        self.u_n_gamma = self.u_star.copy()
        self.project_grad(2)
        self.grad_pres_int = self.u_tmp.copy()
        self.u_tmp = -g * dt * self.u_tmp
        self.u_n_gamma = self.u_n_gamma + self.u_tmp
        self.u_n_minus_1 = self.u_n.copy()
        self.stage = 2
        self.interpolate_velocity()
        self.diffusion_step()
        self.u_tmp = (1.0 - g) * dt * self.grad_pres_int
        self.u_star = self.u_star + self.u_tmp
        self.projection_step()
        self.u_n = self.u_star.copy()
        self.project_grad(1)
        self.u_tmp = (g - 1.0) * dt * self.u_tmp
        self.u_n = self.u_n + self.u_tmp
        return float(np.abs(self.u_n).max())
