"""Host-side mirror of NavierStokesProjection<dim> (NS.cc:2068-3020): the TR-BDF2 projection time loop that drives the
device operators and solvers.  Every linear-algebra object lives on the device (libdgx); this class only sequences the
calls exactly as NavierStokesProjection::run does (NS.cc:2928-3010) and records the Krylov iteration counts, which the
reference only prints through deallog.
"""
from __future__ import annotations

import os

from . import (GAMMA, OP_MASS, OP_PRESSURE, OP_VELOCITY, MatrixFree, Multigrid, NavierStokesProjectionOperator, compute_lift_and_drag,
               estimate_vorticity_indicator, interpolate_velocity, output_results, refine_and_coarsen_fixed_number, solve_cg, solve_gmres)


class NavierStokesProjection:
    def __init__(self, ctx, mesh, degree_p=1, Re=100.0, dt=5e-3, max_its=10000, eps=1e-8, direct_mass_inverse=False):
        """Data_Storage values as parameter-file.prm:1-39 (Reynolds, dt, max_iterations, eps).  direct_mass_inverse: replace
        project_grad's CG on the DG mass matrix by its exact tensor-product inverse (not the reference's algorithm: opt-in)."""
        self.direct_mass_inverse = direct_mass_inverse
        self.ctx, self.mesh, self.dim = ctx, mesh, mesh.dim
        self.degree_p, self.Re, self.dt, self.max_its, self.eps = degree_p, Re, dt, max_its, eps
        self.mf = MatrixFree(ctx, mesh)                                                  # NS.cc:2329
        dv = degree_p + 1
        self.vel_op = NavierStokesProjectionOperator(self.mf, OP_VELOCITY, dv, Re, dt)   # navier_stokes_matrix, NS_stage 1
        self.pres_op = NavierStokesProjectionOperator(self.mf, OP_PRESSURE, degree_p, Re, dt)   # NS_stage 2
        self.mass_op = NavierStokesProjectionOperator(self.mf, OP_MASS, dv, Re, dt)      # NS_stage 3
        self.mg = Multigrid(ctx, mesh, degree_p, Re, dt)                                 # mg_matrices, NS.cc:2351-2379
        v = lambda: self.vel_op.initialize_dof_vector()
        p = lambda: self.pres_op.initialize_dof_vector()
        self.u_n, self.u_n_minus_1, self.u_extr, self.u_star, self.u_n_gamma, self.u_tmp, self.grad_pres_int, self.rhs_u = (v() for _ in range(8))
        self.pres_n, self.pres_int, self.rhs_p = p(), p(), p()                            # NS.cc:2330-2341
        self.TR_BDF2_stage = 1
        self.iterations = []          # (name, iterations) of every solve, in order
        self.time, self.n = 0.0, 1

    def set_dt(self, dt):
        self.dt = dt
        for op in (self.vel_op, self.pres_op, self.mass_op):
            op.set_dt(dt)
        self.mg.set_dt(dt)

    def _set_stage(self, stage):                                                          # NS.cc:2940-2943, 2967-2969
        self.TR_BDF2_stage = stage
        for op in (self.vel_op, self.pres_op, self.mass_op):
            op.set_TR_BDF2_stage(stage)
        self.mg.set_TR_BDF2_stage(stage)

    def interpolate_velocity(self):                                                       # NS.cc:2403-2418
        g = GAMMA
        if self.TR_BDF2_stage == 1:
            interpolate_velocity(self.u_extr, 1, self.u_n, self.u_n_minus_1)
            self.u_tmp.equ(g / (2.0 * (1.0 - g)), self.u_n_minus_1)      # side effect of NS.cc:2409: u_tmp is the initial guess of
        else:                                                             # the next project_grad solve (NS.cc:2576)
            interpolate_velocity(self.u_extr, 2, self.u_n_gamma, self.u_n)
            self.u_tmp.equ((1.0 - g) / g, self.u_n)

    def diffusion_step(self):                                                             # NS.cc:2424-2464
        op = self.vel_op
        if self.TR_BDF2_stage == 1:
            op.vmult_rhs_velocity(self.rhs_u, [self.u_n, self.u_extr, self.pres_n])
        else:
            op.vmult_rhs_velocity(self.rhs_u, [self.u_n, self.u_n_gamma, self.pres_int, self.u_extr])
        op.set_u_extr(self.u_extr)
        self.u_star.copy_from(self.u_extr)
        tol = self.eps * self.rhs_u.l2_norm()
        op.compute_diagonal()
        its, _ = solve_gmres(op, self.u_star, self.rhs_u, "jacobi", self.max_its, tol)
        self.iterations.append((f"gmres_velocity_stage{self.TR_BDF2_stage}", its))

    def projection_step(self):                                                            # NS.cc:2470-2548
        op = self.pres_op
        old = self.pres_n if self.TR_BDF2_stage == 1 else self.pres_int
        op.vmult_rhs_pressure(self.rhs_p, [self.u_star, old])
        tol = self.eps * self.rhs_p.l2_norm()
        self.mg.setup(tol, self.max_its)                                                  # NS.cc:2490-2529 rebuilt per solve
        new = self.pres_int if self.TR_BDF2_stage == 1 else self.pres_n
        new.copy_from(old)                                                                # initial guess, NS.cc:2541 / 2545
        its, _ = solve_cg(op, new, self.rhs_p, self.mg, self.max_its, tol)
        self.iterations.append((f"cg_pressure_stage{self.TR_BDF2_stage}", its))

    def project_grad(self, flag):                                                         # NS.cc:2554-2577
        assert flag in (1, 2)
        self.mass_op.vmult_grad_p_projection(self.rhs_u, self.pres_n if flag == 1 else self.pres_int)
        if self.direct_mass_inverse:
            self.mass_op.mass_inverse(self.u_tmp, self.rhs_u)
            self.iterations.append((f"direct_mass_flag{flag}", 0))
            return
        tol = 1e-12 * self.rhs_u.l2_norm()
        its, _ = solve_cg(self.mass_op, self.u_tmp, self.rhs_u, None, self.max_its, tol)
        self.iterations.append((f"cg_mass_flag{flag}", its))

    def advance(self):
        """one pass of the while-loop body of NavierStokesProjection::run (NS.cc:2934-2986)"""
        g, dt = GAMMA, self.dt
        self.time += dt
        self.n += 1
        self._set_stage(1)
        self.interpolate_velocity()
        self.diffusion_step()
        self.project_grad(1)
        self.u_tmp.equ(g * dt, self.u_tmp)
        self.u_star.add(1.0, self.u_tmp)
        self.projection_step()
        self.u_n_gamma.equ(1.0, self.u_star)
        self.project_grad(2)
        self.grad_pres_int.equ(1.0, self.u_tmp)
        self.u_tmp.equ(-g * dt, self.u_tmp)
        self.u_n_gamma.add(1.0, self.u_tmp)
        self.u_n_minus_1.copy_from(self.u_n)
        self._set_stage(2)
        self.interpolate_velocity()
        self.diffusion_step()
        self.u_tmp.equ((1.0 - g) * dt, self.grad_pres_int)
        self.u_star.add(1.0, self.u_tmp)
        self.projection_step()
        self.u_n.equ(1.0, self.u_star)
        self.project_grad(1)
        self.u_tmp.equ((g - 1.0) * dt, self.u_tmp)
        self.u_n.add(1.0, self.u_tmp)
        return self.u_n.linfty_norm()                                                     # get_maximal_velocity, NS.cc:2988

    def compute_lift_and_drag(self, boundary_id=4):                                       # NS.cc:2642-2708
        """returns (lift, drag) of the faces with `boundary_id` (the reference appends them to lift.dat / drag.dat)"""
        return compute_lift_and_drag(self.u_n, self.pres_n, self.Re, boundary_id)

    def output_results(self, step, saving_dir="."):                                       # NS.cc:2608-2633
        path = os.path.join(saving_dir, f"solution-{step:05d}.vtu")
        output_results(self.u_n, self.pres_n, self.u_n_minus_1, path)
        return path

    def refinement_flags(self, top_fraction=0.01, bottom_fraction=0.3):                    # NS.cc:2727-2771
        """first half of refine_mesh: the vorticity indicator of u_n per cell (device) and the refine / coarsen flags
        refine_and_coarsen_fixed_number would set (+1 / -1 / 0).  Executing them needs hanging-node faces: not built yet."""
        eta = estimate_vorticity_indicator(self.u_n, self.degree_p + 2)
        return eta, refine_and_coarsen_fixed_number(eta, top_fraction, bottom_fraction)

    def run(self, T, output_interval=0, saving_dir=".", verbose=False, boundary_id=4):
        """NavierStokesProjection::run (NS.cc:2928-3010) without mesh refinement: returns the (lift, drag) history"""
        forces = []
        if output_interval:
            self.output_results(1, saving_dir)
        while abs(T - self.time) > 1e-10:
            max_vel = self.advance()
            if verbose and self.ctx.rank == 0:
                print(f"Step = {self.n} Time = {self.time}\nMaximal velocity = {max_vel}")
            forces.append(self.compute_lift_and_drag(boundary_id))
            if output_interval and self.n % output_interval == 0:
                self.output_results(self.n, saving_dir)
            if T - self.time < self.dt and T - self.time > 1e-10:                         # NS.cc:2996-3002
                self.set_dt(T - self.time)
        if output_interval and self.n % output_interval != 0:
            self.output_results(self.n, saving_dir)
        return forces
