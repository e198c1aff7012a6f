"""projection_step's linear solve (NS.cc:2470-2548) restated on the oracle classes.

Test infrastructure only.  Mirrors, in order: SolverControl(max_its, eps*||rhs||) NS.cc:2486;
level operators fp32 with set_dt / set_NS_stage(2) NS.cc:2353-2379; smoother data NS.cc:2501-2516
(range 15, degree 3, 10 eigenvalue CG iterations on levels > 0; level 0 is served by mg_coarse);
MGCoarseGridIterativeSolver with plain CG sharing the outer SolverControl NS.cc:2519-2529;
V-cycle preconditioner NS.cc:2531-2537; outer CG NS.cc:2541-2546.
"""
from __future__ import annotations

import numpy as np

from .mesh import CartesianMesh
from .operators import PressureOperator
from .solvers import Chebyshev, Multigrid, SolverControl, Transfer, solver_cg


def build_hierarchy(mesh: CartesianMesh):
    levels = [mesh]
    while levels[0].level > 0:
        levels.insert(0, levels[0].coarser())
    return levels


class PressureMGSolver:
    def __init__(self, mesh, degree, dt, stage=1, level_dtype=np.float32, op_class=PressureOperator):
        """op_class(mesh, degree, dt, dtype): PressureOperator on a CartesianMesh, PressureOperatorGeneral on a GeneralMesh (the
        hierarchy, smoothers, transfers and solvers do not depend on the geometry)"""
        self.levels = build_hierarchy(mesh)
        self.op = op_class(mesh, degree, dt, np.float64)
        self.op.set_TR_BDF2_stage(stage)
        self.level_ops = []
        for m in self.levels:
            lop = op_class(m, degree, dt, level_dtype)
            lop.set_TR_BDF2_stage(stage)
            self.level_ops.append(lop)
        self.level_dtype = np.dtype(level_dtype).type
        self.degree = degree

    def setup(self):
        self.smoothers, self.transfers = [None], [None]
        for l, lop in enumerate(self.level_ops):
            lop.compute_diagonal()                                   # NS.cc:2514
            if l > 0:
                self.smoothers.append(Chebyshev(lop.vmult, lop.inverse_diagonal, degree=3,
                                                smoothing_range=15.0, eig_cg_n_iterations=10))
                self.transfers.append(Transfer(getattr(self.levels[l], "cart", self.levels[l]), self.degree, self.level_dtype))

    def solve(self, x, rhs, eps=1e-8, max_its=10000):
        self.setup()
        control = SolverControl(max_its, eps * float(np.sqrt(rhs @ rhs)))   # NS.cc:2486
        self.coarse_iterations = []

        def coarse(b):
            xc = np.zeros_like(b)
            cc = SolverControl(control.max_steps, control.tol)       # same tolerance object
            self.coarse_iterations.append(solver_cg(self.level_ops[0].vmult, xc, b, None, cc))
            return xc

        mg = Multigrid([lop.vmult for lop in self.level_ops], self.smoothers, self.transfers, coarse,
                       self.level_dtype)
        its = solver_cg(self.op.vmult, x, rhs, precond=mg.vmult, control=control)
        self.control = control
        return its
