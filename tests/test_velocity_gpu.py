"""GPU parity: velocity operator (NS_stage 1: SIPG viscous + centred/Lax-Friedrichs convective flux with the
advecting field u_extr, NS.cc:933-1223), its probed diagonal (NS.cc:1431-1850) and the Jacobi-GMRES solve of
diffusion_step (NS.cc:2450-2463) through the C ABI vs the oracle.  fp64, rel. L2 1e-12."""
import numpy as np
import pytest

from oracle.operators import VelocityOperator
from oracle.solvers import SolverControl, solver_gmres
from util import make_meshes, rel_l2, smooth_plus_noise

pytestmark = pytest.mark.gpu

CASES = [
    (2, [4, 2], 2, [False, False], [1, 2, 3]),            # channel: ids 0 inflow, 1 outflow (mirror), 2/3 walls
    (2, [1, 1], 3, [True, True], [2, 4]),
    (2, [3, 1], 1, [False, True], [2]),
    (3, [1, 1, 1], 2, [True, True, True], [1, 2, 3]),
    (3, [2, 1, 1], 1, [False, False, False], [2, 3]),      # every boundary kind in 3-D
    (3, [1, 1, 1], 0, [True, True, True], [2]),
]


def _setup(dgx, ctx, dim, cells0, nref, periodic, degree, stage, Re=100.0, dt=5e-3):
    gm, om = make_meshes(dgx, dim, cells0, nref, periodic)
    mf = dgx.MatrixFree(ctx, gm)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_VELOCITY, degree, Re, dt)
    op.set_TR_BDF2_stage(stage)
    oop = VelocityOperator(om, degree, Re, dt)
    oop.set_TR_BDF2_stage(stage)
    ue = smooth_plus_noise(om, oop.sh.nodes, seed=11, ncomp=dim, noise=0.4) + 0.3
    u = smooth_plus_noise(om, oop.sh.nodes, seed=12, ncomp=dim, noise=0.5)
    vue = op.initialize_dof_vector()
    vue.from_host(ue)
    op.set_u_extr(vue)
    oop.set_u_extr(ue)
    return op, oop, u, om


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}d-{c[1]}-r{c[2]}-{''.join('p' if x else 'b' for x in c[3])}")
def test_velocity_vmult_matches_oracle(dgx, gpu_ctx, case):
    dim, cells0, nref, periodic, degrees = case
    for degree in degrees:
        for stage in (1, 2):
            op, oop, u, om = _setup(dgx, gpu_ctx, dim, cells0, nref, periodic, degree, stage)
            src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
            src.from_host(u)
            op.vmult(dst, src)
            ref = oop.vmult(u)
            assert rel_l2(dst.to_host(), ref) < 1e-12, (degree, stage)
            op.vmult_add(dst, src)
            assert rel_l2(dst.to_host(), 2 * ref) < 1e-12


def test_velocity_requires_u_extr(dgx, gpu_ctx):
    gm, om = make_meshes(dgx, 2, [1, 1], 2, [True, True])
    mf = dgx.MatrixFree(gpu_ctx, gm)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_VELOCITY, 2, 100.0, 5e-3)
    a, b = op.initialize_dof_vector(), op.initialize_dof_vector()
    with pytest.raises(dgx.DgxError):
        op.vmult(a, b)


@pytest.mark.parametrize("case", [CASES[0], CASES[3], CASES[4]], ids=["2d-channel", "3d-periodic", "3d-box"])
def test_velocity_diagonal_matches_reference_probing(dgx, gpu_ctx, case):
    dim, cells0, nref, periodic, degrees = case
    degree = degrees[1]
    op, oop, u, om = _setup(dgx, gpu_ctx, dim, cells0, nref, periodic, degree, 1)
    op.compute_diagonal()
    d = op.get_matrix_diagonal_inverse().to_host()
    assert rel_l2(d, oop.compute_diagonal()) < 1e-12
    # the tensorised form (default) equals literal probing with n^dim operator applications (variant 1)
    op.set_variant(1)
    op.compute_diagonal()
    assert rel_l2(op.get_matrix_diagonal_inverse().to_host(), d) < 1e-12


def test_diffusion_step_solve(dgx, gpu_ctx):
    """Jacobi-preconditioned GMRES on the velocity operator, tolerance eps * ||rhs|| (NS.cc:2450-2463)."""
    op, oop, u, om = _setup(dgx, gpu_ctx, 2, [4, 2], 2, [False, False], 2, 1)
    op.compute_diagonal()
    dinv = oop.compute_diagonal()
    rhs = oop.vmult(u) + 0.05 * np.random.default_rng(3).standard_normal(u.size)
    tol = 1e-8 * np.linalg.norm(rhs)
    vb, vx = op.initialize_dof_vector(), op.initialize_dof_vector()
    vb.from_host(rhs)
    x0 = 0.9 * u
    vx.from_host(x0)
    its, res = dgx.solve_gmres(op, vx, vb, "jacobi", 10000, tol)
    x = x0.copy()
    oits = solver_gmres(oop.vmult, x, rhs, lambda r: dinv * r, SolverControl(10000, tol))
    assert abs(its - oits) <= 1, (its, oits)
    assert rel_l2(vx.to_host(), x) < 1e-7
