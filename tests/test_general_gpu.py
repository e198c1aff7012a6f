"""GPU parity of the pressure operator on MappingQ1 (non-Cartesian) levels -- csrc/pressure_general.cu through the C ABI -- against
the dense oracle (oracle/general.py): operator, vmult_add, probed diagonal, Chebyshev smoother, multigrid-preconditioned CG."""
import numpy as np
import pytest

from oracle.general import GeneralMesh, PressureOperatorGeneral
from oracle.mesh import CartesianMesh
from oracle.projection import PressureMGSolver
from oracle.solvers import Chebyshev, SolverControl, solver_cg
from test_general_cpu import wavy
from util import rel_l2

pytestmark = pytest.mark.gpu


def meshes(dgx, dim, cells0, nref, periodic, bids, fn, hi=None):
    lo, hi = [0.0] * dim, [1.0] * dim if hi is None else hi
    cm = CartesianMesh(dim, cells0, nref, lo, hi, periodic, bids)
    gm = dgx.Mesh(dim, cells0, nref, lo, hi, periodic, bids)
    gm.transform(fn, lo, hi, cells0)
    return gm, GeneralMesh(cm, fn)


CASES = [
    (2, [2, 1], 2, [False, False], [1, 0, 1, 3], [1, 2, 3, 4]),
    (2, [1, 1], 3, [True, True], None, [1, 2, 5]),
    (2, [3, 2], 1, [True, False], [0, 1, 1, 1], [2, 6]),
    (3, [1, 1, 1], 2, [True, True, True], None, [1, 2, 3]),
    (3, [2, 1, 1], 1, [False, True, False], [1, 1, 2, 3, 1, 5], [1, 2, 4]),
    (3, [1, 1, 2], 1, [False, False, False], [1, 1, 1, 1, 1, 1], [3, 5]),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}d-{c[1]}-r{c[2]}-{''.join('p' if x else 'b' for x in c[3])}")
def test_general_operator_matches_oracle(dgx, gpu_ctx, case):
    dim, cells0, nref, periodic, bids, degrees = case
    gm, om = meshes(dgx, dim, cells0, nref, periodic, bids, wavy(0.03))
    assert gm.is_general()
    dt = 5e-3
    for dtype, tol in ((dgx.F64, 1e-12), (dgx.F32, 2e-5)):
        mf = dgx.MatrixFree(gpu_ctx, gm, dtype=dtype)
        for degree in degrees:
            op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, dt)
            oop = PressureOperatorGeneral(om, degree, dt)
            u = np.random.default_rng(degree).standard_normal(oop.n_dofs)
            if dtype == dgx.F32:
                u = u.astype(np.float32).astype(np.float64)
            src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
            src.from_host(u)
            for stage in (1, 2):
                op.set_TR_BDF2_stage(stage); oop.set_TR_BDF2_stage(stage)
                op.vmult(dst, src)
                ref = oop.vmult(u)
                assert rel_l2(dst.to_host(), ref) < tol, (dtype, degree, stage)
            op.vmult_add(dst, src)
            assert rel_l2(dst.to_host(), 2 * ref) < tol
            if degree <= 2:
                op.compute_diagonal()
                assert rel_l2(op.get_matrix_diagonal_inverse().to_host(), oop.compute_diagonal()) < tol * 10, (dtype, degree)
            assert mf.general_metric_bytes(degree) > 0
        with pytest.raises(dgx.DgxError):   # only the pressure operator has the metric path
            mop = dgx.NavierStokesProjectionOperator(mf, dgx.OP_MASS, 2, 100.0, dt)
            a, b = mop.initialize_dof_vector(), mop.initialize_dof_vector()
            mop.vmult(a, b)


def test_identity_vertices_reproduce_the_cartesian_kernels(dgx, gpu_ctx):
    """same mesh once as a Cartesian box (Kronecker kernels) and once with explicit vertex positions (metric kernel)"""
    args = (3, [1, 1, 1], 2, [0.0] * 3, [1.0, 1.5, 0.75], [True, False, True])
    bids = [0, 1, 1, 1, 4, 5]
    ga, gb = dgx.Mesh(*args, bids), dgx.Mesh(*args, bids)
    gb.transform(lambda X: X, args[3], args[4], args[1])
    for degree in (1, 3, 4):
        res = []
        for g in (ga, gb):
            mf = dgx.MatrixFree(gpu_ctx, g)
            op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, 5e-3)
            src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
            src.from_host(np.random.default_rng(7).standard_normal(src.locally_owned_size()))
            op.vmult(dst, src)
            op.compute_diagonal()
            res.append((dst.to_host(), op.get_matrix_diagonal_inverse().to_host()))
        assert rel_l2(res[1][0], res[0][0]) < 1e-12 and rel_l2(res[1][1], res[0][1]) < 1e-11


def test_chebyshev_and_cg_on_a_deformed_mesh(dgx, gpu_ctx):
    gm, om = meshes(dgx, 2, [2, 2], 2, [False, True], [1, 0, 2, 3], wavy(0.03))
    mf = dgx.MatrixFree(gpu_ctx, gm)
    degree, dt = 2, 5e-3
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, dt)
    oop = PressureOperatorGeneral(om, degree, dt)
    op.compute_diagonal()
    dinv = oop.compute_diagonal()
    cheb = dgx.PreconditionChebyshev(op)
    est = cheb.estimate_eigenvalues()
    oc = Chebyshev(oop.vmult, dinv)
    oc.estimate_eigenvalues(oop.n_dofs, np.float64)
    assert abs(est["max_eig"] - oc.max_eig_est) < 1e-8 * oc.max_eig_est
    b = np.random.default_rng(2).standard_normal(oop.n_dofs)
    vb, vx = op.initialize_dof_vector(), op.initialize_dof_vector()
    vb.from_host(b)
    cheb.vmult(vx, vb)                       # fused operator + update launches of the metric kernel
    assert rel_l2(vx.to_host(), oc.vmult(b)) < 1e-10
    x0 = vx.to_host()
    cheb.step(vx, vb)
    assert rel_l2(vx.to_host(), oc.step(x0, b)) < 1e-10
    tol = 1e-10 * np.linalg.norm(b)
    vx.set(0.0)
    its, _ = dgx.solve_cg(op, vx, vb, "jacobi", 10000, tol)
    x = np.zeros_like(b)
    oits = solver_cg(oop.vmult, x, b, lambda r: dinv * r, SolverControl(10000, tol))
    assert abs(its - oits) <= 1 and rel_l2(vx.to_host(), x) < 1e-8


@pytest.mark.parametrize("dim,nref,degree", [(2, 3, 2), (3, 2, 1)])
def test_mg_cg_on_a_deformed_mesh(dgx, gpu_ctx, dim, nref, degree):
    """projection_step's solve (NS.cc:2486-2546) with every level a MappingQ1 mesh: fp64 CG, fp32 V-cycle with the metric kernel"""
    gm, om = meshes(dgx, dim, [1] * dim, nref, [True] * dim, None, wavy(0.02))
    dt, eps = 5e-3, 1e-8
    mf = dgx.MatrixFree(gpu_ctx, gm)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, dt)
    mg = dgx.Multigrid(gpu_ctx, gm, degree, 100.0, dt)
    osolver = PressureMGSolver(om, degree, dt, 1, op_class=PressureOperatorGeneral)
    rhs = np.random.default_rng(4).standard_normal(osolver.op.n_dofs)
    rhs -= rhs.mean()
    tol = eps * np.linalg.norm(rhs)
    mg.setup(tol, 10000)
    vb, vx = op.initialize_dof_vector(), op.initialize_dof_vector()
    vb.from_host(rhs)
    its, res = dgx.solve_cg(op, vx, vb, mg, 10000, tol)
    x = np.zeros_like(rhs)
    oits = osolver.solve(x, rhs, eps)
    assert abs(its - oits) <= 1, (its, oits)
    assert rel_l2(vx.to_host(), x) < 1e-4
    assert np.linalg.norm(osolver.op.vmult(vx.to_host()) - rhs) <= 1.5 * tol
