"""GPU parity of the solver layer (transfer, Chebyshev, V-cycle, CG, GMRES) against the oracle's restatement of
the deal.II algorithms the reference configures in NS.cc:2450-2547, 2574-2576.  Tolerances: fp32 levels 1e-5
relative L2 (BASELINE.json), fp64 1e-12; Krylov iteration counts within +-1."""
import numpy as np
import pytest

from oracle.operators import MassOperator, PressureOperator
from oracle.projection import PressureMGSolver
from oracle.solvers import Chebyshev, SolverControl, Transfer, solver_cg, solver_gmres
from util import make_meshes, rel_l2, smooth_plus_noise

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,cells0,nref", [(2, [2, 1], 2), (3, [1, 1, 1], 2), (3, [1, 2, 1], 1)])
@pytest.mark.parametrize("np_dtype", [np.float32, np.float64])
def test_transfer(dgx, gpu_ctx, dim, cells0, nref, np_dtype):
    gm, om = make_meshes(dgx, dim, cells0, nref, [True] * dim)
    dt = dgx.F64 if np_dtype == np.float64 else dgx.F32
    tol = 1e-13 if np_dtype == np.float64 else 1e-6
    mff, mfc = dgx.MatrixFree(gpu_ctx, gm, nref, dt), dgx.MatrixFree(gpu_ctx, gm, nref - 1, dt)
    for degree in (1, 2, 3, 4):
        tr = Transfer(om, degree, np_dtype)
        n = degree + 1
        rng = np.random.default_rng(5)
        uc = rng.standard_normal(om.coarser().n_cells * n ** dim).astype(np_dtype)
        uf = rng.standard_normal(om.n_cells * n ** dim).astype(np_dtype)
        vf, vc = mff.initialize_dof_vector(degree), mfc.initialize_dof_vector(degree)
        vc.from_host(uc)
        vf.from_host(uf)
        dgx.transfer_prolongate_and_add(vf, vc)
        assert rel_l2(vf.to_host(), uf + tr.prolongate(uc)) < tol
        vf.from_host(uf)
        vc.from_host(uc)
        dgx.transfer_restrict_and_add(vc, vf)
        ref = uc.copy()
        tr.restrict_and_add(ref, uf)
        assert rel_l2(vc.to_host(), ref) < tol
        # adjointness on the device: <P c, f> = <c, P^T f>
        vf.set(0.0)
        vc.from_host(uc)
        dgx.transfer_prolongate_and_add(vf, vc)
        w = mff.initialize_dof_vector(degree)
        w.from_host(uf)
        lhs = vf.dot(w)
        vc2 = mfc.initialize_dof_vector(degree)
        dgx.transfer_restrict_and_add(vc2, w)
        rhs = vc2.dot(vc)
        assert abs(lhs - rhs) <= 20 * tol * max(abs(lhs), 1.0)


@pytest.mark.parametrize("np_dtype", [np.float32, np.float64])
def test_chebyshev(dgx, gpu_ctx, np_dtype):
    dim, cells0, nref, degree, dt = 3, [1, 1, 1], 2, 2, 5e-3
    gm, om = make_meshes(dgx, dim, cells0, nref, [True] * dim, hi=[1.0, 1.0, 1.0])
    dd = dgx.F64 if np_dtype == np.float64 else dgx.F32
    tol = 1e-11 if np_dtype == np.float64 else 2e-5
    mf = dgx.MatrixFree(gpu_ctx, gm, -1, dd)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, dt)
    op.compute_diagonal()
    oop = PressureOperator(om, degree, dt, np_dtype)
    oop.compute_diagonal()
    ch = dgx.PreconditionChebyshev(op, 3, 15.0, 10)
    och = Chebyshev(oop.vmult, oop.inverse_diagonal, 3, 15.0, 10)
    est = ch.estimate_eigenvalues()
    och.estimate_eigenvalues(oop.n_dofs, np_dtype)
    assert abs(est["max_eig"] - och.max_eig_est) < (1e-9 if np_dtype == np.float64 else 2e-4) * och.max_eig_est
    assert abs(est["theta"] - och.theta) < (1e-9 if np_dtype == np.float64 else 2e-4) * och.theta
    b = smooth_plus_noise(om, oop.sh.nodes, noise=0.2).astype(np_dtype)
    x0 = np.random.default_rng(2).standard_normal(b.size).astype(np_dtype)
    vb, vx = op.initialize_dof_vector(), op.initialize_dof_vector()
    vb.from_host(b)
    ch.vmult(vx, vb)
    assert rel_l2(vx.to_host(), och.vmult(b)) < tol
    vx.from_host(x0)
    ch.step(vx, vb)
    assert rel_l2(vx.to_host(), och.step(x0, b)) < tol


@pytest.mark.parametrize("np_dtype", [np.float32, np.float64])
@pytest.mark.parametrize("degree", [3, 4])
def test_chebyshev_through_marching_kernel(dgx, gpu_ctx, degree, np_dtype):
    """the smoother's fused operator + update launches on the marching kernel (variant 2 = marching kernel or fail):
    even n stages the update vectors with bulk copies, odd n loads them in the Z warps"""
    dim, cells0, nref, dt = 3, [1, 1, 1], 3, 5e-3
    gm, om = make_meshes(dgx, dim, cells0, nref, [True] * dim, hi=[1.0, 1.0, 1.0])
    dd = dgx.F64 if np_dtype == np.float64 else dgx.F32
    tol = 1e-11 if np_dtype == np.float64 else 2e-5
    mf = dgx.MatrixFree(gpu_ctx, gm, -1, dd)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, dt)
    op.compute_diagonal()
    op.set_variant(2)
    oop = PressureOperator(om, degree, dt, np_dtype)
    oop.compute_diagonal()
    ch = dgx.PreconditionChebyshev(op, 3, 15.0, 10)
    och = Chebyshev(oop.vmult, oop.inverse_diagonal, 3, 15.0, 10)
    ch.estimate_eigenvalues()
    och.estimate_eigenvalues(oop.n_dofs, np_dtype)
    b = smooth_plus_noise(om, oop.sh.nodes, noise=0.2).astype(np_dtype)
    x0 = np.random.default_rng(2).standard_normal(b.size).astype(np_dtype)
    vb, vx = op.initialize_dof_vector(), op.initialize_dof_vector()
    vb.from_host(b)
    l0 = dgx.kernel_launch_count()
    ch.vmult(vx, vb)
    assert rel_l2(vx.to_host(), och.vmult(b)) < tol
    vx.from_host(x0)
    ch.step(vx, vb)
    assert rel_l2(vx.to_host(), och.step(x0, b)) < tol
    assert dgx.kernel_launch_count() - l0 <= 8      # 2 + 3 fused launches, first-iteration updates; no separate update passes


def test_cg_on_mass_matrix(dgx, gpu_ctx):
    """project_grad's solver: unpreconditioned CG on the DG mass matrix, tol 1e-12 ||rhs|| (NS.cc:2574-2576)."""
    gm, om = make_meshes(dgx, 2, [3, 2], 2, [False, False])
    mf = dgx.MatrixFree(gpu_ctx, gm)
    degree = 2
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_MASS, degree)
    oop = MassOperator(om, degree)
    rhs = smooth_plus_noise(om, oop.sh.nodes, ncomp=2, noise=0.3)
    tol = 1e-12 * np.linalg.norm(rhs)
    vb, vx = op.initialize_dof_vector(), op.initialize_dof_vector()
    vb.from_host(rhs)
    its, res = dgx.solve_cg(op, vx, vb, None, 10000, tol)
    x = np.zeros_like(rhs)
    oits = solver_cg(oop.vmult, x, rhs, None, SolverControl(10000, tol))
    assert abs(its - oits) <= 1
    assert rel_l2(vx.to_host(), x) < 1e-10
    with pytest.raises(dgx.NoConvergence):
        vx.set(0.0)
        dgx.solve_cg(op, vx, vb, None, 2, tol)


def test_gmres_jacobi(dgx, gpu_ctx):
    """diffusion_step's solver configuration (Jacobi-preconditioned GMRES, abs tol eps ||rhs||, NS.cc:2450-2463)
    exercised on the pressure operator until the velocity operator kernels land."""
    gm, om = make_meshes(dgx, 2, [2, 2], 2, [False, True])
    mf = dgx.MatrixFree(gpu_ctx, gm)
    degree, dt = 3, 5e-3
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, dt)
    oop = PressureOperator(om, degree, dt)
    op.compute_diagonal()
    dinv = oop.compute_diagonal()
    rhs = smooth_plus_noise(om, oop.sh.nodes, noise=0.3)
    tol = 1e-8 * np.linalg.norm(rhs)
    vb, vx = op.initialize_dof_vector(), op.initialize_dof_vector()
    vb.from_host(rhs)
    its, res = dgx.solve_gmres(op, vx, vb, "jacobi", 10000, tol)
    x = np.zeros_like(rhs)
    oits = solver_gmres(oop.vmult, x, rhs, lambda r: dinv * r, SolverControl(10000, tol))
    assert abs(its - oits) <= 1, (its, oits)
    assert rel_l2(vx.to_host(), x) < 1e-6
    assert np.linalg.norm(oop.vmult(vx.to_host()) - rhs) < 1e-5 * np.linalg.norm(rhs)


@pytest.mark.parametrize("dim,cells0,nref,degree", [(2, [1, 1], 3, 2), (3, [1, 1, 1], 2, 2), (3, [1, 1, 1], 3, 1)])
def test_mg_cg_pressure_solve(dgx, gpu_ctx, dim, cells0, nref, degree):
    """projection_step's linear solve (NS.cc:2486-2546): fp64 CG preconditioned by the fp32 V-cycle."""
    gm, om = make_meshes(dgx, dim, cells0, nref, [True] * dim, hi=[1.0] * dim)
    dt, eps = 5e-3, 1e-8
    for stage in (1, 2):
        mf = dgx.MatrixFree(gpu_ctx, gm)
        op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, dt)
        op.set_TR_BDF2_stage(stage)
        mg = dgx.Multigrid(gpu_ctx, gm, degree, 100.0, dt)
        mg.set_TR_BDF2_stage(stage)
        osolver = PressureMGSolver(om, degree, dt, stage)
        rhs = smooth_plus_noise(om, osolver.op.sh.nodes, noise=0.1)
        rhs -= rhs.mean()
        tol = eps * np.linalg.norm(rhs)
        mg.setup(tol, 10000)
        vb, vx = op.initialize_dof_vector(), op.initialize_dof_vector()
        vb.from_host(rhs)
        # one V-cycle agrees with the oracle's
        osolver.setup()
        its, res = dgx.solve_cg(op, vx, vb, mg, 10000, tol)
        x = np.zeros_like(rhs)
        oits = osolver.solve(x, rhs, eps)
        assert abs(its - oits) <= 1, (its, oits)
        assert rel_l2(vx.to_host(), x) < 1e-4
        assert np.linalg.norm(osolver.op.vmult(vx.to_host()) - rhs) <= 1.5 * tol
        assert len(mg.coarse_iterations()) == its
        # a second setup with the same (stage, dt) reuses the eigenvalue estimates: same iteration count, same solution
        assert not mg.last_setup_was_cached()
        mg.setup(tol, 10000)
        assert mg.last_setup_was_cached()
        x1 = vx.to_host()
        vx.set(0.0)
        its2, _ = dgx.solve_cg(op, vx, vb, mg, 10000, tol)
        assert its2 == its and np.array_equal(vx.to_host(), x1)


@pytest.mark.parametrize("dim", [2, 3])
def test_solution_transfer_after_global_refinement(dgx, gpu_ctx, dim):
    """SolutionTransfer::interpolate for the all-cells-refined case of refine_mesh (NS.cc:2781-2822): the embedding of a DG field into
    the children is exact, so a polynomial velocity field (degree <= p per variable) interpolated on level L and transferred equals its
    interpolation on level L + 1; vector-valued spaces are component-blocked per cell."""
    from oracle.dg1d import gauss_lobatto_nodes
    from oracle.mesh import CartesianMesh
    args = (dim, [2, 1, 1][:dim], 2, [0.0] * dim, [1.0, 1.5, 0.75][:dim], [False] * dim)
    gm = dgx.Mesh(*args)
    fine, coarse = dgx.MatrixFree(gpu_ctx, gm, 2), dgx.MatrixFree(gpu_ctx, gm, 1)
    for degree, ncomp in ((2, dim), (1, 1), (3, dim)):
        nodes = np.asarray(gauss_lobatto_nodes(degree + 1), dtype=np.float64)

        def field(level):
            om = CartesianMesh(args[0], args[1], level, args[3], args[4], args[5])
            pts = om.dof_points(nodes)
            x = [pts[..., d] for d in range(dim)]
            comps = [(1 + c) * x[0] ** degree - x[1] ** (degree - 1) * (x[dim - 1] + 0.5 * c) + 0.25 * c for c in range(ncomp)]   # degree <= p per variable
            return np.stack(comps, axis=1).reshape(-1)
        vc, vf = coarse.initialize_dof_vector(degree, ncomp), fine.initialize_dof_vector(degree, ncomp)
        vc.from_host(field(1))
        vf.set(7.0)                                   # overwritten, not accumulated into
        dgx.solution_transfer_refine(vf, vc)
        ref = field(2)
        assert np.max(np.abs(vf.to_host() - ref)) < 1e-13 * max(1.0, np.max(np.abs(ref))), (degree, ncomp)
        # restriction is the transpose, also for vector-valued spaces: <P c, f> = <c, P^T f>
        rng = np.random.default_rng(degree)
        c_h, f_h = rng.standard_normal(vc.locally_owned_size()), rng.standard_normal(vf.locally_owned_size())
        vc.from_host(c_h); vf.set(0.0)
        dgx.transfer_prolongate_and_add(vf, vc)
        lhs = float(vf.to_host() @ f_h)
        vf.from_host(f_h); vc.set(0.0)
        dgx.transfer_restrict_and_add(vc, vf)
        assert abs(lhs - float(vc.to_host() @ c_h)) < 1e-11 * abs(lhs)
