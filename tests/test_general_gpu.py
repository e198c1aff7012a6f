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
        with pytest.raises(dgx.DgxError):   # the velocity operator has no metric path yet: loud error, never a silent Cartesian answer
            vop = dgx.NavierStokesProjectionOperator(mf, dgx.OP_VELOCITY, 2, 100.0, dt)
            a, b = vop.initialize_dof_vector(), vop.initialize_dof_vector()
            vop.vmult(a, b)


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


def test_o_grid_around_a_cylinder(dgx, gpu_ctx, tmp_path):
    """A structured O-grid around a cylinder (periodic in the angular direction, geometry periodic up to a ROTATION): the pressure
    operator with the cylinder as a no-flux boundary (id 4) and a Dirichlet outer boundary (id 1), the lift / drag integrals over
    the cylinder (NS.cc:2642-2708, boundary id 4 as in the reference) and the output data on the curved mesh."""
    from oracle.diagnostics import lift_and_drag_general, patch_data_general
    from test_diagnostics_cpu import annulus, read_vtu
    dim, cells0, nref, periodic, bids = 2, [1, 1], 4, [False, True], [4, 1, 2, 3]
    gm, om = meshes(dgx, dim, cells0, nref, periodic, bids, annulus(0.5, 2.0))
    mf = dgx.MatrixFree(gpu_ctx, gm)
    for degree in (1, 2, 3):
        op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, 5e-3)
        oop = PressureOperatorGeneral(om, degree, 5e-3)
        u = np.random.default_rng(degree).standard_normal(oop.n_dofs)
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        src.from_host(u)
        op.vmult(dst, src)
        assert rel_l2(dst.to_host(), oop.vmult(u)) < 1e-12
    rng = np.random.default_rng(11)
    dp, Re = 1, 20.0
    nv, npn = om.n_cells * 2 * (dp + 2) ** 2, om.n_cells * (dp + 1) ** 2
    u, uo, p = rng.standard_normal(nv), rng.standard_normal(nv), rng.standard_normal(npn)
    vu, vuo, vp = mf.initialize_dof_vector(dp + 1, 2), mf.initialize_dof_vector(dp + 1, 2), mf.initialize_dof_vector(dp, 1)
    vu.from_host(u); vuo.from_host(uo); vp.from_host(p)
    lift, drag = dgx.compute_lift_and_drag(vu, vp, Re)
    rl, rd = lift_and_drag_general(om, dp + 1, dp, u, p, Re)
    assert abs(lift - rl) < 1e-11 * max(1.0, abs(rl)) and abs(drag - rd) < 1e-11 * max(1.0, abs(rd))
    out = dgx.output_patch_data(vu, vp, vuo)
    ref = patch_data_general(om, dp + 1, dp, u, p, uo)
    assert np.max(np.abs(out - ref)) < 1e-10 * np.max(np.abs(ref))
    # divergence theorem: u = 0, p = x  =>  |drag| = area of the 16-gon that is the cylinder on this mesh, lift = 0
    X = om.cell_vertices()
    vp.from_host(X[:, :, 0].reshape(-1)); vu.set(0.0)
    lift, drag = dgx.compute_lift_and_drag(vu, vp, Re)
    assert abs(drag + 0.5 * 16 * 0.25 * np.sin(2 * np.pi / 16)) < 1e-12 and abs(lift) < 1e-12
    path = str(tmp_path / "ring.vtu")
    dgx.output_results(vu, vp, vuo, path)
    npnt, ncell, arr = read_vtu(path)
    assert ncell == om.n_cells and np.allclose(arr["points"][:, :2], X.reshape(-1, 2), atol=0) and np.allclose(arr["p"], X[:, :, 0].reshape(-1), atol=0)


@pytest.mark.parametrize("dim", [2, 3])
def test_diagnostics_on_a_deformed_mesh(dgx, gpu_ctx, dim):
    from oracle.diagnostics import lift_and_drag_general, patch_data_general
    bids = [4, 1, 4, 3] if dim == 2 else [0, 4, 2, 3, 4, 5]
    gm, om = meshes(dgx, dim, [2, 1, 1][:dim], 1, [False] * dim, bids, wavy(0.04))
    rng = np.random.default_rng(3)
    for dtype in (dgx.F64, dgx.F32):
        mf = dgx.MatrixFree(gpu_ctx, gm, dtype=dtype)
        for dp in (1, 2):
            nv, npn = om.n_cells * dim * (dp + 2) ** dim, om.n_cells * (dp + 1) ** dim
            u, uo, p = rng.standard_normal(nv), rng.standard_normal(nv), rng.standard_normal(npn)
            if dtype == dgx.F32:
                u, uo, p = (a.astype(np.float32).astype(np.float64) for a in (u, uo, p))
            vu, vuo, vp = mf.initialize_dof_vector(dp + 1, dim), mf.initialize_dof_vector(dp + 1, dim), mf.initialize_dof_vector(dp, 1)
            vu.from_host(u); vuo.from_host(uo); vp.from_host(p)
            lift, drag = dgx.compute_lift_and_drag(vu, vp, 9.0)
            rl, rd = lift_and_drag_general(om, dp + 1, dp, u, p, 9.0)
            assert abs(lift - rl) < 1e-11 * max(1.0, abs(rl)) and abs(drag - rd) < 1e-11 * max(1.0, abs(rd)), (dtype, dp)
            ref = patch_data_general(om, dp + 1, dp, u, p, uo)
            assert np.max(np.abs(dgx.output_patch_data(vu, vp, vuo) - ref)) < 1e-10 * np.max(np.abs(ref))


@pytest.mark.parametrize("p", [1, 2])
def test_metric_kernel_equals_exact_forms_on_sheared_cells(dgx, gpu_ctx, p):
    """the CUDA metric kernel against tests/golden/exact_sheared.npz (exact rational integrals of the weak forms on 2 parallelograms,
    derived without oracle code): no oracle in this test"""
    from test_general_cpu import SHEARED, sheared_meshes
    args, shear = sheared_meshes()
    gm = dgx.Mesh(*args)
    gm.transform(shear, args[3], args[4], args[1])
    mf = dgx.MatrixFree(gpu_ctx, gm)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, p, 100.0, float(SHEARED["dt"]))
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    for stage in (1, 2):
        op.set_TR_BDF2_stage(stage)
        ref = SHEARED[f"pressure_q{p}_stage{stage}"]
        A = np.empty_like(ref)
        for j in range(ref.shape[0]):
            e = np.zeros(ref.shape[0]); e[j] = 1.0
            src.from_host(e)
            op.vmult(dst, src)
            A[:, j] = dst.to_host()
        assert np.abs(A - ref).max() <= 1e-12 * np.abs(ref).max(), (p, stage)


@pytest.mark.parametrize("dim", [2, 3])
def test_mass_operator_on_a_deformed_mesh(dgx, gpu_ctx, dim):
    """NS_stage 3 on MappingQ1 cells: vmult, the exact inverse, and project_grad's CG solve (unfused path) against the oracle"""
    from oracle.general import MassOperatorGeneral
    gm, om = meshes(dgx, dim, [2, 1, 1][:dim], 1, [False, True, False][:dim], None, wavy(0.04))
    for dtype, tol in ((dgx.F64, 1e-13), (dgx.F32, 1e-5)):
        mf = dgx.MatrixFree(gpu_ctx, gm, dtype=dtype)
        for dv in (2, 3):
            op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_MASS, dv, 100.0, 5e-3)
            oop = MassOperatorGeneral(om, dv)
            u = np.random.default_rng(dv).standard_normal(oop.n_dofs)
            if dtype == dgx.F32:
                u = u.astype(np.float32).astype(np.float64)
            src, dst, back = (op.initialize_dof_vector() for _ in range(3))
            src.from_host(u)
            op.vmult(dst, src)
            assert rel_l2(dst.to_host(), oop.vmult(u)) < tol
            op.mass_inverse(back, dst)
            assert rel_l2(back.to_host(), u) < tol * 100
            if dtype == dgx.F64:
                x = op.initialize_dof_vector()
                its, _ = dgx.solve_cg(op, x, dst, None, 1000, 1e-12 * np.linalg.norm(oop.vmult(u)))
                assert its > 1 and rel_l2(x.to_host(), u) < 1e-9
    with pytest.raises(dgx.DgxError):   # the right-hand-side forms are Cartesian-only so far: loud error, never a silent wrong answer
        pop = mf.initialize_dof_vector(1, 1)
        op.vmult_grad_p_projection(dst, pop)


def test_metric_tables_can_be_released_and_are_rebuilt(dgx, gpu_ctx):
    gm, om = meshes(dgx, 2, [2, 1], 2, [False, True], None, wavy(0.03))
    mf = dgx.MatrixFree(gpu_ctx, gm)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, 2, 100.0, 5e-3)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    src.from_host(np.random.default_rng(0).standard_normal(src.locally_owned_size()))
    op.vmult(dst, src)
    first = dst.to_host()
    assert mf.general_metric_bytes(2) > 0
    mf.release_general_metric()
    assert mf.general_metric_bytes(2) == 0
    op.vmult(dst, src)
    assert np.array_equal(dst.to_host(), first) and mf.general_metric_bytes(2) > 0
