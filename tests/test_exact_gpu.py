"""The CUDA operators against the exact-rational element matrices of tests/golden/make_exact.py (derived from the weak forms
of NS.cc alone), and the order of convergence of the SIPG discretisation on a manufactured solution: p + 1 in L2.  Neither
test runs any oracle code."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "exact_matrices.npz"))


def matrix_of(op, n):
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    A = np.empty((n, n))
    for j in range(n):
        e = np.zeros(n)
        e[j] = 1.0
        src.from_host(e)
        op.vmult(dst, src)
        A[:, j] = dst.to_host()
    return A


@pytest.mark.parametrize("periodic", [True, False], ids=["periodic", "box"])
@pytest.mark.parametrize("p", [1, 2])
def test_pressure_cuda_equals_exact_forms(dgx, gpu_ctx, p, periodic):
    mesh = dgx.Mesh(2, [2, 1], 0, [0.0, 0.0], [2 * float(G["hx"]), float(G["hy"])], [periodic, periodic])
    mf = dgx.MatrixFree(gpu_ctx, mesh)
    for stage in (1, 2):
        op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, p, float(G["Re"]), float(G["dt"]))
        op.set_TR_BDF2_stage(stage)
        ref = G[f"pressure_q{p}_{'periodic' if periodic else 'box'}_stage{stage}"]
        A = matrix_of(op, ref.shape[0])
        assert np.abs(A - ref).max() <= 1e-12 * np.abs(ref).max(), (p, periodic, stage)


@pytest.mark.parametrize("p", [1, 2])
def test_velocity_cuda_equals_exact_forms(dgx, gpu_ctx, p):
    mesh = dgx.Mesh(2, [2, 1], 0, [0.0, 0.0], [2 * float(G["hx"]), float(G["hy"])], [True, True])
    mf = dgx.MatrixFree(gpu_ctx, mesh)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_VELOCITY, p, float(G["Re"]), float(G["dt"]))
    op.set_TR_BDF2_stage(1)
    ue = op.initialize_dof_vector()
    ue.from_host(np.tile(np.repeat(G["ue"], (p + 1) ** 2), 2))
    op.set_u_extr(ue)
    ref = G[f"velocity_q{p}_stage1"]
    A = matrix_of(op, ref.shape[0])
    assert np.abs(A - ref).max() <= 1e-12 * np.abs(ref).max()


def gauss_lobatto(n):
    from numpy.polynomial import legendre as leg
    if n == 2:
        return np.array([0.0, 1.0])
    c = np.zeros(n); c[-1] = 1.0
    return np.concatenate(([0.0], (np.sort(leg.legroots(leg.legder(c))) + 1) / 2, [1.0]))


@pytest.mark.parametrize("p", [1, 2, 3])
def test_sipg_order_of_convergence(dgx, gpu_ctx, p):
    """Manufactured solution on the periodic unit square: (-Laplace + kappa) u = f with u = sin(2 pi x) cos(2 pi y), solved with
    the device CG + multigrid on 8^2, 16^2, 32^2 cells.  The L2 error must fall with order p + 1 (SIPG with an adequate penalty);
    a wrong penalty scale, flux sign or face normal breaks the order or the convergence of CG."""
    dt = 5e-3
    kappa = 1.0 / (1.0e6 * (2 - np.sqrt(2)) * dt * (2 - np.sqrt(2)) * dt)
    errs = []
    for R in (3, 4, 5):
        mesh = dgx.Mesh(2, [1, 1], R, [0.0, 0.0], [1.0, 1.0], [True, True])
        mf = dgx.MatrixFree(gpu_ctx, mesh)
        op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, p, 100.0, dt)
        mass = dgx.NavierStokesProjectionOperator(mf, dgx.OP_MASS, p, 100.0, dt)   # NS_stage 3: M with `dim` components
        nc, h, n = 2 ** R, 1.0 / 2 ** R, p + 1
        x1 = gauss_lobatto(n)
        coords = mesh.cell_coords().astype(np.float64)
        Xp = (coords[:, None, None, 0] + x1[None, None, :]) * h + 0 * x1[None, :, None]
        Yp = (coords[:, None, None, 1] + x1[None, :, None]) * h + 0 * x1[None, None, :]
        uex = np.sin(2 * np.pi * Xp) * np.cos(2 * np.pi * Yp)
        f = (8 * np.pi ** 2 + kappa) * uex                      # interpolated right-hand side: its error is O(h^(p+1)) as well
        # rhs_i = int phi_i f = (M f)_i through the mass operator (2 components in 2-D: use the first)
        fv, mv = mass.initialize_dof_vector(), mass.initialize_dof_vector()
        f2 = np.stack([f, np.zeros_like(f)], axis=1).reshape(-1)
        fv.from_host(f2)
        mass.vmult(mv, fv)
        rhs = mv.to_host().reshape(nc * nc, 2, n * n)[:, 0, :].reshape(-1)
        b, x = op.initialize_dof_vector(), op.initialize_dof_vector()
        b.from_host(rhs)
        mg = dgx.Multigrid(gpu_ctx, mesh, p, 100.0, dt)
        tol = 1e-11 * b.l2_norm()
        mg.setup(tol, 10000)
        dgx.solve_cg(op, x, b, mg, 10000, tol)
        e = x.to_host().reshape(nc * nc, n, n) - uex
        # discrete L2 norm through the mass operator: e^T M e
        ev, me = mass.initialize_dof_vector(), mass.initialize_dof_vector()
        ev.from_host(np.stack([e, np.zeros_like(e)], axis=1).reshape(-1))
        mass.vmult(me, ev)
        errs.append(np.sqrt(ev.dot(me)))
    orders = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert orders[-1] > p + 1 - 0.35, (p, errs, orders)
    assert errs[-1] < errs[0] / 2 ** (2 * (p + 1) - 1.0), (p, errs)
