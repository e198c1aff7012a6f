"""Pins for the oracle itself (parity is unpinned by the reference): analytic
known-answer tests of the 1-D tables, the weak forms and the solver stack, plus the C port."""
import numpy as np
import pytest

from oracle import cpu_port
from oracle.dg1d import gauss_legendre, gauss_lobatto_nodes, shape
from oracle.mesh import CartesianMesh
from oracle.operators import GAMMA, MassOperator, PressureOperator, VelocityOperator
from oracle.projection import PressureMGSolver
from oracle.solvers import SolverControl, Transfer, solver_cg, solver_gmres


def test_quadrature_and_nodes_known_values():
    x, w = gauss_legendre(2)
    assert np.allclose(np.asarray(x, float), [0.5 - 0.5 / np.sqrt(3), 0.5 + 0.5 / np.sqrt(3)], atol=1e-16)
    assert np.allclose(np.asarray(w, float), [0.5, 0.5], atol=1e-16)
    x, w = gauss_legendre(3)
    assert np.allclose(np.asarray(x, float), [0.5 - 0.5 * np.sqrt(0.6), 0.5, 0.5 + 0.5 * np.sqrt(0.6)], atol=1e-16)
    assert np.allclose(np.asarray(w, float), [5 / 18, 8 / 18, 5 / 18], atol=1e-16)
    assert np.allclose(np.asarray(gauss_lobatto_nodes(3), float), [0, 0.5, 1])
    assert np.allclose(np.asarray(gauss_lobatto_nodes(4), float), [0, 0.5 - 0.5 / np.sqrt(5), 0.5 + 0.5 / np.sqrt(5), 1], atol=1e-16)
    for nq in range(1, 10):     # exactness: degree 2nq-1
        x, w = gauss_legendre(nq)
        for k in range(2 * nq):
            assert abs(float(np.sum(w * x ** k)) - 1.0 / (k + 1)) < 1e-15


@pytest.mark.parametrize("p", range(1, 9))
def test_shape_tables(p):
    sh = shape(p, p + 1)
    assert np.allclose(sh.S.sum(1), 1, atol=1e-14) and np.allclose(sh.D.sum(1), 0, atol=1e-12)
    n = p + 1
    # even-odd symmetry
    assert np.allclose(sh.S[::-1, ::-1], sh.S, atol=1e-15) and np.allclose(sh.D[::-1, ::-1], -sh.D, atol=1e-13)
    # nodal at the end points
    assert np.allclose(sh.f[0], np.eye(n)[0]) and np.allclose(sh.f[1], np.eye(n)[-1])
    # derivative of x^k reproduced
    for k in range(p + 1):
        assert np.allclose(sh.D @ sh.nodes ** k, k * sh.xq ** max(k - 1, 0) if k else 0, atol=1e-12)
    # embedding reproduces polynomials on the children
    for c in (0, 1):
        for k in range(p + 1):
            assert np.allclose(sh.P[c] @ sh.nodes ** k, ((sh.nodes + c) / 2) ** k, atol=1e-14)


@pytest.mark.parametrize("dim", [2, 3])
def test_pressure_form_identities(dim):
    m = CartesianMesh(dim, [2, 1, 1][:dim], 1, [0] * dim, [1.0, 0.5, 0.75][:dim], [True] * dim)
    for p in (1, 2, 3):
        op = PressureOperator(m, p, 5e-3)
        rng = np.random.default_rng(p)
        u, v = rng.standard_normal(op.n_dofs), rng.standard_normal(op.n_dofs)
        Au, Av = op.vmult(u), op.vmult(v)
        assert abs(v @ Au - u @ Av) < 1e-12 * abs(v @ Au)                    # symmetry
        one = np.ones(op.n_dofs)
        M1 = MassOperator(m, p, ncomp=1).vmult(one)
        assert np.allclose(op.vmult(one), M1 / op.coeff(), atol=1e-13 * np.abs(M1).max() / op.coeff() * 1e3)
        assert u @ Au > 0
        # consistency: a(u, v) for a smooth periodic u equals (-lap u + kappa u, v) up to discretisation error
    # polynomial exactness of the cell term: u = x^2 in one cell interior => -lap u = -2 tested against constants
    m1 = CartesianMesh(dim, [1] * dim, 0, [0] * dim, [1.0] * dim, [False] * dim, boundary_ids=[0, 0, 2, 3, 4, 5])
    op = PressureOperator(m1, 2, 5e-3)
    x = m1.dof_points(np.asarray(op.sh.nodes))[..., 0].reshape(-1)
    Au = op.vmult(x ** 2)
    # sum_i (A u)_i = a(u, 1) = kappa * int u (no id-1 faces, Neumann elsewhere)
    assert abs(Au.sum() - (1.0 / op.coeff()) / 3.0) < 1e-12


def test_stage_coefficients():
    m = CartesianMesh(2, [1, 1], 1, [0, 0], [1, 1], [True, True])
    op = PressureOperator(m, 1, 5e-3)
    assert op.coeff() == 1.0e6 * GAMMA * 5e-3 * GAMMA * 5e-3
    op.set_TR_BDF2_stage(2)
    assert op.coeff() == 1.0e6 * (1.0 - GAMMA) * 5e-3 * (1.0 - GAMMA) * 5e-3


def test_reference_diagonal_quirks():
    """the reference's probing differs from the true diagonal on interior faces (both sides get the unit
    dof, NS.cc:1921-1922) and uses 2*sigma on id-1 boundaries (NS.cc:1997)."""
    m = CartesianMesh(2, [3, 1], 0, [0, 0], [3, 1], [False, False])
    op = PressureOperator(m, 1, 5e-3)
    dinv = op.compute_diagonal()
    A = np.stack([op.vmult(e) for e in np.eye(op.n_dofs)], 1)
    true = np.diag(A)
    got = 1.0 / dinv
    assert not np.allclose(got, true)
    # cell 1 has no boundary in x: only the interior-face quirk acts there; cell 2 touches id 1
    assert np.all(got > 0)


def test_velocity_reduces_to_helmholtz_without_convection():
    """with u_extr = 0 the velocity operator is alpha*M + (c/Re)*SIPG per component."""
    m = CartesianMesh(2, [2, 2], 0, [0, 0], [1, 1], [True, True])
    p = 2
    vop = VelocityOperator(m, p, Re=10.0, dt=1e-2)
    rng = np.random.default_rng(0)
    u = rng.standard_normal(vop.n_dofs)
    Au = vop.vmult(u).reshape(m.n_cells, 2, -1)
    pop = PressureOperator(m, p, 1e-2)
    mass = MassOperator(m, p, ncomp=1)
    alpha, c = 1.0 / (GAMMA * 1e-2), 0.5
    uc = u.reshape(m.n_cells, 2, -1)
    for comp in range(2):
        s = np.ascontiguousarray(uc[:, comp]).reshape(-1)
        lap = pop.vmult(s) - mass.vmult(s) / pop.coeff()
        ref = alpha * mass.vmult(s) + c / 10.0 * lap
        assert np.allclose(Au[:, comp].reshape(-1), ref, rtol=1e-11, atol=1e-11)


def test_velocity_convective_term_conserves_constants():
    """a_u(u, 1) for periodic data: the SIPG and Lax-Friedrichs jump terms vanish against constants and
    the convective volume term cancels => sum_i (A u)_i = alpha * int u per component."""
    m = CartesianMesh(2, [2, 2], 1, [0, 0], [1, 1], [True, True])
    vop = VelocityOperator(m, 2, Re=100.0, dt=5e-3)
    rng = np.random.default_rng(1)
    vop.set_u_extr(rng.standard_normal(vop.n_dofs))
    u = rng.standard_normal(vop.n_dofs)
    Au = vop.vmult(u).reshape(m.n_cells, 2, -1)
    Mu = MassOperator(m, 2).vmult(u).reshape(m.n_cells, 2, -1)
    alpha = 1.0 / (GAMMA * 5e-3)
    for comp in range(2):
        assert abs(Au[:, comp].sum() - alpha * Mu[:, comp].sum()) < 1e-9 * abs(alpha * Mu[:, comp].sum()) + 1e-9


def test_transfer_is_transpose_and_exact():
    m = CartesianMesh(3, [1, 1, 1], 2, [0, 0, 0], [1, 1, 1], [True] * 3)
    tr = Transfer(m, 2, np.float64)
    rng = np.random.default_rng(0)
    nc, nf = tr.coarse.n_cells * 27, m.n_cells * 27
    uc, uf = rng.standard_normal(nc), rng.standard_normal(nf)
    lhs = tr.prolongate(uc) @ uf
    rhs = uc @ tr.restrict_and_add(np.zeros(nc), uf)
    assert abs(lhs - rhs) < 1e-12 * abs(lhs)
    # a global quadratic is reproduced exactly
    ptsc = tr.coarse.dof_points(np.asarray(shape(2, 3).nodes)).reshape(-1, 3)
    ptsf = m.dof_points(np.asarray(shape(2, 3).nodes)).reshape(-1, 3)
    f = lambda x: x[:, 0] ** 2 + x[:, 1] * x[:, 2] - 0.3 * x[:, 2] ** 2
    assert np.allclose(tr.prolongate(f(ptsc)), f(ptsf), atol=1e-13)


def test_cg_and_gmres_on_spd_system():
    rng = np.random.default_rng(0)
    B = rng.standard_normal((40, 40))
    A = B @ B.T + 40 * np.eye(40)
    b = rng.standard_normal(40)
    x = np.zeros(40)
    its = solver_cg(lambda v: A @ v, x, b, None, SolverControl(100, 1e-10))
    assert np.linalg.norm(A @ x - b) <= 1e-10 and 0 < its <= 40
    x2 = np.zeros(40)
    d = 1.0 / np.diag(A)
    its2 = solver_gmres(lambda v: A @ v, x2, b, lambda r: d * r, SolverControl(100, 1e-10))
    assert np.linalg.norm(d * (A @ x2 - b)) <= 1e-9 and 0 < its2 <= 40


def test_mg_cg_converges_fast():
    m = CartesianMesh(2, [1, 1], 3, [0, 0], [1, 1], [True, True])
    s = PressureMGSolver(m, 2, 5e-3)
    pts = m.dof_points(np.asarray(s.op.sh.nodes))
    rhs = MassOperator(m, 2, ncomp=1).vmult(np.prod(np.cos(2 * np.pi * pts), axis=-1).reshape(-1))
    x = np.zeros_like(rhs)
    its = s.solve(x, rhs)
    assert its <= 10
    assert np.linalg.norm(s.op.vmult(x) - rhs) <= 1e-8 * np.linalg.norm(rhs)


@pytest.mark.parametrize("case", [(3, [1, 1, 1], 2, [1, 1, 1], 4), (2, [4, 2], 1, [0, 0], 2), (3, [2, 1, 1], 1, [0, 0, 0], 3),
                                  (3, [1, 1, 1], 0, [1, 1, 1], 2)])
def test_c_port_matches_numpy_oracle(case):
    dim, c0, L, per, deg = case
    m = CartesianMesh(dim, c0, L, [0] * dim, [1.0, 1.5, 0.75][:dim], per)
    op = PressureOperator(m, deg, 5e-3)
    u = np.random.default_rng(0).standard_normal(op.n_dofs)
    a, b = op.vmult(u), cpu_port.pressure_vmult(m, deg, 5e-3, u)
    assert np.linalg.norm(a - b) <= 1e-13 * np.linalg.norm(a)


def test_cpu_kronecker_port_matches_oracle():
    """oracle/cpu_kron.c (the kernels' Kronecker form on the CPU, used as an extra timed baseline) equals the quadrature-point oracle"""
    from oracle import cpu_port
    from oracle.mesh import CartesianMesh
    from oracle.operators import PressureOperator
    m = CartesianMesh(3, [1, 2, 1], 1, [0, 0, 0], [1.0, 1.5, 0.75], [True] * 3)
    for degree in (1, 2, 4):
        op = PressureOperator(m, degree, 5e-3)
        u = np.random.default_rng(degree).standard_normal(op.n_dofs)
        got = cpu_port.kron_vmult(m, degree, 5e-3, u)
        ref = op.vmult(u)
        assert np.linalg.norm(got - ref) <= 1e-12 * np.linalg.norm(ref)


def test_time_step_converges_to_the_analytic_taylor_green_vortex():
    """Physics pin of the whole restatement (both RHS forms, the three operators, GMRES / MG-CG / mass CG, the TR-BDF2
    projection sequence of NS.cc:2934-2986): on [0, 2 pi]^2 periodic the Navier-Stokes equations have the exact solution
    u = (sin x cos y, -cos x sin y) exp(-2 t / Re), p = (cos 2x + cos 2y)/4 exp(-4 t / Re).  With the reference's degrees
    (velocity Q2 / pressure Q1) the oracle's velocity error at t = 0.2 must fall under mesh refinement and be below 1 % on
    16^2 cells."""
    from oracle.timestep import OracleProjection
    Re, L, T = 10.0, 2 * np.pi, 0.2
    errs, perrs = [], []
    for nref, nsteps in ((2, 4), (3, 4), (4, 8)):
        om = CartesianMesh(2, [1, 1], nref, [0.0, 0.0], [L, L], [True, True])
        orc = OracleProjection(om, 1, Re, T / nsteps)
        pts = om.dof_points(np.asarray(shape(2, 3).nodes))
        x, y = pts[..., 0], pts[..., 1]
        exact = lambda t: np.stack([np.sin(x) * np.cos(y), -np.cos(x) * np.sin(y)], axis=1).reshape(-1) * np.exp(-2 * t / Re)
        pp = om.dof_points(np.asarray(shape(1, 2).nodes))
        pex = lambda t: (0.25 * (np.cos(2 * pp[..., 0]) + np.cos(2 * pp[..., 1]))).reshape(-1) * np.exp(-4 * t / Re)
        orc.u_n, orc.u_n_minus_1, orc.pres_n = exact(0.0).copy(), exact(0.0).copy(), pex(0.0).copy()
        for _ in range(nsteps):
            orc.advance()
        errs.append(np.linalg.norm(orc.u_n - exact(T)) / np.linalg.norm(exact(T)))
        pe = pex(T)
        perrs.append(np.linalg.norm((orc.pres_n - orc.pres_n.mean()) - (pe - pe.mean())) / np.linalg.norm(pe - pe.mean()))
    assert errs[1] < 0.6 * errs[0] and errs[2] < 0.25 * errs[1], errs
    assert errs[2] < 1e-2, errs
    assert perrs[2] < 0.1 and perrs[2] < perrs[1], perrs
