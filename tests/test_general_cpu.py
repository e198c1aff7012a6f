"""CPU pins of the oracle's MappingQ1 pressure operator (oracle/general.py, following NS.cc:1230-1326 with deal.II's general-mapping
semantics): it must reduce to the Cartesian restatement (itself pinned by the exact-rational element matrices), be invariant under
rigid motions, stay symmetric on a deformed mesh and reproduce linear fields."""
import os

import numpy as np
import pytest

from oracle.dg1d import gauss_lobatto_nodes
from oracle.general import GeneralMesh, PressureOperatorGeneral
from oracle.mesh import CartesianMesh
from oracle.operators import PressureOperator


def wavy(amp):
    def fn(X):
        x = np.array(X, dtype=np.float64)
        out = x.copy()
        out[..., 0] = x[..., 0] + amp * np.sin(2 * np.pi * x[..., 1]) * (1 + 0.5 * np.cos(2 * np.pi * x[..., 0]))
        out[..., 1] = x[..., 1] + amp * np.sin(2 * np.pi * x[..., 0])
        if x.shape[-1] == 3:
            out[..., 1] += 0.5 * amp * np.sin(2 * np.pi * x[..., 2])
            out[..., 2] = x[..., 2] + amp * np.sin(2 * np.pi * (x[..., 0] + x[..., 1]))
        return out
    return fn


def rng_vec(n, seed=0):
    return np.random.default_rng(seed).standard_normal(n)


@pytest.mark.parametrize("dim,periodic,bids", [(2, [False, False], [1, 0, 1, 3]), (2, [True, True], None), (3, [False, True, False], [1, 1, 2, 3, 1, 5])])
def test_identity_map_reduces_to_the_cartesian_operator(dim, periodic, bids):
    cm = CartesianMesh(dim, [2, 1, 1][:dim], 1, [0.0] * dim, [1.0, 1.5, 0.75][:dim], periodic, bids)
    for degree in (1, 2):
        ref = PressureOperator(cm, degree, 5e-3)
        gen = PressureOperatorGeneral(GeneralMesh(cm, lambda X: X), degree, 5e-3)
        u = rng_vec(ref.n_dofs, degree)
        a, b = gen.vmult(u), ref.vmult(u)
        assert np.linalg.norm(a - b) < 1e-12 * np.linalg.norm(b)
    dref = ref.compute_diagonal()
    assert np.linalg.norm(gen.compute_diagonal() - dref) < 1e-12 * np.linalg.norm(dref)


def test_rigid_motion_and_scaling():
    cm = CartesianMesh(2, [2, 2], 1, [0.0, 0.0], [1.0, 1.0], [False, False], [1, 0, 1, 1])
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    base = PressureOperatorGeneral(GeneralMesh(cm, lambda X: X), 2, 5e-3)
    rot = PressureOperatorGeneral(GeneralMesh(cm, lambda X: X @ R.T + np.array([3.0, -1.0])), 2, 5e-3)
    u = rng_vec(base.n_dofs, 3)
    assert np.linalg.norm(rot.vmult(u) - base.vmult(u)) < 1e-12 * np.linalg.norm(base.vmult(u))
    cm2 = CartesianMesh(2, [2, 2], 1, [0.0, 0.0], [2.0, 2.0], [False, False], [1, 0, 1, 1])
    scaled = PressureOperatorGeneral(GeneralMesh(cm, lambda X: 2.0 * X), 2, 5e-3)
    ref2 = PressureOperator(cm2, 2, 5e-3)
    assert np.linalg.norm(scaled.vmult(u) - ref2.vmult(u)) < 1e-12 * np.linalg.norm(ref2.vmult(u))


@pytest.mark.parametrize("dim", [2, 3])
def test_symmetric_and_reproduces_linear_fields_on_a_deformed_mesh(dim):
    periodic = [False] * dim
    cm = CartesianMesh(dim, [1] * dim, 2 if dim == 2 else 1, [0.0] * dim, [1.0] * dim, periodic, [1] * (2 * dim))
    gm = GeneralMesh(cm, wavy(0.04))
    degree = 2 if dim == 2 else 1
    op = PressureOperatorGeneral(gm, degree, 5e-3)
    N = op.n_dofs
    A = np.stack([op.vmult(e) for e in np.eye(N)], axis=1)
    assert np.max(np.abs(A - A.T)) < 1e-11 * np.max(np.abs(A))          # SIPG with theta_p = 1
    assert np.min(np.linalg.eigvalsh(0.5 * (A + A.T))) > 0
    if dim == 2:
        # u = a . x is in the mapped Q_p space; on cells without boundary faces the Laplace part vanishes: (A u)_i = 1/coeff int phi_i u
        nodes = np.asarray(gauss_lobatto_nodes(degree + 1), dtype=np.float64)
        X = op.X                                                         # [c, 4, 2]
        a, b = np.meshgrid(nodes, nodes, indexing="xy")                  # a: x (fastest), b: y
        N0, N1, N2, N3 = (1 - a) * (1 - b), a * (1 - b), (1 - a) * b, a * b
        pts = (N0[None, ..., None] * X[:, None, None, 0] + N1[None, ..., None] * X[:, None, None, 1] +
               N2[None, ..., None] * X[:, None, None, 2] + N3[None, ..., None] * X[:, None, None, 3])          # [c, j, i, 2]
        u = (0.3 + 1.7 * pts[..., 0] - 0.9 * pts[..., 1]).reshape(-1)
        r = op.vmult(u).reshape(cm.n_cells, -1)
        mass = (np.einsum("cq,qi,cq->ci", u.reshape(cm.n_cells, -1) @ op.B.T, op.B, op.JxW)) / op.coeff()
        interior = np.all(cm.nbr >= 0, axis=1)
        assert interior.sum() == 4
        assert np.max(np.abs(r[interior] - mass[interior])) < 1e-11 * np.max(np.abs(r))


@pytest.mark.parametrize("dim", [2, 3])
def test_metric_tables_of_the_library_match_the_oracle_geometry(dgx, dim):
    """MatrixFree::reinit on a MappingQ1 level is host logic (csrc/pressure_general.cu::build_metric): its tables -- J^-1 J^-T JxW per cell
    point, J^-1 n JxW_f of both sides and the penalty per face point -- against the oracle's independently computed Jacobians."""
    periodic, bids = ([False, True], [1, 0, 2, 3]) if dim == 2 else ([True, False, False], [0, 1, 1, 3, 4, 1])
    cells0, nref, lo, hi = [2, 1, 1][:dim], 1, [0.0] * dim, [1.0] * dim
    cm = CartesianMesh(dim, cells0, nref, lo, hi, periodic, bids)
    om = GeneralMesh(cm, wavy(0.03))
    gm = dgx.Mesh(dim, cells0, nref, lo, hi, periodic, bids)
    gm.transform(wavy(0.03), lo, hi, cells0)
    ctx = dgx.Context(0)                                  # host-only context when there is no device
    mf = dgx.MatrixFree(ctx, gm)
    degree = 2
    cell, face = mf.general_metric_tables(degree)
    op = PressureOperatorGeneral(om, degree, 5e-3)
    G = np.einsum("cqek,cqfk,cq->cqef", op.invJ, op.invJ, op.JxW)
    pairs = [(0, 0), (1, 1), (0, 1)] if dim == 2 else [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]
    for k, (e, f) in enumerate(pairs):
        assert np.allclose(cell[:, k], G[:, :, e, f], rtol=1e-12, atol=1e-14)
    assert np.allclose(cell[:, len(pairs)], op.JxW, rtol=1e-13)
    for fno in range(2 * dim):
        d, s = fno // 2, fno % 2
        fe = op.face[(d, s)]
        nb = cm.nbr[:, fno]
        used = (nb >= 0) | (nb == -2)
        own = np.einsum("cqek,cqk,cq->cqe", fe["invJ"], fe["normal"], fe["JxW"])
        assert np.allclose(face[used, fno, 0:dim], np.moveaxis(own, 2, 1)[used], rtol=1e-12, atol=1e-14)
        assert np.all(face[~used, fno] == 0)
        fo = op.face[(d, 1 - s)]
        inner = nb >= 0
        nbr_side = np.einsum("cqek,cqk,cq->cqe", fo["invJ"][nb[inner]], fe["normal"][inner], fe["JxW"][inner])
        assert np.allclose(face[inner, fno, dim:2 * dim], np.moveaxis(nbr_side, 2, 1), rtol=1e-12, atol=1e-14)
        sigma = np.where(inner, 0.5 * (fe["norm"][:, 0] + fo["norm"][np.maximum(nb, 0), 0]), fe["norm"][:, 0]) * op.C_p
        assert np.allclose(face[used, fno, 2 * dim], (sigma[:, None] * fe["JxW"])[used], rtol=1e-12)
    ctx.close()


SHEARED = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "exact_sheared.npz"))


def sheared_meshes():
    hx, hy, s = float(SHEARED["hx"]), float(SHEARED["hy"]), float(SHEARED["shear"])
    args = (2, [2, 1], 0, [0.0, 0.0], [2 * hx, hy], [False, False])          # colorized ids: 1 = the upper x face
    shear = lambda X: np.stack([X[..., 0] + s * X[..., 1], X[..., 1]], axis=-1)
    return args, shear


@pytest.mark.parametrize("p", [1, 2])
def test_oracle_equals_exact_forms_on_sheared_cells(p):
    """tests/golden/exact_sheared.npz: the 2-cell operator on parallelograms as exact rational integrals of the weak forms
    (tests/golden/make_exact_sheared.py: no oracle code, no quadrature, no metric tables)"""
    args, shear = sheared_meshes()
    for stage in (1, 2):
        op = PressureOperatorGeneral(GeneralMesh(CartesianMesh(*args), shear), p, float(SHEARED["dt"]))
        op.set_TR_BDF2_stage(stage)
        ref = SHEARED[f"pressure_q{p}_stage{stage}"]
        A = np.stack([op.vmult(e) for e in np.eye(ref.shape[0])], axis=1)
        assert np.abs(A - ref).max() <= 1e-12 * np.abs(ref).max(), (p, stage)
    K = SHEARED[f"pressure_q{p}_K"]
    assert np.abs(K - K.T).max() < 1e-13 * np.abs(K).max()


def test_general_mass_operator_pins():
    """identity map == the Cartesian mass operator; 1^T M 1 = area of the deformed mesh (det J of a bilinear cell is linear: exact)"""
    from oracle.general import MassOperatorGeneral
    from oracle.operators import MassOperator
    cm = CartesianMesh(2, [2, 1], 1, [0.0, 0.0], [1.0, 1.5], [False, False])
    ref, gen = MassOperator(cm, 2), MassOperatorGeneral(GeneralMesh(cm, lambda X: X), 2)
    u = rng_vec(ref.n_dofs, 1)
    assert np.linalg.norm(gen.vmult(u) - ref.vmult(u)) < 1e-13 * np.linalg.norm(ref.vmult(u))
    gm = GeneralMesh(cm, wavy(0.05))
    op = MassOperatorGeneral(gm, 2, ncomp=1)
    X = gm.cell_vertices()[:, [0, 1, 3, 2]]                               # counter-clockwise
    area = 0.5 * np.abs(np.sum(X[:, :, 0] * np.roll(X[:, :, 1], -1, axis=1) - np.roll(X[:, :, 0], -1, axis=1) * X[:, :, 1], axis=1)).sum()
    one = np.ones(op.n_dofs)
    assert abs(one @ op.vmult(one) - area) < 1e-13 * area
    assert np.linalg.norm(op.inverse(op.vmult(u[:op.n_dofs])) - u[:op.n_dofs]) < 1e-12 * np.linalg.norm(u[:op.n_dofs])
