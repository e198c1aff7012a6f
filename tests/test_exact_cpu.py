"""The oracle against element matrices derived independently of it: tests/golden/exact_matrices.npz holds the pressure
and velocity operators of a 2 x 1 mesh as exact rational integrals of the reference's weak forms (tests/golden/make_exact.py:
no oracle code, no quadrature, no tensor-product evaluation).  A wrong SIPG penalty scale, flux sign, normal convention or
dof ordering in the oracle fails here."""
import os

import numpy as np
import pytest

from oracle.mesh import CartesianMesh
from oracle.operators import PressureOperator, VelocityOperator

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "exact_matrices.npz"))


def mesh(periodic):
    return CartesianMesh(2, [2, 1], 0, [0.0, 0.0], [2 * float(G["hx"]), float(G["hy"])], [periodic, periodic])


def matrix_of(apply, n):
    A = np.empty((n, n))
    for j in range(n):
        e = np.zeros(n)
        e[j] = 1.0
        A[:, j] = apply(e)
    return A


@pytest.mark.parametrize("periodic", [True, False], ids=["periodic", "box"])
@pytest.mark.parametrize("p", [1, 2])
def test_pressure_oracle_equals_exact_forms(p, periodic):
    for stage in (1, 2):
        op = PressureOperator(mesh(periodic), p, float(G["dt"]), np.float64)
        op.set_TR_BDF2_stage(stage)
        ref = G[f"pressure_q{p}_{'periodic' if periodic else 'box'}_stage{stage}"]
        A = matrix_of(op.vmult, ref.shape[0])
        assert np.abs(A - ref).max() <= 1e-12 * np.abs(ref).max(), (p, periodic, stage)


@pytest.mark.parametrize("p", [1, 2])
def test_velocity_oracle_equals_exact_forms(p):
    m = mesh(True)
    op = VelocityOperator(m, p, float(G["Re"]), float(G["dt"]))
    op.set_TR_BDF2_stage(1)
    n1 = (p + 1) ** 2
    ue = np.tile(np.repeat(G["ue"], n1), m.n_cells)            # constant field, component-blocked per cell
    op.set_u_extr(ue)
    ref = G[f"velocity_q{p}_stage1"]
    A = matrix_of(op.vmult, ref.shape[0])
    assert np.abs(A - ref).max() <= 1e-12 * np.abs(ref).max()


def test_generator_reproduces_committed_file(tmp_path):
    """the committed npz is what make_exact.py writes (the generator is part of the fixture)"""
    import importlib.util
    import math
    spec = importlib.util.spec_from_file_location("make_exact", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_exact.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from fractions import Fraction as Fr
    K = mod.to_np(mod.pressure_matrix(Fr(1, 2), Fr(1, 3), 1, True, Fr(0)))
    assert np.array_equal(K, G["pressure_q1_periodic_K"])
    # sanity of the exact forms themselves: symmetric, constants in the kernel of the periodic Laplacian part
    assert np.abs(K - K.T).max() < 1e-15 and np.abs(K @ np.ones(8)).max() < 1e-14
    assert math.isclose(float(np.ones(8) @ G["pressure_q1_periodic_M"] @ np.ones(8)), 2 * 0.5 / 3, rel_tol=1e-13)   # area of the mesh
