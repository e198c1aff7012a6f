"""GPU: device vector operations used by the Krylov / Chebyshev / multigrid layers vs numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype_name", ["F64", "F32"])
def test_vector_ops(dgx, gpu_ctx, dtype_name):
    dt = getattr(dgx, dtype_name)
    npdt = np.float64 if dt == dgx.F64 else np.float32
    tol = 1e-13 if dt == dgx.F64 else 1e-5
    gm = dgx.Mesh(3, [1, 1, 1], 3, [0, 0, 0], [1, 1, 1], [True] * 3)
    mf = dgx.MatrixFree(gpu_ctx, gm, -1, dt)
    a, b, c = (mf.initialize_dof_vector(3) for _ in range(3))
    n = a.locally_owned_size()
    assert n == 512 * 64 and a.size() == n
    rng = np.random.default_rng(3)
    ha, hb = rng.standard_normal(n).astype(npdt), rng.standard_normal(n).astype(npdt)
    a.from_host(ha)
    b.from_host(hb)
    assert np.array_equal(a.to_host(), ha)
    assert abs(a.dot(b) - float(ha.astype(np.float64) @ hb.astype(np.float64))) < tol * n
    assert abs(a.l2_norm() - np.linalg.norm(ha.astype(np.float64))) < tol * np.sqrt(n) * 10
    assert abs(a.linfty_norm() - np.abs(ha).max()) == 0
    assert abs(a.mean_value() - ha.astype(np.float64).mean()) < tol
    c.equ(2.5, a)
    assert np.allclose(c.to_host(), npdt(2.5) * ha, rtol=tol, atol=0)
    c.add(-1.5, b)
    hc = npdt(2.5) * ha + npdt(-1.5) * hb
    assert np.allclose(c.to_host(), hc, rtol=10 * tol, atol=10 * tol)
    c.sadd(0.5, 2.0, a)
    hc = npdt(0.5) * hc + npdt(2.0) * ha
    assert np.allclose(c.to_host(), hc, rtol=10 * tol, atol=10 * tol)
    c.scale(3.0)
    hc = hc * npdt(3.0)
    c.scale(b)
    hc = hc * hb
    assert np.allclose(c.to_host(), hc, rtol=10 * tol, atol=10 * tol)
    r = c.add_and_dot(0.25, a, b)
    hc = hc + npdt(0.25) * ha
    assert abs(r - float(hc.astype(np.float64) @ hb.astype(np.float64))) < 10 * tol * n
    r = c.add_and_dot(-0.5, b, c)       # aliasing form used by SolverCG: r.add_and_dot(-alpha, v, r)
    hc = hc + npdt(-0.5) * hb
    assert abs(r - float(hc.astype(np.float64) @ hc.astype(np.float64))) < 10 * tol * n
    c.set(1.25)
    assert np.all(c.to_host() == npdt(1.25))


@pytest.mark.parametrize("dtype_name", ["F64", "F32"])
def test_cg_update_and_gram_schmidt(dgx, gpu_ctx, dtype_name):
    """fused SolverCG update and the SolverGMRES Gram-Schmidt sweep (incl. more than 8 basis vectors, 2 kernel batches)"""
    dt = getattr(dgx, dtype_name)
    npdt = np.float64 if dt == dgx.F64 else np.float32
    tol = 1e-13 if dt == dgx.F64 else 2e-5
    gm = dgx.Mesh(3, [1, 1, 1], 3, [0, 0, 0], [1, 1, 1], [True] * 3)
    mf = dgx.MatrixFree(gpu_ctx, gm, -1, dt)
    rng = np.random.default_rng(11)
    x, p, r, v = (mf.initialize_dof_vector(2) for _ in range(4))
    n = x.locally_owned_size()
    hx, hp, hr, hv = (rng.standard_normal(n).astype(npdt) for _ in range(4))
    for d, h in ((x, hx), (p, hp), (r, hr), (v, hv)):
        d.from_host(h)
    rr = x.cg_update(0.75, p, r, -0.25, v)
    ex, er = hx + npdt(0.75) * hp, hr + npdt(-0.25) * hv
    assert np.allclose(x.to_host(), ex, rtol=10 * tol, atol=10 * tol)
    assert np.allclose(r.to_host(), er, rtol=10 * tol, atol=10 * tol)
    assert abs(rr - float(er.astype(np.float64) @ er.astype(np.float64))) < 10 * tol * n
    basis, hb = [], []
    for j in range(11):
        b = mf.initialize_dof_vector(2)
        hb.append(rng.standard_normal(n).astype(npdt))
        b.from_host(hb[-1])
        basis.append(b)
    w = mf.initialize_dof_vector(2)
    hw = rng.standard_normal(n).astype(npdt)
    w.from_host(hw)
    h = w.orthogonalize(basis)
    eh = [float(hw.astype(np.float64) @ b.astype(np.float64)) for b in hb]
    assert np.allclose(h, eh, rtol=0, atol=10 * tol * n)
    ew = hw.astype(np.float64) - sum(hj * b.astype(np.float64) for hj, b in zip(eh, hb))
    assert np.allclose(w.to_host(), ew, rtol=0, atol=200 * tol)


def test_copy_convert(dgx, gpu_ctx):
    gm = dgx.Mesh(2, [2, 2], 2, [0, 0], [1, 1], [True] * 2)
    mfd = dgx.MatrixFree(gpu_ctx, gm, -1, dgx.F64)
    mff = dgx.MatrixFree(gpu_ctx, gm, -1, dgx.F32)
    d, f = mfd.initialize_dof_vector(2), mff.initialize_dof_vector(2)
    x = np.random.default_rng(0).standard_normal(d.locally_owned_size())
    d.from_host(x)
    f.copy_from(d)                       # copy_to_mg: double -> float
    assert np.array_equal(f.to_host(), x.astype(np.float32))
    d.set(0.0)
    d.copy_from(f)                       # copy_from_mg: float -> double
    assert np.array_equal(d.to_host(), x.astype(np.float32).astype(np.float64))


def test_interpolate_velocity_bit_exact(dgx, gpu_ctx):
    """fused interpolate_velocity == equ, equ, -= of the reference (NS.cc:2403-2418), bit for bit"""
    gm = dgx.Mesh(2, [2, 2], 2, [0, 0], [1, 1], [True] * 2)
    mf = dgx.MatrixFree(gpu_ctx, gm)
    a, b, e, tmp, ref = (mf.initialize_dof_vector(3, 2) for _ in range(5))
    rng = np.random.default_rng(9)
    ha, hb = rng.standard_normal(a.locally_owned_size()), rng.standard_normal(a.locally_owned_size())
    a.from_host(ha)
    b.from_host(hb)
    g = 2.0 - np.sqrt(2.0)
    for stage, (c1, c2) in {1: (1.0 + g / (2.0 * (1.0 - g)), g / (2.0 * (1.0 - g))), 2: (1.0 + (1.0 - g) / g, (1.0 - g) / g)}.items():
        dgx.interpolate_velocity(e, stage, a, b)
        ref.equ(c1, a)
        tmp.equ(c2, b)
        ref.add(-1.0, tmp)
        assert np.array_equal(e.to_host(), ref.to_host())
        assert np.array_equal(e.to_host(), c1 * ha - c2 * hb)
