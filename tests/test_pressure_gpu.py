"""GPU parity: pressure operator / diagonal / mass operator through the C ABI vs the oracle.
Tolerances are BASELINE.json's: rel. L2 1e-12 in fp64, 1e-5 in fp32."""
import numpy as np
import pytest

from oracle.operators import MassOperator, PressureOperator
from util import make_meshes, rel_l2, smooth_plus_noise

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-12, np.float32: 1e-5}

CASES = [
    # dim, cells0, nref, periodic, degrees
    (2, [1, 1], 3, [True, True], [1, 2, 3, 4, 5, 6, 7, 8]),
    (2, [4, 2], 2, [False, False], [1, 2, 4]),           # channel: ids 0 inflow, 1 outflow, 2/3 walls
    (2, [3, 1], 1, [False, True], [2, 3]),
    (3, [1, 1, 1], 2, [True, True, True], [1, 2, 3, 4, 5, 6, 7, 8]),
    (3, [1, 1, 1], 0, [True, True, True], [2, 4]),       # one cell: its own neighbour
    (3, [2, 1, 1], 1, [False, False, False], [1, 3, 4]), # all boundary kinds incl. id 1
    (3, [1, 2, 1], 1, [True, False, True], [2]),
]


def _run(dgx, ctx, dim, cells0, nref, periodic, degree, np_dtype, stage=1, variant=0):
    gm, om = make_meshes(dgx, dim, cells0, nref, periodic)
    dt = 5e-3
    mf = dgx.MatrixFree(ctx, gm, -1, dgx.F64 if np_dtype == np.float64 else dgx.F32)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, dt)
    op.set_TR_BDF2_stage(stage)
    op.set_variant(variant)
    oop = PressureOperator(om, degree, dt, np.float64)
    oop.set_TR_BDF2_stage(stage)
    u = smooth_plus_noise(om, oop.sh.nodes, noise=0.3)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    src.from_host(u)
    op.vmult(dst, src)
    ref = oop.vmult(u.astype(np_dtype).astype(np.float64))
    return dst.to_host(), ref, op, oop, src, dst, u


@pytest.mark.parametrize("np_dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}d-{c[1]}-r{c[2]}-{''.join('p' if x else 'b' for x in c[3])}")
def test_vmult_matches_oracle(dgx, gpu_ctx, case, np_dtype):
    dim, cells0, nref, periodic, degrees = case
    for degree in degrees:
        for stage in (1, 2):
            got, ref, *_ = _run(dgx, gpu_ctx, dim, cells0, nref, periodic, degree, np_dtype, stage)
            err = rel_l2(got, ref)
            assert err < TOL[np_dtype], (degree, stage, err)


def test_vmult_add_and_Tvmult(dgx, gpu_ctx):
    got, ref, op, oop, src, dst, u = _run(dgx, gpu_ctx, 3, [1, 1, 1], 2, [True, True, False], 3, np.float64)
    op.vmult_add(dst, src)
    assert rel_l2(dst.to_host(), 2 * ref) < 1e-12
    op.Tvmult(dst, src)
    assert rel_l2(dst.to_host(), ref) < 1e-12
    assert op.m() == oop.m()
    with pytest.raises(dgx.DgxError):
        op.vmult(src, src)


@pytest.mark.parametrize("np_dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", [CASES[1], CASES[3], CASES[5], CASES[4]], ids=["2d-channel", "3d-periodic", "3d-box", "3d-1cell"])
def test_inverse_diagonal_matches_reference_probing(dgx, gpu_ctx, case, np_dtype):
    dim, cells0, nref, periodic, degrees = case
    for degree in degrees[:3]:
        if dim == 3 and degree > 4:
            continue
        got, ref, op, oop, *_ = _run(dgx, gpu_ctx, dim, cells0, nref, periodic, degree, np_dtype)
        op.compute_diagonal()
        d = op.get_matrix_diagonal_inverse().to_host()
        dref = oop.compute_diagonal()
        assert rel_l2(d, dref) < TOL[np_dtype], degree


def test_el_returns_the_diagonal_entry(dgx, gpu_ctx):
    """MatrixFreeOperators::Base::el(i, i) = 1 / inverse diagonal entry; m() == n()"""
    gm, om = make_meshes(dgx, 2, [2, 1], 1, [False, False])
    mf = dgx.MatrixFree(gpu_ctx, gm)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, 2, 100.0, 5e-3)
    with pytest.raises(dgx.DgxError):
        op.el(0)
    op.compute_diagonal()
    dinv = PressureOperator(om, 2, 5e-3).compute_diagonal()
    for i in (0, 7, op.m() - 1):
        assert abs(op.el(i, i) - 1.0 / dinv[i]) < 1e-12 / abs(dinv[i])
    assert op.n() == op.m() and op.get_matrix_free() is mf
    with pytest.raises(dgx.DgxError):
        op.el(op.m())


def test_mass_operator(dgx, gpu_ctx):
    for dim, cells0, nref in [(2, [2, 1], 2), (3, [1, 1, 2], 1)]:
        gm, om = make_meshes(dgx, dim, cells0, nref, [False] * dim)
        mf = dgx.MatrixFree(gpu_ctx, gm)
        for degree in (1, 2, 3, 5):
            op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_MASS, degree)
            oop = MassOperator(om, degree)
            u = smooth_plus_noise(om, oop.sh.nodes, ncomp=dim, noise=0.5)
            src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
            src.from_host(u)
            op.vmult(dst, src)
            assert rel_l2(dst.to_host(), oop.vmult(u)) < 1e-12


def test_mass_inverse(dgx, gpu_ctx):
    """exact tensor-product inverse of the DG mass operator: M^-1 (M u) == u"""
    for dim, cells0, nref in [(2, [2, 1], 2), (3, [1, 1, 2], 1)]:
        gm, om = make_meshes(dgx, dim, cells0, nref, [False] * dim)
        mf = dgx.MatrixFree(gpu_ctx, gm)
        for degree in (1, 3, 4):
            op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_MASS, degree)
            oop = MassOperator(om, degree)
            u = smooth_plus_noise(om, oop.sh.nodes, ncomp=dim, noise=0.5)
            a, b, c = (op.initialize_dof_vector() for _ in range(3))
            a.from_host(u)
            op.vmult(b, a)
            op.mass_inverse(c, b)
            assert rel_l2(c.to_host(), u) < 1e-12


def test_vmult_host_entry(dgx, gpu_ctx):
    got, ref, op, oop, src, dst, u = _run(dgx, gpu_ctx, 3, [1, 1, 1], 2, [True, True, True], 4, np.float64)
    out = np.empty_like(u)
    op.vmult_host(out, u)
    assert rel_l2(out, ref) < 1e-12


@pytest.mark.parametrize("case", [([1, 1, 1], 4, [True] * 3, 2), ([2, 1, 1], 4, [False, True, False], 3), ([1, 1, 1], 4, [True] * 3, 4)],
                         ids=["periodic-Q2", "boundaries-Q3", "periodic-Q4"])
def test_vmult_host_entry_pipelined(dgx, gpu_ctx, case):
    """upload, operator and download of the host entry point run chunk by chunk (three streams, slab-ordered upload); the
    result must equal the device-resident vmult on the same vector.  Chunks are forced small here (32 cells) so that the
    4096-cell meshes have 128+ of them; production chunks are >= 1 MiB."""
    import os
    os.environ["DGX_HOST_CHUNK_LOG2"] = "5"
    cells0, nref, periodic, degree = case
    hi = [float(c) for c in cells0]
    mesh = dgx.Mesh(3, cells0, nref, [0, 0, 0], hi, periodic, None if all(periodic) else [0, 1, 2, 2, 2, 2])
    mf = dgx.MatrixFree(gpu_ctx, mesh)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, 5e-3)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    n = src.locally_owned_size()
    u = np.random.default_rng(7).standard_normal(n)
    src.from_host(u)
    op.vmult(dst, src)
    ref = dst.to_host()
    for _ in range(2):                      # second call reuses the cached chunk plan
        out = np.full_like(u, np.nan)
        op.vmult_host(out, u)
        assert rel_l2(out, ref) < 1e-13
    del os.environ["DGX_HOST_CHUNK_LOG2"]


def test_properties_at_scale(dgx, gpu_ctx):
    """size-independent properties on a mesh the oracle cannot do quickly (32^3 Q4):
    symmetry <Au,v> = <u,Av>, constants map to kappa*M*1, positivity."""
    gm = dgx.Mesh(3, [1, 1, 1], 5, [0, 0, 0], [1, 1, 1], [True] * 3)
    mf = dgx.MatrixFree(gpu_ctx, gm)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, 4, 100.0, 5e-3)
    mass = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, 4, 100.0, 5e-3)
    u, v, au, av = (op.initialize_dof_vector() for _ in range(4))
    rng = np.random.default_rng(7)
    n = u.locally_owned_size()
    u.from_host(rng.standard_normal(n))
    v.from_host(rng.standard_normal(n))
    op.vmult(au, u)
    op.vmult(av, v)
    a, b = v.dot(au), u.dot(av)
    assert abs(a - b) <= 1e-12 * abs(a)
    assert u.dot(au) > 0
    one = op.initialize_dof_vector()
    one.set(1.0)
    op.vmult(au, one)
    # A*1 = kappa * M * 1 and 1^T M 1 = volume = 1
    kappa = 1.0 / (1.0e6 * (2 - np.sqrt(2)) * 5e-3 * (2 - np.sqrt(2)) * 5e-3)
    assert abs(one.dot(au) - kappa) <= 1e-10 * kappa


FAST_CASES = [
    # cells0, nref  (fully periodic 3-D boxes: the marching kernel, variant 2 = fail if it does not apply)
    ([1, 1, 1], 3),
    ([2, 1, 1], 3),
    ([1, 1, 2], 3),
    ([1, 1, 1], 4),
]


@pytest.mark.parametrize("np_dtype", [np.float64, np.float32])
@pytest.mark.parametrize("degree", [2, 3, 4])
@pytest.mark.parametrize("case", FAST_CASES, ids=lambda c: f"{c[0]}-r{c[1]}")
def test_marching_kernel_matches_oracle(dgx, gpu_ctx, case, degree, np_dtype):
    cells0, nref = case
    for stage in (1, 2):
        got, ref, op, oop, src, dst, u = _run(dgx, gpu_ctx, 3, cells0, nref, [True] * 3, degree, np_dtype, stage, variant=2)
        err = rel_l2(got, ref)
        assert err < TOL[np_dtype], (degree, stage, err)
    # vmult_add through the same kernel
    op.vmult_add(dst, src)
    assert rel_l2(dst.to_host(), 2 * ref) < TOL[np_dtype]
    # and the generic kernel agrees (variant 1)
    op.set_variant(1)
    op.vmult(dst, src)
    assert rel_l2(dst.to_host(), ref) < TOL[np_dtype]


@pytest.mark.parametrize("gen", ["v1", "v2", "v2eo", "v2r", "v2eor", "v2eol", "v3", "v3x5", "v3x6", "v3b2", "v3b32", "v3zf"])
@pytest.mark.parametrize("case", FAST_CASES + [([1, 2, 1], 4)], ids=lambda c: f"{c[0]}-r{c[1]}")
def test_marching_generations_match_oracle(dgx, gpu_ctx, case, gen, monkeypatch):
    """Q4 fp64 through every generation of the marching kernel (DGX_MARCH): v1 = round-1 kernel (named barriers), v2 =
    mbarrier hand-offs + shuffled traces + bulk stores, v2eo = v2 with even-odd 1-D contractions, v2r / v2eor = streaming XY
    role (v2eor is the default), v2eol = XY role in lockstep batches."""
    monkeypatch.setenv("DGX_MARCH", gen)
    cells0, nref = case
    for stage in (1, 2):
        got, ref, op, oop, src, dst, u = _run(dgx, gpu_ctx, 3, cells0, nref, [True] * 3, 4, np.float64, stage, variant=2)
        err = rel_l2(got, ref)
        assert err < 1e-12, (gen, stage, err)
    for _ in range(3):      # repeated launches on the same buffers: no state is left behind in shared / global memory
        op.vmult(dst, src)
    assert rel_l2(dst.to_host(), ref) < 1e-12


def test_marching_generations_agree_at_scale(dgx, gpu_ctx, monkeypatch):
    """64^3 Q4 fp64 (32.8 M dofs, 2048 tile columns of 16 planes = several waves of CTAs): every marching generation against the generic kernel"""
    gm = dgx.Mesh(3, [1, 1, 1], 6, [0, 0, 0], [1.0, 1.5, 0.75], [True] * 3)
    mf = dgx.MatrixFree(gpu_ctx, gm)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, 4, 100.0, 5e-3)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    n = src.locally_owned_size()
    src.from_host(np.random.default_rng(11).standard_normal(n))
    op.set_variant(1)
    op.vmult(dst, src)
    ref = dst.to_host()
    op.set_variant(2)
    for gen in ("v1", "v2", "v2eo", "v2r", "v2eor", "v2eol", "v3", "v3x5", "v3x6", "v3b2", "v3b32", "v3zf"):
        monkeypatch.setenv("DGX_MARCH", gen)
        dst.set(0.0)
        op.vmult(dst, src)
        assert rel_l2(dst.to_host(), ref) < 1e-13, gen


def test_headline_size_matches_oracle_by_periodic_tiling(dgx, gpu_ctx):
    """Oracle parity AT the BASELINE size (3-D Q4 fp64, 128^3 cells, 262 M dofs; the headline kernel, variant 2 = marching or an error).
    The operator on a uniform periodic mesh commutes with translations by whole cells, so an input that repeats a 16^3-cell pattern
    8 x 8 x 8 times must produce the tiled result of the 16^3 periodic problem with the SAME cell size, which the oracle computes."""
    from oracle.mesh import CartesianMesh
    degree, dt, rs, rb = 4, 5e-3, 4, 7
    ND = (degree + 1) ** 3
    small = CartesianMesh(3, [1, 1, 1], rs, [0, 0, 0], [2.0 ** (rs - rb)] * 3, [True] * 3)      # h = 1/128
    oop = PressureOperator(small, degree, dt)
    us = smooth_plus_noise(small, oop.sh.nodes, noise=0.3).reshape(small.n_cells, ND)
    ref_s = oop.vmult(us.reshape(-1)).reshape(small.n_cells, ND)
    gm = dgx.Mesh(3, [1, 1, 1], rb, [0, 0, 0], [1, 1, 1], [True] * 3)
    tile = small.cell_index(gm.cell_coords().astype(np.int64) % (1 << rs))                         # pattern cell of every big-mesh cell
    mf = dgx.MatrixFree(gpu_ctx, gm)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, dt)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    src.from_host(us[tile].reshape(-1))
    ref = ref_s[tile].reshape(-1)
    nref = np.linalg.norm(ref)
    for variant in (2, 1):
        op.set_variant(variant)
        dst.set(0.0)
        op.vmult(dst, src)
        got = dst.to_host()
        got -= ref
        assert np.linalg.norm(got) < 1e-12 * nref, variant
