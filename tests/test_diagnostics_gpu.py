"""GPU parity of compute_lift_and_drag (NS.cc:2642-2708) and of the output data of output_results (NS.cc:2608-2633 with
PostprocessorVorticity NS.cc:131-158) through the C ABI vs the oracle (and the closed forms the oracle is pinned with)."""
import numpy as np
import pytest

from oracle.diagnostics import lift_and_drag, patch_data, patch_points
from test_diagnostics_cpu import fields_2d, fields_3d, read_vtu
from util import make_meshes, smooth_plus_noise
from oracle.dg1d import gauss_lobatto_nodes

pytestmark = pytest.mark.gpu

CASES = [
    (2, [2, 1], 2, [False, False], [0, 1, 4, 3], 1),
    (2, [3, 2], 1, [True, False], [0, 1, 4, 4], 2),
    (3, [1, 1, 1], 2, [False, False, False], [0, 4, 2, 3, 5, 5], 1),
    (3, [2, 1, 1], 1, [False, True, False], [4, 4, 2, 3, 4, 1], 2),
]


def _random_fields(om, dp, seed):
    nv = np.asarray(gauss_lobatto_nodes(dp + 2), dtype=np.float64)
    npn = np.asarray(gauss_lobatto_nodes(dp + 1), dtype=np.float64)
    u = smooth_plus_noise(om, nv, seed=seed, ncomp=om.dim, noise=0.5)
    uo = smooth_plus_noise(om, nv, seed=seed + 1, ncomp=om.dim, noise=0.5)
    p = smooth_plus_noise(om, npn, seed=seed + 2, ncomp=1, noise=0.5)
    return u, uo, p


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}d-{c[1]}-Q{c[5] + 1}Q{c[5]}")
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_lift_drag_and_patches_match_oracle(dgx, gpu_ctx, case, dtype):
    dim, cells0, nref, periodic, bids, dp = case
    Re = 13.0
    gm, om = make_meshes(dgx, dim, cells0, nref, periodic, hi=[2.0, 1.0, 0.9][:dim], boundary_ids=bids)
    mf = dgx.MatrixFree(gpu_ctx, gm, dtype=dgx.F64 if dtype == "f64" else dgx.F32)
    u, uo, p = _random_fields(om, dp, 5)
    vu, vuo, vp = mf.initialize_dof_vector(dp + 1, dim), mf.initialize_dof_vector(dp + 1, dim), mf.initialize_dof_vector(dp, 1)
    vu.from_host(u); vuo.from_host(uo); vp.from_host(p)
    if dtype == "f32":   # the oracle sees the rounded inputs
        u, uo, p = (a.astype(np.float32).astype(np.float64) for a in (u, uo, p))
    tol = 1e-12 if dtype == "f64" else 1e-12   # the device accumulates in double for both precisions
    lift, drag = dgx.compute_lift_and_drag(vu, vp, Re, 4)
    rl, rd = lift_and_drag(om, dp + 1, dp, u, p, Re, 4)
    scale = max(abs(rl), abs(rd), 1.0)
    assert abs(lift - rl) < tol * scale * 10 and abs(drag - rd) < tol * scale * 10
    assert dgx.compute_lift_and_drag(vu, vp, Re, 77) == (0.0, 0.0)   # no face carries this id
    out = dgx.output_patch_data(vu, vp, vuo)
    ref = patch_data(om, dp + 1, dp, u, p, uo)
    assert out.shape == ref.shape
    assert np.max(np.abs(out - ref)) < 1e-11 * max(1.0, np.max(np.abs(ref)))


def test_closed_forms_and_vtu(dgx, gpu_ctx, tmp_path):
    Re = 7.0
    gm, om = make_meshes(dgx, 2, [2, 1], 2, [False, False], hi=[2.0, 1.0], boundary_ids=[0, 1, 4, 3])
    mf = dgx.MatrixFree(gpu_ctx, gm)
    u, p = fields_2d(om)
    vu, vuo, vp = mf.initialize_dof_vector(2, 2), mf.initialize_dof_vector(2, 2), mf.initialize_dof_vector(1, 1)
    vu.from_host(u); vuo.from_host(0.5 * u); vp.from_host(p)
    lift, drag = dgx.compute_lift_and_drag(vu, vp, Re)
    assert abs(drag - 2.0 / Re) < 1e-13 and abs(lift - (6.0 / Re - 6.0)) < 1e-13
    path = str(tmp_path / "solution-00001.vtu")
    dgx.output_results(vu, vp, vuo, path)
    npnt, ncell, arr = read_vtu(path)
    X = patch_points(om).reshape(-1, 2)
    assert ncell == om.n_cells and np.allclose(arr["points"][:, :2], X, atol=0)
    assert np.allclose(arr["vorticity"], -X[:, 1] - X[:, 0], atol=1e-12)   # d u1/dx - d u0/dy
    assert np.allclose(arr["p"], 3 * X[:, 0] + X[:, 1], atol=1e-13) and np.allclose(arr["v_old"][:, :2], 0.5 * arr["v"][:, :2], atol=1e-15)
    gm3, om3 = make_meshes(dgx, 3, [1, 1, 1], 2, [False] * 3, hi=[1.0, 1.0, 1.0], boundary_ids=[0, 4, 2, 3, 5, 5])
    mf3 = dgx.MatrixFree(gpu_ctx, gm3)
    u3, p3 = fields_3d(om3)
    w, q = mf3.initialize_dof_vector(2, 3), mf3.initialize_dof_vector(1, 1)
    w.from_host(u3); q.from_host(p3)
    lift, drag = dgx.compute_lift_and_drag(w, q, 3.0)
    assert abs(drag + (2.0 / 3.0 - 1.5)) < 1e-13 and abs(lift + 0.5 / 3.0) < 1e-13
    with pytest.raises(dgx.DgxError):
        dgx.compute_lift_and_drag(w, vp, 3.0)   # vectors of different MatrixFree objects


def test_run_loop_reports_forces_and_writes_output(dgx, gpu_ctx, tmp_path):
    """NavierStokesProjection::run (NS.cc:2928-3010): lift / drag after every step, output every output_interval steps and at the end"""
    from dgx.navier_stokes import NavierStokesProjection
    from oracle.mesh import CartesianMesh
    args = (2, [2, 1], 2, [0.0, 0.0], [1.0, 1.0], [False, True])
    gm, om = dgx.Mesh(*args), CartesianMesh(*args)
    ns = NavierStokesProjection(gpu_ctx, gm, degree_p=1, Re=100.0, dt=2e-3)
    pts = om.dof_points(np.asarray(gauss_lobatto_nodes(3), dtype=np.float64)) * 2 * np.pi
    u0 = np.stack([np.sin(pts[..., 0]) * np.cos(pts[..., 1]), -np.cos(pts[..., 0]) * np.sin(pts[..., 1])], axis=1).reshape(-1)
    ns.u_n.from_host(u0); ns.u_n_minus_1.from_host(u0)
    forces = ns.run(T=4e-3, output_interval=2, saving_dir=str(tmp_path), boundary_id=0)
    assert len(forces) == 2 and np.isfinite(np.asarray(forces)).all()
    lift, drag = forces[-1]
    rl, rd = lift_and_drag(om, 2, 1, ns.u_n.to_host(), ns.pres_n.to_host(), 100.0, 0)
    assert abs(lift - rl) < 1e-11 * max(1.0, abs(rl)) and abs(drag - rd) < 1e-11 * max(1.0, abs(rd)) and max(abs(rl), abs(rd)) > 1e-6
    eta, flags = ns.refinement_flags(0.1, 0.3)          # first half of refine_mesh on the current u_n
    assert eta.shape == (om.n_cells,) and np.all(eta >= 0) and (flags == 1).sum() >= 3 and (flags == -1).sum() >= 9
    assert np.allclose(eta, __import__("oracle.diagnostics", fromlist=["x"]).vorticity_indicator(om, 2, ns.u_n.to_host()), rtol=2e-6)
    files = sorted(p.name for p in tmp_path.iterdir())
    assert files == ["solution-00001.vtu", "solution-00002.vtu", "solution-00003.vtu"]
    assert read_vtu(str(tmp_path / files[2]))[1] == om.n_cells


def test_vorticity_indicator_matches_oracle(dgx, gpu_ctx):
    """refine_mesh's cell-wise indicator (NS.cc:2727-2752) on the device: Cartesian 2-D / 3-D, deformed cells, the O-grid; float32 results"""
    from oracle.diagnostics import vorticity_indicator
    from oracle.general import GeneralMesh
    from oracle.mesh import CartesianMesh
    from test_diagnostics_cpu import annulus, rigid_rotation
    from test_general_cpu import wavy
    rng = np.random.default_rng(8)
    for dim, cells0, nref, periodic, fn in ((2, [2, 1], 2, [False, True], None), (3, [1, 1, 1], 2, [True] * 3, None),
                                            (2, [2, 1], 1, [False, False], wavy(0.04)), (3, [2, 1, 1], 1, [False] * 3, wavy(0.04)),
                                            (2, [1, 1], 3, [False, True], annulus(0.5, 2.0))):
        lo, hi = [0.0] * dim, [1.0] * dim
        cm = CartesianMesh(dim, cells0, nref, lo, hi, periodic)
        gm = dgx.Mesh(dim, cells0, nref, lo, hi, periodic)
        if fn is not None:
            gm.transform(fn, lo, hi, cells0)
        om = GeneralMesh(cm, fn) if fn is not None else cm
        mf = dgx.MatrixFree(gpu_ctx, gm)
        for dv in (2, 3):
            u = rng.standard_normal(cm.n_cells * dim * (dv + 1) ** dim)
            v = mf.initialize_dof_vector(dv, dim)
            v.from_host(u)
            eta = dgx.estimate_vorticity_indicator(v)
            ref = vorticity_indicator(om, dv, u)
            assert eta.dtype == np.float32 and np.allclose(eta, ref, rtol=2e-6), (dim, dv, fn is not None)
    # closed form: rigid rotation on the O-grid, vorticity 2 c everywhere
    g = GeneralMesh(cm, annulus(0.5, 2.0))
    v = mf.initialize_dof_vector(2, 2)
    v.from_host(rigid_rotation(g, 2, 0.7))
    eta = dgx.estimate_vorticity_indicator(v)
    flags = dgx.refine_and_coarsen_fixed_number(eta, 0.1, 0.3)
    assert (flags == 1).sum() >= 6 and (flags == -1).sum() >= 19 and np.all(eta[flags == 1].min() >= eta[flags == -1].max())
