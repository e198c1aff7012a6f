"""CPU tests of the oracle's lift / drag and output data (oracle/diagnostics.py, following NS.cc:2608-2708) against closed
forms, and of the host-only VTU writer behind the C ABI (dgx_write_vtu_piece: no device needed)."""
import ctypes as C
import xml.etree.ElementTree as ET

import numpy as np

from oracle.dg1d import gauss_lobatto_nodes
from oracle.diagnostics import lift_and_drag, patch_data, patch_points
from oracle.mesh import CartesianMesh


def interp(om, degree, funcs):
    nodes = np.asarray(gauss_lobatto_nodes(degree + 1), dtype=np.float64)
    pts = om.dof_points(nodes)
    x = [pts[..., d] for d in range(om.dim)]
    return np.stack([f(*x) for f in funcs], axis=1).reshape(-1)


def fields_2d(om, dv=2, dp=1):
    u = interp(om, dv, [lambda x, y: y * y + x + x * y, lambda x, y: x * y + 2 * y])
    p = interp(om, dp, [lambda x, y: 3 * x + y])
    return u, p


def fields_3d(om, dv=2, dp=1):
    u = interp(om, dv, [lambda x, y, z: x * x + y * z, lambda x, y, z: x * y, lambda x, y, z: x * z * z])
    p = interp(om, dp, [lambda x, y, z: x + y])
    return u, p


def test_lift_and_drag_closed_form_2d():
    Re = 7.0
    om = CartesianMesh(2, [2, 1], 2, [0, 0], [2, 1], [False, False], [0, 1, 4, 3])   # lower y face carries id 4
    u, p = fields_2d(om)
    lift, drag = lift_and_drag(om, 2, 1, u, p, Re)
    # on y = 0 the reference's normal is (0, 1): forces = (1/Re du0/dy, 1/Re du1/dy - p) = (x / Re, (x + 2) / Re - 3 x)
    assert abs(drag - 2.0 / Re) < 1e-13
    assert abs(lift - (6.0 / Re - 6.0)) < 1e-13


def test_lift_and_drag_closed_form_3d():
    Re = 3.0
    om = CartesianMesh(3, [1, 1, 1], 2, [0, 0, 0], [1, 1, 1], [False] * 3, [0, 4, 2, 3, 5, 5])   # upper x face carries id 4
    u, p = fields_3d(om)
    lift, drag = lift_and_drag(om, 2, 1, u, p, Re)
    # on x = 1 the reference's normal is (-1, 0, 0): forces = -(1/Re du_c/dx - p delta_c0) = -(2/Re - 1 - y, y/Re, z^2/Re)
    assert abs(drag + (2.0 / Re - 1.5)) < 1e-13
    assert abs(lift + 0.5 / Re) < 1e-13


def test_patch_data_closed_form():
    om = CartesianMesh(3, [1, 2, 1], 1, [0, 0, 0], [1, 2, 1], [False] * 3)
    u, p = fields_3d(om)
    uo = 0.5 * u
    out = patch_data(om, 2, 1, u, p, uo)
    X = patch_points(om)
    x, y, z = X[..., 0], X[..., 1], X[..., 2]
    assert np.allclose(out[..., 0], x * x + y * z, atol=1e-13) and np.allclose(out[..., 3], x + y, atol=1e-13)
    assert np.allclose(out[..., 4:7], 0.5 * out[..., 0:3], atol=1e-14)
    # curl of (x^2 + y z, x y, x z^2) = (0 - 0, y - z^2, y - z)
    assert np.allclose(out[..., 7], 0.0, atol=1e-12)
    assert np.allclose(out[..., 8], y - z * z, atol=1e-12)
    assert np.allclose(out[..., 9], y - z, atol=1e-12)
    om2 = CartesianMesh(2, [2, 1], 1, [0, 0], [2, 1], [False, False])
    u2, p2 = fields_2d(om2)
    out2 = patch_data(om2, 2, 1, u2, p2, u2)
    X2 = patch_points(om2)
    # d u1/dx - d u0/dy = y - (2 y + x)
    assert np.allclose(out2[..., 5], -X2[..., 1] - X2[..., 0], atol=1e-12)


def read_vtu(path):
    root = ET.parse(path).getroot()
    piece = root.find("UnstructuredGrid/Piece")
    arrays = {}
    for da in piece.iter("DataArray"):
        name = da.get("Name", "points")
        arr = np.array(da.text.split(), dtype=np.float64)
        nc = int(da.get("NumberOfComponents", "1"))
        arrays[name] = arr.reshape(-1, nc) if nc > 1 else arr
    return int(piece.get("NumberOfPoints")), int(piece.get("NumberOfCells")), arrays


def test_vtu_writer_round_trip(dgx, tmp_path):
    om = CartesianMesh(2, [2, 1], 1, [0, 0], [2, 1], [False, False])
    u, p = fields_2d(om)
    data = patch_data(om, 2, 1, u, p, 2 * u)
    pts = patch_points(om)
    path = str(tmp_path / "piece.vtu")
    dgx.lib.dgx_write_vtu_piece.argtypes = [C.c_char_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p]
    rc = dgx.lib.dgx_write_vtu_piece(path.encode(), 2, om.n_cells, np.ascontiguousarray(pts).ctypes.data, np.ascontiguousarray(data).ctypes.data)
    assert rc == 0
    npnt, ncell, arr = read_vtu(path)
    assert npnt == om.n_cells * 4 and ncell == om.n_cells
    assert np.array_equal(arr["points"][:, :2], pts.reshape(-1, 2)) and np.all(arr["points"][:, 2] == 0)
    flat = data.reshape(-1, data.shape[-1])
    assert np.array_equal(arr["v"][:, :2], flat[:, 0:2]) and np.array_equal(arr["p"], flat[:, 2])
    assert np.array_equal(arr["v_old"][:, :2], flat[:, 3:5]) and np.array_equal(arr["vorticity"], flat[:, 5])
    assert np.array_equal(arr["connectivity"].reshape(-1, 4)[1], [4, 5, 7, 6]) and np.all(arr["types"] == 9)
    assert np.array_equal(arr["offsets"], 4 * np.arange(1, om.n_cells + 1))


def annulus(r0, r1):
    """(r, theta) in [0, 1]^2 -> ring around the origin: a structured O-grid around a cylinder, periodic in theta (direction 1)"""
    def fn(X):
        x = np.array(X, dtype=np.float64)
        r, th = r0 + (r1 - r0) * x[..., 0], 2 * np.pi * x[..., 1]      # (r, theta) counter-clockwise: det J = 2 pi r (r1 - r0) > 0
        out = x.copy()
        out[..., 0], out[..., 1] = r * np.cos(th), r * np.sin(th)
        return out
    return fn


def test_general_diagnostics_reduce_to_cartesian_and_obey_the_divergence_theorem():
    from oracle.diagnostics import lift_and_drag_general, patch_data_general
    from oracle.general import GeneralMesh
    Re = 7.0
    om = CartesianMesh(2, [2, 1], 2, [0, 0], [2, 1], [False, False], [0, 1, 4, 3])
    u, p = fields_2d(om)
    gm = GeneralMesh(om, lambda X: X)
    a, b = lift_and_drag_general(gm, 2, 1, u, p, Re), lift_and_drag(om, 2, 1, u, p, Re)
    assert abs(a[0] - b[0]) < 1e-13 and abs(a[1] - b[1]) < 1e-13
    assert np.max(np.abs(patch_data_general(gm, 2, 1, u, p, u) - patch_data(om, 2, 1, u, p, u))) < 1e-12
    # cylinder of radius 0.5 as the inner boundary (id 4) of an O-grid with 16 cells around: with u = 0 and p = x the force is
    # oint p n dS = (area of the 16-gon, 0) by the divergence theorem (the reference's normal points INTO the fluid domain's hole)
    ring = CartesianMesh(2, [1, 1], 4, [0, 0], [1, 1], [False, True], [4, 1, 2, 3])
    g = GeneralMesh(ring, annulus(0.5, 2.0))
    nodes = np.asarray(gauss_lobatto_nodes(2), dtype=np.float64)
    X = g.cell_vertices()                                                # Q1 pressure dofs sit at the vertices
    px = X[:, :, 0].reshape(-1)
    lift, drag = lift_and_drag_general(g, 2, 1, np.zeros(ring.n_cells * 2 * 9), px, Re)
    area = 0.5 * 16 * 0.25 * np.sin(2 * np.pi / 16)
    assert abs(drag + area) < 1e-12 and abs(lift) < 1e-12                # forces = (-p I)(-n_out) = p n_out, n_out = -e_r (into the cylinder)


def mapped_dof_points(gm, degree):
    """physical support points of a scalar FE_DGQ(degree) field on a MappingQ1 mesh: [cell, (k,) j, i, dim]"""
    nodes = np.asarray(gauss_lobatto_nodes(degree + 1), dtype=np.float64)
    X, dim = gm.cell_vertices(), gm.dim
    grids = np.meshgrid(*([nodes] * dim), indexing="ij")              # axes slowest..fastest
    xi = [grids[dim - 1 - d] for d in range(dim)]                      # xi[d] = reference coordinate d at every dof
    pts = 0
    for v in range(1 << dim):
        N = 1.0
        for d in range(dim):
            N = N * (xi[d] if (v >> d) & 1 else 1.0 - xi[d])
        pts = pts + N[None, ..., None] * X[:, v].reshape((gm.n_cells,) + (1,) * dim + (dim,))
    return pts


def rigid_rotation(gm, degree, c):
    pts = mapped_dof_points(gm, degree)
    comps = [-c * pts[..., 1], c * pts[..., 0]] + ([0.3 * pts[..., 0]] if gm.dim == 3 else [])
    return np.stack(comps, axis=1).reshape(-1)


def test_vorticity_indicator_and_refinement_flags(dgx):
    """refine_mesh's indicator (NS.cc:2727-2752) for the rigid rotation u = c (-y, x): vorticity 2 c, eta_K = diameter^2 4 c^2 |K| on
    Cartesian and deformed cells; refine_and_coarsen_fixed_number of the library (host logic) == the oracle's, ties included"""
    from oracle.diagnostics import refine_and_coarsen_fixed_number, vorticity_indicator
    from oracle.general import GeneralMesh
    from test_general_cpu import wavy
    c = 0.7
    cm = CartesianMesh(2, [2, 1], 2, [0, 0], [2, 1], [False, False])
    eta = vorticity_indicator(cm, 2, rigid_rotation(GeneralMesh(cm, lambda X: X), 2, c))
    assert np.allclose(eta, (cm.h[0] ** 2 + cm.h[1] ** 2) * 4 * c * c * cm.h[0] * cm.h[1], rtol=1e-6)
    for dim in (2, 3):
        cmd = CartesianMesh(dim, [2, 1, 1][:dim], 1, [0.0] * dim, [1.0] * dim, [False] * dim)
        gm = GeneralMesh(cmd, wavy(0.05))
        eta = vorticity_indicator(gm, 2, rigid_rotation(gm, 2, c))
        X = gm.cell_vertices()
        NV = 1 << dim
        diam2 = np.max(np.stack([np.sum((X[:, v] - X[:, NV - 1 - v]) ** 2, axis=-1) for v in range(NV // 2)]), axis=0)
        from oracle.general import MassOperatorGeneral
        vol = MassOperatorGeneral(gm, 1, ncomp=1).M.sum(axis=(1, 2))
        assert np.allclose(eta, diam2 * 4 * c * c * vol, rtol=1e-5)
    rng = np.random.default_rng(0)
    ind = rng.random(1000).astype(np.float32)
    ind[10:20] = ind[5]                                            # ties
    import ctypes as C
    for top, bottom in ((0.01, 0.3), (0.2, 0.2), (0.0, 0.5), (0.5, 0.0)):
        flags = dgx.refine_and_coarsen_fixed_number(ind, top, bottom)
        assert np.array_equal(flags, refine_and_coarsen_fixed_number(ind, top, bottom))
        assert (flags == 1).sum() >= int(top * 1000) and (flags == -1).sum() >= min(int(bottom * 1000), 1000 - (flags == 1).sum()) - 10


def test_vtu_writer_round_trip_3d(dgx, tmp_path):
    om = CartesianMesh(3, [1, 2, 1], 0, [0, 0, 0], [1, 2, 1], [False] * 3)
    u, p = fields_3d(om)
    data = patch_data(om, 2, 1, u, p, 3 * u)
    pts = patch_points(om)
    path = str(tmp_path / "piece3d.vtu")
    dgx.lib.dgx_write_vtu_piece.argtypes = [C.c_char_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p]
    assert dgx.lib.dgx_write_vtu_piece(path.encode(), 3, om.n_cells, np.ascontiguousarray(pts).ctypes.data, np.ascontiguousarray(data).ctypes.data) == 0
    npnt, ncell, arr = read_vtu(path)
    flat = data.reshape(-1, data.shape[-1])
    assert npnt == om.n_cells * 8 and ncell == om.n_cells and np.all(arr["types"] == 12)
    assert np.array_equal(arr["points"], pts.reshape(-1, 3)) and np.array_equal(arr["v"], flat[:, 0:3]) and np.array_equal(arr["p"], flat[:, 3])
    assert np.array_equal(arr["v_old"], flat[:, 4:7]) and np.array_equal(arr["vorticity"], flat[:, 7:10])
    assert np.array_equal(arr["connectivity"].reshape(-1, 8)[1], 8 + np.array([0, 1, 3, 2, 4, 5, 7, 6]))
