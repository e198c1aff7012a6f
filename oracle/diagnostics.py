"""Lift / drag and output data of the reference's time loop (oracle; test infrastructure only).

* ``lift_and_drag`` follows NavierStokesProjection::compute_lift_and_drag,
  navier_stokes_TRBDF2_DG.cc:2642-2708: FEFaceValues with QGauss<dim-1>(degree_p + 2) on the faces with
  boundary id 4, full velocity gradient and pressure value at every face quadrature point,
  ``forces = (1/Re grad u - p I) * (-n) * JxW``, drag = forces[0], lift = forces[1].
* ``patch_data`` follows output_results, NS.cc:2608-2633, with PostprocessorVorticity, NS.cc:131-158:
  ``build_patches(MappingQ1, 1)`` evaluates every field at the 2^dim vertices of every cell (deal.II vertex order: x fastest);
  the vorticity is formed from ``solution_gradients[q][component][derivative]``.

Both are written with dense point evaluations (every basis function at every point), i.e. not the way the CUDA
kernels index the face layer / the line through the vertex.
"""
from __future__ import annotations

import numpy as np

from .dg1d import gauss_lobatto_nodes, lagrange_eval, shape
from .feeval import FaceEval
from .mesh import CartesianMesh


def lift_and_drag(mesh: CartesianMesh, degree_v: int, degree_p: int, u, p, Re: float, boundary_id: int = 4):
    dim = mesh.dim
    nq = degree_p + 2                                                   # NS.cc:2643
    shv, shp = shape(degree_v, nq), shape(degree_p, nq)
    nv, npp = degree_v + 1, degree_p + 1
    U = np.asarray(u, dtype=np.float64).reshape((mesh.n_cells, dim) + (nv,) * dim)
    P = np.asarray(p, dtype=np.float64).reshape((mesh.n_cells, 1) + (npp,) * dim)
    drag = lift = 0.0
    for f in range(2 * dim):
        d, s = f // 2, f % 2
        cells = np.nonzero(mesh.nbr[:, f] == -1 - boundary_id)[0]       # NS.cc:2671
        if len(cells) == 0:
            continue
        fv, fp = FaceEval(dim, mesh.h, shv, d, s), FaceEval(dim, mesh.h, shp, d, s)
        grad = fv.gradients(U[cells])                                   # [N, c, k, face points]: d u_c / d x_k
        pres = fp.values(P[cells])[:, 0]                                # [N, face points]
        normal = np.zeros(dim)
        normal[d] = -(2 * s - 1)                                        # NS.cc:2679: minus the outward normal
        stress = grad / Re                                              # NS.cc:2684
        for c in range(dim):
            stress[:, c, c] -= pres                                     # NS.cc:2686
        forces = np.einsum("nck...,k->nc...", stress, normal) * fv.JxW  # NS.cc:2688
        drag += float(forces[:, 0].sum())
        lift += float(forces[:, 1].sum())
    return lift, drag


def _point_tables(degree, pts):
    nodes = gauss_lobatto_nodes(degree + 1)
    L, D = lagrange_eval(nodes, np.asarray(pts, dtype=np.longdouble))
    return np.asarray(L, dtype=np.float64), np.asarray(D, dtype=np.float64)


def patch_data(mesh: CartesianMesh, degree_v: int, degree_p: int, u_n, pres_n, u_n_minus_1):
    """[n_cells, 2^dim vertices, fields]: v (dim), p, v_old (dim), vorticity (1 in 2-D, 3 in 3-D)."""
    dim = mesh.dim
    nv, npp = degree_v + 1, degree_p + 1
    Lv, Dv = _point_tables(degree_v, [0.0, 1.0])
    Lp, _ = _point_tables(degree_p, [0.0, 1.0])
    U = np.asarray(u_n, dtype=np.float64).reshape((mesh.n_cells, dim) + (nv,) * dim)
    Uo = np.asarray(u_n_minus_1, dtype=np.float64).reshape((mesh.n_cells, dim) + (nv,) * dim)
    P = np.asarray(pres_n, dtype=np.float64).reshape((mesh.n_cells,) + (npp,) * dim)
    nvort = 1 if dim == 2 else 3
    out = np.zeros((mesh.n_cells, 1 << dim, 2 * dim + 1 + nvort))

    def at_vertex(T, tabs):
        """contract the trailing spatial axes (slowest..fastest = z, y, x) of T with one 1-D row per direction"""
        for d in range(dim - 1, -1, -1):        # first contraction removes the z (or y) axis = first spatial axis
            T = np.tensordot(T, tabs[d], axes=([T.ndim - 1 - d], [0]))
        return T

    for v in range(1 << dim):
        s = [(v >> d) & 1 for d in range(dim)]
        val_rows = [Lv[s[d]] for d in range(dim)]
        out[:, v, 0:dim] = at_vertex(U, val_rows)
        out[:, v, dim] = at_vertex(P, [Lp[s[d]] for d in range(dim)])
        out[:, v, dim + 1:2 * dim + 1] = at_vertex(Uo, val_rows)
        grad = np.zeros((mesh.n_cells, dim, dim))
        for k in range(dim):
            rows = [Dv[s[d]] / mesh.h[d] if d == k else Lv[s[d]] for d in range(dim)]
            grad[:, :, k] = at_vertex(U, rows)
        if dim == 2:
            out[:, v, 2 * dim + 1] = grad[:, 1, 0] - grad[:, 0, 1]      # NS.cc:146-148
        else:
            out[:, v, 2 * dim + 1] = grad[:, 2, 1] - grad[:, 1, 2]      # NS.cc:150-155
            out[:, v, 2 * dim + 2] = grad[:, 0, 2] - grad[:, 2, 0]
            out[:, v, 2 * dim + 3] = grad[:, 1, 0] - grad[:, 0, 1]
    return out


def patch_points(mesh: CartesianMesh):
    """[n_cells, 2^dim, dim] vertex coordinates of every cell, deal.II vertex order"""
    org = mesh.cell_origin()
    out = np.zeros((mesh.n_cells, 1 << mesh.dim, mesh.dim))
    for v in range(1 << mesh.dim):
        for d in range(mesh.dim):
            out[:, v, d] = org[:, d] + ((v >> d) & 1) * mesh.h[d]
    return out


# ---- the same two functions on a MappingQ1 (non-Cartesian) mesh: dense point evaluation with physical gradients ----------------
def _dense_tables(degree, pts_per_dir):
    """B[q, i], dBref[q, i, e] for a tensor grid of reference points (pts_per_dir[d] = 1-D points of direction d), x fastest"""
    from .general import _tensor_tables
    nodes = gauss_lobatto_nodes(degree + 1)
    tabs = [tuple(np.asarray(a, np.float64) for a in lagrange_eval(nodes, np.asarray(p, dtype=np.longdouble))) for p in pts_per_dir]
    return _tensor_tables(tabs)


def lift_and_drag_general(gmesh, degree_v: int, degree_p: int, u, p, Re: float, boundary_id: int = 4):
    """compute_lift_and_drag (NS.cc:2642-2708) with FEFaceValues on MappingQ1 cells: normal_vector, JxW and the velocity gradients
    J^-T grad_xi u at the points of QGauss<dim-1>(degree_p + 2)"""
    from .dg1d import gauss_legendre
    from .general import _points, q1_jacobian
    dim = gmesh.dim
    nq = degree_p + 2
    xq, wq = (np.asarray(a, np.float64) for a in gauss_legendre(nq))
    X = gmesh.cell_vertices()
    U = np.asarray(u, np.float64).reshape(gmesh.n_cells, dim, -1)
    P = np.asarray(p, np.float64).reshape(gmesh.n_cells, -1)
    drag = lift = 0.0
    for f in range(2 * dim):
        d, s = f // 2, f % 2
        cells = np.nonzero(gmesh.nbr[:, f] == -1 - boundary_id)[0]
        if len(cells) == 0:
            continue
        per_dir = [np.array([float(s)]) if e == d else xq for e in range(dim)]
        Bv, dBv = _dense_tables(degree_v, per_dir)
        Bp, _ = _dense_tables(degree_p, per_dir)
        pts = _points(per_dir)
        wf = np.ones(1)
        for e in range(dim):
            if e != d:
                wf = np.kron(wq, wf)
        invJ, det = q1_jacobian(X[cells], pts)
        gxi = invJ[:, :, d, :]
        norm = np.linalg.norm(gxi, axis=-1)
        n_out = (2 * s - 1) * gxi / norm[..., None]
        JxW = det * norm * wf
        grad = np.einsum("qie,cqek,cmi->cqmk", dBv, invJ, U[cells])      # d u_m / d x_k
        pres = P[cells] @ Bp.T
        stress = grad / Re
        for m in range(dim):
            stress[:, :, m, m] -= pres
        forces = np.einsum("cqmk,cqk,cq->cm", stress, -n_out, JxW)        # NS.cc:2679-2688
        drag += float(forces[:, 0].sum())
        lift += float(forces[:, 1].sum())
    return lift, drag


def patch_data_general(gmesh, degree_v: int, degree_p: int, u_n, pres_n, u_n_minus_1):
    """output_results' patch data on MappingQ1 cells: values at the cell vertices, vorticity from J^-T grad_xi u there"""
    from .general import _points, q1_jacobian
    dim = gmesh.dim
    per_dir = [np.array([0.0, 1.0])] * dim
    Bv, dBv = _dense_tables(degree_v, per_dir)
    Bp, _ = _dense_tables(degree_p, per_dir)
    pts = _points(per_dir)                                               # vertex v = bits (x fastest): deal.II vertex order
    X = gmesh.cell_vertices()
    invJ, _ = q1_jacobian(X, pts)
    U = np.asarray(u_n, np.float64).reshape(gmesh.n_cells, dim, -1)
    Uo = np.asarray(u_n_minus_1, np.float64).reshape(gmesh.n_cells, dim, -1)
    P = np.asarray(pres_n, np.float64).reshape(gmesh.n_cells, -1)
    nvort = 1 if dim == 2 else 3
    out = np.zeros((gmesh.n_cells, 1 << dim, 2 * dim + 1 + nvort))
    out[:, :, 0:dim] = np.einsum("qi,cmi->cqm", Bv, U)
    out[:, :, dim] = P @ Bp.T
    out[:, :, dim + 1:2 * dim + 1] = np.einsum("qi,cmi->cqm", Bv, Uo)
    grad = np.einsum("qie,cqek,cmi->cqmk", dBv, invJ, U)
    if dim == 2:
        out[:, :, 2 * dim + 1] = grad[:, :, 1, 0] - grad[:, :, 0, 1]
    else:
        out[:, :, 2 * dim + 1] = grad[:, :, 2, 1] - grad[:, :, 1, 2]
        out[:, :, 2 * dim + 2] = grad[:, :, 0, 2] - grad[:, :, 2, 0]
        out[:, :, 2 * dim + 3] = grad[:, :, 1, 0] - grad[:, :, 0, 1]
    return out


def vorticity_indicator(gmesh_or_mesh, degree_v: int, u, nq=None):
    """refine_mesh's indicator (NS.cc:2727-2752): eta_K = diameter(K)^2 sum_q (d u_1/d x_0 - d u_0/d x_1)^2 JxW with QGauss<dim>(nq),
    nq = degree_p + 2 = degree_v + 1; dense point evaluation with physical gradients.  Accepts a CartesianMesh or a GeneralMesh."""
    from .dg1d import gauss_legendre
    from .general import GeneralMesh, _points, q1_jacobian
    gm = gmesh_or_mesh if isinstance(gmesh_or_mesh, GeneralMesh) else GeneralMesh(gmesh_or_mesh, lambda X: X)
    dim = gm.dim
    nq = degree_v + 1 if nq is None else nq
    xq, wq = (np.asarray(a, np.float64) for a in gauss_legendre(nq))
    _B, dB = _dense_tables(degree_v, [xq] * dim)
    w = np.ones(1)
    for _d in range(dim):
        w = np.kron(wq, w)
    X = gm.cell_vertices()
    invJ, det = q1_jacobian(X, _points([xq] * dim))
    U = np.asarray(u, np.float64).reshape(gm.n_cells, dim, -1)
    grad = np.einsum("qie,cqek,cmi->cqmk", dB, invJ, U)                  # d u_m / d x_k
    vort = grad[:, :, 1, 0] - grad[:, :, 0, 1]                           # NS.cc:2745
    NV = 1 << dim
    diam2 = np.max(np.stack([np.sum((X[:, v] - X[:, NV - 1 - v]) ** 2, axis=-1) for v in range(NV // 2)]), axis=0)
    return (diam2 * np.einsum("cq,cq->c", vort * vort, det * w)).astype(np.float32)


def refine_and_coarsen_fixed_number(indicators, top_fraction=0.01, bottom_fraction=0.3):
    """GridRefinement::refine_and_coarsen_fixed_number, serial semantics: refine the int(top_fraction N) largest (>= threshold),
    coarsen the int(bottom_fraction N) smallest (<= threshold)"""
    ind = np.asarray(indicators, dtype=np.float32)
    flags = np.zeros(ind.size, dtype=np.int8)
    nr, ncs = int(top_fraction * ind.size), int(bottom_fraction * ind.size)
    srt = np.sort(ind)
    if nr > 0:
        flags[ind >= srt[-nr]] = 1
    if ncs > 0:
        flags[(ind <= srt[ncs - 1]) & (flags == 0)] = -1
    return flags
