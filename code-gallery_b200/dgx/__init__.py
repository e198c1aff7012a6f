"""ctypes binding of libdgx.so (the C ABI declared in include/dgx.h).

The classes mirror the reference's operator surface for this path -- ``MatrixFree``,
``LinearAlgebra::distributed::Vector`` and ``NavierStokesProjectionOperator`` with its
``MatrixFreeOperators::Base`` members (navier_stokes_TRBDF2_DG.cc:242-2060) -- with the same
method names, argument meaning and error behaviour (exceptions instead of return codes).
There is no CPU fallback: if the CUDA library is missing the import fails.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DGX_LIB_PATH", os.path.join(_HERE, "libdgx.so"))   # override: kernel tuning experiments
if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback)")
lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

F64, F32 = 0, 1
OP_VELOCITY, OP_PRESSURE, OP_MASS = 1, 2, 3
ERR_NO_CONVERGENCE = -4
NCCL_ID_BYTES = 128
_NP = {F64: np.float64, F32: np.float32}

c_ll, c_int, c_dbl, c_vp = C.c_longlong, C.c_int, C.c_double, C.c_void_p
lib.dgx_last_error.restype = C.c_char_p
lib.dgx_version.restype = C.c_char_p
lib.dgx_kernel_launch_count.restype = c_ll
lib.dgx_ctx_get_stream.restype = c_vp
lib.dgx_vec_data.restype = c_vp
for _name in ("dgx_mf_general_metric_bytes", "dgx_mesh_n_cells", "dgx_mf_n_owned_cells", "dgx_mf_n_ghost_cells", "dgx_mf_first_owned_cell",
              "dgx_vec_size", "dgx_vec_ghost_size", "dgx_vec_global_size", "dgx_op_m"):
    getattr(lib, _name).restype = c_ll


class DgxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dgx error {code}: {msg}")
        self.code = code


class NoConvergence(DgxError):
    """SolverControl::NoConvergence"""


def _check(rc):
    if rc != 0:
        msg = lib.dgx_last_error().decode()
        raise (NoConvergence if rc == ERR_NO_CONVERGENCE else DgxError)(rc, msg)


def kernel_launch_count() -> int:
    return int(lib.dgx_kernel_launch_count())


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    _check(lib.dgx_nccl_unique_id(buf))
    return buf.raw


class Context:
    """One per GPU / rank (replaces MPI_InitFinalize + MPI_COMM_WORLD, NS.cc:3038, 2216)."""

    def __init__(self, device=0, rank=0, nranks=1, nccl_id: bytes | None = None):
        self.h = c_vp()
        idbuf = C.create_string_buffer(nccl_id, NCCL_ID_BYTES) if nccl_id is not None else None
        _check(lib.dgx_ctx_create(c_int(device), c_int(rank), c_int(nranks), idbuf, C.byref(self.h)))
        self.rank, self.nranks, self.device = rank, nranks, device

    def set_stream(self, cuda_stream_ptr):
        _check(lib.dgx_ctx_set_stream(self.h, c_vp(cuda_stream_ptr)))

    def stream(self):
        return lib.dgx_ctx_get_stream(self.h)

    def synchronize(self):
        _check(lib.dgx_ctx_synchronize(self.h))

    def timer_start(self):
        _check(lib.dgx_ctx_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = c_dbl()
        _check(lib.dgx_ctx_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def close(self):
        if self.h:
            lib.dgx_ctx_destroy(self.h)
            self.h = c_vp()


class Mesh:
    """subdivided_hyper_rectangle + refine_global(n_refine) (NS.cc:2261-2270)."""

    def __init__(self, dim, cells0, n_refine, lo, hi, periodic, boundary_ids=None):
        self.dim, self.n_refine = dim, n_refine
        c0 = (c_int * 3)(*(list(cells0) + [1, 1, 1])[:3])
        lo_ = (c_dbl * 3)(*(list(lo) + [0, 0, 0])[:3])
        hi_ = (c_dbl * 3)(*(list(hi) + [1, 1, 1])[:3])
        mask = sum((1 << d) for d in range(min(dim, len(periodic))) if periodic[d])
        bids = None if boundary_ids is None else (c_int * 6)(*(list(boundary_ids) + [0] * 6)[:6])
        self.h = c_vp()
        _check(lib.dgx_mesh_create_cartesian(dim, c0, n_refine, lo_, hi_, mask, bids, C.byref(self.h)))

    def n_levels(self):
        return lib.dgx_mesh_n_levels(self.h)

    def set_vertices(self, level, coords):
        """vertex positions of `level` (lattice order, x fastest, dim doubles each): the level becomes a MappingQ1 mesh"""
        arr = np.ascontiguousarray(coords, dtype=np.float64)
        _check(lib.dgx_mesh_set_vertices(self.h, c_int(level), arr.ctypes.data_as(c_vp), c_ll(arr.size)))

    def transform(self, fn, lo, hi, cells0):
        """GridTools::transform on every level of the hierarchy: x = fn(X) for the Cartesian vertex lattice X[..., dim]"""
        for level in range(self.n_refine + 1):
            axes = [lo[d] + (hi[d] - lo[d]) / (cells0[d] << level) * np.arange((cells0[d] << level) + 1) for d in range(self.dim)]
            grids = np.meshgrid(*axes[::-1], indexing="ij")
            self.set_vertices(level, fn(np.stack(grids[::-1], axis=-1)))

    def is_general(self, level=-1):
        return bool(lib.dgx_mesh_is_general(self.h, c_int(level)))

    def n_cells(self, level=-1):
        return int(lib.dgx_mesh_n_cells(self.h, level))

    def cell_coords(self, level=-1, first=0, count=None):
        count = self.n_cells(level) - first if count is None else count
        out = np.empty((count, self.dim), dtype=np.int32)
        _check(lib.dgx_mesh_cell_coords(self.h, level, c_ll(first), c_ll(count), out.ctypes.data_as(c_vp)))
        return out

    def neighbors(self, level=-1, first=0, count=None):
        count = self.n_cells(level) - first if count is None else count
        out = np.empty((count, 2 * self.dim), dtype=np.int64)
        _check(lib.dgx_mesh_neighbors(self.h, level, c_ll(first), c_ll(count), out.ctypes.data_as(c_vp)))
        return out


class MatrixFree:
    """MatrixFree<dim, Number>::reinit for one level (NS.cc:2329 / 2374-2375)."""

    def __init__(self, ctx: Context, mesh: Mesh, level=-1, dtype=F64):
        self.ctx, self.mesh, self.dtype = ctx, mesh, dtype
        self.level = mesh.n_refine if level < 0 else level
        self.h = c_vp()
        _check(lib.dgx_mf_create(ctx.h, mesh.h, level, dtype, C.byref(self.h)))
        self.dim = mesh.dim

    def n_owned_cells(self):
        return int(lib.dgx_mf_n_owned_cells(self.h))

    def n_ghost_cells(self):
        return int(lib.dgx_mf_n_ghost_cells(self.h))

    def first_owned_cell(self):
        return int(lib.dgx_mf_first_owned_cell(self.h))

    def local_neighbors(self):
        out = np.empty((2 * self.dim, self.n_owned_cells()), dtype=np.int32)
        _check(lib.dgx_mf_local_neighbors(self.h, out.ctypes.data_as(c_vp)))
        return out

    def host_pipeline_plan(self, bytes_per_cell, chunk_log2=0):
        """schedule of the pipelined host entry point (dgx_op_vmult_host): (uploads [n,2], groups [m,3] = begin, end, need);
        empty arrays when the call would not be pipelined"""
        n_up, n_grp = c_int(), c_int()
        _check(lib.dgx_mf_host_pipeline_plan(self.h, c_ll(bytes_per_cell), c_int(chunk_log2), c_int(0), None, None, None, None, None,
                                             C.byref(n_up), C.byref(n_grp)))
        cap = max(n_up.value, n_grp.value, 1)
        ub, ue, gb, ge = (np.zeros(cap, dtype=np.int64) for _ in range(4))
        gn = np.zeros(cap, dtype=np.int32)
        _check(lib.dgx_mf_host_pipeline_plan(self.h, c_ll(bytes_per_cell), c_int(chunk_log2), c_int(cap), ub.ctypes.data_as(c_vp), ue.ctypes.data_as(c_vp),
                                             gb.ctypes.data_as(c_vp), ge.ctypes.data_as(c_vp), gn.ctypes.data_as(c_vp), C.byref(n_up), C.byref(n_grp)))
        ups = np.stack([ub[:n_up.value], ue[:n_up.value]], axis=1)
        grps = np.stack([gb[:n_grp.value], ge[:n_grp.value], gn[:n_grp.value].astype(np.int64)], axis=1)
        return ups, grps

    def ghost_cells(self):
        out = np.empty(self.n_ghost_cells(), dtype=np.int64)
        _check(lib.dgx_mf_ghost_cells(self.h, out.ctypes.data_as(c_vp)))
        return out

    def peers(self):
        """ghost-exchange plan: list of (rank, send_cells[local], recv_first, recv_count)"""
        out = []
        for i in range(lib.dgx_mf_n_peers(self.h)):
            rank, ns, rf, rc = c_int(), c_ll(), c_ll(), c_ll()
            _check(lib.dgx_mf_peer_info(self.h, i, C.byref(rank), C.byref(ns), C.byref(rf), C.byref(rc)))
            cells = np.empty(ns.value, dtype=np.int32)
            if ns.value:
                _check(lib.dgx_mf_peer_send_cells(self.h, i, cells.ctypes.data_as(c_vp)))
            out.append((rank.value, cells, rf.value, rc.value))
        return out

    def initialize_dof_vector(self, degree, ncomp=1):
        return Vector(self, degree, ncomp)

    def general_metric_tables(self, degree):
        """host copy (fp64) of the metric tables of a non-Cartesian level: (cell[c, ns + 1, n^dim], face[c, 2 dim, 2 dim + 1, n^(dim-1)])"""
        nc, nf = c_ll(), c_ll()
        _check(lib.dgx_mf_general_metric_tables(self.h, c_int(degree), None, c_ll(0), None, c_ll(0), C.byref(nc), C.byref(nf)))
        cell, face = np.empty(nc.value), np.empty(nf.value)
        _check(lib.dgx_mf_general_metric_tables(self.h, c_int(degree), cell.ctypes.data_as(c_vp), nc, face.ctypes.data_as(c_vp), nf, None, None))
        n, dim, cells = degree + 1, self.dim, self.n_owned_cells()
        return cell.reshape(cells, dim * (dim + 1) // 2 + 1, n ** dim), face.reshape(cells, 2 * dim, 2 * dim + 1, n ** (dim - 1))

    def release_general_metric(self):
        """free the metric tables of a non-Cartesian level (rebuilt by the next operator application)"""
        _check(lib.dgx_mf_release_general_metric(self.h))

    def general_metric_bytes(self, degree):
        """bytes of per-quadrature-point metric data a degree-`degree` operator streams per application on a non-Cartesian level"""
        return int(lib.dgx_mf_general_metric_bytes(self.h, c_int(degree)))


class Vector:
    """LinearAlgebra::distributed::Vector<Number> on the device."""

    def __init__(self, mf: MatrixFree, degree, ncomp=1, _handle=None):
        self.mf, self.degree, self.ncomp = mf, degree, ncomp
        self._borrowed = _handle is not None
        self.h = _handle if _handle is not None else c_vp()
        if _handle is None:
            _check(lib.dgx_vec_create(mf.h, degree, ncomp, C.byref(self.h)))
        self.dtype = _NP[lib.dgx_vec_dtype(self.h)]

    def __del__(self):
        if getattr(self, "h", None) and not self._borrowed and lib is not None:
            lib.dgx_vec_destroy(self.h)
            self.h = None

    def locally_owned_size(self):
        return int(lib.dgx_vec_size(self.h))

    def size(self):
        return int(lib.dgx_vec_global_size(self.h))

    def data_ptr(self):
        return lib.dgx_vec_data(self.h)

    def from_host(self, arr):
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        _check(lib.dgx_vec_copy_from_host(self.h, arr.ctypes.data_as(c_vp), c_ll(arr.size)))
        return self

    def to_host(self):
        out = np.empty(self.locally_owned_size(), dtype=self.dtype)
        _check(lib.dgx_vec_copy_to_host(self.h, out.ctypes.data_as(c_vp), c_ll(out.size)))
        return out

    def set(self, value):
        _check(lib.dgx_vec_set(self.h, c_dbl(value)))

    def copy_from(self, src: "Vector"):
        _check(lib.dgx_vec_copy(self.h, src.h))

    def equ(self, a, src):
        _check(lib.dgx_vec_equ(self.h, c_dbl(a), src.h))

    def add(self, a, src):
        _check(lib.dgx_vec_add(self.h, c_dbl(a), src.h))

    def sadd(self, s, a, src):
        _check(lib.dgx_vec_sadd(self.h, c_dbl(s), c_dbl(a), src.h))

    def scale(self, a):
        if isinstance(a, Vector):
            _check(lib.dgx_vec_scale_by(self.h, a.h))
        else:
            _check(lib.dgx_vec_scale(self.h, c_dbl(a)))

    def dot(self, other):
        out = c_dbl()
        _check(lib.dgx_vec_dot(self.h, other.h, C.byref(out)))
        return out.value

    def l2_norm(self):
        out = c_dbl()
        _check(lib.dgx_vec_l2_norm(self.h, C.byref(out)))
        return out.value

    def linfty_norm(self):
        out = c_dbl()
        _check(lib.dgx_vec_linfty_norm(self.h, C.byref(out)))
        return out.value

    def mean_value(self):
        out = c_dbl()
        _check(lib.dgx_vec_mean_value(self.h, C.byref(out)))
        return out.value

    def add_and_dot(self, a, v, w):
        out = c_dbl()
        _check(lib.dgx_vec_add_and_dot(self.h, c_dbl(a), v.h, w.h, C.byref(out)))
        return out.value

    def cg_update(self, a, p, r, b, v):
        """self += a*p ; r += b*v ; returns r.r  (the vector work of one SolverCG iteration in one pass)"""
        out = c_dbl()
        _check(lib.dgx_vec_cg_update(self.h, c_dbl(a), p.h, r.h, c_dbl(b), v.h, C.byref(out)))
        return out.value

    def orthogonalize(self, basis):
        """one classical Gram-Schmidt sweep against `basis` (SolverGMRES): returns h[j] = self . basis[j], self -= sum_j h[j] basis[j]"""
        k = len(basis)
        arr = (C.c_void_p * k)(*[v.h for v in basis])
        h = (c_dbl * k)()
        _check(lib.dgx_vec_orthogonalize(self.h, arr, C.c_int(k), h))
        return [h[j] for j in range(k)]

    def interpolate(self, a, x, b, y):
        """dst = a*x - b*y (interpolate_velocity, NS.cc:2403-2418)"""
        _check(lib.dgx_vec_interpolate(self.h, c_dbl(a), x.h, c_dbl(b), y.h))

    def update_ghost_values(self):
        _check(lib.dgx_vec_update_ghost_values(self.h))

    def zero_out_ghost_values(self):
        _check(lib.dgx_vec_zero_out_ghost_values(self.h))

    def compress(self):
        _check(lib.dgx_vec_compress_add(self.h))


class NavierStokesProjectionOperator:
    """One NS_stage of the reference operator (NS.cc:242-2060) behind the Base surface."""

    def __init__(self, mf: MatrixFree, NS_stage, degree, Re=1.0, dt=1.0):
        self.mf, self.kind, self.degree = mf, NS_stage, degree
        self.ncomp = 1 if NS_stage == OP_PRESSURE else mf.dim
        self.h = c_vp()
        _check(lib.dgx_op_create(mf.h, NS_stage, degree, c_dbl(Re), c_dbl(dt), C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib.dgx_op_destroy(self.h)
            self.h = None

    def initialize_dof_vector(self):
        return Vector(self.mf, self.degree, self.ncomp)

    def set_dt(self, dt):
        _check(lib.dgx_op_set_dt(self.h, c_dbl(dt)))

    def set_TR_BDF2_stage(self, stage):
        _check(lib.dgx_op_set_TR_BDF2_stage(self.h, stage))

    def set_u_extr(self, src):
        _check(lib.dgx_op_set_u_extr(self.h, src.h))

    def set_variant(self, variant):
        _check(lib.dgx_op_set_variant(self.h, variant))

    def m(self):
        return int(lib.dgx_op_m(self.h))

    def vmult(self, dst, src):
        _check(lib.dgx_op_vmult(self.h, dst.h, src.h))

    def vmult_add(self, dst, src):
        _check(lib.dgx_op_vmult_add(self.h, dst.h, src.h))

    def Tvmult(self, dst, src):
        _check(lib.dgx_op_Tvmult(self.h, dst.h, src.h))

    def Tvmult_add(self, dst, src):
        _check(lib.dgx_op_Tvmult_add(self.h, dst.h, src.h))

    def mass_inverse(self, dst, src):
        """dst = M^-1 src (exact tensor-product inverse of the DG mass matrix; NS_stage 3 only)"""
        _check(lib.dgx_op_mass_inverse(self.h, dst.h, src.h))

    def compute_diagonal(self):
        _check(lib.dgx_op_compute_diagonal(self.h))

    def get_matrix_diagonal_inverse(self):
        h = c_vp()
        _check(lib.dgx_op_get_matrix_diagonal_inverse(self.h, C.byref(h)))
        return Vector(self.mf, self.degree, self.ncomp, _handle=h)

    def el(self, i, j=None):
        """Base::el(i, i): diagonal entry of locally owned row i"""
        assert j is None or j == i, "only the diagonal is available (MatrixFreeOperators::Base::el)"
        v = c_dbl()
        _check(lib.dgx_op_el(self.h, c_ll(i), C.byref(v)))
        return v.value

    def n(self):
        return self.m()

    def get_matrix_free(self):
        return self.mf

    def precondition_Jacobi(self, dst, src, omega=1.0):
        _check(lib.dgx_op_precondition_Jacobi(self.h, dst.h, src.h, c_dbl(omega)))

    def _srcs(self, srcs):
        arr = (c_vp * len(srcs))(*[v.h for v in srcs])
        return arr, len(srcs)

    def vmult_rhs_velocity(self, dst, srcs):
        """NS.cc:788: stage 1 srcs = [u_n, u_extr, pres_n]; stage 2 srcs = [u_n, u_n_gamma, pres_int, u_extr]"""
        arr, n = self._srcs(srcs)
        _check(lib.dgx_op_vmult_rhs_velocity(self.h, dst.h, arr, n))

    def vmult_rhs_pressure(self, dst, srcs):
        """NS.cc:915: srcs = [u_star, pres_old]"""
        arr, n = self._srcs(srcs)
        _check(lib.dgx_op_vmult_rhs_pressure(self.h, dst.h, arr, n))

    def vmult_grad_p_projection(self, dst, src):
        """NS.cc:1363: dst = int grad(p) . phi on the velocity space"""
        _check(lib.dgx_op_vmult_grad_p_projection(self.h, dst.h, src.h))

    def vmult_host(self, dst: np.ndarray, src: np.ndarray):
        """End-to-end entry: host buffers, H2D + vmult + D2H inside."""
        assert dst.flags.c_contiguous and src.flags.c_contiguous and dst.dtype == src.dtype
        _check(lib.dgx_op_vmult_host(self.h, dst.ctypes.data_as(c_vp), src.ctypes.data_as(c_vp), c_ll(src.size)))


# ---- solver layer (deal.II objects wired up in NS.cc:2424-2577) ------------------------------------------
PRECOND_IDENTITY, PRECOND_JACOBI, PRECOND_MG = 0, 1, 2


def transfer_prolongate_and_add(dst_fine: Vector, src_coarse: Vector):
    """MGTransferMatrixFree::prolongate_and_add (NS.cc:2490-2491)"""
    _check(lib.dgx_transfer_prolongate_and_add(dst_fine.h, src_coarse.h))


def transfer_restrict_and_add(dst_coarse: Vector, src_fine: Vector):
    """MGTransferMatrixFree::restrict_and_add"""
    _check(lib.dgx_transfer_restrict_and_add(dst_coarse.h, src_fine.h))


def solution_transfer_refine(dst_fine: Vector, src_coarse: Vector):
    """SolutionTransfer::interpolate after refine_global (the all-cells-refined case of refine_mesh, NS.cc:2781-2822)"""
    _check(lib.dgx_solution_transfer_refine(dst_fine.h, src_coarse.h))


class PreconditionChebyshev:
    """PreconditionChebyshev<Op, Vector> with the operator's inverse diagonal (NS.cc:2493-2516)."""

    def __init__(self, op: NavierStokesProjectionOperator, degree=3, smoothing_range=15.0, eig_cg_n_iterations=10):
        self.op = op
        self.h = c_vp()
        _check(lib.dgx_cheb_create(op.h, degree, c_dbl(smoothing_range), eig_cg_n_iterations, C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib.dgx_cheb_destroy(self.h)
            self.h = None

    def estimate_eigenvalues(self):
        _check(lib.dgx_cheb_estimate(self.h))
        v = [c_dbl() for _ in range(4)]
        _check(lib.dgx_cheb_get_estimates(self.h, *[C.byref(x) for x in v]))
        return {"min_eig": v[0].value, "max_eig": v[1].value, "theta": v[2].value, "delta": v[3].value}

    def vmult(self, dst, src):
        _check(lib.dgx_cheb_vmult(self.h, dst.h, src.h))

    def step(self, x, b):
        _check(lib.dgx_cheb_step(self.h, x.h, b.h))


class Multigrid:
    """Multigrid V-cycle + PreconditionMG + coarse CG for the pressure operator (NS.cc:2490-2537)."""

    def __init__(self, ctx: Context, mesh: Mesh, degree, Re=1.0, dt=1.0, level_dtype=F32):
        self.ctx, self.mesh = ctx, mesh
        self.h = c_vp()
        _check(lib.dgx_mg_create(ctx.h, mesh.h, degree, level_dtype, c_dbl(Re), c_dbl(dt), C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib.dgx_mg_destroy(self.h)
            self.h = None

    def n_levels(self):
        return lib.dgx_mg_n_levels(self.h)

    def set_dt(self, dt):
        _check(lib.dgx_mg_set_dt(self.h, c_dbl(dt)))

    def set_TR_BDF2_stage(self, stage):
        _check(lib.dgx_mg_set_TR_BDF2_stage(self.h, stage))

    def setup(self, coarse_abs_tol, coarse_max_it=10000):
        _check(lib.dgx_mg_setup(self.h, c_dbl(coarse_abs_tol), coarse_max_it))

    def cache_estimates(self, enable=True):
        """reuse the Chebyshev eigenvalue estimates for an unchanged (stage, dt) (default) or recompute per setup like the reference"""
        _check(lib.dgx_mg_cache_estimates(self.h, 1 if enable else 0, None))

    def last_setup_was_cached(self):
        c = c_int()
        _check(lib.dgx_mg_cache_estimates(self.h, -1, C.byref(c)))
        return bool(c.value)

    def vmult(self, dst, src):
        _check(lib.dgx_mg_vmult(self.h, dst.h, src.h))

    def coarse_iterations(self):
        n = c_int()
        buf = (c_int * 4096)()
        _check(lib.dgx_mg_coarse_iterations(self.h, buf, 4096, C.byref(n)))
        return list(buf[: min(n.value, 4096)])


def _solve(fn, op, x, b, precond, max_it, abs_tol, *extra):
    kind, handle = PRECOND_IDENTITY, None
    if isinstance(precond, Multigrid):
        kind, handle = PRECOND_MG, precond.h
    elif isinstance(precond, NavierStokesProjectionOperator):
        kind, handle = PRECOND_JACOBI, precond.h
    elif precond == "jacobi":
        kind = PRECOND_JACOBI
    its, res = c_int(), c_dbl()
    rc = fn(op.h, x.h, b.h, kind, handle, max_it, c_dbl(abs_tol), *extra, C.byref(its), C.byref(res))
    if rc == ERR_NO_CONVERGENCE:
        e = NoConvergence(rc, lib.dgx_last_error().decode())
        e.last_step, e.last_residual = its.value, res.value
        raise e
    _check(rc)
    return its.value, res.value


def solve_cg(op, x, b, precond=None, max_it=10000, abs_tol=0.0):
    """SolverCG<Vector>::solve(op, x, b, precond) with SolverControl(max_it, abs_tol); returns (iterations, residual)."""
    return _solve(lib.dgx_solve_cg, op, x, b, precond, max_it, abs_tol)


def solve_gmres(op, x, b, precond=None, max_it=10000, abs_tol=0.0, max_basis=30):
    """SolverGMRES<Vector>::solve with deal.II defaults (left preconditioning, max_n_tmp_vectors = 30)."""
    return _solve(lib.dgx_solve_gmres, op, x, b, precond, max_it, abs_tol, c_int(max_basis))


# ---- diagnostics and output of the time loop (NS.cc:2608-2708), evaluated on the device --------------------
def compute_lift_and_drag(u: Vector, p: Vector, Re, boundary_id=4):
    """NavierStokesProjection::compute_lift_and_drag (NS.cc:2642-2708); returns (lift, drag) summed over all ranks"""
    lift, drag = c_dbl(), c_dbl()
    _check(lib.dgx_compute_lift_and_drag(u.h, p.h, c_dbl(Re), c_int(boundary_id), C.byref(lift), C.byref(drag)))
    return lift.value, drag.value


def estimate_vorticity_indicator(u: Vector, n_q_points_1d=None):
    """refine_mesh's cell-wise indicator (NS.cc:2727-2752): diameter^2 int (d u_1/d x_0 - d u_0/d x_1)^2, float32 per owned cell"""
    nq = u.degree + 1 if n_q_points_1d is None else n_q_points_1d          # degree_p + 2 = degree_v + 1
    out = np.empty(u.mf.n_owned_cells(), dtype=np.float32)
    _check(lib.dgx_estimate_vorticity_indicator(u.h, c_int(nq), out.ctypes.data_as(c_vp), c_ll(out.size)))
    return out


def refine_and_coarsen_fixed_number(indicators, top_fraction=0.01, bottom_fraction=0.3):
    """GridRefinement::refine_and_coarsen_fixed_number (NS.cc:2771; one rank): +1 refine / -1 coarsen / 0 per cell"""
    ind = np.ascontiguousarray(indicators, dtype=np.float32)
    flags = np.zeros(ind.size, dtype=np.int8)
    _check(lib.dgx_refine_and_coarsen_fixed_number(ind.ctypes.data_as(c_vp), c_ll(ind.size), c_dbl(top_fraction), c_dbl(bottom_fraction),
                                                   flags.ctypes.data_as(c_vp)))
    return flags


def output_patch_data(u_n: Vector, pres_n: Vector, u_n_minus_1: Vector):
    """what DataOut::build_patches(MappingQ1, 1) collects in output_results (NS.cc:2608-2630):
    [owned cells, 2^dim vertices, fields] with fields = v (dim), p, v_old (dim), vorticity (1 in 2-D, 3 in 3-D)"""
    dim = u_n.mf.dim
    nf = lib.dgx_output_n_fields(dim)
    out = np.empty((u_n.mf.n_owned_cells(), 1 << dim, nf), dtype=np.float64)
    _check(lib.dgx_output_patch_data(u_n.h, pres_n.h, u_n_minus_1.h, out.ctypes.data_as(c_vp), c_ll(out.size)))
    return out


def output_results(u_n: Vector, pres_n: Vector, u_n_minus_1: Vector, path):
    """NavierStokesProjection::output_results (NS.cc:2608-2633): writes `path` (.vtu; one piece per rank + .pvtu on several ranks)"""
    _check(lib.dgx_output_results(u_n.h, pres_n.h, u_n_minus_1.h, os.fsencode(path)))


GAMMA = 2.0 - 2.0 ** 0.5      # NS.cc:396, 2210


def interpolate_velocity(u_extr: Vector, TR_BDF2_stage: int, u_a: Vector, u_b: Vector):
    """NavierStokesProjection::interpolate_velocity (NS.cc:2403-2418): stage 1: (u_a, u_b) = (u_n, u_n_minus_1),
    stage 2: (u_a, u_b) = (u_n_gamma, u_n)."""
    g = GAMMA
    if TR_BDF2_stage == 1:
        u_extr.interpolate(1.0 + g / (2.0 * (1.0 - g)), u_a, g / (2.0 * (1.0 - g)), u_b)
    else:
        u_extr.interpolate(1.0 + (1.0 - g) / g, u_a, (1.0 - g) / g, u_b)
