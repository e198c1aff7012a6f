"""ctypes wrapper of oracle/cpu_vmult.c (C restatement of the oracle; CPU baseline only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

from .dg1d import shape
from .mesh import CartesianMesh
from .operators import GAMMA

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
MAXN = 9


class _Params(C.Structure):
    _fields_ = [("dim", C.c_int), ("n", C.c_int), ("S", C.c_double * (MAXN * MAXN)), ("D", C.c_double * (MAXN * MAXN)),
                ("w", C.c_double * MAXN), ("f", (C.c_double * MAXN) * 2), ("g", (C.c_double * MAXN) * 2),
                ("h", C.c_double * 3), ("kappa", C.c_double), ("C", C.c_double)]


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libcpu_vmult.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE])
        _LIB = C.CDLL(path)
        _LIB.cpu_num_threads.restype = C.c_int
    return _LIB


def make_params(mesh: CartesianMesh, degree: int, dt: float, stage: int = 1) -> _Params:
    sh = shape(degree, degree + 1)
    n = degree + 1
    P = _Params()
    P.dim, P.n = mesh.dim, n
    for q in range(n):
        for i in range(n):
            P.S[q * n + i] = sh.S[q, i]
            P.D[q * n + i] = sh.D[q, i]
        P.w[q] = sh.w[q]
    for s in (0, 1):
        for i in range(n):
            P.f[s][i] = sh.f[s, i]
            P.g[s][i] = sh.g[s, i]
    for d in range(mesh.dim):
        P.h[d] = mesh.h[d]
    coeff = 1.0e6 * GAMMA * dt * GAMMA * dt if stage == 1 else 1.0e6 * (1.0 - GAMMA) * dt * (1.0 - GAMMA) * dt
    P.kappa = 1.0 / coeff
    P.C = float(n * n)
    return P


def pressure_vmult(mesh: CartesianMesh, degree: int, dt: float, src: np.ndarray, stage: int = 1) -> np.ndarray:
    lib = _lib()
    P = make_params(mesh, degree, dt, stage)
    nbr = np.ascontiguousarray(mesh.nbr, dtype=np.int64)
    coords = np.ascontiguousarray(mesh.coords, dtype=np.int32)
    src = np.ascontiguousarray(src, dtype=np.float64)
    dst = np.empty_like(src)
    lib.cpu_pressure_vmult(C.byref(P), C.c_longlong(mesh.n_cells), nbr.ctypes.data_as(C.c_void_p),
                           coords.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p))
    return dst


def time_vmult(degree: int, refine: int, dtype: str = "f64", steps: int = 3):
    """times the C port on a (2^refine)^3 periodic mesh with all host threads."""
    lib = _lib()
    mesh = CartesianMesh(3, [1, 1, 1], refine, [0, 0, 0], [1, 1, 1], [True] * 3)
    P = make_params(mesh, degree, 5e-3)
    nbr = np.ascontiguousarray(mesh.nbr, dtype=np.int64)
    coords = np.ascontiguousarray(mesh.coords, dtype=np.int32)
    n_dofs = mesh.n_cells * (degree + 1) ** 3
    src = np.sin(np.arange(n_dofs) * 1e-3)
    dst = np.empty_like(src)
    args = (C.byref(P), C.c_longlong(mesh.n_cells), nbr.ctypes.data_as(C.c_void_p), coords.ctypes.data_as(C.c_void_p),
            src.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p))
    lib.cpu_pressure_vmult(*args)           # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        lib.cpu_pressure_vmult(*args)
    el = (time.perf_counter() - t0) / steps
    cores = int(lib.cpu_num_threads())
    return {"n_dofs": n_dofs, "ms_per_step": el * 1e3, "dofs_per_s": n_dofs / el, "cores": cores, "kind": "port",
            "sample": f"C/OpenMP restatement of the reference algorithm (not deal.II), fp64, {2**refine}^3 cells Q{degree} "
                      f"= {n_dofs} DoFs, {steps} applications, {cores} threads"}


# ---- the Kronecker-factored cell-centric algorithm of the CUDA kernels, on the CPU (measurement infrastructure) --------
class _KronParams(C.Structure):
    _fields_ = [("n", C.c_int), ("Kt", (C.c_double * (MAXN * MAXN)) * 3), ("a", ((C.c_double * MAXN) * 2) * 3), ("b", ((C.c_double * MAXN) * 2) * 3),
                ("g", (C.c_double * MAXN) * 2), ("gn", (C.c_double * MAXN) * 2), ("Mh", (C.c_double * (MAXN * MAXN)) * 3),
                ("kappa", C.c_double), ("C", C.c_double)]


_KLIB = None


def _klib():
    global _KLIB
    if _KLIB is None:
        path = os.path.join(_HERE, "libcpu_kron.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE])
        _KLIB = C.CDLL(path)
        _KLIB.cpu_kron_num_threads.restype = C.c_int
    return _KLIB


def make_kron_params(mesh: CartesianMesh, degree: int, dt: float, stage: int = 1) -> _KronParams:
    sh = shape(degree, degree + 1)
    n = degree + 1
    M = np.asarray(sh.mass_ld(), dtype=np.float64)
    K = np.asarray(sh.stiff_ld(), dtype=np.float64)
    Minv = np.linalg.inv(M)
    f, g = np.asarray(sh.F_ld, dtype=np.float64), np.asarray(sh.G_ld, dtype=np.float64)
    P = _KronParams()
    P.n = n
    ns = (-1.0, 1.0)
    for d in range(3):
        h = float(mesh.h[d])
        Kt = Minv @ K / h ** 2
        for i in range(n):
            for j in range(n):
                P.Kt[d][i * n + j] = Kt[i, j]
                P.Mh[d][i * n + j] = h * M[i, j]
        for s in (0, 1):
            av, bv = Minv @ f[s] / h ** 2, ns[s] * (Minv @ g[s]) / h ** 2
            for i in range(n):
                P.a[d][s][i] = av[i]
                P.b[d][s][i] = bv[i]
    for s in (0, 1):
        for i in range(n):
            P.g[s][i] = ns[s] * g[s][i]
            P.gn[s][i] = ns[s] * g[1 - s][i]
    coeff = 1.0e6 * GAMMA * dt * GAMMA * dt if stage == 1 else 1.0e6 * (1.0 - GAMMA) * dt * (1.0 - GAMMA) * dt
    P.kappa = 1.0 / coeff
    P.C = float(n * n)
    return P


def kron_vmult(mesh: CartesianMesh, degree: int, dt: float, src: np.ndarray, stage: int = 1) -> np.ndarray:
    lib = _klib()
    P = make_kron_params(mesh, degree, dt, stage)
    nbr = np.ascontiguousarray(mesh.nbr, dtype=np.int64)
    src = np.ascontiguousarray(src, dtype=np.float64)
    dst = np.empty_like(src)
    lib.cpu_kron_vmult(C.byref(P), C.c_longlong(mesh.n_cells), nbr.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p),
                       dst.ctypes.data_as(C.c_void_p))
    return dst          # (h_d M) already carries the cell volume


def time_kron(degree: int, refine: int, steps: int = 3):
    lib = _klib()
    mesh = CartesianMesh(3, [1, 1, 1], refine, [0, 0, 0], [1, 1, 1], [True] * 3)
    P = make_kron_params(mesh, degree, 5e-3)
    nbr = np.ascontiguousarray(mesh.nbr, dtype=np.int64)
    n_dofs = mesh.n_cells * (degree + 1) ** 3
    src = np.sin(np.arange(n_dofs) * 1e-3)
    dst = np.empty_like(src)
    args = (C.byref(P), C.c_longlong(mesh.n_cells), nbr.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p))
    lib.cpu_kron_vmult(*args)
    t0 = time.perf_counter()
    for _ in range(steps):
        lib.cpu_kron_vmult(*args)
    el = (time.perf_counter() - t0) / steps
    cores = int(lib.cpu_kron_num_threads())
    return {"n_dofs": n_dofs, "ms_per_step": el * 1e3, "dofs_per_s": n_dofs / el, "cores": cores,
            "what": f"the CUDA kernels' Kronecker-factored cell-centric algorithm in C/OpenMP (NOT the reference's algorithm), fp64, "
                    f"{2**refine}^3 cells Q{degree} = {n_dofs} DoFs, {cores} threads"}


# ---- vectorised baseline: the reference algorithm executed the way deal.II executes it (oracle/cpu_simd.c) ---------------
class _SimdParams(C.Structure):
    _fields_ = [("n", C.c_int), ("S", C.c_double * 25), ("Dh", C.c_double * 25), ("w", C.c_double * 5), ("f", (C.c_double * 5) * 2),
                ("g", (C.c_double * 5) * 2), ("h", C.c_double * 3), ("kappa", C.c_double), ("C", C.c_double)]


_SLIB = None


def host_threads() -> int:
    """logical CPUs this process may run on (torchrun's OMP_NUM_THREADS=1 is deliberately ignored)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _slib():
    global _SLIB
    if _SLIB is None:
        flags = ""
        try:
            with open("/proc/cpuinfo") as fh:
                for ln in fh:
                    if ln.startswith("flags"):
                        flags = ln
                        break
        except OSError:
            pass
        name = "libcpu_simd_v4.so" if " avx512f " in flags + " " and " avx512dq " in flags + " " else "libcpu_simd_v3.so"
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE])
        _SLIB = C.CDLL(path)
        _SLIB.simd_setup.restype = C.c_void_p
        _SLIB.simd_setup.argtypes = [C.c_void_p]
        _SLIB.simd_free.argtypes = [C.c_void_p]
        _SLIB.simd_pressure_vmult.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _SLIB.isa = "AVX-512" if name.endswith("v4.so") else "AVX2"
    return _SLIB


def make_simd_params(h, degree: int, dt: float, stage: int = 1) -> _SimdParams:
    if degree != 4:
        raise ValueError("oracle/cpu_simd.c is compiled for Q4 (the BASELINE workload)")
    sh = shape(degree, degree + 1)
    n = degree + 1
    S = np.asarray(sh.S_ld, dtype=np.longdouble)
    Dh = np.asarray(sh.D_ld, dtype=np.longdouble) @ np.linalg.inv(np.asarray(S, dtype=np.float64)).astype(np.longdouble)
    P = _SimdParams()
    P.n = n
    for q in range(n):
        for i in range(n):
            P.S[q * n + i] = float(S[q, i])
            P.Dh[q * n + i] = float(Dh[q, i])
        P.w[q] = float(sh.wq_ld[q])
    for s in (0, 1):
        for i in range(n):
            P.f[s][i] = float(sh.F_ld[s, i])
            P.g[s][i] = float(sh.G_ld[s, i])
    for d in range(3):
        P.h[d] = float(h[d])
    coeff = 1.0e6 * GAMMA * dt * GAMMA * dt if stage == 1 else 1.0e6 * (1.0 - GAMMA) * dt * (1.0 - GAMMA) * dt
    P.kappa = 1.0 / coeff
    P.C = float(n * n)
    return P


def simd_vmult(mesh: CartesianMesh, degree: int, dt: float, src: np.ndarray, stage: int = 1, threads: int = 0) -> np.ndarray:
    """dst = A src on a fully periodic 3-D mesh whose cell count is a multiple of the SIMD width"""
    lib = _slib()
    P = make_simd_params(mesh.h, degree, dt, stage)
    T = lib.simd_setup(C.byref(P))
    nbr = np.ascontiguousarray(mesh.nbr, dtype=np.int64)
    assert (nbr >= 0).all() and mesh.n_cells % lib.simd_vl() == 0
    src = np.ascontiguousarray(src, dtype=np.float64)
    dst = np.empty_like(src)
    lib.simd_pressure_vmult(T, mesh.n_cells, nbr.ctypes.data, src.ctypes.data, dst.ctypes.data, threads or host_threads())
    lib.simd_free(T)
    return dst


def periodic_cube_neighbors(refine: int) -> np.ndarray:
    """nbr[cell, f] of the (2^refine)^3 periodic cube in Morton order without the oracle mesh object (128^3 = 2 M cells)"""
    L = refine
    n = 1 << L
    idx = np.arange(n ** 3, dtype=np.int64)

    def compact(v):  # every third bit
        out = np.zeros_like(v)
        for b in range(L):
            out |= ((v >> (3 * b)) & 1) << b
        return out

    def spread(v):
        out = np.zeros_like(v)
        for b in range(L):
            out |= ((v >> b) & 1) << (3 * b)
        return out

    c = [compact(idx >> d) for d in range(3)]
    nbr = np.empty((n ** 3, 6), dtype=np.int64)
    for d in range(3):
        others = sum(spread(c[e]) << e for e in range(3) if e != d)
        for s in (0, 1):
            nbr[:, 2 * d + s] = others | (spread((c[d] + (2 * s - 1)) % n) << d)
    return nbr


def time_simd(degree: int, refine: int, steps: int = 3, threads: int = 0, src: np.ndarray | None = None):
    """times the vectorised baseline on the (2^refine)^3 periodic cube with `threads` (default: all) host threads"""
    lib = _slib()
    threads = threads or host_threads()
    n_cells = (1 << refine) ** 3
    h = [1.0 / (1 << refine)] * 3
    P = make_simd_params(h, degree, 5e-3)
    T = lib.simd_setup(C.byref(P))
    nbr = periodic_cube_neighbors(refine)
    n_dofs = n_cells * (degree + 1) ** 3
    if src is None:
        src = np.sin(np.arange(n_dofs) * 1e-3)
    dst = np.empty_like(src)
    args = (T, n_cells, nbr.ctypes.data, src.ctypes.data, dst.ctypes.data, threads)
    lib.simd_pressure_vmult(*args)          # warm-up (also first touch of dst)
    t0 = time.perf_counter()
    for _ in range(steps):
        lib.simd_pressure_vmult(*args)
    el = (time.perf_counter() - t0) / steps
    lib.simd_free(T)
    return {"n_dofs": n_dofs, "ms_per_step": el * 1e3, "dofs_per_s": n_dofs / el, "cores": threads, "kind": "port",
            "checksum": float(np.abs(dst[:: max(1, n_dofs // 4096)]).sum()),
            "sample": f"reference algorithm (NS.cc:1230-1292) as deal.II executes it, restated in C (not deal.II): {lib.isa}, {lib.simd_vl()} cells per "
                      f"SIMD batch, even-odd sum factorisation, compile-time degree, face-centric, {threads} threads; fp64, {2**refine}^3 cells "
                      f"Q{degree} = {n_dofs} DoFs, {steps} applications"}
