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
