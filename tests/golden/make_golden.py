#!/usr/bin/env python
"""Generates tests/golden/*.npz: seeded inputs and the ORACLE's outputs for the hot path (operators, diagonals, right-hand
sides, Chebyshev estimates, Krylov iteration counts of a TR-BDF2 step).

The reference ships no golden vectors for NavierStokes_TRBDF2_DG and cannot be built here, so these fixtures
pin the ORACLE (oracle/, the CPU restatement of NS.cc's forms), not deal.II: `-m "not gpu"` tests check that the oracle
still reproduces them (a regression pin of the checker), `-m gpu` tests compare the CUDA path with them through the C ABI.
Regenerate with `python tests/golden/make_golden.py` only when an oracle change is intended, and say why in the commit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from oracle.dg1d import shape
from oracle.mesh import CartesianMesh
from oracle.operators import MassOperator, PressureOperator, VelocityOperator
from oracle.rhs import RhsForms
from oracle.solvers import Chebyshev
from oracle.timestep import OracleProjection

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (dim, cells0, nref, periodic, boundary_ids, hi)
MESHES = {
    "channel2d": (2, [4, 2], 1, [False, False], None, [2.0, 1.0]),
    "periodic2d": (2, [1, 1], 2, [True, True], None, [1.0, 1.5]),
    "periodic3d": (3, [1, 1, 1], 1, [True, True, True], None, [1.0, 1.5, 0.75]),
    "box3d": (3, [2, 1, 1], 1, [False, True, False], None, [2.0, 1.0, 1.0]),
}


def mesh_of(name):
    dim, cells0, nref, periodic, bids, hi = MESHES[name]
    return CartesianMesh(dim, cells0, nref, [0.0] * dim, hi, periodic, bids)


def rand(n, seed):
    return np.random.default_rng(seed).standard_normal(n)


def pressure_cases():
    out = {}
    for mname, degrees in (("channel2d", (1, 2)), ("periodic2d", (3,)), ("periodic3d", (2, 4)), ("box3d", (1, 3))):
        om = mesh_of(mname)
        for p in degrees:
            for stage in (1, 2):
                op = PressureOperator(om, p, 5e-3, np.float64)
                op.set_TR_BDF2_stage(stage)
                u = rand(op.n_dofs, 100 + p)
                op.compute_diagonal()
                key = f"{mname}_Q{p}_s{stage}"
                out[key + "_u"], out[key + "_Au"], out[key + "_invdiag"] = u, op.vmult(u), op.inverse_diagonal
    return out


def velocity_cases():
    out = {}
    for mname, degrees in (("channel2d", (2,)), ("periodic2d", (2, 3)), ("periodic3d", (2, 3)), ("box3d", (2,))):
        om = mesh_of(mname)
        for pv in degrees:
            for stage in (1, 2):
                op = VelocityOperator(om, pv, 100.0, 5e-3)
                op.set_TR_BDF2_stage(stage)
                u, ue = rand(op.n_dofs, 200 + pv), rand(op.n_dofs, 300 + pv)
                op.set_u_extr(ue)
                key = f"{mname}_Q{pv}_s{stage}"
                out[key + "_u"], out[key + "_ue"], out[key + "_Au"] = u, ue, op.vmult(u)
                mo = MassOperator(om, pv)
                out[key + "_Mu"] = mo.vmult(u)
    return out


def rhs_cases():
    out = {}
    for mname, dp in (("channel2d", 1), ("periodic3d", 1), ("box3d", 1), ("periodic2d", 2)):
        om = mesh_of(mname)
        f = RhsForms(om, dp, 100.0, 5e-3)
        nv = om.n_cells * om.dim * (dp + 2) ** om.dim
        npd = om.n_cells * (dp + 1) ** om.dim
        a, b, c, p = rand(nv, 1), rand(nv, 2), rand(nv, 3), rand(npd, 4)
        key = f"{mname}_p{dp}"
        out[key + "_a"], out[key + "_b"], out[key + "_c"], out[key + "_p"] = a, b, c, p
        out[key + "_rhs_u_s1"] = f.rhs_velocity(1, [a, b, p])
        out[key + "_rhs_u_s2"] = f.rhs_velocity(2, [a, b, p, c])
        out[key + "_rhs_p_s1"] = f.rhs_pressure(1, [a, p])
        out[key + "_rhs_p_s2"] = f.rhs_pressure(2, [a, p])
        out[key + "_rhs_gradp"] = f.rhs_grad_p(p)
    return out


def solver_cases():
    out = {}
    om = mesh_of("periodic3d")
    for p in (2, 3):
        op = PressureOperator(om, p, 5e-3, np.float32)
        op.compute_diagonal()
        ch = Chebyshev(op.vmult, op.inverse_diagonal, 3, 15.0, 10)
        ch.estimate_eigenvalues(op.n_dofs, np.float32)
        out[f"cheb_Q{p}_max_eig"], out[f"cheb_Q{p}_theta"], out[f"cheb_Q{p}_delta"] = (np.float64(ch.max_eig_est), np.float64(ch.theta), np.float64(ch.delta))
    # two TR-BDF2 steps of the 2-D periodic Taylor-Green vortex, velocity Q2 / pressure Q1 (the reference's degrees)
    om = CartesianMesh(2, [1, 1], 3, [0.0, 0.0], [1.0, 1.0], [True, True])
    dp = 1
    orc = OracleProjection(om, dp, 100.0, 2e-3)
    pts = om.dof_points(np.asarray(shape(dp + 1, dp + 2).nodes)) * 2 * np.pi
    u0 = np.stack([np.sin(pts[..., 0]) * np.cos(pts[..., 1]), -np.cos(pts[..., 0]) * np.sin(pts[..., 1])], axis=1).reshape(-1)
    ppts = om.dof_points(np.asarray(shape(dp, dp + 1).nodes)) * 2 * np.pi
    p0 = (0.25 * (np.cos(2 * ppts[..., 0]) + np.cos(2 * ppts[..., 1]))).reshape(-1)
    orc.u_n, orc.u_n_minus_1, orc.pres_n = u0.copy(), u0.copy(), p0.copy()
    vmax = [orc.advance() for _ in range(2)]
    out["tgv2d_u0"], out["tgv2d_p0"] = u0, p0
    out["tgv2d_iterations"] = np.array([its for _, its in orc.iterations], dtype=np.int64)
    out["tgv2d_names"] = np.array([name for name, _ in orc.iterations])
    out["tgv2d_vmax"] = np.array(vmax)
    out["tgv2d_u_n"], out["tgv2d_pres_n"] = orc.u_n, orc.pres_n
    return out


# MappingQ1 (non-Cartesian) levels: name -> (dim, cells0, nref, periodic, boundary_ids, deformation amplitude | "ring")
GENERAL_MESHES = {
    "wavy2d": (2, [2, 1], 2, [False, True], [1, 4, 2, 3], 0.03),
    "wavy3d": (3, [2, 1, 1], 1, [False, True, False], [1, 1, 2, 3, 4, 5], 0.03),
    "ring": (2, [1, 1], 3, [False, True], [4, 1, 2, 3], "ring"),
}


def general_map(kind):
    """the vertex maps of the fixtures (the GPU test applies the same function through Mesh.transform)"""
    if kind == "ring":      # O-grid around a cylinder of radius 0.5, periodic in the angle
        def ring(X):
            x = np.array(X, dtype=np.float64)
            r, th = 0.5 + 1.5 * x[..., 0], 2 * np.pi * x[..., 1]
            out = x.copy()
            out[..., 0], out[..., 1] = r * np.cos(th), r * np.sin(th)
            return out
        return ring

    def wavy(X):
        x = np.array(X, dtype=np.float64)
        out = x.copy()
        out[..., 0] = x[..., 0] + kind * np.sin(2 * np.pi * x[..., 1]) * (1 + 0.5 * np.cos(2 * np.pi * x[..., 0]))
        out[..., 1] = x[..., 1] + kind * np.sin(2 * np.pi * x[..., 0])
        if x.shape[-1] == 3:
            out[..., 1] += 0.5 * kind * np.sin(2 * np.pi * x[..., 2])
            out[..., 2] = x[..., 2] + kind * np.sin(2 * np.pi * (x[..., 0] + x[..., 1]))
        return out
    return wavy


def general_cases():
    """pressure operator + probed diagonal, mass operator, lift / drag and output data on MappingQ1 meshes (oracle/general.py,
    oracle/diagnostics.py)"""
    from oracle.diagnostics import lift_and_drag_general, patch_data_general
    from oracle.general import GeneralMesh, MassOperatorGeneral, PressureOperatorGeneral
    out = {}
    for mname, (dim, cells0, nref, periodic, bids, kind) in GENERAL_MESHES.items():
        gm = GeneralMesh(CartesianMesh(dim, cells0, nref, [0.0] * dim, [1.0] * dim, periodic, bids), general_map(kind))
        for p in (1, 2):
            for stage in (1, 2):
                op = PressureOperatorGeneral(gm, p, 5e-3)
                op.set_TR_BDF2_stage(stage)
                u = rand(op.n_dofs, 300 + p)
                key = f"{mname}_Q{p}_s{stage}"
                out[key + "_u"], out[key + "_Au"] = u, op.vmult(u)
                if stage == 1:
                    out[key + "_invdiag"] = op.compute_diagonal()
            mass = MassOperatorGeneral(gm, p + 1)
            v = rand(mass.n_dofs, 310 + p)
            vo, pr = rand(mass.n_dofs, 320 + p), rand(gm.n_cells * (p + 1) ** dim, 330 + p)
            key = f"{mname}_Q{p}"
            out[key + "_v"], out[key + "_Mv"], out[key + "_vold"], out[key + "_p"] = v, mass.vmult(v), vo, pr
            out[key + "_liftdrag"] = np.array(lift_and_drag_general(gm, p + 1, p, v, pr, 50.0, 4))
            out[key + "_patches"] = patch_data_general(gm, p + 1, p, v, pr, vo)
    return out


def build():
    return {"pressure": pressure_cases(), "velocity": velocity_cases(), "rhs": rhs_cases(), "solvers": solver_cases(), "general": general_cases()}


if __name__ == "__main__":
    for name, data in build().items():
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print(path, os.path.getsize(path) // 1024, "KiB", len(data), "arrays")
