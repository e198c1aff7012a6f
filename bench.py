#!/usr/bin/env python
"""bench.py -- DG matrix-free vmult DoF/s (3D Q4 fp64), BASELINE.json's metric.

One "step" = one application of the SIPG pressure operator (NS.cc:1398-1421, NS_stage 2) to a
device-resident vector on the C2 workload of SURVEY.md 8d: unit cube, refine_global(7) = 128^3 cells,
FE_DGQ(4), periodic, fp64 (262,144,000 DoFs, 2.1 GB per vector => inputs exceed the 126 MB L2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--refine R] [--degree P] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = DoFs processed per second with src/dst resident in HBM;
`e2e` = the same through dgx_op_vmult_host (pinned host buffers, H2D + vmult + D2H timed; the three legs of a call are
  pipelined brick by brick on any number of ranks, see DESIGN.md "Host entry point");
`roofline` = algorithmic bytes (16 B/DoF) / kernel time vs MEASURED_PEAKS.json;
`cpu_baseline` = the reference algorithm as deal.II executes it (SIMD over cells, even-odd, compile-time degree), restated in C
  (oracle/cpu_simd.c), timed on all host cores on the SAME 128^3 workload (3 applications).
Extra keys outside the timed region: `mg_solve` (MG-CG pressure solve), `trbdf2_step`, `degree_sweep` (Q1..Q8 x fp64/fp32),
`general_mapping` (pressure operator and MG-CG on MappingQ1 cells), `marching_vs_generic_rel_l2`, `multi_gpu_parity` (N > 1).
`--impl reference` times that CPU implementation alone on the arm's config (rank 0 only, all host threads).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "code-gallery_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np

METRIC = "DG matrix-free vmult DoF/s (3D Q4 fp64)"
UNIT = "DoF/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--refine", type=int, default=7, help="refine_global level per GPU-count 1 (128^3 cells)")
    ap.add_argument("--degree", type=int, default=4)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--impl", default="dgx", choices=["dgx", "reference"])
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-refine", type=int, default=-1, help="refinement of the CPU baseline's cube (default: the workload's own, 128^3)")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-solve", action="store_true", help="skip the (untimed-region) MG-CG pressure solve report")
    ap.add_argument("--solve-refine", type=int, default=7, help="cells per GPU of the MG-CG report: (2^r)^3 (7: 8 GPUs = 256^3 Q3, 9 levels = BASELINE config 4)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the (untimed-region) Q1..Q8 x fp64/fp32 sweep (BASELINE config 3)")
    ap.add_argument("--no-general", action="store_true", help="skip the (untimed-region) non-Cartesian (MappingQ1) operator report")
    ap.add_argument("--solve-degree", type=int, default=3)
    ap.add_argument("--no-step", action="store_true", help="skip the (untimed-region) TR-BDF2 step report")
    ap.add_argument("--step-refine", type=int, default=5)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe): one nvidia-smi process
    streaming every 50 ms, started before and killed after the timed steps."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.samples = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.device), "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
            self.samples = [[x.strip() for x in ln.split(",")] for ln in out.strip().splitlines() if ln.strip()]
        except Exception:
            pass

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def cpu_baseline(degree, refine, dtype, steps=3):
    """The reference algorithm vectorised the way deal.II does it (oracle/cpu_simd.c: AVX2 / AVX-512 over cells, even-odd,
    compile-time degree, face-centric) on ALL host threads, whatever OMP_NUM_THREADS says (torchrun sets it to 1); other
    degrees than Q4 fall back to the scalar port (oracle/cpu_vmult.c)."""
    from oracle import cpu_port
    if degree == 4 and dtype == "f64":
        return cpu_port.time_simd(degree, refine, steps)
    os.environ["OMP_NUM_THREADS"] = str(cpu_port.host_threads())
    return cpu_port.time_vmult(degree, min(refine, 5), dtype, steps)


def layout_for(world, refine):
    """weak scaling with (2^refine)^3 cells per GPU on ONE periodic box cut into contiguous chunks of the hierarchical (p4est /
    Morton) order: 2 -> two cubes side by side, 4 -> 2 x 2 x 1 cubes, 8 -> one cube refined once more (octants, 6 face neighbours)"""
    table = {1: ([1, 1, 1], refine), 2: ([2, 1, 1], refine), 4: ([2, 2, 1], refine), 8: ([1, 1, 1], refine + 1)}
    if world in table:
        return table[world]
    return [world, 1, 1], refine


def workload_string(degree, world, refine, n_global, dtype):
    cells0, r = layout_for(world, refine)
    n = [c << r for c in cells0]
    return (f"3D DG SIPG pressure-operator vmult, Q{degree}, {n[0]}x{n[1]}x{n[2]} cells ({2**refine}^3 per GPU) periodic Cartesian, "
            f"{n_global} DoFs {dtype}")


def pressure_solve_report(dgx, ctx, args, world=1):
    """one CG solve preconditioned by the fp32 Chebyshev/V-cycle on a periodic box of `world` cubes, one per rank (weak
    scaling; not part of the timed region)"""
    p = args.solve_degree
    cells0, R = layout_for(world, args.solve_refine)
    mesh = dgx.Mesh(3, cells0, R, [0, 0, 0], [float(c) for c in cells0], [True, True, True])
    mf = dgx.MatrixFree(ctx, mesh)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, p, 100.0, 5e-3)
    mg = dgx.Multigrid(ctx, mesh, p, 100.0, 5e-3)
    x, b = op.initialize_dof_vector(), op.initialize_dof_vector()
    n = b.locally_owned_size()
    rhs = np.sin(np.arange(n) * 7e-4) + 1e-2 * np.random.default_rng(5).uniform(-1, 1, n)
    rhs -= rhs.mean()
    b.from_host(rhs)
    tol = 1e-8 * b.l2_norm()
    out = {}
    cold_ms = None
    for rep in range(5):            # passes 0, 1 warm up (NCCL connections, pools); 2 = estimates recomputed (reference behaviour);
        mg.cache_estimates(rep >= 3)   # 3 = estimates cached, cycle captured; 4 = cached, replayed CUDA graph (steady state of a time loop)
        if rep == 3:
            mg.setup(tol, 10000)    # fills the cache (untimed)
        x.set(0.0)
        ctx.synchronize()
        l0 = dgx.kernel_launch_count()
        ctx.timer_start()
        mg.setup(tol, 10000)        # the reference rebuilds diagonals / eigenvalue estimates per solve (NS.cc:2490-2517)
        its, res = dgx.solve_cg(op, x, b, mg, 10000, tol)
        ms = ctx.timer_stop()
        nc = [c << R for c in cells0]
        out = {"workload": f"pressure MG-CG, 3D Q{p}, {nc[0]}x{nc[1]}x{nc[2]} cells periodic ({2**args.solve_refine}^3 per GPU), {n * world} DoFs on {world} GPU(s), "
                           f"fp64 outer / fp32 V-cycle, tol 1e-8*||rhs||",
               "ms_per_solve": ms, "cg_iterations": its, "levels": mg.n_levels(), "final_residual": res,
               "gpu_launches": dgx.kernel_launch_count() - l0, "includes": "level diagonals + Chebyshev eigenvalue estimates + solve"}
        if rep == 2:
            cold_ms, cold_launches = ms, out["gpu_launches"]
    out["ms_per_solve"] = cold_ms
    out["gpu_launches"] = cold_launches
    out["ms_per_solve_cached_estimates"] = ms
    out["note"] = ("ms_per_solve recomputes the eigenvalue estimates like NS.cc:2501-2517; the cached variant reuses them for an unchanged (stage, dt) "
                   "and replays the V-cycle as one CUDA graph")
    return out


def degree_sweep_report(dgx, ctx, hbm_gbs, steps=5):
    """BASELINE config 3: the pressure-operator vmult for Q1..Q8 in fp64 and fp32 on the periodic cubes of SURVEY 8d C3
    (not part of the timed region); which kernel ran is reported per entry."""
    refine = {1: 8, 2: 8, 3: 7, 4: 7, 5: 7, 6: 7, 7: 6, 8: 6}
    out = []
    for p in range(1, 9):
        mesh = dgx.Mesh(3, [1, 1, 1], refine[p], [0, 0, 0], [1, 1, 1], [True] * 3)
        for dt, name, nb in ((dgx.F64, "f64", 8), (dgx.F32, "f32", 4)):
            mf = dgx.MatrixFree(ctx, mesh, -1, dt)
            op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, p, 100.0, 5e-3)
            src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
            n = src.locally_owned_size()
            src.from_host(np.sin(np.arange(n, dtype=np.float64) * 1e-3).astype(np.float64 if dt == dgx.F64 else np.float32))
            for _ in range(3):
                op.vmult(dst, src)
            ctx.synchronize()
            ctx.timer_start()
            for _ in range(steps):
                op.vmult(dst, src)
            ms = ctx.timer_stop() / steps
            gb = 2 * nb * n / (ms * 1e-3) / 1e9
            out.append({"degree": p, "dtype": name, "cells": f"{2 ** refine[p]}^3", "dofs": n, "ms": round(ms, 4),
                        "gdofs": round(n / (ms * 1e-3) / 1e9, 2), "hbm_frac": round(gb / hbm_gbs, 3),
                        "kernel": "marching" if 2 <= p <= 4 else ("q1 (one thread per cell)" if p == 1 else "generic")})
            del src, dst, op, mf
    return out


def general_mapping_report(dgx, ctx, hbm_gbs, refine=6, steps=10):
    """SURVEY 8f rank 2, the regime with per-quadrature-point metric data: the pressure operator on a smoothly deformed periodic
    cube (every cell a MappingQ1 image of the unit cell; csrc/pressure_general.cu), Q4 in fp64 and fp32.  Algorithmic bytes =
    src + dst + the metric tables the kernel streams.  Not part of the timed region.  `identity_vs_cartesian_rel_l2`: the same
    kernel on undeformed vertex positions against the Cartesian (Kronecker) kernels."""
    def wavy(X):
        x = np.array(X, dtype=np.float64)
        out = x.copy()
        a = 0.2 / (1 << refine) * 4
        out[..., 0] = x[..., 0] + a * np.sin(2 * np.pi * x[..., 1]) * (1 + 0.5 * np.cos(2 * np.pi * x[..., 0]))
        out[..., 1] = x[..., 1] + a * np.sin(2 * np.pi * x[..., 0]) + 0.5 * a * np.sin(2 * np.pi * x[..., 2])
        out[..., 2] = x[..., 2] + a * np.sin(2 * np.pi * (x[..., 0] + x[..., 1]))
        return out
    degree, out = 4, {"workload": f"3D pressure operator Q4 on a deformed periodic cube, {1 << refine}^3 MappingQ1 cells", "entries": []}
    lat = lambda r: np.stack(np.meshgrid(*[np.arange((1 << r) + 1) / (1 << r)] * 3, indexing="ij")[::-1], axis=-1)
    for dt, name, es in ((dgx.F64, "f64", 8), (dgx.F32, "f32", 4)):
        mesh = dgx.Mesh(3, [1, 1, 1], refine, [0, 0, 0], [1, 1, 1], [True] * 3)
        mesh.set_vertices(refine, wavy(lat(refine)))
        mf = dgx.MatrixFree(ctx, mesh, -1, dt)
        op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, 5e-3)
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        n = src.locally_owned_size()
        src.from_host(np.sin(np.arange(n, dtype=np.float64) * 1e-3))
        for _ in range(3):
            op.vmult(dst, src)
        ctx.synchronize()
        ctx.timer_start()
        for _ in range(steps):
            op.vmult(dst, src)
        ms = ctx.timer_stop() / steps
        bpd = 2 * es + mf.general_metric_bytes(degree) / n
        out["entries"].append({"dtype": name, "dofs": n, "ms": round(ms, 4), "gdofs": round(n / ms * 1e-6, 2), "algorithmic_bytes_per_dof": round(bpd, 1),
                               "GBs": round(bpd * n / ms * 1e-6, 1), "hbm_frac": round(bpd * n / ms * 1e-6 / hbm_gbs, 3), "kernel": "pressure_general_kernel"})
        mf.release_general_metric()      # 4 GB of tables (fp64) that the rest of the run does not need
        del src, dst, op, mf, mesh
    r = 4
    res = []
    for general in (False, True):
        mesh = dgx.Mesh(3, [1, 1, 1], r, [0, 0, 0], [1, 1, 1], [True] * 3)
        if general:
            mesh.set_vertices(r, lat(r))
        mf = dgx.MatrixFree(ctx, mesh)
        op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, 5e-3)
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        src.from_host(np.sin(np.arange(src.locally_owned_size(), dtype=np.float64) * 1e-2))
        op.vmult(dst, src)
        res.append(dst.to_host())
    out["identity_vs_cartesian_rel_l2"] = float(np.linalg.norm(res[1] - res[0]) / np.linalg.norm(res[0]))
    # projection_step's solver (NS.cc:2486-2546) with every multigrid level a MappingQ1 mesh: fp64 CG, fp32 V-cycle on the metric kernel
    r, deg = 5, 3
    mesh = dgx.Mesh(3, [1, 1, 1], r, [0, 0, 0], [1, 1, 1], [True] * 3)
    mesh.transform(wavy, [0, 0, 0], [1, 1, 1], [1, 1, 1])
    mf = dgx.MatrixFree(ctx, mesh)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, deg, 100.0, 5e-3)
    mg = dgx.Multigrid(ctx, mesh, deg, 100.0, 5e-3)
    x, b = op.initialize_dof_vector(), op.initialize_dof_vector()
    n = b.locally_owned_size()
    g = np.sin(np.arange(n) * 3e-3) + 0.1 * np.random.default_rng(1).standard_normal(n)
    g -= g.mean()
    b.from_host(g)
    tol = 1e-8 * float(np.linalg.norm(g))
    its, times = None, []
    for _ in range(4):
        x.set(0.0)
        ctx.synchronize()
        ctx.timer_start()
        mg.setup(tol, 10000)
        its, _res = dgx.solve_cg(op, x, b, mg, 10000, tol)
        times.append(round(ctx.timer_stop(), 2))
    out["mg_cg"] = {"workload": f"pressure MG-CG, Q{deg}, {1 << r}^3 deformed cells, {n} DoFs, {mg.n_levels()} MappingQ1 levels, tol 1e-8*||rhs||",
                    "cg_iterations": its, "ms_per_solve_first": times[0], "ms_per_solve": times[-1],
                    "note": "first solve: level diagonals probed with n^dim operator applications per level (as NS.cc:2018-2060) + eigenvalue estimates; "
                            "later solves with the same (stage, dt): cached diagonals / estimates, single-CTA coarse CG, V-cycle replayed as a CUDA graph"}
    return out


def trbdf2_step_report(dgx, ctx, args, world=1):
    """time per TR-BDF2 step (NS.cc:2934-2986) of the 3-D periodic Taylor-Green vortex, velocity Q3 / pressure Q2, on `world`
    GPUs ((2^step_refine)^3 cells per GPU; box = one 2 pi cube per coarse cell, so h and dt do not change with `world`)"""
    from dgx.navier_stokes import NavierStokesProjection
    dp = 2
    cells0, R = layout_for(world, args.step_refine)
    L = 2 * np.pi
    per = 2 ** (R - args.step_refine)          # 2 when the cube itself is refined once more (8 GPUs): keep h fixed
    hi = [L * c * per for c in cells0]
    mesh = dgx.Mesh(3, cells0, R, [0, 0, 0], hi, [True, True, True])
    h = L / 2 ** args.step_refine
    dt = 0.2 * h / (dp + 2)
    ns = NavierStokesProjection(ctx, mesh, dp, 1600.0, dt)
    ns_direct = NavierStokesProjection(ctx, mesh, dp, 1600.0, dt, direct_mass_inverse=True)
    # TGV initial state interpolated at the dof support points (Gauss-Lobatto nodes), owned cells in the library's order
    def nodes(n):
        from numpy.polynomial import legendre as leg
        if n == 2:
            return np.array([0.0, 1.0])
        c = np.zeros(n); c[-1] = 1.0
        return np.concatenate(([0.0], (np.sort(leg.legroots(leg.legder(c))) + 1) / 2, [1.0]))
    first, count = ns.mf.first_owned_cell(), ns.mf.n_owned_cells()
    coords = mesh.cell_coords(-1, first, count).astype(np.float64)       # [owned cells, 3]
    def field(n, f):
        x1 = nodes(n)
        X = (coords[:, None, None, None, 0] + x1[None, None, None, :]) * h
        Y = (coords[:, None, None, None, 1] + x1[None, None, :, None]) * h
        Z = (coords[:, None, None, None, 2] + x1[None, :, None, None]) * h
        return f(X + 0 * Y + 0 * Z, Y + 0 * X + 0 * Z, Z + 0 * X + 0 * Y)
    nv, npn = dp + 2, dp + 1
    u = np.stack([field(nv, lambda x, y, z: np.sin(x) * np.cos(y) * np.cos(z)), field(nv, lambda x, y, z: -np.cos(x) * np.sin(y) * np.cos(z)),
                  field(nv, lambda x, y, z: 0 * x)], axis=1).reshape(-1)
    p = field(npn, lambda x, y, z: (np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * z) + 2) / 16).reshape(-1)
    for s_ in (ns, ns_direct):
        s_.u_n.from_host(u); s_.u_n_minus_1.from_host(u); s_.pres_n.from_host(p)
    out = {}
    nc = [c << R for c in cells0]
    for rep in range(3):            # step 1 warms up (allocations, graph capture); steps 2 and 3 are reported (3 = steady state)
        ns.iterations.clear()
        ctx.synchronize()
        l0 = dgx.kernel_launch_count()
        ctx.timer_start()
        vmax = ns.advance()
        ms = ctx.timer_stop()
        out = {"workload": f"TR-BDF2 step, 3D periodic Taylor-Green vortex, {nc[0]}x{nc[1]}x{nc[2]} cells ({2**args.step_refine}^3 per GPU) on {world} GPU(s), "
                           f"velocity Q{dp+1} ({u.size * world} DoFs) / pressure Q{dp} ({p.size * world} DoFs), Re 1600, dt {dt:.3e}",
               "ms_per_step": ms, "max_velocity": vmax, "iterations": dict((k + f"#{i}", v) for i, (k, v) in enumerate(ns.iterations)),
               "gpu_launches": dgx.kernel_launch_count() - l0,
               "note": "2 GMRES + 2 MG-CG + 3 mass CG solves (reference algorithm), tensorised velocity diagonals; third step of the run"}
    for rep in range(2):
        ctx.synchronize()
        ctx.timer_start()
        ns_direct.advance()
        out["ms_per_step_direct_mass_inverse"] = ctx.timer_stop()
    return out


def multi_gpu_parity(dgx, ctx, local_rank, world, dist, torch):
    """N-rank results against the single-GPU ones on small problems, run at the start of every N > 1 bench: the distributed Q4
    vmult (marching kernel with ghost planes / halo) and the MG-CG solve (replicated coarse levels, gather, CUDA graph)."""
    solo = dgx.Context(local_rank)
    out = {}
    cells0, R = layout_for(world, 3)
    hi = [float(c) for c in cells0]
    mesh = dgx.Mesh(3, cells0, R, [0, 0, 0], hi, [True] * 3)
    # operator
    mfd, mfs = dgx.MatrixFree(ctx, mesh), dgx.MatrixFree(solo, mesh)
    opd, ops = (dgx.NavierStokesProjectionOperator(m, dgx.OP_PRESSURE, 4, 100.0, 5e-3) for m in (mfd, mfs))
    N = mesh.n_cells() * 125
    g = np.random.default_rng(42).standard_normal(N)
    lo, n = mfd.first_owned_cell() * 125, mfd.n_owned_cells() * 125
    xs, ys, xd, yd = ops.initialize_dof_vector(), ops.initialize_dof_vector(), opd.initialize_dof_vector(), opd.initialize_dof_vector()
    xs.from_host(g); xd.from_host(g[lo:lo + n])
    ops.vmult(ys, xs); opd.vmult(yd, xd)
    ref, got = ys.to_host()[lo:lo + n], yd.to_host()
    e = torch.tensor([float(np.sum((ref - got) ** 2)), float(np.sum(ref ** 2))], dtype=torch.float64, device="cuda")
    dist.all_reduce(e)
    out["vmult_rel_l2"] = float(torch.sqrt(e[0] / e[1]))
    # MG-CG
    its, sols = {}, {}
    for tag, c, mf in (("multi", ctx, mfd), ("solo", solo, mfs)):
        op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, 3, 100.0, 5e-3)
        mg = dgx.Multigrid(c, mesh, 3, 100.0, 5e-3)
        Np = mesh.n_cells() * 64
        gp = np.sin(np.arange(Np) * 3e-3) + 0.1 * np.random.default_rng(1).standard_normal(Np)
        gp -= gp.mean()
        lo2, n2 = mf.first_owned_cell() * 64, mf.n_owned_cells() * 64
        x, b = op.initialize_dof_vector(), op.initialize_dof_vector()
        b.from_host(gp[lo2:lo2 + n2])
        tol = 1e-8 * float(np.linalg.norm(gp))
        for rep in range(3):
            mg.setup(tol, 10000)
            x.set(0.0)
            its[tag], _ = dgx.solve_cg(op, x, b, mg, 10000, tol)
        if tag == "multi":
            sols[tag], rng_multi = x.to_host(), (lo2, n2)
        else:
            sols[tag] = x.to_host()[rng_multi[0]:rng_multi[0] + rng_multi[1]]
    e = torch.tensor([float(np.sum((sols["multi"] - sols["solo"]) ** 2)), float(np.sum(sols["solo"] ** 2))], dtype=torch.float64, device="cuda")
    dist.all_reduce(e)
    out.update({"mg_cg_its": its["multi"], "mg_cg_its_single_gpu": its["solo"], "mg_cg_solution_rel_l2": float(torch.sqrt(e[0] / e[1])),
                "what": f"{cells0[0] << R}x{cells0[1] << R}x{cells0[2] << R} cells on {world} ranks vs the same problem on one GPU: Q4 fp64 vmult, Q3 MG-CG (3rd solve = replayed graph)"})
    solo.close()
    return out


def run_reference(args):
    """the reference's own CPU path for this operator on the arm's config, all host threads, rank 0 only"""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if rank != 0:
        return
    t0 = time.time()
    refine = args.cpu_refine if args.cpu_refine >= 0 else args.refine
    steps = max(1, min(args.steps, 20))          # one step = one vmult of the whole cube on the host: ~0.5 s at 128^3
    res = cpu_baseline(args.degree, refine, args.dtype, steps=steps)
    ndofs = res["n_dofs"]
    same = refine == args.refine and world == 1
    line = {
        "impl": "reference", "metric": METRIC, "value": res["dofs_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload_string(args.degree, 1, refine, ndofs, args.dtype) if same else
                   workload_string(args.degree, 1, refine, ndofs, args.dtype) + f" (one GPU's share of the {world}-GPU workload: the host cores are not scaled with N)",
                   "l2": "inputs (2 vectors) exceed the host caches", "partition": "one thread per host core, contiguous chunks of the cell order"},
        "cpu_baseline": {"value": res["dofs_per_s"], "unit": UNIT, "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": res["dofs_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    # libraries (NCCL's version banner) may print to stdout: keep stdout clean for the ONE JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import dgx

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    # host side of the end-to-end leg: run on (and first-touch the pinned buffers from) the cores next to this rank's GPU; the
    # launcher does not pin ranks, and 8 ranks pulling 2 x 2.1 GB per call through the other socket's memory controller do not scale
    numa = None
    all_cores = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        numa = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        pass
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(dgx.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(dgx.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())
    ctx = dgx.Context(local_rank, rank, world, nccl_id)

    parity = None
    if world > 1:
        try:
            parity = multi_gpu_parity(dgx, ctx, local_rank, world, dist, torch)
        except Exception as exc:  # noqa: BLE001
            parity = {"error": str(exc)}
    # weak scaling: per-GPU work fixed at (2^refine)^3 cells of ONE periodic box in contiguous Morton chunks (layout_for)
    cells0, refine_g = layout_for(world, args.refine)
    mesh = dgx.Mesh(3, cells0, refine_g, [0, 0, 0], [float(c) for c in cells0], [True, True, True])
    dt = dgx.F64 if args.dtype == "f64" else dgx.F32
    mf = dgx.MatrixFree(ctx, mesh, -1, dt)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, args.degree, 100.0, 5e-3)
    op.set_variant(args.variant)
    src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
    n_local = src.locally_owned_size()
    n_global = src.size()
    npdt = np.float64 if dt == dgx.F64 else np.float32
    # synthetic input: smooth field + noise, seed 1234, generated in chunks
    rng = np.random.default_rng(1234 + rank)
    host = torch.empty(n_local, dtype=torch.float64 if dt == dgx.F64 else torch.float32, pin_memory=True)
    hnp = host.numpy()
    chunk = 1 << 24
    for o in range(0, n_local, chunk):
        m = min(chunk, n_local - o)
        hnp[o:o + m] = np.sin(np.arange(o, o + m) * 1e-3) + 1e-3 * rng.uniform(-1, 1, m)
    src.from_host(hnp)

    def barrier():
        ctx.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        op.vmult(dst, src)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)                       # let the sampler come up; the GPU idles meanwhile
    for _ in range(3):                    # re-warm after the pause (untimed)
        op.vmult(dst, src)
    barrier()
    launches0 = dgx.kernel_launch_count()
    ctx.timer_start()
    for _ in range(args.steps):
        op.vmult(dst, src)
    ms = ctx.timer_stop()
    barrier()
    launches = dgx.kernel_launch_count() - launches0
    sampler.stop()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = n_global / (ms_per_step * 1e-3)

    # the timed kernel's output on the full workload against the generic kernel (variant 1) on the same input
    check = None
    if args.variant == 0:
        try:
            ref_v = op.initialize_dof_vector()
            op.set_variant(1)
            op.vmult(ref_v, src)
            op.set_variant(0)
            nrm = ref_v.l2_norm()
            ref_v.add(-1.0, dst)
            check = ref_v.l2_norm() / nrm
            del ref_v
        except Exception as exc:  # noqa: BLE001
            check = str(exc)

    # end to end through the host-buffer entry point (pinned host memory)
    e2e = None
    if not args.no_e2e:
        hout = torch.empty_like(host, pin_memory=True)
        houtnp = hout.numpy()
        op.vmult_host(houtnp, hnp)
        barrier()
        ksteps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        ctx.timer_start()
        for _ in range(ksteps):
            op.vmult_host(houtnp, hnp)
        ems = ctx.timer_stop()
        barrier()
        if world > 1:
            t = torch.tensor([ems], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        bytes_each = n_local * host.element_size()
        e2e = {"value": n_global / (ems / ksteps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": bytes_each,
               "d2h_bytes_per_step": bytes_each, "ms_per_step": ems / ksteps, "steps": ksteps,
               "checksum": float(np.abs(houtnp[:: max(1, n_local // 4096)]).sum())}

    peaks, which = measured_peaks()
    bytes_per_dof = 2 * host.element_size()
    achieved = (n_local * bytes_per_dof) / (ms_per_step * 1e-3) / 1e9
    marching = args.variant != 1 and args.degree == 4 and args.refine >= 3
    gen = os.environ.get("DGX_MARCH", "v3")
    kernel = ("pressure_march_kernel" if gen == "v1" else "pressure_march2_kernel" if gen.startswith("v2") else "pressure_march3_kernel") \
        if marching and args.dtype == "f64" else ("pressure_march3_kernel" if marching else "pressure_generic_kernel")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": None, "peak_source": which, "algorithmic_bytes_per_dof": bytes_per_dof,
                "kernel": kernel + " (1 launch per step)"}
    # DRAM bytes per launch of the same kernel on the same workload from the committed ncu --set full capture
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            ent = json.load(open(tr)).get(f"{kernel}:Q{args.degree}:{args.dtype}:r{args.refine}")
            if ent and world == 1:
                roofline["traffic"] = ent["dram_bytes_per_launch"]
                roofline["traffic_source"] = ent["source"]
        except Exception:
            pass
    # the operator is also bound by the FP64 pipe: report the executed-flop fraction beside the HBM one
    try:
        pk = json.load(open(os.path.join(ROOT, "profiles", "r1_peaks.json")))
        # executed FP64 flops per dof (FMA = 2), counted from the SASS instruction mix of the ncu captures; v3: 44.6 FP64 instructions
        flop_per_dof = {"pressure_march_kernel": 96.0, "pressure_march2_kernel": 84.0, "pressure_march3_kernel": 65.5, "pressure_generic_kernel": 112.0}[kernel]
        key = "fp64_fma_tflops" if args.dtype == "f64" else "fp32_fma_tflops"
        tf = n_local * flop_per_dof / (ms_per_step * 1e-3) / 1e12
        roofline["fma_pipe"] = {"achieved_tflops": tf, "peak_tflops": pk[key], "frac": tf / pk[key], "flop_per_dof": flop_per_dof,
                                "peak_source": "tools/fp64_peak.cu on this pool's B200 (profiles/r1_peaks.json)"}
    except Exception:
        pass

    # MG-preconditioned pressure solve (projection_step's solver, NS.cc:2486-2546), reported beside the vmult metric
    mg_solve = None
    if not args.no_solve:
        try:
            mg_solve = pressure_solve_report(dgx, ctx, args, world)
        except Exception as exc:  # noqa: BLE001
            mg_solve = {"error": str(exc)}

    step = None
    if not args.no_step:
        try:
            step = trbdf2_step_report(dgx, ctx, args, world)
        except Exception as exc:  # noqa: BLE001
            step = {"error": str(exc)}

    sweep = None
    if world == 1 and not args.no_sweep:
        try:
            sweep = degree_sweep_report(dgx, ctx, peaks["hbm_gbs"])
        except Exception as exc:  # noqa: BLE001
            sweep = {"error": str(exc)}

    general = None
    if world == 1 and not args.no_general:
        try:
            general = general_mapping_report(dgx, ctx, peaks["hbm_gbs"])
        except Exception as exc:  # noqa: BLE001
            general = {"error": str(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            os.sched_setaffinity(0, all_cores)   # the CPU leg uses every host core
            res = cpu_baseline(args.degree, args.cpu_refine if args.cpu_refine >= 0 else args.refine, args.dtype, args.cpu_steps)
            cpu = {"value": res["dofs_per_s"], "unit": UNIT, "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]}
            try:    # extra: the kernels' own (Kronecker-factored) algorithm on the host cores, so the ratio is not inflated by the port
                from oracle import cpu_port
                os.environ["OMP_NUM_THREADS"] = str(cpu_port.host_threads())
                k = cpu_port.time_kron(args.degree, 6)
                cpu["kronecker_algorithm_on_cpu"] = {"value": k["dofs_per_s"], "unit": UNIT, "cores": k["cores"], "what": k["what"]}
            except Exception as exc:  # noqa: BLE001
                cpu["kronecker_algorithm_on_cpu"] = {"error": str(exc)}
        except Exception as exc:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload_string(args.degree, world, args.refine, n_global, args.dtype),
                       "l2": "inputs (2 vectors) exceed the 126 MB L2; no flush needed",
                       "input": "sin(1e-3 i) + 1e-3 U(-1,1), seed 1234 + rank (smooth field + noise as SURVEY 8d C2, generated per dof index)",
                       "partition": f"{world} contiguous chunks of the hierarchical (Morton) cell order of one box" if world > 1 else "single GPU"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary(),
            "mg_solve": mg_solve, "trbdf2_step": step, "marching_vs_generic_rel_l2": check, "multi_gpu_parity": parity,
            "degree_sweep": sweep, "general_mapping": general,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
