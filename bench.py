#!/usr/bin/env python
"""bench.py -- DG matrix-free vmult DoF/s (3D Q4 fp64), BASELINE.json's metric.

One "step" = one application of the SIPG pressure operator (NS.cc:1398-1421, NS_stage 2) to a
device-resident vector on the C2 workload of SURVEY.md 8d: unit cube, refine_global(7) = 128^3 cells,
FE_DGQ(4), periodic, fp64 (262,144,000 DoFs, 2.1 GB per vector => inputs exceed the 126 MB L2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--refine R] [--degree P] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = DoFs processed per second with src/dst resident in HBM;
`e2e` = the same through dgx_op_vmult_host (pinned host buffers, H2D + vmult + D2H timed; on one rank the three legs
  of a call are pipelined brick by brick, see DESIGN.md "Host entry point");
`roofline` = algorithmic bytes (16 B/DoF) / kernel time vs MEASURED_PEAKS.json;
`cpu_baseline` = the oracle's C port timed on the host cores on a bounded sample.
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
    ap.add_argument("--cpu-refine", type=int, default=5)
    ap.add_argument("--no-solve", action="store_true", help="skip the (untimed-region) MG-CG pressure solve report")
    ap.add_argument("--solve-refine", type=int, default=6)
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
    """Oracle C port (oracle/cpu_vmult.c) on all host cores; falls back to the numpy oracle."""
    from oracle import cpu_port
    return cpu_port.time_vmult(degree, refine, dtype, steps)


def pressure_solve_report(dgx, ctx, args, world=1):
    """one CG solve preconditioned by the fp32 Chebyshev/V-cycle on a periodic box of `world` cubes, one per rank (weak
    scaling; not part of the timed region)"""
    R, p = args.solve_refine, args.solve_degree
    mesh = dgx.Mesh(3, [world, 1, 1], R, [0, 0, 0], [float(world), 1, 1], [True, True, True])
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
    for rep in range(3):            # pass 0 warms up; pass 1 = estimates recomputed (reference behaviour); pass 2 = estimates cached
        mg.cache_estimates(rep == 2)
        if rep == 2:
            mg.setup(tol, 10000)    # fills the cache (untimed)
        x.set(0.0)
        ctx.synchronize()
        l0 = dgx.kernel_launch_count()
        ctx.timer_start()
        mg.setup(tol, 10000)        # the reference rebuilds diagonals / eigenvalue estimates per solve (NS.cc:2490-2517)
        its, res = dgx.solve_cg(op, x, b, mg, 10000, tol)
        ms = ctx.timer_stop()
        out = {"workload": f"pressure MG-CG, 3D Q{p}, {world}x{2**R}^3 cells periodic, {n * world} DoFs on {world} GPU(s), fp64 outer / fp32 V-cycle, tol 1e-8*||rhs||",
               "ms_per_solve": ms, "cg_iterations": its, "levels": mg.n_levels(), "final_residual": res,
               "gpu_launches": dgx.kernel_launch_count() - l0, "includes": "level diagonals + Chebyshev eigenvalue estimates + solve"}
        if rep == 1:
            cold_ms = ms
    out["ms_per_solve"] = cold_ms
    out["ms_per_solve_cached_estimates"] = ms
    out["note"] = "ms_per_solve recomputes the eigenvalue estimates like NS.cc:2501-2517; the cached variant reuses them for an unchanged (stage, dt)"
    return out


def trbdf2_step_report(dgx, ctx, args):
    """time per TR-BDF2 step (NS.cc:2934-2986) of the 3-D periodic Taylor-Green vortex, velocity Q3 / pressure Q2"""
    from dgx.navier_stokes import NavierStokesProjection
    R, dp = args.step_refine, 2
    L = 2 * np.pi
    mesh = dgx.Mesh(3, [1, 1, 1], R, [0, 0, 0], [L, L, L], [True, True, True])
    h = L / 2 ** R
    dt = 0.2 * h / (dp + 2)
    ns = NavierStokesProjection(ctx, mesh, dp, 1600.0, dt)
    ns_direct = NavierStokesProjection(ctx, mesh, dp, 1600.0, dt, direct_mass_inverse=True)
    # TGV initial state interpolated at the dof support points (Gauss-Lobatto nodes), cells in the library's order
    def nodes(n):
        from numpy.polynomial import legendre as leg
        if n == 2:
            return np.array([0.0, 1.0])
        c = np.zeros(n); c[-1] = 1.0
        return np.concatenate(([0.0], (np.sort(leg.legroots(leg.legder(c))) + 1) / 2, [1.0]))
    coords = mesh.cell_coords().astype(np.float64)                       # [cells, 3]
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
    for rep in range(2):
        ns.iterations.clear()
        ctx.synchronize()
        l0 = dgx.kernel_launch_count()
        ctx.timer_start()
        vmax = ns.advance()
        ms = ctx.timer_stop()
        out = {"workload": f"TR-BDF2 step, 3D periodic Taylor-Green vortex, {2**R}^3 cells, velocity Q{dp+1} ({u.size} DoFs) / pressure Q{dp} ({p.size} DoFs), Re 1600, dt {dt:.3e}",
               "ms_per_step": ms, "max_velocity": vmax, "iterations": dict((k + f"#{i}", v) for i, (k, v) in enumerate(ns.iterations)),
               "gpu_launches": dgx.kernel_launch_count() - l0,
               "note": "2 GMRES + 2 MG-CG + 3 mass CG solves (reference algorithm), tensorised velocity diagonals"}
    for rep in range(2):
        ctx.synchronize()
        ctx.timer_start()
        ns_direct.advance()
        out["ms_per_step_direct_mass_inverse"] = ctx.timer_stop()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    res = cpu_baseline(args.degree, args.cpu_refine, args.dtype, steps=max(1, args.steps))
    ndofs = res["n_dofs"]
    line = {
        "impl": "reference", "metric": METRIC, "value": res["dofs_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"3D SIPG pressure vmult Q{args.degree}, {2**args.cpu_refine}^3 cells periodic "
                               f"(bounded sample of the 128^3 workload), {ndofs} DoFs", "timing": "inputs exceed host caches"},
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
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(dgx.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(dgx.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())
    ctx = dgx.Context(local_rank, rank, world, nccl_id)

    # weak scaling: per-GPU work fixed at (2^refine)^3 cells; N GPUs => N times the cells along x
    cells0 = [world, 1, 1]
    mesh = dgx.Mesh(3, cells0, args.refine, [0, 0, 0], [float(world), 1, 1], [True, True, True])
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
    kernel = "pressure_march_kernel" if marching else "pressure_generic_kernel"
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
        flop_per_dof = {"pressure_march_kernel": 96.0, "pressure_generic_kernel": 112.0}[kernel]
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
    if not args.no_step and world == 1:
        try:
            step = trbdf2_step_report(dgx, ctx, args)
        except Exception as exc:  # noqa: BLE001
            step = {"error": str(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            res = cpu_baseline(args.degree, args.cpu_refine, args.dtype)
            cpu = {"value": res["dofs_per_s"], "unit": UNIT, "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]}
            try:    # extra: the kernels' own (Kronecker-factored) algorithm on the host cores, so the ratio is not inflated by the port
                from oracle import cpu_port
                k = cpu_port.time_kron(args.degree, args.cpu_refine + 1)
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
            "config": {"workload": f"3D DG SIPG pressure-operator vmult, Q{args.degree}, {world}x{2**args.refine}^3 cells periodic Cartesian, "
                                   f"{n_global} DoFs {args.dtype}", "l2": "inputs (2 vectors) exceed the 126 MB L2; no flush needed",
                       "partition": f"{world} contiguous Morton chunks" if world > 1 else "single GPU"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary(),
            "mg_solve": mg_solve, "trbdf2_step": step,
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
