#!/usr/bin/env python
"""Multi-GPU correctness: the distributed pressure / velocity vmult (contiguous Morton chunks, NCCL ghost import) must equal
the single-GPU result on the same global vector.  Launch with torchrun, one rank per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_multi_gpu.py
Every rank also builds the whole problem on its own GPU (single-rank context) as the reference result."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "code-gallery_b200")]
import numpy as np
import torch
import torch.distributed as dist
import dgx

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
idt = torch.zeros(dgx.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(dgx.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
ctx = dgx.Context(local, rank, world, bytes(idt.cpu().numpy().tobytes()))
solo = dgx.Context(local)
report = {}
def wavy(X):   # smooth deformation of the box, periodic up to a translation: the MappingQ1 metric path (pressure_general.cu)
    x = np.array(X, dtype=np.float64); out = x.copy()
    out[..., 0] = x[..., 0] + 0.03 * np.sin(2 * np.pi * x[..., 1]); out[..., 1] = x[..., 1] + 0.03 * np.sin(2 * np.pi * x[..., 2])
    out[..., 2] = x[..., 2] + 0.03 * np.sin(2 * np.pi * x[..., 0])
    return out

for name, cells0, nref, periodic, kind, degree in [
        ("pressure Q4 periodic slab (marching kernel)", [world, 1, 1], 3, [True] * 3, dgx.OP_PRESSURE, 4),
        ("pressure Q3 periodic cube (Morton chunks)", [1, 1, 1], 3, [True] * 3, dgx.OP_PRESSURE, 3),
        ("pressure Q2 box with boundaries", [2, 1, 1], 2, [False] * 3, dgx.OP_PRESSURE, 2),
        ("pressure Q1 periodic cube (one-thread-per-cell kernel)", [1, 1, 1], 4, [True] * 3, dgx.OP_PRESSURE, 1),
        ("pressure Q1 box with boundaries (one-thread-per-cell kernel)", [2, 1, 1], 3, [False] * 3, dgx.OP_PRESSURE, 1),
        ("pressure Q3 deformed periodic cube (metric kernel)", [1, 1, 1], 2, [True] * 3, dgx.OP_PRESSURE, 3),
        ("pressure Q2 deformed box with boundaries (metric kernel)", [2, 1, 1], 2, [False, True, False], dgx.OP_PRESSURE, 2),
        ("velocity Q3 periodic cube", [1, 1, 1], 2, [True] * 3, dgx.OP_VELOCITY, 3)]:
    hi = [float(c) for c in cells0]
    mesh = dgx.Mesh(3, cells0, nref, [0, 0, 0], hi, periodic, [1, 1, 2, 3, 1, 5] if "deformed box" in name else None)
    if "deformed" in name:
        mesh.transform(wavy, [0, 0, 0], hi, cells0)
    errs = []
    for dt in (dgx.F64, dgx.F32):
        if kind == dgx.OP_VELOCITY and dt == dgx.F32:
            continue
        mfd, mfs = dgx.MatrixFree(ctx, mesh, -1, dt), dgx.MatrixFree(solo, mesh, -1, dt)
        opd, ops = (dgx.NavierStokesProjectionOperator(m, kind, degree, 100.0, 5e-3) for m in (mfd, mfs))
        ncomp = 3 if kind == dgx.OP_VELOCITY else 1
        dpc = ncomp * (degree + 1) ** 3
        N = mesh.n_cells() * dpc
        rng = np.random.default_rng(42)
        g = rng.standard_normal(N).astype(np.float64 if dt == dgx.F64 else np.float32)
        ge = rng.standard_normal(N).astype(g.dtype)
        lo, n = mfd.first_owned_cell() * dpc, mfd.n_owned_cells() * dpc
        xs, ys, xd, yd = ops.initialize_dof_vector(), ops.initialize_dof_vector(), opd.initialize_dof_vector(), opd.initialize_dof_vector()
        xs.from_host(g); xd.from_host(g[lo:lo + n])
        if kind == dgx.OP_VELOCITY:
            es, ed = ops.initialize_dof_vector(), opd.initialize_dof_vector()
            es.from_host(ge); ed.from_host(ge[lo:lo + n])
            ops.set_u_extr(es); opd.set_u_extr(ed)
        if "marching" in name:
            opd.set_variant(2)      # marching kernel or an error: at N > 1 this is the path with the ghost import in flight
        ops.vmult(ys, xs); opd.vmult(yd, xd)
        if "marching" in name:
            opd.vmult(yd, xd); opd.vmult(yd, xd)   # both mailbox slots, sequence numbers 2 and 3
        ref, got = ys.to_host()[lo:lo + n].astype(np.float64), yd.to_host().astype(np.float64)
        e = torch.tensor([float(np.sum((ref - got) ** 2)), float(np.sum(ref ** 2)), abs(xd.dot(xd) - float(g.astype(np.float64) @ g.astype(np.float64)))],
                         dtype=torch.float64, device="cuda")
        dist.all_reduce(e[:2])
        errs.append({"dtype": "f64" if dt == dgx.F64 else "f32", "rel_l2": float(torch.sqrt(e[0] / e[1])), "dot_abs_err": float(e[2])})
    report[name] = errs
# host entry point across ranks (dgx_op_vmult_host: rank-boundary cells first, ghost import, then the upload / compute / download pipeline)
os.environ["DGX_HOST_CHUNK_LOG2"] = "5"      # 32-cell bricks so that these small meshes are pipelined
for name, cells0, nref, periodic, degree in [("host entry Q4 periodic cube", [1, 1, 1], 4, [True] * 3, 4),
                                             ("host entry Q2 box with boundaries", [2, 1, 1], 3, [False] * 3, 2)]:
    mesh = dgx.Mesh(3, cells0, nref, [0, 0, 0], [float(c) for c in cells0], periodic)
    mfd, mfs = dgx.MatrixFree(ctx, mesh), dgx.MatrixFree(solo, mesh)
    opd, ops = (dgx.NavierStokesProjectionOperator(m, dgx.OP_PRESSURE, degree, 100.0, 5e-3) for m in (mfd, mfs))
    dpc = (degree + 1) ** 3
    g = np.random.default_rng(3).standard_normal(mesh.n_cells() * dpc)
    lo, n = mfd.first_owned_cell() * dpc, mfd.n_owned_cells() * dpc
    ups, grps = mfd.host_pipeline_plan(dpc * 8)
    xs, ys = ops.initialize_dof_vector(), ops.initialize_dof_vector()
    xs.from_host(g)
    ops.set_variant(1)
    ops.vmult(ys, xs)
    ref = ys.to_host()[lo:lo + n]
    hin, hout = np.ascontiguousarray(g[lo:lo + n]), np.zeros(n)
    worst = 0.0
    for rep in range(3):
        hout[:] = 0
        opd.vmult_host(hout, hin)
        worst = max(worst, float(np.sum((hout - ref) ** 2)))
    e = torch.tensor([worst, float(np.sum(ref ** 2))], dtype=torch.float64, device="cuda")
    dist.all_reduce(e)
    report[name] = [{"dtype": "f64", "rel_l2": float(torch.sqrt(e[0] / e[1])), "pipelined_groups": int(len(grps))}]
os.environ.pop("DGX_HOST_CHUNK_LOG2", None)

# lift / drag face integrals (compute_lift_and_drag, NS.cc:2642-2708) summed over the ranks == the single-GPU value
mesh = dgx.Mesh(3, [2, 1, 1], 2, [0, 0, 0], [2.0, 1.0, 1.0], [False] * 3)   # boundary id 4 = the lower z face
rng = np.random.default_rng(5)
ug, pg = rng.standard_normal(mesh.n_cells() * 3 * 64), rng.standard_normal(mesh.n_cells() * 27)
forces = {}
for tag, c in (("multi", ctx), ("solo", solo)):
    mf = dgx.MatrixFree(c, mesh)
    fc, nc = mf.first_owned_cell(), mf.n_owned_cells()
    u, p = mf.initialize_dof_vector(3, 3), mf.initialize_dof_vector(2, 1)
    u.from_host(ug[fc * 192:(fc + nc) * 192]); p.from_host(pg[fc * 27:(fc + nc) * 27])
    forces[tag] = dgx.compute_lift_and_drag(u, p, 100.0, 4)
report["lift_and_drag"] = [{"dtype": "f64", "multi": forces["multi"], "solo": forces["solo"],
                            "rel_l2": max(abs(a - b) / max(abs(b), 1e-300) for a, b in zip(forces["multi"], forces["solo"]))}]

# MG-preconditioned pressure solve across ranks (projection_step's solver): same iteration count and solution as on one GPU.
# Two layouts: one coarse cube per rank, and ONE cube cut into contiguous Morton chunks (p4est-style; the coarse levels are
# replicated on all ranks, DGX_MG_REPLICATE_MAX=64 forces two distributed levels + the gather).
mg_report = {}
for layout, cells0, nref, hi in (("row_of_cubes", [world, 1, 1], 3, [float(world), 1, 1]), ("one_cube_morton_chunks", [1, 1, 1], 4, [1.0, 1.0, 1.0])):
  for rep_max in ("default", "64"):
    if rep_max == "64":
        os.environ["DGX_MG_REPLICATE_MAX"] = "64"
    else:
        os.environ.pop("DGX_MG_REPLICATE_MAX", None)
    degree = 3
    mesh = dgx.Mesh(3, cells0, nref, [0, 0, 0], hi, [True] * 3)
    res = {}
    sols = {}
    for tag, c in (("multi", ctx), ("solo", solo)):
        mf = dgx.MatrixFree(c, mesh)
        op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, 5e-3)
        mg = dgx.Multigrid(c, mesh, degree, 100.0, 5e-3)
        dpc = (degree + 1) ** 3
        N = mesh.n_cells() * dpc
        g = np.sin(np.arange(N) * 3e-3) + 0.1 * np.random.default_rng(1).standard_normal(N)
        g -= g.mean()
        lo, n = mf.first_owned_cell() * dpc, mf.n_owned_cells() * dpc
        x, b = op.initialize_dof_vector(), op.initialize_dof_vector()
        b.from_host(g[lo:lo + n])
        tol = 1e-8 * float(np.linalg.norm(g))
        its_all = []
        for rep in range(3):          # 3 solves: eager cycle, captured cycle, replayed graph
            mg.setup(tol, 10000)
            x.set(0.0)
            its, r = dgx.solve_cg(op, x, b, mg, 10000, tol)
            its_all.append(its)
        res[tag] = its_all
        sols[tag] = (x.to_host(), lo, n)
    xm, lo, n = sols["multi"]
    xs = sols["solo"][0][lo:lo + n]
    e = torch.tensor([float(np.sum((xm - xs) ** 2)), float(np.sum(xs ** 2))], dtype=torch.float64, device="cuda")
    dist.all_reduce(e)
    mg_report[f"{layout}/replicate_max={rep_max}"] = {"iterations_multi": res["multi"], "iterations_solo": res["solo"], "solution_rel_l2": float(torch.sqrt(e[0] / e[1]))}
os.environ.pop("DGX_MG_REPLICATE_MAX", None)

# one TR-BDF2 step of the Taylor-Green vortex across ranks: Krylov iteration counts and velocity equal to the single-GPU run
from dgx.navier_stokes import NavierStokesProjection
L, R, dp = 2 * np.pi, 3, 2
mesh = dgx.Mesh(3, [1, 1, 1], R, [0, 0, 0], [L, L, L], [True] * 3)
h = L / 2 ** R
step = {}
for tag, c in (("multi", ctx), ("solo", solo)):
    ns = NavierStokesProjection(c, mesh, dp, 1600.0, 0.2 * h / (dp + 2))
    nv, npn = (dp + 2) ** 3 * 3, (dp + 1) ** 3
    rng = np.random.default_rng(9)
    ug = np.sin(np.arange(mesh.n_cells() * nv) * 1e-2) + 0.05 * rng.standard_normal(mesh.n_cells() * nv)
    pg = np.cos(np.arange(mesh.n_cells() * npn) * 2e-2)
    fc, nc = ns.mf.first_owned_cell(), ns.mf.n_owned_cells()
    ns.u_n.from_host(ug[fc * nv:(fc + nc) * nv]); ns.u_n_minus_1.from_host(ug[fc * nv:(fc + nc) * nv]); ns.pres_n.from_host(pg[fc * npn:(fc + nc) * npn])
    for _ in range(2):
        vmax = ns.advance()
    step[tag] = (list(ns.iterations), ns.u_n.to_host(), fc * nv, nc * nv, vmax)
um, lo, n = step["multi"][1], step["multi"][2], step["multi"][3]
us = step["solo"][1][lo:lo + n]
e = torch.tensor([float(np.sum((um - us) ** 2)), float(np.sum(us ** 2))], dtype=torch.float64, device="cuda")
dist.all_reduce(e)
step_report = {"iterations_multi": step["multi"][0], "iterations_solo": step["solo"][0], "u_rel_l2": float(torch.sqrt(e[0] / e[1])),
               "max_velocity": [step["multi"][4], step["solo"][4]]}

cells0, nref, degree = [world, 1, 1], 3, 3
mesh = dgx.Mesh(3, cells0, nref, [0, 0, 0], [float(world), 1, 1], [True] * 3)
res = {}
sols = {}
for tag, c in (("multi", ctx), ("solo", solo)):
    mf = dgx.MatrixFree(c, mesh)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, 5e-3)
    mg = dgx.Multigrid(c, mesh, degree, 100.0, 5e-3)
    dpc = (degree + 1) ** 3
    N = mesh.n_cells() * dpc
    g = np.sin(np.arange(N) * 3e-3) + 0.1 * np.random.default_rng(1).standard_normal(N)
    g -= g.mean()
    lo, n = mf.first_owned_cell() * dpc, mf.n_owned_cells() * dpc
    x, b = op.initialize_dof_vector(), op.initialize_dof_vector()
    b.from_host(g[lo:lo + n])
    tol = 1e-8 * float(np.linalg.norm(g))
    mg.setup(tol, 10000)
    its, r = dgx.solve_cg(op, x, b, mg, 10000, tol)
    res[tag] = its
    sols[tag] = (x.to_host(), lo, n)
xm, lo, n = sols["multi"]
xs = sols["solo"][0][lo:lo + n]
e = torch.tensor([float(np.sum((xm - xs) ** 2)), float(np.sum(xs ** 2))], dtype=torch.float64, device="cuda")
dist.all_reduce(e)
mg_rel = float(torch.sqrt(e[0] / e[1]))
if rank == 0:
    print(json.dumps({"world": world, "checks": report, "mg_cg": {"iterations_multi": res["multi"], "iterations_solo": res["solo"], "solution_rel_l2": mg_rel},
                      "mg_cg_layouts": mg_report, "trbdf2_step": step_report}))
    if abs(res["multi"] - res["solo"]) > 1 or mg_rel > 1e-5:
        report["mg_cg"] = [{"dtype": "f64", "rel_l2": 1.0}]
    for k, v in mg_report.items():
        if max(abs(a - b) for a, b in zip(v["iterations_multi"], v["iterations_solo"])) > 1 or v["solution_rel_l2"] > 1e-5:
            report["mg_cg " + k] = [{"dtype": "f64", "rel_l2": 1.0}]
    if any(abs(a[1] - b[1]) > 1 for a, b in zip(step_report["iterations_multi"], step_report["iterations_solo"])) or step_report["u_rel_l2"] > 1e-8:
        report["trbdf2_step"] = [{"dtype": "f64", "rel_l2": 1.0}]
    bad = [k for k, v in report.items() for x in v if x["rel_l2"] > (1e-12 if x["dtype"] == "f64" else 1e-5)]
    print("MULTI-GPU CHECK", "FAILED " + str(bad) if bad else "PASSED")
dist.destroy_process_group()
