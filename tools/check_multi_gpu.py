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
for name, cells0, nref, periodic, kind, degree in [
        ("pressure Q4 periodic slab (marching kernel)", [world, 1, 1], 3, [True] * 3, dgx.OP_PRESSURE, 4),
        ("pressure Q3 periodic cube (Morton chunks)", [1, 1, 1], 3, [True] * 3, dgx.OP_PRESSURE, 3),
        ("pressure Q2 box with boundaries", [2, 1, 1], 2, [False] * 3, dgx.OP_PRESSURE, 2),
        ("velocity Q3 periodic cube", [1, 1, 1], 2, [True] * 3, dgx.OP_VELOCITY, 3)]:
    hi = [float(c) for c in cells0]
    mesh = dgx.Mesh(3, cells0, nref, [0, 0, 0], hi, periodic)
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
        ops.vmult(ys, xs); opd.vmult(yd, xd)
        ref, got = ys.to_host()[lo:lo + n].astype(np.float64), yd.to_host().astype(np.float64)
        e = torch.tensor([float(np.sum((ref - got) ** 2)), float(np.sum(ref ** 2)), abs(xd.dot(xd) - float(g.astype(np.float64) @ g.astype(np.float64)))],
                         dtype=torch.float64, device="cuda")
        dist.all_reduce(e[:2])
        errs.append({"dtype": "f64" if dt == dgx.F64 else "f32", "rel_l2": float(torch.sqrt(e[0] / e[1])), "dot_abs_err": float(e[2])})
    report[name] = errs
# MG-preconditioned pressure solve across ranks (projection_step's solver): same iteration count and solution as on one GPU
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
    print(json.dumps({"world": world, "checks": report, "mg_cg": {"iterations_multi": res["multi"], "iterations_solo": res["solo"], "solution_rel_l2": mg_rel}}))
    if abs(res["multi"] - res["solo"]) > 1 or mg_rel > 1e-5:
        report["mg_cg"] = [{"dtype": "f64", "rel_l2": 1.0}]
    bad = [k for k, v in report.items() for x in v if x["rel_l2"] > (1e-12 if x["dtype"] == "f64" else 1e-5)]
    print("MULTI-GPU CHECK", "FAILED " + str(bad) if bad else "PASSED")
dist.destroy_process_group()
