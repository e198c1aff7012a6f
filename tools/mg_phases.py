#!/usr/bin/env python
"""Phase timings of the MG-preconditioned pressure solve on N GPUs (torchrun) or one: setup (diagonals + eigenvalue estimates)
and solve, with recomputed and cached estimates.   [torchrun ...] tools/mg_phases.py [refine_per_gpu] [degree]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "code-gallery_b200")]
import numpy as np
import torch
import dgx
from bench import layout_for

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
nccl_id = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(dgx.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(dgx.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    nccl_id = bytes(idt.cpu().numpy().tobytes())
ctx = dgx.Context(local, rank, world, nccl_id)
R0 = int(sys.argv[1]) if len(sys.argv) > 1 else 6
p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cells0, R = layout_for(world, R0)
mesh = dgx.Mesh(3, cells0, R, [0, 0, 0], [float(c) for c in cells0], [True] * 3)
mf = dgx.MatrixFree(ctx, mesh)
op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, p, 100.0, 5e-3)
mg = dgx.Multigrid(ctx, mesh, p, 100.0, 5e-3)
x, b = op.initialize_dof_vector(), op.initialize_dof_vector()
n = b.locally_owned_size()
rhs = np.sin(np.arange(n) * 7e-4) + 1e-2 * np.random.default_rng(5).uniform(-1, 1, n)
rhs -= rhs.mean()
b.from_host(rhs)
tol = 1e-8 * b.l2_norm()
rows = []
for rep, cached in enumerate([False, False, False, True, True, True]):
    mg.cache_estimates(cached)
    x.set(0.0)
    ctx.synchronize()
    l0 = dgx.kernel_launch_count()
    t0 = time.perf_counter()
    mg.setup(tol, 10000)
    ctx.synchronize()
    t1 = time.perf_counter()
    its, res = dgx.solve_cg(op, x, b, mg, 10000, tol)
    ctx.synchronize()
    t2 = time.perf_counter()
    rows.append({"rep": rep, "cached": cached, "setup_ms": (t1 - t0) * 1e3, "solve_ms": (t2 - t1) * 1e3, "its": its, "launches": dgx.kernel_launch_count() - l0})
if rank == 0:
    print(json.dumps({"world": world, "cells": [c << R for c in cells0], "degree": p, "levels": mg.n_levels(), "phases": rows}))
if world > 1:
    dist.destroy_process_group()
