#!/usr/bin/env python
"""BASELINE.json config 3: DG SIPG vmult for Q1..Q8 in fp64 and fp32 on periodic cubes (meshes of SURVEY.md 8d C3),
per-degree roofline characterisation.  Prints one JSON line per (degree, dtype); run on the B200."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "code-gallery_b200")]
import numpy as np

import dgx

REFINE = {1: 8, 2: 8, 3: 7, 4: 7, 5: 7, 6: 7, 7: 6, 8: 6}
hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
ctx = dgx.Context(0)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for p in range(1, 9):
    mesh = dgx.Mesh(3, [1, 1, 1], REFINE[p], [0, 0, 0], [1, 1, 1], [True] * 3)
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
        print(json.dumps({"degree": p, "dtype": name, "cells": f"{2**REFINE[p]}^3", "dofs": n, "ms": round(ms, 4),
                          "gdofs": round(n / (ms * 1e-3) / 1e9, 2), "algorithmic_GBs": round(gb, 1), "hbm_frac_of_measured": round(gb / hbm, 3)}), flush=True)
        del src, dst, op, mf
ctx.close()
