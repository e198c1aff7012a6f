#!/usr/bin/env python
"""Throughput of the metric-path pressure operator (csrc/pressure_general.cu) on a deformed periodic 3-D mesh:
python tools/time_general.py [refine=6] [reps=20] -> one JSON line per (degree, dtype): GDoF/s, algorithmic bytes per DoF
(src + dst + metric tables), achieved GB/s and fraction of the measured HBM peak."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "code-gallery_b200")]
import dgx

def wavy(amp):
    def fn(X):
        x = np.array(X, dtype=np.float64); out = x.copy()
        out[..., 0] = x[..., 0] + amp * np.sin(2 * np.pi * x[..., 1]) * (1 + 0.5 * np.cos(2 * np.pi * x[..., 0]))
        out[..., 1] = x[..., 1] + amp * np.sin(2 * np.pi * x[..., 0]) + 0.5 * amp * np.sin(2 * np.pi * x[..., 2])
        out[..., 2] = x[..., 2] + amp * np.sin(2 * np.pi * (x[..., 0] + x[..., 1]))
        return out
    return fn

R = int(sys.argv[1]) if len(sys.argv) > 1 else 6
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
degrees = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2, 3, 4]
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6537.3) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6537.3
ctx = dgx.Context(0)
for degree in degrees:
    for dtype, name, es in ((dgx.F64, "f64", 8), (dgx.F32, "f32", 4)):
        r = R if degree <= 3 else min(R, 6)
        mesh = dgx.Mesh(3, [1, 1, 1], r, [0, 0, 0], [1, 1, 1], [True] * 3)
        mesh.set_vertices(r, wavy(0.2 / (1 << r) * 4)(np.stack(np.meshgrid(*[np.arange((1 << r) + 1) / (1 << r)] * 3, indexing="ij")[::-1], axis=-1)))
        mf = dgx.MatrixFree(ctx, mesh, dtype=dtype)
        op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, degree, 100.0, 5e-3)
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        n = src.locally_owned_size()
        src.from_host(np.sin(1e-3 * np.arange(n)))
        for _ in range(3): op.vmult(dst, src)
        ctx.synchronize(); ctx.timer_start()
        for _ in range(reps): op.vmult(dst, src)
        ms = ctx.timer_stop() / reps
        mb = mf.general_metric_bytes(degree)
        bpd = 2 * es + mb / n
        print(json.dumps({"degree": degree, "dtype": name, "cells": f"{1 << r}^3", "dofs": n, "ms": round(ms, 4), "gdofs": round(n / ms * 1e-6, 2),
                          "bytes_per_dof": round(bpd, 1), "GBs": round(bpd * n / ms * 1e-6, 1), "hbm_frac": round(bpd * n / ms * 1e-6 / peak, 3),
                          "checksum": float(np.abs(dst.to_host()[:1000]).sum())}), flush=True)
        del op, src, dst, mf, mesh
