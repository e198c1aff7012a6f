#!/usr/bin/env python
"""Device time of the fp32 level operator and of one Chebyshev smoother application (degree 3 = 3 fused operator
launches with a non-zero initial guess) on a periodic cube.  usage: time_cheb.py REFINE DEGREE [f32|f64]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "code-gallery_b200")]
import numpy as np
import dgx
R = int(sys.argv[1]) if len(sys.argv) > 1 else 6
p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dt = dgx.F64 if len(sys.argv) > 3 and sys.argv[3] == "f64" else dgx.F32
ctx = dgx.Context(0)
mesh = dgx.Mesh(3, [1, 1, 1], R, [0, 0, 0], [1, 1, 1], [True] * 3)
mf = dgx.MatrixFree(ctx, mesh, -1, dt)
op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, p, 100.0, 5e-3)
op.compute_diagonal()
ch = dgx.PreconditionChebyshev(op)
est = ch.estimate_eigenvalues()
x, b, y = (op.initialize_dof_vector() for _ in range(3))
n = x.locally_owned_size()
b.from_host(np.sin(np.arange(n) * 7e-4).astype(np.float32 if dt == dgx.F32 else np.float64))
def t(f, reps=5):
    f(); ctx.synchronize(); ctx.timer_start()
    for _ in range(reps): f()
    return ctx.timer_stop() / reps
ms_v = t(lambda: op.vmult(y, b))
ms_s = t(lambda: ch.step(x, b))
ms_0 = t(lambda: ch.vmult(x, b))
print(json.dumps({"dofs": n, "estimates": est, "vmult ms": round(ms_v, 4), "vmult GDoF/s": round(n / ms_v / 1e6, 1),
                  "cheb step (3 fused launches) ms": round(ms_s, 4), "per fused launch GDoF/s": round(3 * n / ms_s / 1e6, 1),
                  "cheb vmult (zero guess) ms": round(ms_0, 4)}))
