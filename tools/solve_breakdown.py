#!/usr/bin/env python
"""Where the MG-CG pressure solve spends its time (device timers around the pieces); run on the B200."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "code-gallery_b200")]
import numpy as np
import dgx
R, p = int(sys.argv[1]) if len(sys.argv) > 1 else 6, int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = dgx.Context(0)
mesh = dgx.Mesh(3, [1, 1, 1], R, [0, 0, 0], [1, 1, 1], [True] * 3)
mf = dgx.MatrixFree(ctx, mesh)
op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, p, 100.0, 5e-3)
mg = dgx.Multigrid(ctx, mesh, p, 100.0, 5e-3)
x, b, y = (op.initialize_dof_vector() for _ in range(3))
n = b.locally_owned_size()
rhs = np.sin(np.arange(n) * 7e-4) + 1e-2 * np.random.default_rng(5).uniform(-1, 1, n); rhs -= rhs.mean()
b.from_host(rhs)
tol = 1e-8 * float(np.linalg.norm(rhs))
def timed(f, reps=3):
    f(); ctx.synchronize()
    t0 = time.perf_counter(); ctx.timer_start()
    for _ in range(reps): f()
    ms = ctx.timer_stop() / reps
    return round(ms, 3), round((time.perf_counter() - t0) * 1e3 / reps, 3)
out = {"dofs": n}
out["setup (device ms, wall ms)"] = timed(lambda: mg.setup(tol, 10000))
out["fine fp64 vmult"] = timed(lambda: op.vmult(y, b), 10)
out["V-cycle (PreconditionMG::vmult)"] = timed(lambda: mg.vmult(y, b), 5)
out["dot"] = timed(lambda: b.dot(y), 10)
out["add"] = timed(lambda: y.add(0.5, b), 10)
def solve():
    x.set(0.0)
    return dgx.solve_cg(op, x, b, mg, 10000, tol)
out["solve only"] = timed(solve, 2)
out["iterations"] = solve()[0]
print(json.dumps(out, indent=1))
