#!/usr/bin/env python
"""Device time of the individual operators of the TR-BDF2 step (3-D periodic cube, velocity Q3 / pressure Q2)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "code-gallery_b200")]
import numpy as np
import dgx
R = int(sys.argv[1]) if len(sys.argv) > 1 else 5
only = sys.argv[2] if len(sys.argv) > 2 else None
ctx = dgx.Context(0)
mesh = dgx.Mesh(3, [1, 1, 1], R, [0, 0, 0], [1, 1, 1], [True] * 3)
mf = dgx.MatrixFree(ctx, mesh)
dp = 2
vop = dgx.NavierStokesProjectionOperator(mf, dgx.OP_VELOCITY, dp + 1, 1600.0, 1e-2)
pop = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, dp, 1600.0, 1e-2)
mop = dgx.NavierStokesProjectionOperator(mf, dgx.OP_MASS, dp + 1, 1600.0, 1e-2)
u, v, w = (vop.initialize_dof_vector() for _ in range(3))
p, q = pop.initialize_dof_vector(), pop.initialize_dof_vector()
n = u.locally_owned_size()
u.from_host(np.sin(np.arange(n) * 1e-3)); v.from_host(np.cos(np.arange(n) * 2e-3))
p.from_host(np.sin(np.arange(p.locally_owned_size()) * 1e-3))
vop.set_u_extr(v)
def t(f, reps=5):
    f(); ctx.synchronize(); ctx.timer_start()
    for _ in range(reps): f()
    return ctx.timer_stop() / reps
res = {"velocity_dofs": n}
ops = {"velocity vmult": lambda: vop.vmult(w, u), "velocity rhs": lambda: vop.vmult_rhs_velocity(w, [u, v, p]),
       "pressure rhs": lambda: pop.vmult_rhs_pressure(q, [u, p]), "grad_p rhs": lambda: mop.vmult_grad_p_projection(w, p),
       "mass vmult": lambda: mop.vmult(w, u), "pressure vmult": lambda: pop.vmult(q, p), "velocity diagonal": lambda: vop.compute_diagonal()}
for k, f in ops.items():
    if only and only not in k: continue
    ms = t(f, 2 if "diagonal" in k else 5)
    res[k] = {"ms": round(ms, 3), "GDoF/s": round(n / ms / 1e6, 2)}
print(json.dumps(res, indent=1))
