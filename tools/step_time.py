#!/usr/bin/env python
"""One TR-BDF2 step of the periodic 3-D Taylor-Green vortex (same set-up as bench.py's trbdf2_step report).
usage: step_time.py REFINE [STEPS] [direct]   -- run on the B200, plain or under an ncu launch list."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "code-gallery_b200")]
import numpy as np
import dgx
from dgx.navier_stokes import NavierStokesProjection
R = int(sys.argv[1]) if len(sys.argv) > 1 else 5
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
direct = len(sys.argv) > 3 and sys.argv[3] == "direct"
dp, L = 2, 2 * np.pi
ctx = dgx.Context(0)
mesh = dgx.Mesh(3, [1, 1, 1], R, [0, 0, 0], [L, L, L], [True, True, True])
h = L / 2 ** R
dt = 0.2 * h / (dp + 2)
ns = NavierStokesProjection(ctx, mesh, dp, 1600.0, dt, direct_mass_inverse=direct)
def nodes(n):
    from numpy.polynomial import legendre as leg
    if n == 2:
        return np.array([0.0, 1.0])
    c = np.zeros(n); c[-1] = 1.0
    return np.concatenate(([0.0], (np.sort(leg.legroots(leg.legder(c))) + 1) / 2, [1.0]))
coords = mesh.cell_coords().astype(np.float64)
def field(n, f):
    x1 = nodes(n)
    X = (coords[:, None, None, None, 0] + x1[None, None, None, :]) * h
    Y = (coords[:, None, None, None, 1] + x1[None, None, :, None]) * h
    Z = (coords[:, None, None, None, 2] + x1[None, :, None, None]) * h
    return f(X + 0 * Y + 0 * Z, Y + 0 * X + 0 * Z, Z + 0 * X + 0 * Y)
u = np.stack([field(dp + 2, lambda x, y, z: np.sin(x) * np.cos(y) * np.cos(z)), field(dp + 2, lambda x, y, z: -np.cos(x) * np.sin(y) * np.cos(z)),
              field(dp + 2, lambda x, y, z: 0 * x)], axis=1).reshape(-1)
p = field(dp + 1, lambda x, y, z: (np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * z) + 2) / 16).reshape(-1)
ns.u_n.from_host(u); ns.u_n_minus_1.from_host(u); ns.pres_n.from_host(p)
out = []
for s in range(steps):
    ns.iterations.clear()
    ctx.synchronize()
    l0 = dgx.kernel_launch_count()
    ctx.timer_start()
    vmax = ns.advance()
    ms = ctx.timer_stop()
    out.append({"step": s, "ms": round(ms, 3), "max_velocity": vmax, "launches": dgx.kernel_launch_count() - l0,
                "iterations": [(k, v) for k, v in ns.iterations]})
print(json.dumps({"cells": f"{2**R}^3", "velocity_dofs": int(u.size), "direct_mass_inverse": direct, "steps": out}))
