import json, os, sys
import numpy as np
sys.path[:0] = ['.', 'code-gallery_b200']
import dgx
ctx = dgx.Context(0)
for p, r in ((1, 8),):
    mesh = dgx.Mesh(3, [1, 1, 1], r, [0, 0, 0], [1, 1, 1], [True] * 3)
    for dt, name, nb in ((dgx.F64, "f64", 8), (dgx.F32, "f32", 4)):
        mf = dgx.MatrixFree(ctx, mesh, -1, dt)
        op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, p, 100.0, 5e-3)
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        n = src.locally_owned_size()
        src.from_host(np.sin(np.arange(n, dtype=np.float64) * 1e-3))
        for _ in range(3): op.vmult(dst, src)
        ctx.synchronize(); ctx.timer_start()
        for _ in range(10): op.vmult(dst, src)
        ms = ctx.timer_stop() / 10
        print(json.dumps({"degree": p, "dtype": name, "ms": round(ms, 4), "gdofs": round(n / ms * 1e-6, 2), "hbm_frac": round(2 * nb * n / ms * 1e-6 / 6537.3, 3), "env": os.environ.get("DGX_NO_Q1_KERNEL", "")}), flush=True)
