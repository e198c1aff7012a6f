#!/usr/bin/env python
"""Q2 pressure vmult (256^3 cells) through the marching kernel selected by DGX_MARCH3_Q2, checked against the generic kernel.
   python tools/q2_variants.py [refine] [steps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "code-gallery_b200")]
import numpy as np
import dgx
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
DEG = int(sys.argv[3]) if len(sys.argv) > 3 else 2
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
ctx = dgx.Context(0)
mesh = dgx.Mesh(3, [1, 1, 1], R, [0, 0, 0], [1, 1, 1], [True] * 3)
for dt, name, nb in ((dgx.F64, "f64", 8), (dgx.F32, "f32", 4)):
    mf = dgx.MatrixFree(ctx, mesh, -1, dt)
    op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, DEG, 100.0, 5e-3)
    src, dst, ref = (op.initialize_dof_vector() for _ in range(3))
    n = src.locally_owned_size()
    src.from_host((np.sin(np.arange(n) * 1e-3) + 1e-3 * np.random.default_rng(3).uniform(-1, 1, n)).astype(np.float64 if dt == dgx.F64 else np.float32))
    op.set_variant(1); op.vmult(ref, src); refn = ref.l2_norm()
    op.set_variant(2)
    for _ in range(3): op.vmult(dst, src)
    ctx.synchronize(); ctx.timer_start()
    for _ in range(steps): op.vmult(dst, src)
    ms = ctx.timer_stop() / steps
    dst.add(-1.0, ref)
    print(json.dumps({"q2": os.environ.get("DGX_MARCH3_Q2", "0"), "dtype": name, "ms": round(ms, 4), "gdofs": round(n / ms / 1e6, 1), "hbm_frac": round(2 * nb * n / (ms * 1e-3) / 1e9 / peak, 3), "rel_l2_vs_generic": dst.l2_norm() / refn}), flush=True)
    del src, dst, ref, op, mf
ctx.close()
