#!/usr/bin/env python
"""Times every generation of the marching kernel (DGX_MARCH = v1 | v2 | v2eo) on the headline workload (3-D Q4 fp64,
2^R cells per direction, periodic) and checks each against the generic kernel.   python tools/march_variants.py [R] [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "code-gallery_b200")]
import numpy as np

import dgx

R = int(sys.argv[1]) if len(sys.argv) > 1 else 7
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
gens = sys.argv[3].split(",") if len(sys.argv) > 3 else ["v1", "v2", "v2eo"]
ctx = dgx.Context(0)
mesh = dgx.Mesh(3, [1, 1, 1], R, [0, 0, 0], [1, 1, 1], [True] * 3)
mf = dgx.MatrixFree(ctx, mesh)
op = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, 4, 100.0, 5e-3)
src, dst, ref = (op.initialize_dof_vector() for _ in range(3))
n = src.locally_owned_size()
rng = np.random.default_rng(3)
h = np.empty(n)
for o in range(0, n, 1 << 24):
    m = min(1 << 24, n - o)
    h[o:o + m] = np.sin(np.arange(o, o + m) * 1e-3) + 1e-3 * rng.uniform(-1, 1, m)
src.from_host(h)
op.set_variant(1)
op.vmult(ref, src)
refn = ref.l2_norm()
op.set_variant(2)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
for gen in gens:
    os.environ["DGX_MARCH"] = gen
    for _ in range(3):
        op.vmult(dst, src)
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(steps):
        op.vmult(dst, src)
    ms = ctx.timer_stop() / steps
    dst.add(-1.0, ref)
    err = dst.l2_norm() / refn
    print(json.dumps({"gen": gen, "refine": R, "ms": ms, "gdofs": n / ms / 1e6, "hbm_frac": n * 16 / (ms * 1e-3) / 1e9 / peak, "rel_l2_vs_generic": err}), flush=True)
ctx.close()
