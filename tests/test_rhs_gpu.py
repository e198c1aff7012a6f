"""GPU parity of the right-hand-side forms (vmult_rhs_velocity NS.cc:788, vmult_rhs_pressure NS.cc:915,
vmult_grad_p_projection NS.cc:1363) through the C ABI vs the oracle's face-centric restatement.  fp64, 1e-12."""
import numpy as np
import pytest

from oracle.rhs import RhsForms
from util import make_meshes, rel_l2, smooth_plus_noise

pytestmark = pytest.mark.gpu

CASES = [
    (2, [4, 2], 2, [False, False], [1, 2]),             # channel: inflow id 0, outflow id 1, walls
    (2, [1, 1], 3, [True, True], [1, 3]),
    (3, [1, 1, 1], 2, [True, True, True], [1, 2]),
    (3, [2, 1, 1], 1, [False, False, False], [1, 2]),
    (3, [1, 2, 1], 1, [True, False, True], [1]),
]


def _fields(om, R, seed):
    dim = om.dim
    u = [smooth_plus_noise(om, R.vel_q0.nodes, seed=seed + k, ncomp=dim, noise=0.4) + 0.2 * k for k in range(3)]
    p = [smooth_plus_noise(om, R.pre_q0.nodes, seed=seed + 10 + k, ncomp=1, noise=0.3) for k in range(2)]
    return u, p


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}d-{c[1]}-r{c[2]}-{''.join('p' if x else 'b' for x in c[3])}")
def test_rhs_forms_match_oracle(dgx, gpu_ctx, case):
    dim, cells0, nref, periodic, degrees_p = case
    Re, dt = 100.0, 5e-3
    for dp in degrees_p:
        gm, om = make_meshes(dgx, dim, cells0, nref, periodic, hi=[2.0, 4.1, 0.9][:dim])
        mf = dgx.MatrixFree(gpu_ctx, gm)
        R = RhsForms(om, dp, Re, dt)
        u, p = _fields(om, R, 3)
        vop = dgx.NavierStokesProjectionOperator(mf, dgx.OP_VELOCITY, dp + 1, Re, dt)
        pop = dgx.NavierStokesProjectionOperator(mf, dgx.OP_PRESSURE, dp, Re, dt)
        mop = dgx.NavierStokesProjectionOperator(mf, dgx.OP_MASS, dp + 1, Re, dt)
        vu = [vop.initialize_dof_vector() for _ in u]
        vp = [pop.initialize_dof_vector() for _ in p]
        for a, b in zip(vu + vp, u + p):
            a.from_host(b)
        dv, dpv = vop.initialize_dof_vector(), pop.initialize_dof_vector()
        for stage in (1, 2):
            vop.set_TR_BDF2_stage(stage)
            pop.set_TR_BDF2_stage(stage)
            if stage == 1:
                vop.vmult_rhs_velocity(dv, [vu[0], vu[1], vp[0]])
                ref = R.rhs_velocity(1, [u[0], u[1], p[0]])
            else:
                vop.vmult_rhs_velocity(dv, [vu[0], vu[1], vp[0], vu[2]])
                ref = R.rhs_velocity(2, [u[0], u[1], p[0], u[2]])
            assert rel_l2(dv.to_host(), ref) < 1e-12, ("velocity", dp, stage)
            pop.vmult_rhs_pressure(dpv, [vu[0], vp[1]])
            assert rel_l2(dpv.to_host(), R.rhs_pressure(stage, [u[0], p[1]])) < 1e-12, ("pressure", dp, stage)
        mop.vmult_grad_p_projection(dv, vp[0])
        assert rel_l2(dv.to_host(), R.rhs_grad_p(p[0])) < 1e-12, ("grad_p", dp)
        with pytest.raises(dgx.DgxError):
            vop.vmult_rhs_velocity(dv, [vu[0], vu[1]])
