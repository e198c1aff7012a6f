// Velocity operator (NS_stage == 1, NS.cc:1399-1406): placeholder until the kernels land.
#include "pressure_op.cuh"

namespace dgx {
int velocity_apply(dgx_op *, dgx_vec *, const dgx_vec *, bool) { set_error("velocity operator not implemented yet"); return DGX_ERR_UNSUPPORTED; }
int velocity_diagonal(dgx_op *, dgx_vec *) { set_error("velocity diagonal not implemented yet"); return DGX_ERR_UNSUPPORTED; }
}  // namespace dgx
