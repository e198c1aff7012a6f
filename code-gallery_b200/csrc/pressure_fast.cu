// Blocked fast path of the pressure operator for large 3-D meshes.
#include "pressure_op.cuh"

namespace dgx {
int pressure_apply_fast(dgx_op *, dgx_vec *, const dgx_vec *, bool) { return DGX_ERR_UNSUPPORTED; }
}  // namespace dgx
