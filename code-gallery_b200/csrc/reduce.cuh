// Warp-level sum that is also valid in the LAST warp of a block whose size is not a multiple of 32 (the cell-batch kernels
// run CPB * n^(dim-1) threads: 175, 180, 189 ...).  __shfl_down_sync with a full mask names lanes that do not exist there and
// returns undefined values; here the mask is the set of lanes that exist and a lane only adds what an existing lane sent.
#pragma once
namespace dgx {
__device__ __forceinline__ double warp_sum_existing(double v) {
  const unsigned m = __activemask();          // all existing lanes of this warp are converged at the call sites
  const int nl = __popc(m), lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_down_sync(m, v, o);
    if (lane + o < nl) v += t;
  }
  return v;
}
}  // namespace dgx
