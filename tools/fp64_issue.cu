// Microbenchmark: FP64 FMA throughput of one SM as a function of resident warps and independent chains per warp.
// Answers "how many warps / how much ILP does the FP64 pipe of a B200 SM need?" (the marching kernel runs 12 warps per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu && ./fp64_issue
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double *out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
void run(int warps, double *d) {
  const int iters = 4096, blocks = 148;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP><<<blocks, warps * 32>>>(d, 16, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k<ILP><<<blocks, warps * 32>>>(d, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)blocks * warps * 32 * iters * 8 * ILP;
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("{\"warps_per_sm\": %d, \"ilp\": %d, \"tflops\": %.2f, \"dfma_per_clk_per_sm\": %.2f, \"cycles_per_dfma_per_warp\": %.2f}\n", warps, ILP,
         2 * fma / (ms * 1e-3) / 1e12, fma / blocks / cycles, cycles / ((double)iters * 8 * ILP));
}

int main() {
  double *d;
  cudaMalloc(&d, 148 * 1024 * sizeof(double));
  for (int w : {4, 8, 12, 16, 24, 32}) {
    run<1>(w, d); run<2>(w, d); run<4>(w, d); run<8>(w, d); run<16>(w, d);
  }
  return 0;
}
