// Microbenchmark for the roofline denominators MEASURED_PEAKS.json does not hold:
// FP64 / FP32 FMA issue peak, FP64 tensor (DMMA m8n8k4) peak, both mixed, shared-memory LDS.64 rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu ; run on the B200.
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int ILP>
__global__ void fma_kernel(T *out, T a, T b, int iters) {
  T acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = (T)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = acc[i] * a + b;
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == (T)123456.789) out[0] = s;
}

// DMMA m8n8k4: D(8x8) += A(8x4) B(4x8); per thread 1 a, 1 b, 2 c
template <int ILP, bool MIX>
__global__ void dmma_kernel(double *out, double a, double b, int iters) {
  double c0[ILP], c1[ILP], f[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x; c1[i] = i; f[i] = i + threadIdx.x; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
      if (MIX) { f[i] = f[i] * a + b; f[i] = f[i] * a + b; f[i] = f[i] * a + b; f[i] = f[i] * a + b; }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i] + f[i];
  if (s == 123456.789) out[0] = s;
}

// DMMA m16n8k16 (sm_90+ shape): D(16x8) += A(16x16) B(16x8); per thread 8 a, 4 b, 4 c -- 2048 FMAs per instruction, 8 x m8n8k4
template <int ILP>
__global__ void dmma16_kernel(double *out, double a, double b, int iters) {
  double c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + i + j;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%4,%4,%4,%4,%4,%4,%4}, {%5,%5,%5,%5}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123456.789) out[0] = s;
}

__global__ void lds_kernel(double *out, int iters) {
  __shared__ double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double s = 0;
  int idx = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) s += sm[(idx + k * 256) & 4095];
    idx = (idx + 1) & 4095;
  }
  if (s == 123456.789) out[0] = s;
}

// streaming read of `n` double2 elements, repeated `reps` times (L2-resident when the buffer is < 126 MB)
__global__ void read_kernel(const double2 *__restrict__ in, size_t n, int reps, double *out) {
  double s = 0;
  for (int r = 0; r < reps; ++r)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      double2 v = __ldcg(in + ((i + (size_t)r * 977) % n));
      s += v.x + v.y;
    }
  if (s == 123456.789) out[0] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  double *out; cudaMalloc(&out, 64);
  const int blocks = sms * 4, threads = 512, iters = 20000;
  {
    float ms = time_ms([&] { fma_kernel<double, 8><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters); });
    printf(", \"fp64_fma_tflops\": %.2f", 2.0 * blocks * threads * 8.0 * iters / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { fma_kernel<float, 8><<<blocks, threads>>>((float *)out, 1.0000001f, 1e-9f, iters); });
    printf(", \"fp32_fma_tflops\": %.2f", 2.0 * blocks * threads * 8.0 * iters / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { dmma_kernel<8, false><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters / 4); });
    printf(", \"fp64_dmma_tflops\": %.2f", 2.0 * 256.0 * blocks * (threads / 32) * 8.0 * (iters / 4) / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { dmma16_kernel<4><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters / 16); });
    printf(", \"fp64_dmma_m16n8k16_tflops\": %.2f", 2.0 * 2048.0 * blocks * (threads / 32) * 4.0 * (iters / 16) / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { dmma_kernel<8, true><<<blocks, threads>>>(out, 1.0000001, 1e-9, iters / 4); });
    double dm = 2.0 * 256.0 * blocks * (threads / 32) * 8.0 * (iters / 4), fm = 2.0 * blocks * threads * 8.0 * 4.0 * (iters / 4);
    printf(", \"mixed_dmma_tflops\": %.2f, \"mixed_fma_tflops\": %.2f", dm / ms * 1e-9, fm / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { lds_kernel<<<blocks, 256>>>(out, 20000); });
    printf(", \"lds64_TBps\": %.2f", 8.0 * blocks * 256.0 * 16.0 * 20000 / ms * 1e-9);
  }
  {
    const size_t bytes = 48ull << 20, n = bytes / 16;   // L2-resident
    double2 *buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    float ms = time_ms([&] { read_kernel<<<sms * 8, 512>>>(buf, n, 20, out); });
    printf(", \"l2_read_TBps\": %.2f", bytes * 20.0 / ms * 1e-9);
    cudaFree(buf);
  }
  {
    const size_t bytes = 4ull << 30, n = bytes / 16;    // DRAM
    double2 *buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    float ms = time_ms([&] { read_kernel<<<sms * 8, 512>>>(buf, n, 1, out); });
    printf(", \"dram_read_TBps\": %.2f", bytes * 1.0 / ms * 1e-9);
    cudaFree(buf);
  }
  printf("}\n");
  return 0;
}
