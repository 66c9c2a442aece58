// Microbenchmark: FP64 peak of DMMA.8x8x4 (mma.sync m8n8k4 f64) vs DFMA on this GPU. nvcc -arch=sm_100a tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void dmma_k(double *out, int iters) {
  double c[NACC][2];
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void dfma_k(double *out, int iters) {
  double c[NACC];
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double *out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      dmma_k<8><<<148 * 2, warps * 32 / 2>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 256 * 8 * (double)iters * (148.0 * warps);
    printf("DMMA  warps/SM=%2d  %.2f TFLOP/s\n", warps, flops / ms / 1e9);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      dfma_k<8><<<148 * 2, warps * 32 / 2>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    cudaEventElapsedTime(&ms, e0, e1);
    flops = 2.0 * 32 * 8 * (double)iters * (148.0 * warps);
    printf("DFMA  warps/SM=%2d  %.2f TFLOP/s\n", warps, flops / ms / 1e9);
  }
  return 0;
}
