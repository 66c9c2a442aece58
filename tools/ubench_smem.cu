// Microbenchmark: is the warp shuffle a second data path beside shared memory?  8 warps per SM (the qreg kernel's consumer
// team) gather 16-byte complex amplitudes either with LDS.128 from a staged tile, with 4 x SHFL.BFLY from registers, or with a
// mix of both, and apply 2 DFMA per amplitude like the real kernel.  Reports cycles per 16-amplitude gather per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_smem tools/ubench_smem.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NLDS, int NSHF>
__global__ void __launch_bounds__(256, 1) k(double2 *out, int iters, long long *cycles) {
  extern __shared__ __align__(16) unsigned char sm[];
  double2 *xs = reinterpret_cast<double2 *>(sm);
  for (int i = threadIdx.x; i < 4096; i += 256) xs[i] = make_double2(i * 1e-3, 1.0 - i * 1e-3);
  __syncthreads();
  double2 acc[16], xr[16];
  for (int u = 0; u < 16; ++u) {
    acc[u] = make_double2(0.0, 0.0);
    xr[u] = xs[threadIdx.x + 256 * u];
  }
  const double w = 1.0 + threadIdx.x * 1e-9;
  const unsigned char *base = sm + threadIdx.x * 16;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < NLDS; ++c) {
      const unsigned char *xt = sm + ((threadIdx.x * 16) ^ (16u << ((it + c) & 7)));
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const double2 v = *reinterpret_cast<const double2 *>(xt + u * 4096);
        acc[u].x = fma(w, v.x, acc[u].x);
        acc[u].y = fma(w, v.y, acc[u].y);
      }
    }
#pragma unroll
    for (int c = 0; c < NSHF; ++c) {
      const int m = 1 << ((it + c) % 5);
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        double2 v;
        v.x = __shfl_xor_sync(0xffffffffu, xr[u].x, m);
        v.y = __shfl_xor_sync(0xffffffffu, xr[u].y, m);
        acc[u].x = fma(w, v.x, acc[u].x);
        acc[u].y = fma(w, v.y, acc[u].y);
      }
    }
  }
  const long long t1 = clock64();
  double2 s = make_double2(0, 0);
  for (int u = 0; u < 16; ++u) s.x += acc[u].x, s.y += acc[u].y;
  out[blockIdx.x * 256 + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  (void)base;
}

template <int NLDS, int NSHF>
void run(const char *name, double2 *out, long long *dcyc) {
  const int iters = 2000;
  cudaFuncSetAttribute(k<NLDS, NSHF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<NLDS, NSHF><<<148, 256, 65536>>>(out, 10, dcyc);
  k<NLDS, NSHF><<<148, 256, 65536>>>(out, iters, dcyc);
  long long c = 0;
  cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)c / iters;
  printf("%-34s %8.1f cycles per iteration = %6.1f per 16-amplitude gather per warp (x8 warps: %6.1f per SM)\n", name, per,
         per / (NLDS + NSHF), per / (NLDS + NSHF));
}

int main() {
  double2 *out;
  long long *dcyc;
  cudaMalloc(&out, 148 * 256 * sizeof(double2));
  cudaMalloc(&dcyc, 8);
  run<4, 0>("4 x LDS.128 gathers", out, dcyc);
  run<0, 4>("4 x shuffle gathers", out, dcyc);
  run<2, 2>("2 x LDS.128 + 2 x shuffle", out, dcyc);
  run<4, 2>("4 x LDS.128 + 2 x shuffle", out, dcyc);
  run<4, 4>("4 x LDS.128 + 4 x shuffle", out, dcyc);
  run<6, 0>("6 x LDS.128", out, dcyc);
  run<3, 3>("3 x LDS.128 + 3 x shuffle", out, dcyc);
  printf("(8 warps x 16 LDS.128 x 4 wavefronts = 512 cycles per gather per SM is the shared-memory floor)\n");
  return 0;
}
