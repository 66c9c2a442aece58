// Microbenchmarks behind the round-2 design of the fused LazySum tile kernel:
//   T1  DRAM copy bandwidth (reference for everything else)
//   T2  L2-resident read / read-modify-write bandwidth as a function of the working set
//   T4  "chained passes": pass A (contiguous tiles, y = 2x) and pass B (window tiles, y += x) over a 2^28-amplitude state,
//       either as two launches (every byte through DRAM twice) or as ONE persistent launch whose tile queue is ordered
//       chunk by chunk (A(c+1) then B(c)), so that B finds x and y of its chunk in L2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_l2 tools/ubench_l2.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e = (x);                                                          \
    if (e != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                    \
    }                                                                             \
  } while (0)

__global__ void copy_k(const double2 *__restrict__ x, double2 *__restrict__ y, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    double2 a = x[i], b = x[i + stride], c = x[i + 2 * stride], d = x[i + 3 * stride];
    y[i] = a;
    y[i + stride] = b;
    y[i + 2 * stride] = c;
    y[i + 3 * stride] = d;
  }
  for (; i < n; i += stride) y[i] = x[i];
}

// every CTA sweeps the whole working set `reps` times (different starting offsets), 8 independent 16-byte loads in flight
__global__ void l2_read_k(const double2 *__restrict__ x, size_t n, int reps, double *sink) {
  double acc = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x + (size_t)r * 977 * 32) % n;
    for (size_t k = 0; k < n / stride; k += 8) {
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        size_t j = i + (k + u) * stride;
        if (j >= n) j -= n;
        v[u] = __ldcg(x + j);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y;
    }
  }
  if (acc == 1.2345) sink[0] = acc;
}
__global__ void l2_rmw_k(double2 *__restrict__ y, size_t n, int reps) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t k = 0; k < n / stride; k += 8) {
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(y + i + (k + u) * stride);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        v[u].x += 1.0;
        __stcg(y + i + (k + u) * stride, v[u]);
      }
    }
  }
}

// ---------------------------------------------------------------- T4
struct ChainP {
  int nbits, cbits, T;  // state bits, chunk bits, tile bits
  int chained;          // 1: chunk-ordered queue with counters; 0: only pass `only`
  int only;             // unchained: 0 = pass A over the whole state, 1 = pass B
  int hint;             // 1: pass B reads/writes with L2 evict_first
  unsigned *queue;      // work counter
  unsigned *done;       // per chunk: finished A tiles
};

__device__ __forceinline__ double2 ld_hint(const double2 *p, unsigned long long pol, int hint) {
  double2 v;
  if (hint)
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  else
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_hint(double2 *p, double2 v, unsigned long long pol, int hint) {
  if (hint)
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
  else
    asm volatile("st.global.cg.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

__device__ __forceinline__ void tileA(const double2 *x, double2 *y, size_t base) {
  // 4096 contiguous amplitudes, 256 threads x 16
  double2 v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = __ldcg(x + base + k * 256 + threadIdx.x);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    v[k].x *= 2.0;
    v[k].y *= 2.0;
    __stcg(y + base + k * 256 + threadIdx.x, v[k]);
  }
}
__device__ __forceinline__ void tileB(const double2 *x, double2 *y, size_t base, int wshift, unsigned long long pol, int hint) {
  // rows of 8 amplitudes (128 B) at stride 2^wshift amplitudes, 512 rows
  const unsigned l = threadIdx.x & 7, w0 = threadIdx.x >> 3;
  double2 v[16], o[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = ld_hint(x + base + ((size_t)(w0 + 32 * k) << wshift) + l, pol, hint);
#pragma unroll
  for (int k = 0; k < 16; ++k) o[k] = ld_hint(y + base + ((size_t)(w0 + 32 * k) << wshift) + l, pol, hint);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    o[k].x += v[k].x;
    o[k].y += v[k].y;
    st_hint(y + base + ((size_t)(w0 + 32 * k) << wshift) + l, o[k], pol, hint);
  }
}

__global__ void __launch_bounds__(256) chain_k(ChainP P, const double2 *__restrict__ x, double2 *__restrict__ y) {
  __shared__ unsigned s_item;
  const unsigned tpc = 1u << (P.cbits - P.T);       // tiles per chunk and pass
  const unsigned C = 1u << (P.nbits - P.cbits);     // chunks
  unsigned long long pol = 0;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const unsigned total = P.chained ? 2u * C * tpc : C * tpc;
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_item = atomicAdd(P.queue, 1u);
    __syncthreads();
    const unsigned item = s_item;
    if (item >= total) break;
    unsigned b = item / tpc, t = item % tpc, c;
    int isB;
    if (P.chained) {
      if (b == 0) { isB = 0; c = 0; }
      else if (b == 2 * C - 1) { isB = 1; c = C - 1; }
      else if (b & 1) { isB = 0; c = (b + 1) / 2; }
      else { isB = 1; c = b / 2 - 1; }
    } else {
      isB = P.only;
      c = b;
    }
    if (!isB) {
      tileA(x, y, ((size_t)c << P.cbits) | ((size_t)t << P.T));
      if (P.chained) {
        __syncthreads();
        if (threadIdx.x == 0) {
          __threadfence();
          atomicAdd(P.done + c, 1u);
        }
      }
    } else {
      if (P.chained) {
        if (threadIdx.x == 0) {
          while (*((volatile unsigned *)(P.done + c)) < tpc) __nanosleep(100);
          __threadfence();
        }
        __syncthreads();
      }
      // window tile: free bits {0,1,2} + the top 9 bits of the chunk; tile id = bits 3 .. cbits-10
      tileB(x, y, ((size_t)c << P.cbits) | ((size_t)t << 3), P.cbits - 9, pol, P.hint);
    }
  }
}

__global__ void check_k(const double2 *x, const double2 *y, size_t n, unsigned long long *bad) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned long long b = 0;
  for (; i < n; i += stride) {
    double2 a = x[i], c = y[i];
    if (c.x != 3.0 * a.x || c.y != 3.0 * a.y) ++b;
  }
  if (b) atomicAdd(bad, b);
}
__global__ void init_k(double2 *x, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] = make_double2((double)(i % 1021) + 1.0, (double)(i % 509) - 7.0);
}

int main(int argc, char **argv) {
  const int nbits = argc > 1 ? atoi(argv[1]) : 28;
  const size_t n = (size_t)1 << nbits;
  double2 *x, *y;
  CK(cudaMalloc(&x, n * sizeof(double2)));
  CK(cudaMalloc(&y, n * sizeof(double2)));
  init_k<<<148 * 8, 256>>>(x, n);
  CK(cudaMemset(y, 0, n * sizeof(double2)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms;
  // ---- T1
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    copy_k<<<148 * 16, 256>>>(x, y, n);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    printf("T1 copy %zu MiB: %.3f ms  %.1f GB/s (read+write)\n", n * 16 >> 20, ms, 2.0 * n * 16 / ms / 1e6);
  }
  // ---- T2
  double *sink;
  CK(cudaMalloc(&sink, 8));
  for (int mb : {8, 16, 32, 48, 64, 80, 96, 112, 128, 192, 256}) {
    const size_t m = (size_t)mb << 20 >> 4;
    const int reps = 40;
    l2_read_k<<<148 * 4, 256>>>(x, m, 2, sink);  // warm
    cudaEventRecord(e0);
    l2_read_k<<<148 * 4, 256>>>(x, m, reps, sink);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    const double rd = (double)m * 16 * reps / ms / 1e6;
    l2_rmw_k<<<148 * 4, 256>>>(y, m, 2);
    cudaEventRecord(e0);
    l2_rmw_k<<<148 * 4, 256>>>(y, m, reps);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    printf("T2 working set %3d MB: read %.0f GB/s   rmw %.0f GB/s (read+write)\n", mb, rd, 2.0 * m * 16 * reps / ms / 1e6);
  }
  // ---- T4
  unsigned *queue, *done;
  CK(cudaMalloc(&queue, 4));
  CK(cudaMalloc(&done, 4 << 16));
  unsigned long long *bad;
  CK(cudaMalloc(&bad, 8));
  for (int ctas : {2, 4, 6}) {
    for (int mode = 0; mode < 8; ++mode) {
      ChainP P;
      P.nbits = nbits;
      P.T = 12;
      P.queue = queue;
      P.done = done;
      int cb = 20;
      P.hint = 0;
      const char *name = "";
      if (mode == 0) { P.chained = 0; name = "two launches (A then B)"; }
      if (mode == 1) { P.chained = 1; cb = 20; name = "chained, chunk 2^20"; }
      if (mode == 2) { P.chained = 1; cb = 20; P.hint = 1; name = "chained, chunk 2^20, evict_first in B"; }
      if (mode == 3) { P.chained = 1; cb = 19; name = "chained, chunk 2^19"; }
      if (mode == 4) { P.chained = 1; cb = 21; name = "chained, chunk 2^21"; }
      if (mode == 5) { P.chained = 1; cb = 21; P.hint = 1; name = "chained, chunk 2^21, evict_first in B"; }
      if (mode == 6) { P.chained = 1; cb = 22; P.hint = 1; name = "chained, chunk 2^22, evict_first in B"; }
      if (mode == 7) { P.chained = 1; cb = 18; name = "chained, chunk 2^18"; }
      P.cbits = cb;
      float best = 1e9f;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemsetAsync(done, 0, 4 << 16));
        CK(cudaMemsetAsync(queue, 0, 4));
        cudaEventRecord(e0);
        if (P.chained) {
          chain_k<<<148 * ctas, 256>>>(P, x, y);
        } else {
          P.only = 0;
          chain_k<<<148 * ctas, 256>>>(P, x, y);
          CK(cudaMemsetAsync(queue, 0, 4));
          P.only = 1;
          chain_k<<<148 * ctas, 256>>>(P, x, y);
        }
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
      }
      CK(cudaMemset(bad, 0, 8));
      check_k<<<148 * 8, 256>>>(x, y, n, bad);
      unsigned long long hb = 0;
      CK(cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost));
      printf("T4 ctas/SM=%d %-44s %.3f ms  (80 B/amp => %.0f GB/s; 32 B/amp => %.0f GB/s)  wrong=%llu\n", ctas, name, best,
             80.0 * n / best / 1e6, 32.0 * n / best / 1e6, hb);
    }
  }
  return 0;
}
