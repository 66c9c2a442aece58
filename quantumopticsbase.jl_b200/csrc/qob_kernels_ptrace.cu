// qob_kernels_ptrace.cu — partial trace of dense operators, kets and bras on the device.
//
// Reference behaviour replaced: src/operators_dense.jl:191-215 (ptrace of DataOperator / Ket / Bra) and the generated loop
// nests _ptrace / _ptrace_ket / _ptrace_bra (:311-383): for every pair of right / left multi-indices that agree on the traced
// subsystems, `result[Jl, Jr] += a[Il, Ir]` (operators) or `+= a[Il]*conj(a[Ir])` (kets; bras conjugate the other factor).
//
//   * operator: one output element per thread; the K = prod(traced dims) source elements of an output sit at
//     `base(jl, jr) + off[k]` with a host-built offset table (K entries, read through L1 by every thread).  Each source element
//     is read at most once: the kernel is a strided gather bound by HBM at 16 B x (Dl*Dr/K + Ml*Mr).
//   * ket / bra: result = Psi Psi^+ with Psi the (kept x traced) reshaping of the state — a 16 x 16-tiled complex rank-K
//     update through shared memory; a long traced extent is split over the grid and the partial tiles are summed by a second
//     kernel in a fixed order (deterministic, no atomics).
#include <algorithm>
#include <cstdio>

#include "qob_internal.h"

#define PT_MAXSUB 32

struct PtGeom {
  int nkeep, ntr;
  long long keep_dim_l[PT_MAXSUB], keep_dim_r[PT_MAXSUB];  // kept subsystems: dimension on the left / right basis
  long long keep_str_l[PT_MAXSUB], keep_str_r[PT_MAXSUB];  // stride of a kept subsystem in the left / right composite index
  long long tr_dim[PT_MAXSUB], tr_str_l[PT_MAXSUB], tr_str_r[PT_MAXSUB];
  long long Ml, Mr, K, Dl;
};

__device__ __forceinline__ long long pt_expand(long long j, int n, const long long *dims, const long long *strides) {
  long long off = 0;
  for (int d = 0; d < n; ++d) {
    const long long q = j / dims[d];
    off += (j - q * dims[d]) * strides[d];
    j = q;
  }
  return off;
}

__global__ void ptrace_op_kernel(const __grid_constant__ PtGeom G, const long long *__restrict__ off, const double2 *__restrict__ a,
                                 double2 *__restrict__ out) {
  const long long total = G.Ml * G.Mr;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const long long jr = o / G.Ml, jl = o - jr * G.Ml;
    const long long base = pt_expand(jl, G.nkeep, G.keep_dim_l, G.keep_str_l) + G.Dl * pt_expand(jr, G.nkeep, G.keep_dim_r, G.keep_str_r);
    double2 acc = make_double2(0.0, 0.0);
    long long k = 0;
    for (; k + 4 <= G.K; k += 4) {  // four independent loads in flight
      const double2 v0 = a[base + __ldg(off + k)], v1 = a[base + __ldg(off + k + 1)], v2 = a[base + __ldg(off + k + 2)],
                    v3 = a[base + __ldg(off + k + 3)];
      acc.x += (v0.x + v1.x) + (v2.x + v3.x);
      acc.y += (v0.y + v1.y) + (v2.y + v3.y);
    }
    for (; k < G.K; ++k) {
      const double2 v = a[base + __ldg(off + k)];
      acc.x += v.x;
      acc.y += v.y;
    }
    out[o] = acc;
  }
}

// out[z][jl, jr] = sum over the k range of split z of Psi[jl, k] * conj(Psi[jr, k])   (BRA: conj(Psi[jl,k]) * Psi[jr,k])
template <bool BRA>
__global__ void __launch_bounds__(256) ptrace_state_kernel(const __grid_constant__ PtGeom G, const double2 *__restrict__ psi,
                                                           double2 *__restrict__ out, long long kchunk) {
  __shared__ double2 A[16][17], B[16][17];
  __shared__ long long offA[16], offB[16], offK[16];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long M = G.Ml;
  const long long jl0 = (long long)blockIdx.x * 16, jr0 = (long long)blockIdx.y * 16;
  if (threadIdx.x < 16) {
    const long long j = jl0 + threadIdx.x;
    offA[threadIdx.x] = j < M ? pt_expand(j, G.nkeep, G.keep_dim_l, G.keep_str_l) : -1;
  } else if (threadIdx.x < 32) {
    const long long j = jr0 + threadIdx.x - 16;
    offB[threadIdx.x - 16] = j < M ? pt_expand(j, G.nkeep, G.keep_dim_l, G.keep_str_l) : -1;
  }
  const long long k_begin = (long long)blockIdx.z * kchunk, k_end = min(G.K, k_begin + kchunk);
  double2 acc = make_double2(0.0, 0.0);
  for (long long k0 = k_begin; k0 < k_end; k0 += 16) {
    __syncthreads();  // previous chunk consumed (first trip: offA / offB visible)
    if (threadIdx.x < 16) {
      const long long k = k0 + threadIdx.x;
      offK[threadIdx.x] = k < k_end ? pt_expand(k, G.ntr, G.tr_dim, G.tr_str_l) : -1;
    }
    __syncthreads();
    {
      const long long ok = offK[tx], oa = offA[ty], ob = offB[ty];
      A[ty][tx] = (ok >= 0 && oa >= 0) ? psi[oa + ok] : make_double2(0.0, 0.0);
      B[ty][tx] = (ok >= 0 && ob >= 0) ? psi[ob + ok] : make_double2(0.0, 0.0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const double2 p = A[tx][kk], q = B[ty][kk];  // out[jl0+tx, jr0+ty]
      if (!BRA) {  // p * conj(q)
        acc.x = fma(p.x, q.x, fma(p.y, q.y, acc.x));
        acc.y = fma(p.y, q.x, fma(-p.x, q.y, acc.y));
      } else {     // conj(p) * q
        acc.x = fma(p.x, q.x, fma(p.y, q.y, acc.x));
        acc.y = fma(p.x, q.y, fma(-p.y, q.x, acc.y));
      }
    }
  }
  const long long jl = jl0 + tx, jr = jr0 + ty;
  if (jl < M && jr < M) out[(long long)blockIdx.z * M * M + jl + M * jr] = acc;
}

__global__ void ptrace_sum_partials_kernel(const double2 *__restrict__ part, double2 *__restrict__ out, long long n, int nsplit) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double2 acc = part[i];
    for (int z = 1; z < nsplit; ++z) {
      const double2 v = part[(long long)z * n + i];
      acc.x += v.x;
      acc.y += v.y;
    }
    out[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------- host
// check_ptrace_arguments / check_indices (src/operators.jl:153-176): indices unique and in range, not all subsystems traced,
// traced subsystems square.
static int pt_geometry(int nsub, const int64_t *dims_l, const int64_t *dims_r, int ntraced, const int32_t *traced, PtGeom &G) {
  if (nsub < 1 || nsub > PT_MAXSUB) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "ptrace: %d subsystems (max %d)", nsub, PT_MAXSUB);
  if (!dims_l || !dims_r || (ntraced > 0 && !traced)) QOB_FAIL(QOB_STATUS_INVALID_ARG, "ptrace: null argument");
  if (ntraced == nsub)
    QOB_FAIL(QOB_STATUS_INVALID_ARG, "Partial trace can't be used to trace out all subsystems - use tr() instead.");
  std::vector<bool> is_tr(nsub, false);
  for (int i = 0; i < ntraced; ++i) {
    const int s = traced[i] - 1;
    if (s < 0 || s >= nsub) QOB_FAIL(QOB_STATUS_INVALID_ARG, "ptrace: index %d out of range 1:%d", traced[i], nsub);
    if (is_tr[s]) QOB_FAIL(QOB_STATUS_INVALID_ARG, "ptrace: index %d given twice", traced[i]);
    is_tr[s] = true;
    if (dims_l[s] != dims_r[s])
      QOB_FAIL(QOB_STATUS_INVALID_ARG, "Partial trace can only be applied onto subsystems that have the same left and right dimension.");
  }
  memset(&G, 0, sizeof G);
  long long sl = 1, sr = 1;
  G.Ml = G.Mr = G.K = 1;
  for (int s = 0; s < nsub; ++s) {
    if (dims_l[s] < 1 || dims_r[s] < 1) QOB_FAIL(QOB_STATUS_INVALID_ARG, "ptrace: non-positive dimension");
    if (is_tr[s]) {
      G.tr_dim[G.ntr] = dims_l[s];
      G.tr_str_l[G.ntr] = sl;
      G.tr_str_r[G.ntr] = sr;
      G.K *= dims_l[s];
      ++G.ntr;
    } else {
      G.keep_dim_l[G.nkeep] = dims_l[s];
      G.keep_dim_r[G.nkeep] = dims_r[s];
      G.keep_str_l[G.nkeep] = sl;
      G.keep_str_r[G.nkeep] = sr;
      G.Ml *= dims_l[s];
      G.Mr *= dims_r[s];
      ++G.nkeep;
    }
    sl *= dims_l[s];
    sr *= dims_r[s];
  }
  G.Dl = sl;
  return QOB_STATUS_OK;
}

int launch_ptrace_op(qob_ctx *ctx, int slot, int nsub, const int64_t *dims_l, const int64_t *dims_r, int ntraced, const int32_t *traced,
                     const void *a, void *result, cudaStream_t s) {
  PtGeom G;
  QOB_TRY(pt_geometry(nsub, dims_l, dims_r, ntraced, traced, G));
  if (t_planning_only) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "planning-only context (no CUDA device): libqob200 has no CPU fallback");
  if (G.K > (1ll << 26)) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "ptrace: traced extent %lld too large", G.K);
  // offsets of the K diagonal source elements relative to the (jl, jr) base
  std::vector<long long> off((size_t)G.K);
  for (long long k = 0; k < G.K; ++k) {
    long long kk = k, ol = 0, orr = 0;
    for (int d = 0; d < G.ntr; ++d) {
      const long long dig = kk % G.tr_dim[d];
      kk /= G.tr_dim[d];
      ol += dig * G.tr_str_l[d];
      orr += dig * G.tr_str_r[d];
    }
    off[(size_t)k] = ol + G.Dl * orr;
  }
  void *d_off = nullptr;
  QOB_TRY(ctx->get_scratch(s, slot, off.size() * sizeof(long long), &d_off));
  QOB_CUDA(cudaMemcpyAsync(d_off, off.data(), off.size() * sizeof(long long), cudaMemcpyHostToDevice, s));
  QOB_CUDA(cudaStreamSynchronize(s));  // `off` is a pageable host vector that dies with this call
  const long long total = G.Ml * G.Mr;
  const int sms = qob_device_sm_count();
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)sms * 16));
  ptrace_op_kernel<<<grid, 256, 0, s>>>(G, (const long long *)d_off, (const double2 *)a, (double2 *)result);
  QOB_LAUNCHED();
  QOB_CUDA(cudaGetLastError());
  return QOB_STATUS_OK;
}

int launch_ptrace_state(qob_ctx *ctx, int slot, int nsub, const int64_t *dims, int ntraced, const int32_t *traced, bool bra, const void *psi,
                        void *result, cudaStream_t s) {
  PtGeom G;
  QOB_TRY(pt_geometry(nsub, dims, dims, ntraced, traced, G));
  if (t_planning_only) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "planning-only context (no CUDA device): libqob200 has no CPU fallback");
  const long long M = G.Ml;
  const long long tiles = (M + 15) / 16;
  if (tiles > 65535) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "ptrace: reduced dimension %lld too large", M);
  const int sms = qob_device_sm_count();
  // split the traced extent so that the grid fills the GPU (partials are summed in a fixed order afterwards)
  long long nsplit = std::max<long long>(1, std::min<long long>((G.K + 255) / 256, (4ll * sms + tiles * tiles - 1) / (tiles * tiles)));
  nsplit = std::min<long long>(nsplit, 65535);
  while (nsplit > 1 && (double)nsplit * (double)M * (double)M * 16.0 > 512.0 * 1024 * 1024) nsplit /= 2;
  long long kchunk = ((G.K + nsplit - 1) / nsplit + 15) / 16 * 16;
  nsplit = (G.K + kchunk - 1) / kchunk;
  double2 *out = (double2 *)result;
  void *part = nullptr;
  if (nsplit > 1) {
    QOB_TRY(ctx->get_scratch(s, slot, (size_t)nsplit * (size_t)M * (size_t)M * sizeof(double2), &part));
    out = (double2 *)part;
  }
  const dim3 grid((unsigned)tiles, (unsigned)tiles, (unsigned)nsplit);
  if (bra) ptrace_state_kernel<true><<<grid, 256, 0, s>>>(G, (const double2 *)psi, out, kchunk);
  else ptrace_state_kernel<false><<<grid, 256, 0, s>>>(G, (const double2 *)psi, out, kchunk);
  QOB_LAUNCHED();
  QOB_CUDA(cudaGetLastError());
  if (nsplit > 1) {
    const long long n = M * M;
    const unsigned g2 = (unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)sms * 8));
    ptrace_sum_partials_kernel<<<g2, 256, 0, s>>>((const double2 *)part, (double2 *)result, n, (int)nsplit);
    QOB_LAUNCHED();
    QOB_CUDA(cudaGetLastError());
  }
  return QOB_STATUS_OK;
}
