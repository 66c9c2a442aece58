// qob_kernels_gather.cu — generic fused "row gather" kernel for a LazySum of LazyTensor terms on
// arbitrary subsystem dimensions, plus the small utility kernels (scale, axpby, fill, reductions).
//
// One launch computes  y = beta*y + alpha * sum_t coef_t * (A_t x)  for every term t, where
// A_t = (x)_k A_{t,k} acts on a few tensor axes (reference: one pass PER TERM,
// src/operators_lazysum.jl:189-200 looping over src/operators_lazytensor.jl:539-557).
// Each thread owns one output element, decomposes its tensor index only on the axes a term
// touches, and walks the product of the factors' CSR rows.  Non-square Eye isometries
// (src/operators_lazytensor.jl:491-514) are ordinary factors here (rows >= min(dl,dr) are empty).
//
// This is the correctness backbone (any dims, dense/CSC/Eye factors, Ket/Bra/left/right with the
// pre/post batch extents); the HBM-roofline kernels for qubit chains live in qob_kernels_qtile.cu,
// dense d>=16 factors go through qob_kernels_axis.cu.
#include <algorithm>
#include <cstdio>

#include "qob_internal.h"

#define GMAXF 4  // max factors per fused term (longer terms are applied factor by factor)

struct GatherParams {
  const int4 *terms;        // nfac, fac_begin, nseg (-1: uniform strides), seg_begin
  const int2 *fac_i;        // dim_out, rowptr offset
  const longlong2 *fac_s;   // stride_out, stride_in
  const longlong3 *segs;    // stride_out, size, stride_in
  const int *rowptr;
  const int *colidx;
  const double2 *vals;
  const double2 *coef;
  int nterms;
  long long d_out, d_in, pre, post;
  double2 alpha, beta;
  int beta_zero;
  int tpo;  // threads per output element (power of two <= 32): small states split the TERMS of the sum over the lanes of a
            // group and reduce with shuffles, which shortens the per-thread chain of dependent loads (launch-latency regime)
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void cfma(double2 &acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}

template <typename IdxT>
__global__ void __launch_bounds__(256) gather_kernel(GatherParams p, const double2 *__restrict__ x,
                                                     double2 *__restrict__ y) {
  const IdxT total = (IdxT)(p.pre * p.d_out * p.post);
  const IdxT pre = (IdxT)p.pre, d_out = (IdxT)p.d_out;
  const unsigned tpo = (unsigned)p.tpo, sub = threadIdx.x & (tpo - 1);
  const unsigned gmask = tpo >= 32 ? 0xffffffffu : (((1u << tpo) - 1u) << ((threadIdx.x & 31u) & ~(tpo - 1)));
  for (IdxT idx = ((IdxT)blockIdx.x * blockDim.x + threadIdx.x) / tpo; idx < total; idx += ((IdxT)gridDim.x * blockDim.x) / tpo) {
    IdxT r = idx % pre, t1 = idx / pre;
    IdxT I = t1 % d_out, c = t1 / d_out;
    const double2 *xb = x + (long long)r + (long long)pre * p.d_in * (long long)c;
    double2 acc = make_double2(0.0, 0.0);
    for (int t = (int)sub; t < p.nterms; t += (int)tpo) {
      const int4 T = p.terms[t];
      IdxT J0;
      const bool uniform = T.z < 0;
      if (uniform) {
        J0 = I;
      } else {
        J0 = 0;
        for (int s = 0; s < T.z; ++s) {
          longlong3 sg = p.segs[T.w + s];
          IdxT q = (I / (IdxT)sg.x) % (IdxT)sg.y;
          J0 += q * (IdxT)sg.z;
        }
      }
      int lo[GMAXF], hi[GMAXF];
      IdxT sin[GMAXF];
      bool empty = false;
#pragma unroll
      for (int f = 0; f < GMAXF; ++f) {
        if (f < T.x) {
          int2 fi = p.fac_i[T.y + f];
          longlong2 fs = p.fac_s[T.y + f];
          IdxT i = (I / (IdxT)fs.x) % (IdxT)fi.x;
          if (uniform) J0 -= i * (IdxT)fs.x;
          lo[f] = p.rowptr[fi.y + (int)i];
          hi[f] = p.rowptr[fi.y + (int)i + 1];
          sin[f] = (IdxT)fs.y;
          empty |= lo[f] >= hi[f];
        }
      }
      if (empty) continue;
      double2 tacc = make_double2(0.0, 0.0);
      if (T.x == 0) {
        tacc = xb[(long long)J0 * pre];
      } else if (T.x == 1) {
        for (int a = lo[0]; a < hi[0]; ++a)
          cfma(tacc, p.vals[a], xb[(long long)(J0 + (IdxT)p.colidx[a] * sin[0]) * pre]);
      } else if (T.x == 2) {
        for (int b = lo[1]; b < hi[1]; ++b) {
          IdxT Jb = J0 + (IdxT)p.colidx[b] * sin[1];
          double2 vb = p.vals[b];
          double2 inner = make_double2(0.0, 0.0);
          for (int a = lo[0]; a < hi[0]; ++a)
            cfma(inner, p.vals[a], xb[(long long)(Jb + (IdxT)p.colidx[a] * sin[0]) * pre]);
          cfma(tacc, vb, inner);
        }
      } else {
        // odometer over up to GMAXF factor rows
        int cur[GMAXF];
#pragma unroll
        for (int f = 0; f < GMAXF; ++f) cur[f] = (f < T.x) ? lo[f] : 0;
        while (true) {
          double2 v = make_double2(1.0, 0.0);
          IdxT J = J0;
#pragma unroll
          for (int f = 0; f < GMAXF; ++f)
            if (f < T.x) {
              v = cmul(v, p.vals[cur[f]]);
              J += (IdxT)p.colidx[cur[f]] * sin[f];
            }
          cfma(tacc, v, xb[(long long)J * pre]);
          int f = 0;
          while (f < T.x) {
            if (++cur[f] < hi[f]) break;
            cur[f] = lo[f];
            ++f;
          }
          if (f == T.x) break;
        }
      }
      cfma(acc, p.coef[t], tacc);
    }
    for (unsigned o = tpo >> 1; o > 0; o >>= 1) {
      acc.x += __shfl_xor_sync(gmask, acc.x, o);
      acc.y += __shfl_xor_sync(gmask, acc.y, o);
    }
    if (sub == 0) {
      double2 out = cmul(p.alpha, acc);
      if (!p.beta_zero) cfma(out, p.beta, y[idx]);
      y[idx] = out;
    }
  }
}

// ------------------------------------------------------------------------------------------
int gather_program_build(GatherProgram &p, const std::vector<int64_t> &dims_out, const std::vector<int64_t> &dims_in,
                         const std::vector<OrientedTerm> &terms) {
  const int n = (int)dims_out.size();
  std::vector<int64_t> so(n), si(n);
  int64_t a = 1, b = 1;
  for (int k = 0; k < n; ++k) {
    so[k] = a;
    si[k] = b;
    a *= dims_out[k];
    b *= dims_in[k];
  }
  p.d_out = a;
  p.d_in = b;
  p.nterms = (int)terms.size();
  p.max_fac = 0;
  bool all_same = true;
  for (int k = 0; k < n; ++k) all_same &= dims_out[k] == dims_in[k];

  std::vector<int32_t> terms_i, fac_i, rowptr, colidx;
  std::vector<int64_t> fac_s, segs;
  std::vector<double2> vals;
  p.coef_of_term.clear();
  p.scalars.clear();
  for (const OrientedTerm &t : terms) {
    // factor list = given factors + explicit rectangular identities on untouched unequal axes
    std::vector<int> axes = t.axes;
    std::vector<HostMat> extra;
    std::vector<const HostMat *> mats;
    for (const HostMat &m : t.mats) mats.push_back(&m);
    if (!all_same) {
      extra.reserve(n);
      for (int k = 0; k < n; ++k) {
        bool used = false;
        for (int ax : t.axes) used |= ax == k;
        if (!used && dims_out[k] != dims_in[k]) {
          HostMat e;
          e.kind = QOB_FACTOR_EYE;
          e.rows = dims_out[k];
          e.cols = dims_in[k];
          extra.push_back(e);
          axes.push_back(k);
        }
      }
      for (size_t i = 0; i < extra.size(); ++i) mats.push_back(&extra[i]);
    }
    if ((int)axes.size() > GMAXF) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "gather program: term with %d factors (max %d)", (int)axes.size(), GMAXF);
    // sort factors by axis
    std::vector<int> order(axes.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int u, int v) { return axes[u] < axes[v]; });
    int fac_begin = (int)fac_i.size() / 2;
    for (int oi : order) {
      const HostMat &m = *mats[oi];
      int k = axes[oi];
      if (m.rows != dims_out[k] || m.cols != dims_in[k])
        QOB_FAIL(QOB_STATUS_DIM_MISMATCH, "factor on axis %d is %lldx%lld, expected %lldx%lld", k + 1, (long long)m.rows,
                 (long long)m.cols, (long long)dims_out[k], (long long)dims_in[k]);
      std::vector<int32_t> rp, ci;
      std::vector<cplx> v;
      m.to_csr(rp, ci, v);
      int off = (int)rowptr.size();
      int voff = (int)vals.size();
      for (int32_t r : rp) rowptr.push_back(r + voff);
      for (int32_t cidx : ci) colidx.push_back(cidx);
      for (cplx z : v) vals.push_back(make_double2(z.real(), z.imag()));
      fac_i.push_back((int32_t)dims_out[k]);
      fac_i.push_back(off);
      fac_s.push_back(so[k]);
      fac_s.push_back(si[k]);
    }
    int nseg = -1, seg_begin = (int)segs.size() / 3;
    if (!all_same) {
      // plain runs between special axes: equal dims, (possibly) different strides
      nseg = 0;
      int k = 0;
      while (k < n) {
        bool special = false;
        for (int ax : axes) special |= ax == k;
        if (special) {
          ++k;
          continue;
        }
        int k0 = k;
        int64_t size = 1;
        while (k < n) {
          bool sp2 = false;
          for (int ax : axes) sp2 |= ax == k;
          if (sp2) break;
          size *= dims_out[k];
          ++k;
        }
        if (size > 1) {
          segs.push_back(so[k0]);
          segs.push_back(size);
          segs.push_back(si[k0]);
          ++nseg;
        }
      }
    }
    terms_i.push_back((int)axes.size());
    terms_i.push_back(fac_begin);
    terms_i.push_back(nseg);
    terms_i.push_back(seg_begin);
    p.max_fac = std::max(p.max_fac, (int)axes.size());
    p.coef_of_term.push_back(t.coef_index);
    p.scalars.push_back(t.scalar);
  }
  // pack: i32 = [terms | fac_i | rowptr | colidx], i64 = [fac_s | segs]
  while (fac_i.size() % 4) fac_i.push_back(0);  // keep the int4 / int2 views 16-byte aligned
  std::vector<int32_t> i32(terms_i);
  p.off_fac_i = i32.size();
  i32.insert(i32.end(), fac_i.begin(), fac_i.end());
  p.off_rowptr = i32.size();
  i32.insert(i32.end(), rowptr.begin(), rowptr.end());
  p.off_colidx = i32.size();
  i32.insert(i32.end(), colidx.begin(), colidx.end());
  if (i32.empty()) i32.push_back(0);
  std::vector<int64_t> i64(fac_s);
  p.off_segs = i64.size();
  i64.insert(i64.end(), segs.begin(), segs.end());
  if (i64.empty()) i64.push_back(0);
  if (vals.empty()) vals.push_back(make_double2(0, 0));
  QOB_TRY(p.d_i32.upload(i32));
  QOB_TRY(p.d_i64.upload(i64));
  QOB_TRY(p.d_vals.upload(vals));
  p.describe = "gather[terms=" + std::to_string(p.nterms) + ",maxfac=" + std::to_string(p.max_fac) + "]";
  return QOB_STATUS_OK;
}

int gather_program_set_coefs(GatherProgram &p, const std::vector<cplx> &coefs, cudaStream_t s) {
  std::vector<double2> c(std::max<size_t>(1, p.coef_of_term.size()));
  for (size_t t = 0; t < p.coef_of_term.size(); ++t) {
    cplx v = p.scalars[t];
    if (p.coef_of_term[t] >= 0) {
      if ((size_t)p.coef_of_term[t] >= coefs.size()) QOB_FAIL(QOB_STATUS_INVALID_ARG, "coefficient index out of range");
      v *= coefs[p.coef_of_term[t]];
    }
    c[t] = make_double2(v.real(), v.imag());
  }
  // pageable-host async copies are staged by the runtime before returning, so `c` may die here
  return p.d_coef.upload_async(c, s);
}

int gather_program_launch(const GatherProgram &p, int64_t pre, int64_t post, cplx alpha, const void *x, cplx beta,
                          void *y, cudaStream_t s) {
  GatherParams g;
  g.terms = reinterpret_cast<const int4 *>(p.d_i32.ptr);
  g.fac_i = reinterpret_cast<const int2 *>(p.d_i32.ptr + p.off_fac_i);
  g.rowptr = p.d_i32.ptr + p.off_rowptr;
  g.colidx = p.d_i32.ptr + p.off_colidx;
  g.fac_s = reinterpret_cast<const longlong2 *>(p.d_i64.ptr);
  g.segs = reinterpret_cast<const longlong3 *>(p.d_i64.ptr + p.off_segs);
  g.vals = p.d_vals.ptr;
  g.coef = p.d_coef.ptr;
  g.nterms = p.nterms;
  g.d_out = p.d_out;
  g.d_in = p.d_in;
  g.pre = pre;
  g.post = post;
  g.alpha = make_double2(alpha.real(), alpha.imag());
  g.beta = make_double2(beta.real(), beta.imag());
  g.beta_zero = (beta == cplx(0.0, 0.0));
  const int64_t total_out = pre * p.d_out * post, total_in = pre * p.d_in * post;
  if (total_out == 0) return QOB_STATUS_OK;
  const int threads = 256;
  int tpo = 1;
  while (tpo < 32 && tpo < p.nterms && total_out * tpo * 2 <= (int64_t)148 * 1024) tpo <<= 1;
  g.tpo = tpo;
  int64_t blocks = (total_out * tpo + threads - 1) / threads;
  if (blocks > (int64_t)1 << 30) blocks = (int64_t)1 << 30;
  if (total_out < ((int64_t)1 << 31) && total_in < ((int64_t)1 << 31))
    gather_kernel<uint32_t><<<(unsigned)blocks, threads, 0, s>>>(g, (const double2 *)x, (double2 *)y);
  else
    gather_kernel<unsigned long long><<<(unsigned)blocks, threads, 0, s>>>(g, (const double2 *)x, (double2 *)y);
  QOB_LAUNCHED();
  QOB_CUDA(cudaGetLastError());
  return QOB_STATUS_OK;
}

// ------------------------------------------------------------------------------------------ utilities
__global__ void scale_kernel(double2 *__restrict__ y, long long n, double2 beta, int zero) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = zero ? make_double2(0.0, 0.0) : cmul(beta, y[i]);
}
int launch_scale(void *y, int64_t n, cplx beta, cudaStream_t s) {
  if (n == 0 || beta == cplx(1.0, 0.0)) return QOB_STATUS_OK;  // _zero_op_mul!: beta == 1 is a no-op
  int64_t blocks = std::min<int64_t>((n + 255) / 256, 148 * 32);
  scale_kernel<<<(unsigned)blocks, 256, 0, s>>>((double2 *)y, n, make_double2(beta.real(), beta.imag()),
                                                beta == cplx(0.0, 0.0));
  QOB_LAUNCHED();
  QOB_CUDA(cudaGetLastError());
  return QOB_STATUS_OK;
}

__global__ void axpby_kernel(const double2 *__restrict__ x, double2 *__restrict__ y, long long n, double2 alpha,
                             double2 beta, int zero) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double2 o = cmul(alpha, x[i]);
    if (!zero) cfma(o, beta, y[i]);
    y[i] = o;
  }
}
int launch_axpby(const void *x, void *y, int64_t n, cplx alpha, cplx beta, cudaStream_t s) {
  if (n == 0) return QOB_STATUS_OK;
  int64_t blocks = std::min<int64_t>((n + 255) / 256, 148 * 32);
  axpby_kernel<<<(unsigned)blocks, 256, 0, s>>>((const double2 *)x, (double2 *)y, n,
                                                make_double2(alpha.real(), alpha.imag()),
                                                make_double2(beta.real(), beta.imag()), beta == cplx(0.0, 0.0));
  QOB_LAUNCHED();
  QOB_CUDA(cudaGetLastError());
  return QOB_STATUS_OK;
}

// counter-based generator — must match oracle/qob_oracle.c:orc_fill_state bit for bit
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double u11(unsigned long long seed, unsigned long long ctr) {
  unsigned long long r = splitmix64(seed ^ splitmix64(ctr));
  return (double)(r >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}
__global__ void fill_kernel(double2 *__restrict__ x, long long offset, long long n, unsigned long long seed,
                            double scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long g = (unsigned long long)(offset + i);
    x[i] = make_double2(scale * u11(seed, 2 * g), scale * u11(seed, 2 * g + 1));
  }
}
int launch_fill_state(void *x, int64_t offset, int64_t n, uint64_t seed, double scale, cudaStream_t s) {
  if (n == 0) return QOB_STATUS_OK;
  int64_t blocks = std::min<int64_t>((n + 255) / 256, 148 * 32);
  fill_kernel<<<(unsigned)blocks, 256, 0, s>>>((double2 *)x, offset, n, seed, scale);
  QOB_LAUNCHED();
  QOB_CUDA(cudaGetLastError());
  return QOB_STATUS_OK;
}

// reductions, deterministic (the same bits on every run): per-block tree in shared memory -> one partial per block and
// component, then one block folds the partials in a fixed order
template <int MODE>  // 0: sum |x|^2 ; 1: sum conj(x)*y
__global__ void reduce_kernel(const double2 *__restrict__ x, const double2 *__restrict__ y, long long n,
                              double *__restrict__ partial) {
  double re = 0.0, im = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double2 a = x[i];
    if (MODE == 0) {
      re = fma(a.x, a.x, re);
      re = fma(a.y, a.y, re);
    } else {
      double2 b = y[i];
      re += a.x * b.x + a.y * b.y;
      im += a.x * b.y - a.y * b.x;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    re += __shfl_down_sync(0xffffffffu, re, o);
    im += __shfl_down_sync(0xffffffffu, im, o);
  }
  __shared__ double sre[8], sim[8];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    sre[w] = re;
    sim[w] = im;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      re += sre[k];
      im += sim[k];
    }
    partial[2 * blockIdx.x] = re;
    partial[2 * blockIdx.x + 1] = im;
  }
}
__global__ void reduce_final_kernel(const double *__restrict__ partial, int nblocks, double *__restrict__ out) {
  double re = 0.0, im = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
    re += partial[2 * i];
    im += partial[2 * i + 1];
  }
  for (int o = 16; o > 0; o >>= 1) {
    re += __shfl_down_sync(0xffffffffu, re, o);
    im += __shfl_down_sync(0xffffffffu, im, o);
  }
  __shared__ double sre[8], sim[8];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    sre[w] = re;
    sim[w] = im;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      re += sre[k];
      im += sim[k];
    }
    out[0] = re;
    out[1] = im;
  }
}
static int reduce_common(int mode, const void *x, const void *y, int64_t n, double *host2, cudaStream_t s) {
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 148 * 16));
  double *d = nullptr;   // [2*blocks partials][2 result], stream-ordered allocation (no device-wide synchronisation)
  QOB_CUDA(cudaMallocAsync(&d, (size_t)(2 * blocks + 2) * sizeof(double), s));
  if (mode == 0)
    reduce_kernel<0><<<(unsigned)blocks, 256, 0, s>>>((const double2 *)x, nullptr, n, d);
  else
    reduce_kernel<1><<<(unsigned)blocks, 256, 0, s>>>((const double2 *)x, (const double2 *)y, n, d);
  QOB_LAUNCHED();
  reduce_final_kernel<<<1, 256, 0, s>>>(d, blocks, d + 2 * blocks);
  QOB_LAUNCHED();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(host2, d + 2 * blocks, 2 * sizeof(double), cudaMemcpyDeviceToHost, s);
  cudaError_t e2 = cudaFreeAsync(d, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  QOB_CUDA(e);
  QOB_CUDA(e2);
  return QOB_STATUS_OK;
}
int launch_norm2(const void *x, int64_t n, double *host_out, cudaStream_t s) {
  double h[2] = {0, 0};
  QOB_TRY(reduce_common(0, x, nullptr, n, h, s));
  *host_out = h[0];
  return QOB_STATUS_OK;
}
int launch_dot(const void *x, const void *y, int64_t n, cplx *host_out, cudaStream_t s) {
  double h[2] = {0, 0};
  QOB_TRY(reduce_common(1, x, y, n, h, s));
  *host_out = cplx(h[0], h[1]);
  return QOB_STATUS_OK;
}
