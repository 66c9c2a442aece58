// qob_kernels_axis.cu — dense single-axis contraction on FP64 tensor cores (DMMA) and the
// sparse x dense kernels.
//
//   y[l, i, r] = alpha * sum_j A[i, j] * x[l, j, r] + beta * y[l, i, r]
//
// replaces `_tp_matmul!` (src/operators_lazytensor.jl:406-428): the reference runs zgemm on a 2-D
// reshape for the first/last axis (:281-301) and permute -> zgemm -> permute for a middle axis
// (:333-404, three extra passes over the state).  Here any axis position is one kernel: the state is
// viewed as (L, d, R), the "GEMM" column index n = l + L*r addresses x[l + L*(j + dr*r)] directly, so
// no permutation is ever materialised.  Complex products are four real m8n8k4 DMMAs on split re/im
// shared-memory planes (tcgen05 has no FP64 kind; `mma.sync ... f64` -> SASS DMMA.8x8x4 is the FP64
// tensor path on sm_100a).  Used for genuinely dense factors (d >= 16, BASELINE config 3, d = 48).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "qob_internal.h"

#define AX_KT 16
#define AX_LD (AX_KT + 4)  // +4 doubles: conflict-free 64-bit fragment loads (ld = 4 mod 16)

struct AxisParams {
  const double *a_re, *a_im;  // [dl_pad][dr_pad] row-major
  int dl, dr, dl_pad, dr_pad;
  long long L, R, N;          // N = L*R
  double2 alpha, beta;
  int beta_zero;
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int MW, int NW>
__global__ void __launch_bounds__(32 * MW * NW)
    axis_dmma_kernel(const __grid_constant__ AxisParams P, const double2 *__restrict__ x, double2 *__restrict__ y) {
  constexpr int TM = 16 * MW, TN = 32 * NW, NT = 32 * MW * NW;
  __shared__ double As_re[TM * AX_LD], As_im[TM * AX_LD];
  __shared__ double Xs_re[TN * AX_LD], Xs_im[TN * AX_LD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp % MW, wn = warp / MW;
  const int g = lane >> 2, t = lane & 3;
  const long long n0 = (long long)blockIdx.x * TN;
  const int m0 = blockIdx.y * TM;
  const bool n_fast = P.L >= 8;  // which index is contiguous in global memory for the x tile

  double cre[2][4][2], cim[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) cre[a][b][0] = cre[a][b][1] = cim[a][b][0] = cim[a][b][1] = 0.0;

  for (int k0 = 0; k0 < P.dr; k0 += AX_KT) {
    // A tile (zero padded on the host to multiples of 16 in both directions)
    for (int e = tid; e < TM * AX_KT; e += NT) {
      int i = e / AX_KT, k = e % AX_KT;
      double re = 0.0, im = 0.0;
      if (m0 + i < P.dl_pad) {
        re = P.a_re[(long long)(m0 + i) * P.dr_pad + k0 + k];
        im = P.a_im[(long long)(m0 + i) * P.dr_pad + k0 + k];
      }
      As_re[i * AX_LD + k] = re;
      As_im[i * AX_LD + k] = im;
    }
    // x tile: element (n, k) <- x[l + L*((k0+k) + dr*r)],  n0+n = l + L*r
    for (int e = tid; e < TN * AX_KT; e += NT) {
      int n, k;
      if (n_fast) {
        n = e % TN;
        k = e / TN;
      } else {
        k = e % AX_KT;
        n = e / AX_KT;
      }
      double2 v = make_double2(0.0, 0.0);
      long long nn = n0 + n;
      if (nn < P.N && k0 + k < P.dr) {
        long long r = nn / P.L, l = nn - r * P.L;
        v = x[l + P.L * ((long long)(k0 + k) + (long long)P.dr * r)];
      }
      Xs_re[n * AX_LD + k] = v.x;
      Xs_im[n * AX_LD + k] = v.y;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < AX_KT; kk += 4) {
      double are[2], aim[2], naim[2];
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        int row = wm * 16 + mb * 8 + g;
        are[mb] = As_re[row * AX_LD + kk + t];
        aim[mb] = As_im[row * AX_LD + kk + t];
        naim[mb] = -aim[mb];
      }
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        int col = wn * 32 + nb * 8 + g;
        double bre = Xs_re[col * AX_LD + kk + t];
        double bim = Xs_im[col * AX_LD + kk + t];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          dmma884(cre[mb][nb][0], cre[mb][nb][1], are[mb], bre);
          dmma884(cim[mb][nb][0], cim[mb][nb][1], are[mb], bim);
        }
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          dmma884(cre[mb][nb][0], cre[mb][nb][1], naim[mb], bim);
          dmma884(cim[mb][nb][0], cim[mb][nb][1], aim[mb], bre);
        }
      }
    }
    __syncthreads();
  }
  // epilogue: thread holds C[row g][cols 2t, 2t+1] of each 8x8 block
#pragma unroll
  for (int mb = 0; mb < 2; ++mb) {
    int i = m0 + wm * 16 + mb * 8 + g;
    if (i >= P.dl) continue;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        long long nn = n0 + wn * 32 + nb * 8 + 2 * t + e;
        if (nn >= P.N) continue;
        long long r = nn / P.L, l = nn - r * P.L;
        long long addr = l + P.L * ((long long)i + (long long)P.dl * r);
        double ar = cre[mb][nb][e], ai = cim[mb][nb][e];
        double2 o = make_double2(P.alpha.x * ar - P.alpha.y * ai, P.alpha.x * ai + P.alpha.y * ar);
        if (!P.beta_zero) {
          double2 yo = y[addr];
          o.x += P.beta.x * yo.x - P.beta.y * yo.y;
          o.y += P.beta.x * yo.y + P.beta.y * yo.x;
        }
        y[addr] = o;
      }
  }
}

// ---- whole-K, software-pipelined variant (dl, dr <= 64: every Fock/NLevel site factor of the BASELINE configs) ----
// Persistent CTA, one per SM.  The factor A (interleaved re/im, row stride == 4 mod 8 complex -> conflict-free LDS.128
// fragment loads that deliver re and im together) stays in shared memory; the (K x 64) slabs of x are double-buffered
// with 16-byte cp.async (zero-fill for the ragged edges), so the DMMA pipe works on tile i while tile i+1 streams in.
// Warp tile 16x16 (MW*4 warps), four real DMMA.8x8x4 per complex 8x8x4 block.
// (l, r) of GEMM column n0 + j (j < 64) given (l0, r0) of column n0: no 64-bit division in the inner loops
// (the per-element `nn / L` divisions cost ~40 % of the kernel: ~120 instructions each on the integer pipe)
__device__ __forceinline__ void split_col(unsigned j, unsigned l0, long long r0, long long L, unsigned &l, long long &r) {
  if (L >= 64) {  // at most one wrap
    const unsigned long long t = (unsigned long long)l0 + j;
    const bool w = t >= (unsigned long long)L;
    l = (unsigned)(w ? t - (unsigned long long)L : t);
    r = r0 + (w ? 1 : 0);
  } else {        // small L: 32-bit division of a number < 128
    const unsigned t = l0 + j, Ls = (unsigned)L;
    const unsigned q = t / Ls;
    l = t - q * Ls;
    r = r0 + q;
  }
}

__device__ __forceinline__ void cp_async16_zfill(void *smem, const void *gmem, bool valid) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(gmem), "r"(sz));
}

template <int MW, int NB>
__global__ void __launch_bounds__(MW * 128, (NB == 1 && MW <= 3) ? 2 : 1)
    axis_dmma_pipe_kernel(const __grid_constant__ AxisParams P, const double2 *__restrict__ x, double2 *__restrict__ y,
                          int kpad, int ldx) {
  // NB = 8-column blocks per warp: NB = 2 -> 64-column slabs, one CTA per SM; NB = 1 -> 32-column slabs, half the
  // shared memory and registers, TWO CTAs per SM so that one CTA's epilogue / barrier waits overlap the other's DMMAs
  constexpr int TM = 16 * MW, TN = 32 * NB, NT = MW * 128;
  extern __shared__ __align__(16) double2 smem_c[];
  double2 *As = smem_c;               // [TM][ldx]
  double2 *Xs0 = As + TM * ldx;       // [2][TN][ldx]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp % MW, wn = warp / MW;  // wn in 0..3: 8*NB columns each
  const int g = lane >> 2, t = lane & 3;
  const bool n_fast = P.L >= 8;

  for (int e = tid; e < TM * kpad; e += NT) {
    const int i = e / kpad, k = e - i * kpad;
    double2 v = make_double2(0.0, 0.0);
    if (i < P.dl_pad && k < P.dr_pad) v = make_double2(P.a_re[(long long)i * P.dr_pad + k], P.a_im[(long long)i * P.dr_pad + k]);
    As[i * ldx + k] = v;
  }
  const long long ntiles = (P.N + TN - 1) / TN;
  // which (column n, row k) elements of a slab this thread stages: independent of the tile, so the divisions happen once
  constexpr int MAXE = (TN * 64 + NT - 1) / NT;
  unsigned short en[MAXE], ek[MAXE];
#pragma unroll
  for (int i = 0; i < MAXE; ++i) {
    const int e = tid + i * NT;
    int n = 0, k = 0x7fff;  // k out of range -> zero fill
    if (e < TN * kpad) {
      if (n_fast) {
        k = e / TN;
        n = e - k * TN;
      } else {
        n = e / kpad;
        k = e - n * kpad;
      }
    }
    en[i] = (unsigned short)n;
    ek[i] = (unsigned short)k;
  }
  const int nelem = (TN * kpad + NT - 1) / NT;
  auto issue = [&](long long tile, int buf) {
    double2 *Xs = Xs0 + (size_t)buf * TN * ldx;
    const long long n0 = tile * TN;
    const long long r0 = n0 / P.L;                       // one 64-bit division per tile (warp-uniform)
    const unsigned l0 = (unsigned)(n0 - r0 * P.L);
#pragma unroll
    for (int i = 0; i < MAXE; ++i) {
      if (i < nelem && ek[i] != 0x7fff) {
        const int n = en[i], k = ek[i];
        const bool valid = tile < ntiles && n0 + n < P.N && k < P.dr;
        long long src = 0;
        if (valid) {
          unsigned l;
          long long r;
          if (P.L == 1) {
            l = 0;
            r = n0 + n;
          } else {
            split_col((unsigned)n, l0, r0, P.L, l, r);
          }
          src = (long long)l + P.L * ((long long)k + (long long)P.dr * r);
        }
        cp_async16_zfill(Xs + n * ldx + k, x + src, valid);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  long long tile = blockIdx.x;
  int buf = 0;
  issue(tile, 0);
  for (; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    issue(tile + gridDim.x, buf ^ 1);                      // prefetch the next slab (an empty group past the end)
    asm volatile("cp.async.wait_group 1;\n" ::);           // this tile's slab has landed
    __syncthreads();
    const double2 *Xs = Xs0 + (size_t)buf * TN * ldx;
    const long long n0 = tile * TN;
    const long long er0 = n0 / P.L;
    const unsigned el0 = (unsigned)(n0 - er0 * P.L);
    double cre[2][NB][2], cim[2][NB][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < NB; ++b) cre[a][b][0] = cre[a][b][1] = cim[a][b][0] = cim[a][b][1] = 0.0;
#pragma unroll 2
    for (int kk = 0; kk < kpad; kk += 4) {
      double2 a[2], b[NB];
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) a[mb] = As[(wm * 16 + mb * 8 + g) * ldx + kk + t];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) b[nb] = Xs[(wn * 8 * NB + nb * 8 + g) * ldx + kk + t];
      // two sweeps over the 8 accumulators: the two updates of one accumulator are 8 DMMAs apart, so no DMMA waits
      // for its predecessor (back-to-back dependent DMMAs halved the pipe utilisation: ncu 50 % -> see profiles/)
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          dmma884(cre[mb][nb][0], cre[mb][nb][1], a[mb].x, b[nb].x);
          dmma884(cim[mb][nb][0], cim[mb][nb][1], a[mb].x, b[nb].y);
        }
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          dmma884(cre[mb][nb][0], cre[mb][nb][1], -a[mb].y, b[nb].y);
          dmma884(cim[mb][nb][0], cim[mb][nb][1], a[mb].y, b[nb].x);
        }
    }
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
      const int i = wm * 16 + mb * 8 + g;
      if (i >= P.dl) continue;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const unsigned j = wn * 8 * NB + nb * 8 + 2 * t + e;
          if (n0 + j >= P.N) continue;
          unsigned l;
          long long r;
          split_col(j, el0, er0, P.L, l, r);
          const long long addr = (long long)l + P.L * ((long long)i + (long long)P.dl * r);
          const double ar = cre[mb][nb][e], ai = cim[mb][nb][e];
          double2 o = make_double2(P.alpha.x * ar - P.alpha.y * ai, P.alpha.x * ai + P.alpha.y * ar);
          if (!P.beta_zero) {
            const double2 yo = y[addr];
            o.x += P.beta.x * yo.x - P.beta.y * yo.y;
            o.y += P.beta.x * yo.y + P.beta.y * yo.x;
          }
          y[addr] = o;
        }
    }
    __syncthreads();  // everyone is done with this buffer before the next prefetch overwrites it
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
}

template <int MW, int NB>
static int launch_axis_pipe(const AxisParams &P, const double2 *xp, double2 *yp, int sms, cudaStream_t s) {
  constexpr int TN = 32 * NB;
  const int kpad = (P.dr + 3) / 4 * 4;
  int ldx = kpad;
  while (ldx % 8 != 4) ++ldx;  // conflict-free LDS.128 fragment loads
  const size_t smem = (size_t)(16 * MW + 2 * TN) * ldx * sizeof(double2);
  {
    // the opt-in is per device: remember the largest size configured on each
    static std::atomic<size_t> configured[QOB_MAX_DEVICES];
    int dev = 0;
    QOB_CUDA(cudaGetDevice(&dev));
    const int slot = dev >= 0 && dev < QOB_MAX_DEVICES ? dev : 0;
    if (dev != slot || smem > configured[slot].load(std::memory_order_relaxed)) {
      QOB_CUDA(cudaFuncSetAttribute(axis_dmma_pipe_kernel<MW, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured[slot].store(smem, std::memory_order_relaxed);
    }
  }
  const int per_sm = (NB == 1 && MW <= 3 && 2 * (smem + 1024) <= 227 * 1024) ? 2 : 1;
  const long long ntiles = (P.N + TN - 1) / TN;
  const long long grid = std::min<long long>(ntiles, (long long)sms * per_sm);
  axis_dmma_pipe_kernel<MW, NB><<<(unsigned)grid, MW * 128, smem, s>>>(P, xp, yp, kpad, ldx);
  return QOB_STATUS_OK;
}

template <int MW>
static int launch_axis_wk(const AxisParams &P, const double2 *xp, double2 *yp, cudaStream_t s) {
  const int sms = qob_device_sm_count();
  static const int nb_env = getenv("QOB_AXIS_NB") ? atoi(getenv("QOB_AXIS_NB")) : 1;
  if (nb_env == 2) return launch_axis_pipe<MW, 2>(P, xp, yp, sms, s);
  return launch_axis_pipe<MW, 1>(P, xp, yp, sms, s);
}

int prepare_axis_matrix(const HostMat &m, AxisMatrixDev &out) {
  out.dl = (int)m.rows;
  out.dr = (int)m.cols;
  out.dl_pad = (out.dl + 15) / 16 * 16;
  out.dr_pad = (out.dr + AX_KT - 1) / AX_KT * AX_KT;
  std::vector<double> h((size_t)2 * out.dl_pad * out.dr_pad, 0.0);
  const size_t plane = (size_t)out.dl_pad * out.dr_pad;
  for (int i = 0; i < out.dl; ++i)
    for (int j = 0; j < out.dr; ++j) {
      cplx v = m.at(i, j);
      h[(size_t)i * out.dr_pad + j] = v.real();
      h[plane + (size_t)i * out.dr_pad + j] = v.imag();
    }
  return out.planes.upload(h);
}

int launch_axis_dense(const AxisMatrixDev &A, int64_t L, int64_t R, cplx alpha, const void *x, cplx beta, void *y,
                      cudaStream_t s) {
  AxisParams P;
  P.a_re = A.planes.ptr;
  P.a_im = A.planes.ptr + (size_t)A.dl_pad * A.dr_pad;
  P.dl = A.dl;
  P.dr = A.dr;
  P.dl_pad = A.dl_pad;
  P.dr_pad = A.dr_pad;
  P.L = L;
  P.R = R;
  P.N = L * R;
  P.alpha = make_double2(alpha.real(), alpha.imag());
  P.beta = make_double2(beta.real(), beta.imag());
  P.beta_zero = beta == cplx(0.0, 0.0);
  if (P.N == 0 || A.dl == 0) return QOB_STATUS_OK;
  const double2 *xp = (const double2 *)x;
  double2 *yp = (double2 *)y;
#define AX_LAUNCH(MW, NW)                                                            \
  do {                                                                               \
    constexpr int TM = 16 * MW, TN = 32 * NW;                                        \
    long long gx = (P.N + TN - 1) / TN;                                              \
    if (gx > 0x7FFFFFFFll) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "axis kernel: grid too large"); \
    dim3 grid((unsigned)gx, (unsigned)((A.dl + TM - 1) / TM));                       \
    axis_dmma_kernel<MW, NW><<<grid, 32 * MW * NW, 0, s>>>(P, xp, yp);               \
  } while (0)
  if (A.dl <= 64 && A.dr <= 64 && !getenv("QOB_AXIS_V1")) {
    if (A.dl <= 16) QOB_TRY(launch_axis_wk<1>(P, xp, yp, s));
    else if (A.dl <= 32) QOB_TRY(launch_axis_wk<2>(P, xp, yp, s));
    else if (A.dl <= 48) QOB_TRY(launch_axis_wk<3>(P, xp, yp, s));
    else QOB_TRY(launch_axis_wk<4>(P, xp, yp, s));
    QOB_LAUNCHED();
    QOB_CUDA(cudaGetLastError());
    return QOB_STATUS_OK;
  }
  // static shared memory must stay under 48 KiB: (TM + TN) * AX_LD * 16 B
  if (A.dl <= 16)
    AX_LAUNCH(1, 4);
  else if (A.dl <= 32)
    AX_LAUNCH(2, 3);
  else if (A.dl <= 48)
    AX_LAUNCH(3, 2);
  else
    AX_LAUNCH(4, 2);
#undef AX_LAUNCH
  QOB_LAUNCHED();
  QOB_CUDA(cudaGetLastError());
  return QOB_STATUS_OK;
}

// ------------------------------------------------------------------------------------------ SpMM
// SparseOperator x dense (src/operators_sparse.jl:199-202 -> gemm!/gemv!, src/sparsematrix.jl:99-238).
// Julia's dense operands are column-major, so the coalesced direction is the ROW index of the dense
// matrices: one thread per output element with lanes running down a column ("thread-per-row-element";
// a warp-per-row split would read B with stride k).  The reference's scatter over CSC columns
// (R[row, j] += v*B[col, j]) becomes a gather over CSR rows: no atomics, no beta pre-pass.
__global__ void __launch_bounds__(256)
    spmm_left_kernel(const int *__restrict__ rowptr, const int *__restrict__ colidx, const double2 *__restrict__ val,
                     long long m, long long k, long long n, double2 alpha, double2 beta, int beta_zero,
                     const double2 *__restrict__ B, double2 *__restrict__ R) {
  const long long total = m * n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long c = idx / m, i = idx - c * m;
    const double2 *b = B + c * k;
    double re = 0.0, im = 0.0;
    for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) {
      double2 v = val[p], w = b[colidx[p]];
      re = fma(v.x, w.x, re);
      re = fma(-v.y, w.y, re);
      im = fma(v.x, w.y, im);
      im = fma(v.y, w.x, im);
    }
    double2 o = make_double2(alpha.x * re - alpha.y * im, alpha.x * im + alpha.y * re);
    if (!beta_zero) {
      double2 yo = R[idx];
      o.x += beta.x * yo.x - beta.y * yo.y;
      o.y += beta.x * yo.y + beta.y * yo.x;
    }
    R[idx] = o;
  }
}

__global__ void __launch_bounds__(256)
    spmm_right_kernel(const int *__restrict__ colptr, const int *__restrict__ rowidx, const double2 *__restrict__ val,
                      long long q, long long m, long long n, double2 alpha, double2 beta, int beta_zero,
                      const double2 *__restrict__ B, double2 *__restrict__ R) {
  const long long total = q * n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long c = idx / q, r = idx - c * q;
    double re = 0.0, im = 0.0;
    for (int p = colptr[c]; p < colptr[c + 1]; ++p) {
      double2 v = val[p], w = B[r + q * (long long)rowidx[p]];
      re = fma(v.x, w.x, re);
      re = fma(-v.y, w.y, re);
      im = fma(v.x, w.y, im);
      im = fma(v.y, w.x, im);
    }
    double2 o = make_double2(alpha.x * re - alpha.y * im, alpha.x * im + alpha.y * re);
    if (!beta_zero) {
      double2 yo = R[idx];
      o.x += beta.x * yo.x - beta.y * yo.y;
      o.y += beta.x * yo.y + beta.y * yo.x;
    }
    R[idx] = o;
  }
}

// Large operands (the bandwidth regime): several outputs per thread.  One output per thread leaves a single 16-byte load in
// flight behind a rowptr -> colidx -> B dependency chain, far below the bytes in flight HBM3e needs; here a thread
// reads its CSR row (CSC column) ONCE and gathers for CPT columns (RPT rows) with independent loads.  blockIdx.x =
// (column group, row block) with the row block fastest, so CTAs running together work on the same few columns of B.
template <int CPT>
__global__ void __launch_bounds__(256)
    spmm_left_multi_kernel(const int *__restrict__ rowptr, const int *__restrict__ colidx, const double2 *__restrict__ val,
                           long long m, long long k, long long n, double2 alpha, double2 beta, int beta_zero,
                           const double2 *__restrict__ B, double2 *__restrict__ R, unsigned row_blocks) {
  const unsigned rb = blockIdx.x % row_blocks, cg = blockIdx.x / row_blocks;
  const long long i = (long long)rb * 256 + threadIdx.x;
  if (i >= m) return;
  const long long c0 = (long long)cg * CPT;
  const int nc = (int)min((long long)CPT, n - c0);
  const int p0 = rowptr[i], p1 = rowptr[i + 1];
  double re[CPT], im[CPT];
#pragma unroll
  for (int u = 0; u < CPT; ++u) re[u] = im[u] = 0.0;
  const double2 *b = B + c0 * k;
  for (int p = p0; p < p1; ++p) {
    const double2 v = val[p];
    const double2 *bp = b + colidx[p];
    double2 w[CPT];
#pragma unroll
    for (int u = 0; u < CPT; ++u)
      if (u < nc) w[u] = bp[(long long)u * k];
#pragma unroll
    for (int u = 0; u < CPT; ++u)
      if (u < nc) {
        re[u] = fma(v.x, w[u].x, re[u]);
        re[u] = fma(-v.y, w[u].y, re[u]);
        im[u] = fma(v.x, w[u].y, im[u]);
        im[u] = fma(v.y, w[u].x, im[u]);
      }
  }
  double2 *r = R + c0 * m + i;
  double2 yo[CPT];
  if (!beta_zero) {
#pragma unroll
    for (int u = 0; u < CPT; ++u)
      if (u < nc) yo[u] = r[(long long)u * m];
  }
#pragma unroll
  for (int u = 0; u < CPT; ++u)
    if (u < nc) {
      double2 o = make_double2(alpha.x * re[u] - alpha.y * im[u], alpha.x * im[u] + alpha.y * re[u]);
      if (!beta_zero) {
        o.x += beta.x * yo[u].x - beta.y * yo[u].y;
        o.y += beta.x * yo[u].y + beta.y * yo[u].x;
      }
      r[(long long)u * m] = o;
    }
}

// `order` (may be null) is the sequence in which the output columns are processed: a Cuthill-McKee ordering of the
// operator's graph, so that the columns of B an output column gathers from are the ones its neighbours in time gather
// from too and come out of L2 (for a Jaynes-Cummings H the partner column is ~D/2 columns = hundreds of MB away).
template <int RPT>
__global__ void __launch_bounds__(256)
    spmm_right_multi_kernel(const int *__restrict__ colptr, const int *__restrict__ rowidx, const double2 *__restrict__ val,
                            long long q, long long m, long long n, double2 alpha, double2 beta, int beta_zero,
                            const double2 *__restrict__ B, double2 *__restrict__ R, unsigned row_blocks,
                            const int *__restrict__ order) {
  const unsigned rb = blockIdx.x % row_blocks, cb = blockIdx.x / row_blocks;
  const long long c = order ? order[cb] : cb;
  const long long r0 = (long long)rb * (256 * RPT) + threadIdx.x;
  const int p0 = colptr[c], p1 = colptr[c + 1];
  double re[RPT], im[RPT];
#pragma unroll
  for (int u = 0; u < RPT; ++u) re[u] = im[u] = 0.0;
  for (int p = p0; p < p1; ++p) {
    const double2 v = val[p];
    const double2 *bp = B + q * (long long)rowidx[p] + r0;
    double2 w[RPT];
#pragma unroll
    for (int u = 0; u < RPT; ++u)
      if (r0 + u * 256 < q) w[u] = bp[u * 256];
#pragma unroll
    for (int u = 0; u < RPT; ++u)
      if (r0 + u * 256 < q) {
        re[u] = fma(v.x, w[u].x, re[u]);
        re[u] = fma(-v.y, w[u].y, re[u]);
        im[u] = fma(v.x, w[u].y, im[u]);
        im[u] = fma(v.y, w[u].x, im[u]);
      }
  }
  double2 *r = R + c * q + r0;
  double2 yo[RPT];
  if (!beta_zero) {
#pragma unroll
    for (int u = 0; u < RPT; ++u)
      if (r0 + u * 256 < q) yo[u] = r[u * 256];
  }
#pragma unroll
  for (int u = 0; u < RPT; ++u)
    if (r0 + u * 256 < q) {
      double2 o = make_double2(alpha.x * re[u] - alpha.y * im[u], alpha.x * im[u] + alpha.y * re[u]);
      if (!beta_zero) {
        o.x += beta.x * yo[u].x - beta.y * yo[u].y;
        o.y += beta.x * yo[u].y + beta.y * yo[u].x;
      }
      r[u * 256] = o;
    }
}

static const int64_t kSpmmFill = 148 * 2048;   // outputs that give every SM a full complement of threads

int launch_spmm_left(const SparseDev &csr, int64_t m, int64_t k, int64_t n, cplx alpha, const void *B, cplx beta,
                     void *R, cudaStream_t s) {
  if (m * n == 0) return QOB_STATUS_OK;
  const double2 a2 = make_double2(alpha.real(), alpha.imag()), b2 = make_double2(beta.real(), beta.imag());
  const int bz = beta == cplx(0.0, 0.0);
  const int cpt = (n >= 4 && m * n >= 4 * kSpmmFill) ? 4 : (n >= 2 && m * n >= 2 * kSpmmFill) ? 2 : 1;
  const int64_t row_blocks = (m + 255) / 256, blocks_multi = row_blocks * ((n + cpt - 1) / cpt);
  if (m >= 256 && cpt > 1 && blocks_multi < ((int64_t)1 << 31)) {
    if (cpt == 4)
      spmm_left_multi_kernel<4><<<(unsigned)blocks_multi, 256, 0, s>>>(csr.ptr.ptr, csr.idx.ptr, csr.val.ptr, m, k, n, a2, b2, bz,
                                                                       (const double2 *)B, (double2 *)R, (unsigned)row_blocks);
    else
      spmm_left_multi_kernel<2><<<(unsigned)blocks_multi, 256, 0, s>>>(csr.ptr.ptr, csr.idx.ptr, csr.val.ptr, m, k, n, a2, b2, bz,
                                                                       (const double2 *)B, (double2 *)R, (unsigned)row_blocks);
  } else {
    int64_t blocks = std::min<int64_t>((m * n + 255) / 256, (int64_t)1 << 30);
    spmm_left_kernel<<<(unsigned)blocks, 256, 0, s>>>(csr.ptr.ptr, csr.idx.ptr, csr.val.ptr, m, k, n, a2, b2, bz,
                                                      (const double2 *)B, (double2 *)R);
  }
  QOB_LAUNCHED();
  QOB_CUDA(cudaGetLastError());
  return QOB_STATUS_OK;
}

int launch_spmm_right(const SparseDev &csc, int64_t q, int64_t m, int64_t n, cplx alpha, const void *B, cplx beta,
                      void *R, cudaStream_t s) {
  if (q * n == 0) return QOB_STATUS_OK;
  const double2 a2 = make_double2(alpha.real(), alpha.imag()), b2 = make_double2(beta.real(), beta.imag());
  const int bz = beta == cplx(0.0, 0.0);
  const int rpt = (q >= 1024 && q * n >= 4 * kSpmmFill) ? 4 : (q >= 512 && q * n >= 2 * kSpmmFill) ? 2 : 1;
  const int64_t row_blocks = (q + 256 * rpt - 1) / (256 * rpt), blocks_multi = row_blocks * n;
  if (rpt > 1 && blocks_multi < ((int64_t)1 << 31)) {
    const int *order = csc.order.n == (size_t)n ? csc.order.ptr : nullptr;
    if (rpt == 4)
      spmm_right_multi_kernel<4><<<(unsigned)blocks_multi, 256, 0, s>>>(csc.ptr.ptr, csc.idx.ptr, csc.val.ptr, q, m, n, a2, b2, bz,
                                                                        (const double2 *)B, (double2 *)R, (unsigned)row_blocks, order);
    else
      spmm_right_multi_kernel<2><<<(unsigned)blocks_multi, 256, 0, s>>>(csc.ptr.ptr, csc.idx.ptr, csc.val.ptr, q, m, n, a2, b2, bz,
                                                                        (const double2 *)B, (double2 *)R, (unsigned)row_blocks, order);
  } else {
    int64_t blocks = std::min<int64_t>((q * n + 255) / 256, (int64_t)1 << 30);
    spmm_right_kernel<<<(unsigned)blocks, 256, 0, s>>>(csc.ptr.ptr, csc.idx.ptr, csc.val.ptr, q, m, n, a2, b2, bz,
                                                       (const double2 *)B, (double2 *)R);
  }
  QOB_LAUNCHED();
  QOB_CUDA(cudaGetLastError());
  return QOB_STATUS_OK;
}
