// qob_kernels_dtile.cu — mixed-radix tile passes: the fused LazySum kernel for LARGE states on subsystems of ANY
// dimension (Fock, spin-1, N-level ... lattices), the counterpart of qob_kernels_qtile.cu for non-qubit systems.
//
// Reference semantics: y = beta*y + alpha * sum_t coef_t (A_t x), A_t = (x)_k A_{t,k} on a few tensor axes
// (src/operators_lazysum.jl:189-200 looping over src/operators_lazytensor.jl:539-557 / :612-751), subsystem 1 fastest.
//
// Idea.  Every site factor is split into its shifted diagonals A^(s)[i] = A[i, i+s]; a term then is a sum of
// "components" that each read ONE source amplitude per output amplitude:
//        y[..i_k..] += coef * prod_k A_k^(s_k)[i_k] * x[..i_k+s_k..]
// (number, destroy, create, sigma+-, transition operators have one diagonal; sigmax has two).  A pass picks a set of
// FREE axes — a low contiguous block (coalesced runs) plus a window of higher axes — whose joint extent fits a
// shared-memory tile; all terms whose off-diagonal factors lie on free axes are accumulated in that pass from the tile,
// factors on the other (fixed) axes are diagonal there and fold into one weight per component and CTA.
// The state is streamed once per pass (read x + read/write y) instead of being gathered from L2 with a mixed-radix
// index decode per term and amplitude (qob_kernels_gather.cu), which is ALU/latency bound at a few % of HBM peak.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <type_traits>

#include "qob_internal.h"

#define DT_MAXF 4        // factors per component (free + fixed each)
#define DT_MAXAX 64      // tensor axes (incl. the batch axis)
#define DT_THREADS 256
#define DT_PTHREADS 512   // persistent variant: one CTA per SM, one sweep per tile
#define DT_U 8           // outputs per thread and sweep
#define DT_TILE_CAP 4096 // amplitudes per tile (64 KiB)

struct DCompDev {
  int delta;                        // source element = output element + delta (tile numbering)
  int coef;                         // index into the per-term coefficient array
  unsigned char nfree, nfixed, pad0, pad1;
  unsigned int f_sh[DT_MAXF];       // position of the factor's digit in the packed digit word of an output element
  unsigned int f_mask[DT_MAXF];     // ... and its mask
  unsigned short f_d[DT_MAXF];      // axis dimension
  unsigned short f_tab[DT_MAXF];    // weight table (entries) indexed by the output digit on a free axis
  unsigned short x_tab[DT_MAXF];    // same for fixed axes
  unsigned char x_slot[DT_MAXF];    // which fixed axis
  int pad2[2];
};
static_assert(sizeof(DCompDev) == 80 && sizeof(DCompDev) % 16 == 0, "component records are copied in 16-byte units");

struct DPassParams {
  const DCompDev *comps;
  const double2 *tables;
  const double2 *coef;
  const uint2 *etab;                // per tile element: {packed free-axis digits, offset from the tile base}
  int ncomp, ntab;
  int n_off, n_dfree;               // components [0, n_off): off-diagonal; [n_off, n_off+n_dfree): diagonal with free
                                    // factors; the rest: diagonal and uniform over the tile
  int tile, R;                      // amplitudes per tile; length of the contiguous low run
  int nfixed;
  unsigned int fx_dim[DT_MAXAX];    // fixed axes: dimension, product of the dimensions below (tile-id radix), stride
  unsigned int fx_below[DT_MAXAX];
  long long fx_stride[DT_MAXAX];
  double2 alpha, beta;
  int mode;                         // 0: y = alpha*acc; 1: y = alpha*acc + beta*y
};

__device__ __forceinline__ double2 dmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void dfma(double2 &acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}

// Per tile element (the same for every tile of a pass, built on the host, read through L1): the digits of the element on
// the free axes packed into one word (sum of ceil(log2 d) <= 24 bits for a tile of <= 4096 amplitudes; a bit-field
// extract per factor lookup) and its offset from the tile base in the state.
// REALT: every weight table of the pass is real (number, destroy, create, sigma+-, ... are): tables are stored as
// 8-byte doubles (half the shared-memory wavefronts of a 16-byte lookup) and the weight product is real arithmetic.
// REALW (needs REALT): the coefficients are real as well, so the whole weight is one double (2 DFMA per gather).
// FULL: the tile is a multiple of the sweep (threads x outputs per thread), no range guards.
// One sweep: NT threads x DT_U outputs per thread, element e_k = sweep + k*NT + tid (pk/eo: its packed digits / offset).
template <bool REALT, bool REALW, bool FULL, int NT>
__device__ __forceinline__ void dt_sweep(const DPassParams &P, const double2 *xs,
                                         const typename std::conditional<REALT, double, double2>::type *tab,
                                         const typename std::conditional<REALW, double, double2>::type *cw,
                                         const DCompDev *comps, unsigned sweep, unsigned tid, long long base,
                                         const unsigned (&pk)[DT_U], const unsigned (&eo)[DT_U], double2 *__restrict__ y) {
  typedef typename std::conditional<REALT, double, double2>::type TabT;
  typedef typename std::conditional<REALW, double, double2>::type WT;
  double2 acc[DT_U];
#pragma unroll
  for (int k = 0; k < DT_U; ++k) acc[k] = make_double2(0.0, 0.0);
  const unsigned e0 = sweep + tid;
  // ---- diagonal components (source = the output element itself): their weights are summed first — one lookup and
  // one multiply-add each — and applied with a single multiply per output; the ones without a free factor are
  // uniform over the tile
  if (P.n_off < P.ncomp) {
    WT wu;
    if constexpr (REALW) wu = 0.0;
    else wu = make_double2(0.0, 0.0);
    for (int c = P.n_off + P.n_dfree; c < P.ncomp; ++c) {
      if constexpr (REALW) wu += cw[c];
      else {
        wu.x += cw[c].x;
        wu.y += cw[c].y;
      }
    }
    WT wd[DT_U];
#pragma unroll
    for (int k = 0; k < DT_U; ++k) wd[k] = wu;
    for (int c = P.n_off; c < P.n_off + P.n_dfree; ++c) {
      const DCompDev &C = comps[c];
      const WT w0 = cw[c];
      const int nfree = C.nfree;
      if (nfree == 1) {   // one lookup (also two neighbouring factors folded into a joint table on the host)
        const unsigned s0 = C.f_sh[0], k0 = C.f_mask[0];
        const TabT *t0 = tab + C.f_tab[0];
#pragma unroll
        for (int k = 0; k < DT_U; ++k) {
          if (FULL || pk[k] != 0xffffffffu) {
            const unsigned q0 = (pk[k] >> s0) & k0;
            if constexpr (REALT) {
              const double t = t0[q0];
              if constexpr (REALW) wd[k] = fma(w0, t, wd[k]);
              else {
                wd[k].x = fma(w0.x, t, wd[k].x);
                wd[k].y = fma(w0.y, t, wd[k].y);
              }
            } else {
              dfma(wd[k], w0, t0[q0]);
            }
          }
        }
      } else if (nfree == 2) {
        const unsigned s0 = C.f_sh[0], k0 = C.f_mask[0], s1 = C.f_sh[1], k1 = C.f_mask[1];
        const TabT *t0 = tab + C.f_tab[0], *t1 = tab + C.f_tab[1];
#pragma unroll
        for (int k = 0; k < DT_U; ++k) {
          if (FULL || pk[k] != 0xffffffffu) {
            const unsigned q0 = (pk[k] >> s0) & k0, q1 = (pk[k] >> s1) & k1;
            if constexpr (REALT) {
              const double t = t0[q0] * t1[q1];
              if constexpr (REALW) wd[k] = fma(w0, t, wd[k]);
              else {
                wd[k].x = fma(w0.x, t, wd[k].x);
                wd[k].y = fma(w0.y, t, wd[k].y);
              }
            } else {
              dfma(wd[k], w0, dmul(t0[q0], t1[q1]));
            }
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < DT_U; ++k) {
          if (FULL || pk[k] != 0xffffffffu) {
            double2 w;
            if constexpr (REALW) w = make_double2(w0, 0.0);
            else w = w0;
#pragma unroll 1
            for (int f = 0; f < nfree; ++f) {
              const unsigned q = (pk[k] >> C.f_sh[f]) & C.f_mask[f];
              if constexpr (REALT) {
                const double t = tab[C.f_tab[f] + q];
                w.x *= t;
                w.y *= t;
              } else {
                w = dmul(w, tab[C.f_tab[f] + q]);
              }
            }
            if constexpr (REALW) wd[k] += w.x;
            else {
              wd[k].x += w.x;
              wd[k].y += w.y;
            }
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < DT_U; ++k) {
      if (FULL || pk[k] != 0xffffffffu) {
        if constexpr (REALW) {
          if (wd[k] != 0.0) {
            const double2 xv = xs[e0 + k * NT];
            acc[k].x = fma(wd[k], xv.x, acc[k].x);
            acc[k].y = fma(wd[k], xv.y, acc[k].y);
          }
        } else {
          if (wd[k].x != 0.0 || wd[k].y != 0.0) dfma(acc[k], wd[k], xs[e0 + k * NT]);
        }
      }
    }
  }
  for (int c = 0; c < P.n_off; ++c) {
    const DCompDev &C = comps[c];
    const WT w0 = cw[c];
    const int delta = C.delta, nfree = C.nfree;
    const double2 *xsrc = xs + (int)e0 + delta;
    if (nfree == 0) {
#pragma unroll
      for (int k = 0; k < DT_U; ++k) {
        if (FULL || pk[k] != 0xffffffffu) {
          const double2 xv = xsrc[k * NT];
          if constexpr (REALW) {
            acc[k].x = fma(w0, xv.x, acc[k].x);
            acc[k].y = fma(w0, xv.y, acc[k].y);
          } else {
            dfma(acc[k], w0, xv);
          }
        }
      }
    } else if (nfree == 1) {   // one lookup (also two neighbouring factors folded into a joint table on the host)
      const unsigned s0 = C.f_sh[0], k0 = C.f_mask[0];
      const TabT *t0 = tab + C.f_tab[0];
#pragma unroll
      for (int k = 0; k < DT_U; ++k) {
        if (FULL || pk[k] != 0xffffffffu) {
          const unsigned q0 = (pk[k] >> s0) & k0;
          if constexpr (REALT) {
            const double t = t0[q0];
            if (t != 0.0) {
              const double2 xv = xsrc[k * NT];
              if constexpr (REALW) {
                const double w = w0 * t;
                acc[k].x = fma(w, xv.x, acc[k].x);
                acc[k].y = fma(w, xv.y, acc[k].y);
              } else {
                dfma(acc[k], make_double2(w0.x * t, w0.y * t), xv);
              }
            }
          } else {
            const double2 w = dmul(w0, t0[q0]);
            if (w.x != 0.0 || w.y != 0.0) dfma(acc[k], w, xsrc[k * NT]);
          }
        }
      }
    } else if (nfree == 2) {
      const unsigned s0 = C.f_sh[0], k0 = C.f_mask[0];
      const unsigned s1 = C.f_sh[1], k1 = C.f_mask[1];
      const TabT *t0 = tab + C.f_tab[0], *t1 = tab + C.f_tab[1];
#pragma unroll
      for (int k = 0; k < DT_U; ++k) {
        if (FULL || pk[k] != 0xffffffffu) {
          const unsigned q0 = (pk[k] >> s0) & k0, q1 = (pk[k] >> s1) & k1;
          if constexpr (REALT) {
            const double t = t0[q0] * t1[q1];
            if (t != 0.0) {
              const double2 xv = xsrc[k * NT];
              if constexpr (REALW) {
                const double w = w0 * t;
                acc[k].x = fma(w, xv.x, acc[k].x);
                acc[k].y = fma(w, xv.y, acc[k].y);
              } else {
                dfma(acc[k], make_double2(w0.x * t, w0.y * t), xv);
              }
            }
          } else {
            const double2 w = dmul(w0, dmul(t0[q0], t1[q1]));
            if (w.x != 0.0 || w.y != 0.0) dfma(acc[k], w, xsrc[k * NT]);
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < DT_U; ++k) {
        if (FULL || pk[k] != 0xffffffffu) {
          double2 w;
          if constexpr (REALW) w = make_double2(w0, 0.0);
          else w = w0;
#pragma unroll 1
          for (int f = 0; f < nfree; ++f) {
            const unsigned q = (pk[k] >> C.f_sh[f]) & C.f_mask[f];
            if constexpr (REALT) {
              const double t = tab[C.f_tab[f] + q];
              w.x *= t;
              w.y *= t;
            } else {
              w = dmul(w, tab[C.f_tab[f] + q]);
            }
          }
          if (w.x != 0.0 || w.y != 0.0) dfma(acc[k], w, xsrc[k * NT]);
        }
      }
    }
  }
  // epilogue in batches of 4: the loads of y first (independent, in flight together), then the stores
#pragma unroll
  for (int kb = 0; kb < DT_U; kb += 4) {
    long long goff[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) goff[j] = (FULL || pk[kb + j] != 0xffffffffu) ? base + (long long)eo[kb + j] : -1;
    if (P.mode) {
      double2 yv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (FULL || goff[j] >= 0) yv[j] = y[goff[j]];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (FULL || goff[j] >= 0) {
          double2 o = dmul(P.alpha, acc[kb + j]);
          dfma(o, P.beta, yv[j]);
          y[goff[j]] = o;
        }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (FULL || goff[j] >= 0) y[goff[j]] = dmul(P.alpha, acc[kb + j]);
    }
  }
}

// the x tile -> shared memory (16-byte cp.async per amplitude, contiguous runs of R amplitudes); the y lines of a
// read-modify-write pass are pulled into L2 at the same time
template <bool FULL, int NT>
__device__ __forceinline__ void dt_issue_loads(const DPassParams &P, double2 *xs, unsigned tid, long long base,
                                               const double2 *__restrict__ x, const double2 *__restrict__ y) {
  for (unsigned e0 = tid; e0 < (unsigned)P.tile; e0 += 8 * NT) {
    unsigned eo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {   // the table lookups first (independent), then the copies
      const unsigned e = e0 + j * NT;
      eo[j] = (FULL || e < (unsigned)P.tile) ? __ldg(&P.etab[e]).y : 0xffffffffu;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned e = e0 + j * NT;
      if (FULL || e < (unsigned)P.tile) {
        const long long off = base + (long long)eo[j];
        const unsigned sa = (unsigned)__cvta_generic_to_shared(xs + e);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(x + off) : "memory");
        if (P.mode && (off & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(y + off));
      }
    }
  }
}

template <bool REALT, bool REALW, bool FULL>
__global__ void __launch_bounds__(DT_THREADS, 3) dtile_kernel(const __grid_constant__ DPassParams P,
                                                              const double2 *__restrict__ x, double2 *__restrict__ y) {
  typedef typename std::conditional<REALT, double, double2>::type TabT;
  typedef typename std::conditional<REALW, double, double2>::type WT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *xs = reinterpret_cast<double2 *>(smem_raw);
  double2 *cw2 = xs + P.tile;
  WT *cw = reinterpret_cast<WT *>(cw2);
  DCompDev *comps = reinterpret_cast<DCompDev *>(cw2 + P.ncomp);
  TabT *tab = reinterpret_cast<TabT *>(comps + P.ncomp);   // 16-byte aligned: records are 80 bytes
  __shared__ unsigned int fdig[DT_MAXAX];
  __shared__ long long s_base;
  const unsigned tid = threadIdx.x;

  // ---- which tile: digits of the fixed axes (one thread per axis), base offset of the tile
  if (tid < 32) {
    long long part = 0;
    for (unsigned a = tid; a < (unsigned)P.nfixed; a += 32) {
      const unsigned dg = (blockIdx.x / P.fx_below[a]) % P.fx_dim[a];
      fdig[a] = dg;
      part += (long long)dg * P.fx_stride[a];
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (tid == 0) s_base = part;
  }
  // ---- operator data of this pass -> shared memory
  for (int i = tid; i < P.ntab; i += DT_THREADS) {
    if constexpr (REALT) tab[i] = P.tables[i].x;
    else tab[i] = P.tables[i];
  }
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(P.comps);
    uint4 *dst = reinterpret_cast<uint4 *>(comps);
    const int n16 = P.ncomp * (int)(sizeof(DCompDev) / 16);
    for (int i = tid; i < n16; i += DT_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const long long base = s_base;
  dt_issue_loads<FULL, DT_THREADS>(P, xs, tid, base, x, y);
  // ---- per-component weight that is uniform over the tile: coefficient x diagonal factors on fixed axes
  for (int c = tid; c < P.ncomp; c += DT_THREADS) {
    const DCompDev &C = comps[c];
    double2 w = P.coef[C.coef];
    for (int f = 0; f < C.nfixed; ++f) {
      if constexpr (REALT) {
        const double t = tab[C.x_tab[f] + fdig[C.x_slot[f]]];
        w.x *= t;
        w.y *= t;
      } else {
        w = dmul(w, tab[C.x_tab[f] + fdig[C.x_slot[f]]]);
      }
    }
    if constexpr (REALW) cw[c] = w.x;
    else cw[c] = w;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  for (unsigned sweep = 0; sweep < (unsigned)P.tile; sweep += DT_THREADS * DT_U) {
    unsigned pk[DT_U], eo[DT_U];
#pragma unroll
    for (int k = 0; k < DT_U; ++k) {
      const unsigned e = sweep + k * DT_THREADS + tid;
      const uint2 t = (FULL || e < (unsigned)P.tile) ? __ldg(&P.etab[e]) : make_uint2(0xffffffffu, 0u);
      pk[k] = t.x;   // 0xffffffff: out of range, no lookups (guards in dt_sweep)
      eo[k] = t.y;
    }
    dt_sweep<REALT, REALW, FULL, DT_THREADS>(P, xs, tab, cw, comps, sweep, tid, base, pk, eo, y);
  }
}

// Persistent variant: one CTA of DT_PTHREADS threads per SM walks the tiles with TWO tile buffers — the copies of tile
// t+1 (and the L2 prefetch of its y lines) are in flight while tile t is computed, so the memory pipe never waits for
// the arithmetic and vice versa.  One sweep covers a tile (DT_PTHREADS x DT_U = 4096), so the per-element table entries
// are loaded once per CTA and live in registers.  Per-tile data (base offset, uniform weights) rotates over three slots:
// the slot of tile t+1 is written while laggard warps may still be reading the slot of tile t-1.
template <bool REALT, bool REALW, bool FULL>
__global__ void __launch_bounds__(DT_PTHREADS, 1) dtile_persist_kernel(const __grid_constant__ DPassParams P,
                                                                       const double2 *__restrict__ x, double2 *__restrict__ y,
                                                                       unsigned ntiles) {
  typedef typename std::conditional<REALT, double, double2>::type TabT;
  typedef typename std::conditional<REALW, double, double2>::type WT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *xs0 = reinterpret_cast<double2 *>(smem_raw);
  double2 *cw2 = xs0 + 2 * (size_t)P.tile;
  DCompDev *comps = reinterpret_cast<DCompDev *>(cw2 + 3 * (size_t)P.ncomp);
  TabT *tab = reinterpret_cast<TabT *>(comps + P.ncomp);
  __shared__ long long s_base[3];
  const unsigned tid = threadIdx.x;

  for (int i = tid; i < P.ntab; i += DT_PTHREADS) {
    if constexpr (REALT) tab[i] = P.tables[i].x;
    else tab[i] = P.tables[i];
  }
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(P.comps);
    uint4 *dst = reinterpret_cast<uint4 *>(comps);
    const int n16 = P.ncomp * (int)(sizeof(DCompDev) / 16);
    for (int i = tid; i < n16; i += DT_PTHREADS) dst[i] = src[i];
  }
  unsigned pk[DT_U], eo[DT_U];
#pragma unroll
  for (int k = 0; k < DT_U; ++k) {
    const unsigned e = k * DT_PTHREADS + tid;
    const uint2 t = (FULL || e < (unsigned)P.tile) ? __ldg(&P.etab[e]) : make_uint2(0xffffffffu, 0u);
    pk[k] = t.x;
    eo[k] = t.y;
  }
  __syncthreads();
  // per-tile data into slot `slot`: base offset (warp 0) and the uniform weight of every component (one thread each,
  // with its own digits of the fixed axes)
  auto decode = [&](unsigned tile, int slot) {
    if (tid < 32) {
      long long part = 0;
      for (unsigned a = tid; a < (unsigned)P.nfixed; a += 32)
        part += (long long)((tile / P.fx_below[a]) % P.fx_dim[a]) * P.fx_stride[a];
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (tid == 0) s_base[slot] = part;
    }
    WT *cw = reinterpret_cast<WT *>(cw2 + (size_t)slot * P.ncomp);
    for (int c = (int)tid - 32; c < P.ncomp; c += DT_PTHREADS - 32) {
      if (c < 0) continue;
      const DCompDev &C = comps[c];
      double2 w = P.coef[C.coef];
      for (int f = 0; f < C.nfixed; ++f) {
        const unsigned a = C.x_slot[f];
        const unsigned dg = (tile / P.fx_below[a]) % P.fx_dim[a];
        if constexpr (REALT) {
          const double t = tab[C.x_tab[f] + dg];
          w.x *= t;
          w.y *= t;
        } else {
          w = dmul(w, tab[C.x_tab[f] + dg]);
        }
      }
      if constexpr (REALW) cw[c] = w.x;
      else cw[c] = w;
    }
  };
  auto issue = [&](int slot, double2 *xs) {
    const long long base = s_base[slot];
#pragma unroll
    for (int k = 0; k < DT_U; ++k) {
      if (FULL || pk[k] != 0xffffffffu) {
        const long long off = base + (long long)eo[k];
        const unsigned sa = (unsigned)__cvta_generic_to_shared(xs + k * DT_PTHREADS + tid);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(x + off) : "memory");
        if (P.mode && (off & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(y + off));
      }
    }
  };
  unsigned tile = blockIdx.x;
  if (tile >= ntiles) return;
  decode(tile, 0);
  __syncthreads();
  issue(0, xs0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  int slot = 0, buf = 0;
  for (; tile < ntiles; tile += gridDim.x) {
    const unsigned next = tile + gridDim.x;
    const int nslot = slot == 2 ? 0 : slot + 1;
    if (next < ntiles) decode(next, nslot);
    __syncthreads();   // slot of the next tile is written; everyone is done with the other tile buffer
    if (next < ntiles) issue(nslot, xs0 + (size_t)(buf ^ 1) * P.tile);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // this thread's copies of the current tile have landed
    __syncthreads();   // ... and everybody else's
    dt_sweep<REALT, REALW, FULL, DT_PTHREADS>(P, xs0 + (size_t)buf * P.tile, tab, reinterpret_cast<const WT *>(cw2 + (size_t)slot * P.ncomp),
                                              comps, 0u, tid, s_base[slot], pk, eo, y);
    slot = nslot;
    buf ^= 1;
  }
}

// ================================================================================ host: planner
struct DComp {           // host form of a component
  int term;
  std::vector<int> axes;                 // factor axes
  std::vector<int> shift;                // s_k: source digit = output digit + s_k
  std::vector<std::vector<cplx>> table;  // per factor: weight by output digit (0 where there is no entry)
};
struct DPassHost {
  std::vector<int> free_axes;            // ascending; the first `nlow` are axes 0..nlow-1
  int nlow = 0;
  int tile = 1, R = 1;
  std::vector<int> terms;
  DevArray<DCompDev> d_comps;
  DevArray<double2> d_tables;
  DevArray<uint2> d_etab;
  DPassParams params;
  bool real_tables = true;
  size_t smem = 0;
  int64_t ntiles = 1;
};
struct DTileProgramHost {
  std::vector<int64_t> dims;
  std::vector<std::unique_ptr<DPassHost>> passes;
  std::vector<int> coef_of_term;
  std::vector<cplx> scalars;
  DevArray<double2> d_coef;
  bool coefs_real = false;   // every coef*scalar currently loaded is real (set by dtile_set_coefs)
  int64_t total = 1;
};


// shifted diagonals of an oriented square factor (rows = output digit, cols = input digit)
static void factor_diagonals(const HostMat &m, std::vector<int> &shifts, std::vector<std::vector<cplx>> &tables) {
  const int64_t d = m.rows;
  std::map<int, std::vector<cplx>> diag;
  auto put = [&](int64_t i, int64_t j, cplx v) {
    auto &t = diag[(int)(j - i)];
    if (t.empty()) t.assign((size_t)d, cplx(0.0, 0.0));
    t[(size_t)i] += v;
  };
  if (m.kind == QOB_FACTOR_CSC) {
    for (int64_t c = 0; c < m.cols; ++c)
      for (int64_t p = m.colptr[c]; p < m.colptr[c + 1]; ++p) put(m.rowidx[p], c, m.vals[p]);
  } else if (m.kind == QOB_FACTOR_DENSE) {
    for (int64_t c = 0; c < m.cols; ++c)
      for (int64_t r = 0; r < m.rows; ++r) {
        const cplx v = m.dense[(size_t)(r + c * m.rows)];
        if (v != cplx(0.0, 0.0)) put(r, c, v);   // zero weights skip their gather in the kernel anyway
      }
  } else {
    for (int64_t i = 0; i < d; ++i) put(i, i, cplx(1.0, 0.0));
  }
  for (auto &kv : diag) {
    shifts.push_back(kv.first);
    tables.push_back(kv.second);
  }
}

int dtile_build(DTileProgram &prog, const std::vector<int64_t> &dims, const std::vector<OrientedTerm> &terms,
                std::vector<int> *declined_out) {
  std::vector<char> declined(terms.size(), 0);   // terms the scheme cannot take: left to the caller (gather kernel)
  auto hp = std::make_shared<DTileProgramHost>();
  DTileProgramHost &H = *hp;
  H.dims = dims;
  const int n = (int)dims.size();
  if (n > DT_MAXAX) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "dtile: more than %d axes", DT_MAXAX);
  {
    std::vector<char> has_factor((size_t)n, 0);   // axes without factors (e.g. the batch) may be long: they are never free
    for (const OrientedTerm &T : terms)
      for (int a : T.axes) has_factor[(size_t)a] = 1;
    for (int a = 0; a < n; ++a) {
      const int64_t d = dims[a];
      if (d < 1 || d > (has_factor[a] ? 4096 : 0x7FFFFFFFll)) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "dtile: axis dimension %lld", (long long)d);
      H.total *= d;
    }
  }
  // ---- components of every term
  std::vector<std::vector<DComp>> comps(terms.size());
  std::vector<std::vector<int>> offaxes(terms.size());
  for (size_t t = 0; t < terms.size(); ++t) {
    const OrientedTerm &T = terms[t];
    H.coef_of_term.push_back(T.coef_index);
    cplx scalar = T.scalar;
    bool decline = T.axes.size() > DT_MAXF;
    std::vector<std::vector<int>> shifts(T.axes.size());
    std::vector<std::vector<std::vector<cplx>>> tables(T.axes.size());
    std::vector<int> axes;
    size_t ncomb = 1;
    for (size_t f = 0; f < T.axes.size(); ++f) {
      const HostMat &m = T.mats[f];
      const int ax = T.axes[f];
      if (decline) break;
      if (m.rows != m.cols || m.rows != dims[ax]) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "dtile: non-square factor");
      if (m.is_square_eye()) continue;
      if (dims[ax] == 1) {  // a 1x1 factor is a scalar
        scalar *= m.at(0, 0);
        continue;
      }
      std::vector<int> sh;
      std::vector<std::vector<cplx>> tb;
      factor_diagonals(m, sh, tb);
      if (sh.empty()) {  // a factor without entries: the term vanishes
        ncomb = 0;
        break;
      }
      bool off = false;
      for (int s : sh) off |= s != 0;
      if (off) offaxes[t].push_back(ax);
      ncomb *= sh.size();
      axes.push_back(ax);
      shifts[axes.size() - 1] = sh;
      tables[axes.size() - 1] = tb;
    }
    H.scalars.push_back(scalar);
    if (decline || ncomb > 64) {   // too many single-gather components (e.g. two dense 8x8 factors): not this kernel's job
      declined[t] = 1;
      offaxes[t].clear();
      continue;
    }
    if (ncomb == 0) continue;
    std::vector<size_t> cur(axes.size(), 0);
    for (size_t it = 0; it < ncomb; ++it) {
      DComp c;
      c.term = (int)t;
      for (size_t f = 0; f < axes.size(); ++f) {
        c.axes.push_back(axes[f]);
        c.shift.push_back(shifts[f][cur[f]]);
        c.table.push_back(tables[f][cur[f]]);
      }
      comps[t].push_back(std::move(c));
      for (size_t f = 0; f < axes.size(); ++f) {
        if (++cur[f] < shifts[f].size()) break;
        cur[f] = 0;
      }
    }
  }
  // ---- low block: the fewest leading axes whose joint extent gives >= 8-amplitude (128-byte) contiguous runs
  int nlow = 0;
  int64_t R = 1;
  while (nlow < n && R < 8) R *= dims[nlow++];
  if (R > DT_TILE_CAP / 2) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "dtile: leading axis too large for a tile");
  auto extent = [&](const std::vector<int> &axs) {
    int64_t e = 1;
    for (int a : axs) e *= dims[a];
    return e;
  };
  auto merged = [&](std::vector<int> a, const std::vector<int> &b) {
    a.insert(a.end(), b.begin(), b.end());
    std::sort(a.begin(), a.end());
    a.erase(std::unique(a.begin(), a.end()), a.end());
    return a;
  };
  std::vector<int> low(nlow);
  std::iota(low.begin(), low.end(), 0);
  std::vector<char> covered(terms.size(), 0);
  std::vector<std::vector<int>> pass_free;
  std::vector<std::vector<int>> pass_terms;
  {  // pass 0: the longest leading run of axes that fits a tile (fully contiguous tiles)
    std::vector<int> fr;
    int64_t e = 1;
    for (int a = 0; a < n && e * dims[a] <= DT_TILE_CAP; ++a) {
      fr.push_back(a);
      e *= dims[a];
    }
    if ((int)fr.size() < nlow) fr = low;
    pass_free.push_back(fr);
    pass_terms.emplace_back();
  }
  auto subset = [](const std::vector<int> &a, const std::vector<int> &of) {
    for (int v : a)
      if (!std::binary_search(of.begin(), of.end(), v)) return false;
    return true;
  };
  for (size_t t = 0; t < terms.size(); ++t) {
    if (declined[t]) {
      covered[t] = 1;
    } else if (subset(offaxes[t], pass_free[0])) {
      covered[t] = 1;
      pass_terms[0].push_back((int)t);
    }
  }
  while (true) {
    // seed: the uncovered term whose highest off-diagonal axis is lowest
    int seed = -1;
    for (size_t t = 0; t < terms.size(); ++t)
      if (!covered[t] && (seed < 0 || offaxes[t].back() < offaxes[seed].back())) seed = (int)t;
    if (seed < 0) break;
    std::vector<int> fr = merged(low, offaxes[seed]);
    if (extent(fr) > DT_TILE_CAP) {
      fr = offaxes[seed];  // give up the coalesced low block for this pass
      std::sort(fr.begin(), fr.end());
      if (extent(fr) > DT_TILE_CAP) {   // this term alone exceeds a tile
        declined[seed] = 1;
        covered[seed] = 1;
        continue;
      }
    }
    // grow with the uncovered terms that come next
    std::vector<int> order;
    for (size_t t = 0; t < terms.size(); ++t)
      if (!covered[t]) order.push_back((int)t);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return offaxes[a].back() < offaxes[b].back(); });
    for (int t : order) {
      std::vector<int> g = merged(fr, offaxes[t]);
      if (extent(g) <= DT_TILE_CAP) fr = g;
    }
    // a small window wastes the CTA: pad the tile with the axes just below the window (more work per CTA, longer
    // strided runs), as long as it stays within the capacity
    for (int a = (fr.size() > (size_t)nlow ? *std::find_if(fr.begin(), fr.end(), [&](int v) { return v >= nlow; }) : n) - 1;
         a >= nlow && extent(fr) * dims[a] <= DT_TILE_CAP; --a)
      fr = merged(fr, {a});
    pass_free.push_back(fr);
    pass_terms.emplace_back();
    for (size_t t = 0; t < terms.size(); ++t)
      if (!covered[t] && subset(offaxes[t], fr)) {
        covered[t] = 1;
        pass_terms.back().push_back((int)t);
      }
  }
  {
    size_t nd = 0;
    for (char d : declined) nd += d;
    if (nd == terms.size() && !terms.empty()) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "dtile: no term fits the scheme");
    if (declined_out) {
      declined_out->clear();
      for (size_t t = 0; t < terms.size(); ++t)
        if (declined[t]) declined_out->push_back((int)t);
    } else if (nd) {
      QOB_FAIL(QOB_STATUS_UNSUPPORTED, "dtile: %zu terms do not fit the scheme", nd);
    }
  }
  // diagonal terms can run in any pass: take them out of pass 0 and hand each to the pass that has the least work so far
  // (passes with few components are memory bound and have issue slots to spare); a leading pass left without terms is
  // dropped
  {
    std::vector<int> diag_terms;
    for (auto it = pass_terms[0].begin(); it != pass_terms[0].end();) {
      if (offaxes[*it].empty()) {
        diag_terms.push_back(*it);
        it = pass_terms[0].erase(it);
      } else {
        ++it;
      }
    }
    std::vector<size_t> load(pass_free.size(), 0);
    for (size_t pi = 0; pi < pass_free.size(); ++pi)
      for (int t : pass_terms[pi]) load[pi] += 2 * comps[t].size();   // an off-diagonal component costs about two diagonal ones
    const bool drop0 = pass_terms[0].empty() && pass_free.size() > 1;
    for (int t : diag_terms) {
      size_t best = drop0 ? 1 : 0;
      for (size_t pi = best; pi < pass_free.size(); ++pi)
        if (load[pi] < load[best]) best = pi;
      pass_terms[best].push_back(t);
      load[best] += comps[t].size();
    }
  }
  // ---- device programs
  std::string text;
  for (size_t pi = 0; pi < pass_free.size(); ++pi) {
    if (pass_terms[pi].empty() && !(pi == 0 && pass_free.size() == 1)) continue;
    auto pp = std::make_unique<DPassHost>();
    DPassHost &Pz = *pp;
    Pz.free_axes = pass_free[pi];
    Pz.terms = pass_terms[pi];
    const std::vector<int> &fr = Pz.free_axes;
    // low run: leading free axes that are exactly axes 0,1,2,...
    int nl = 0;
    int64_t Rr = 1;
    while (nl < (int)fr.size() && fr[nl] == nl) Rr *= dims[nl++];
    Pz.nlow = nl;
    Pz.R = (int)Rr;
    Pz.tile = (int)extent(fr);
    std::vector<int64_t> gstride(n);
    {
      int64_t s = 1;
      for (int a = 0; a < n; ++a) {
        gstride[a] = s;
        s *= dims[a];
      }
    }
    std::vector<int64_t> tstride(n, 0);
    {
      int64_t s = 1;
      for (int a : fr) {
        tstride[a] = s;
        s *= dims[a];
      }
    }
    DPassParams &Q = Pz.params;
    memset(&Q, 0, sizeof(Q));
    std::vector<int> fixed_slot(n, -1);
    {
      int64_t below = 1;
      for (int a = 0; a < n; ++a) {
        if (std::binary_search(fr.begin(), fr.end(), a)) continue;
        if (dims[a] == 1) continue;
        if (below * dims[a] > 0xFFFFFFFFll) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "dtile: more than 2^32 tiles");
        fixed_slot[a] = Q.nfixed;
        Q.fx_dim[Q.nfixed] = (unsigned)dims[a];
        Q.fx_below[Q.nfixed] = (unsigned)below;
        Q.fx_stride[Q.nfixed] = gstride[a];
        ++Q.nfixed;
        below *= dims[a];
      }
      Pz.ntiles = below;
      if (below >= (1ll << 31)) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "dtile: more than 2^31 tiles");
    }
    std::vector<int> ax_shift(n, 0), ax_bits(n, 0);
    {
      int pos = 0;
      for (int a : fr) {
        if (dims[a] == 1) continue;
        int bits = 1;
        while ((1ll << bits) < dims[a]) ++bits;
        ax_shift[a] = pos;
        ax_bits[a] = bits;
        pos += bits;
      }
      if (pos > 31) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "dtile: packed digits do not fit");
    }
    std::vector<uint2> etab((size_t)Pz.tile);
    for (int e = 0; e < Pz.tile; ++e) {
      int64_t rem = e, off = 0;
      unsigned pk = 0;
      for (int a : fr) {
        const int64_t dg = rem % dims[a];
        rem /= dims[a];
        off += dg * gstride[a];
        pk |= (unsigned)dg << ax_shift[a];
      }
      if (off > 0xFFFFFFFFll) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "dtile: tile spans more than 2^32 amplitudes");
      etab[(size_t)e] = make_uint2(pk, (unsigned)off);
    }
    std::vector<DCompDev> recs;
    std::vector<double2> tabs;
    tabs.push_back(make_double2(1.0, 0.0));   // entry 0: the neutral table of an absent second factor
    auto add_table = [&](const std::vector<cplx> &t) {
      // identical tables are shared (the same site operator appears in many terms)
      for (size_t o = 0; o + t.size() <= tabs.size(); ++o) {
        bool same = true;
        for (size_t i = 0; i < t.size() && same; ++i) same = tabs[o + i].x == t[i].real() && tabs[o + i].y == t[i].imag();
        if (same) return (int)o;
      }
      const int o = (int)tabs.size();
      for (cplx v : t) tabs.push_back(make_double2(v.real(), v.imag()));
      return o;
    };
    for (int t : Pz.terms)
      for (const DComp &c : comps[t]) {
        DCompDev r;
        memset(&r, 0, sizeof(r));
        r.coef = t;
        int64_t delta = 0;
        for (size_t f = 0; f < c.axes.size(); ++f) {
          const int ax = c.axes[f];
          const bool is_free = std::binary_search(fr.begin(), fr.end(), ax);
          if (!is_free && c.shift[f] != 0) QOB_FAIL(QOB_STATUS_INVALID_ARG, "dtile: internal planner error");
          const int off = add_table(c.table[f]);
          if (off + (int)dims[ax] > 65535) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "dtile: operator tables too large");
          if (is_free) {
            const int k = r.nfree++;
            r.f_sh[k] = (unsigned)ax_shift[ax];
            r.f_mask[k] = (1u << ax_bits[ax]) - 1u;
            r.f_d[k] = (unsigned short)dims[ax];
            r.f_tab[k] = (unsigned short)off;
            delta += (int64_t)c.shift[f] * tstride[ax];
          } else {
            const int k = r.nfixed++;
            r.x_tab[k] = (unsigned short)off;
            r.x_slot[k] = (unsigned char)fixed_slot[ax];
          }
        }
        r.delta = (int)delta;
        if (r.nfree == 2) {
          // two free factors whose digits are neighbours in the packed word (sites i, i+1 of a bond): one joint table
          // indexed by both digits at once — one lookup instead of two lookups and a multiply
          int fa[2], nf = 0;
          for (size_t f = 0; f < c.axes.size(); ++f)
            if (std::binary_search(fr.begin(), fr.end(), c.axes[f]) && nf < 2) fa[nf++] = (int)f;
          const int a0 = c.axes[fa[0]], a1 = c.axes[fa[1]];
          const int b0 = ax_bits[a0], b1 = ax_bits[a1];
          if (ax_shift[a1] == ax_shift[a0] + b0 && b0 + b1 <= 8) {
            std::vector<cplx> joint((size_t)1 << (b0 + b1), cplx(0.0, 0.0));
            for (int64_t i1 = 0; i1 < dims[a1]; ++i1)
              for (int64_t i0 = 0; i0 < dims[a0]; ++i0)
                joint[(size_t)(i0 + (i1 << b0))] = c.table[fa[0]][(size_t)i0] * c.table[fa[1]][(size_t)i1];
            const int off = add_table(joint);
            if (off + (int)joint.size() <= 65535) {
              r.nfree = 1;
              r.f_sh[0] = (unsigned)ax_shift[a0];
              r.f_mask[0] = (1u << (b0 + b1)) - 1u;
              r.f_tab[0] = (unsigned short)off;
              r.f_sh[1] = r.f_mask[1] = 0;
              r.f_tab[1] = 0;
            }
          }
        }
        bool diag = true;
        for (int sft : c.shift) diag &= sft == 0;
        r.pad0 = diag ? (r.nfree ? 1 : 2) : 0;   // ordering class
        recs.push_back(r);
      }
    std::stable_sort(recs.begin(), recs.end(), [](const DCompDev &a, const DCompDev &b) { return a.pad0 < b.pad0; });
    for (const DCompDev &r : recs) {
      Q.n_off += r.pad0 == 0;
      Q.n_dfree += r.pad0 == 1;
    }
    if (recs.size() > 2048 || tabs.size() > 2048) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "dtile: too many components in a pass");
    // components with few free factors first is not needed; keep term order (deterministic accumulation order)
    Q.ncomp = (int)recs.size();
    Q.ntab = (int)tabs.size();
    Q.tile = Pz.tile;
    Q.R = Pz.R;
    if (recs.empty()) recs.push_back(DCompDev{});
    if (tabs.empty()) tabs.push_back(make_double2(0.0, 0.0));
    QOB_TRY(Pz.d_comps.upload(recs));
    QOB_TRY(Pz.d_tables.upload(tabs));
    QOB_TRY(Pz.d_etab.upload(etab));
    Q.comps = Pz.d_comps.ptr;
    Q.tables = Pz.d_tables.ptr;
    Q.etab = Pz.d_etab.ptr;
    for (const double2 &v : tabs) Pz.real_tables &= v.y == 0.0;
    Pz.smem = (size_t)Pz.tile * 16 + (size_t)Q.ntab * 16 + (size_t)Q.ncomp * (16 + sizeof(DCompDev));
    if (Pz.smem > 200 * 1024) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "dtile: pass needs %zu bytes of shared memory", Pz.smem);
    char buf[256];
    std::string fa;
    for (int a : fr) fa += (fa.empty() ? "" : ",") + std::to_string(a + 1);
    snprintf(buf, sizeof(buf), " {free axes:%s tile:%d run:%d terms:%zu components:%d tables:%dB%s}", fa.c_str(), Pz.tile, Pz.R,
             Pz.terms.size(), Q.ncomp, Q.ntab * (Pz.real_tables ? 8 : 16), Pz.real_tables ? " real" : "");
    text += buf;
    H.passes.push_back(std::move(pp));
  }
  {
    std::vector<double2> c(std::max<size_t>(1, terms.size()), make_double2(0.0, 0.0));
    QOB_TRY(H.d_coef.upload(c));
  }
  prog.h = hp;
  prog.npasses = (int)H.passes.size();
  prog.describe = "dtile[axes=" + std::to_string(n) + ",passes=" + std::to_string(prog.npasses) + "]" + text;
  return QOB_STATUS_OK;
}

int dtile_set_coefs(DTileProgram &prog, const std::vector<cplx> &coefs, cudaStream_t s) {
  DTileProgramHost &H = *prog.h;
  std::vector<double2> c(std::max<size_t>(1, H.coef_of_term.size()), make_double2(0.0, 0.0));
  for (size_t t = 0; t < H.coef_of_term.size(); ++t) {
    cplx v = H.scalars[t];
    if (H.coef_of_term[t] >= 0) {
      if ((size_t)H.coef_of_term[t] >= coefs.size()) QOB_FAIL(QOB_STATUS_INVALID_ARG, "coefficient index out of range");
      v *= coefs[H.coef_of_term[t]];
    }
    c[t] = make_double2(v.real(), v.imag());
  }
  H.coefs_real = true;
  for (const double2 &v : c) H.coefs_real &= v.y == 0.0;
  return H.d_coef.upload_async(c, s);
}

int dtile_launch(const DTileProgram &prog, cplx alpha, const void *x, cplx beta, void *y, cudaStream_t s) {
  DTileProgramHost &H = *prog.h;
  if (t_planning_only) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "planning-only context (no CUDA device): libqob200 has no CPU fallback");
  // the shared-memory opt-in is per device: once per (process, device)
  static std::once_flag once[QOB_MAX_DEVICES];
  static cudaError_t attr_errs[QOB_MAX_DEVICES];
  int cur_dev = 0;
  QOB_CUDA(cudaGetDevice(&cur_dev));
  if (cur_dev < 0 || cur_dev >= QOB_MAX_DEVICES) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "device ordinal %d out of range", cur_dev);
  cudaError_t &attr_err = attr_errs[cur_dev];
  std::call_once(once[cur_dev], [&attr_err] {
    attr_err = cudaSuccess;
    auto set = [&attr_err](const void *f) {
      if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    };
    set((const void *)dtile_kernel<true, true, true>);
    set((const void *)dtile_kernel<true, true, false>);
    set((const void *)dtile_kernel<true, false, true>);
    set((const void *)dtile_kernel<true, false, false>);
    set((const void *)dtile_kernel<false, false, true>);
    set((const void *)dtile_kernel<false, false, false>);
    set((const void *)dtile_persist_kernel<true, true, true>);
    set((const void *)dtile_persist_kernel<true, true, false>);
    set((const void *)dtile_persist_kernel<true, false, true>);
    set((const void *)dtile_persist_kernel<true, false, false>);
    set((const void *)dtile_persist_kernel<false, false, true>);
    set((const void *)dtile_persist_kernel<false, false, false>);
  });
  if (attr_err != cudaSuccess) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  // QOB_DTILE_PERSIST: 0 = one CTA per tile, 1 (default) = persistent CTAs when there are enough tiles, 2 = always
  const char *pm = getenv("QOB_DTILE_PERSIST");
  const int persist_mode = pm && *pm ? atoi(pm) : 1;
  int dev = 0, sm_count = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  bool first = true;
  for (auto &pp : H.passes) {
    DPassHost &Pz = *pp;
    DPassParams P = Pz.params;
    P.coef = H.d_coef.ptr;
    P.alpha = make_double2(alpha.real(), alpha.imag());
    const cplx b = first ? beta : cplx(1.0, 0.0);
    P.beta = make_double2(b.real(), b.imag());
    P.mode = (b == cplx(0.0, 0.0)) ? 0 : 1;
    first = false;
    const unsigned grid = (unsigned)Pz.ntiles;
    const double2 *xp = (const double2 *)x;
    double2 *yp = (double2 *)y;
    // persistent double-buffered variant: worth it when every SM gets several tiles
    const size_t psmem = Pz.smem + (size_t)Pz.tile * 16 + (size_t)Pz.params.ncomp * 32;
    const bool persist = persist_mode != 0 && Pz.tile <= DT_PTHREADS * DT_U && psmem <= 220 * 1024 &&
                         (persist_mode == 2 || Pz.ntiles >= 8ll * sm_count);
    const bool full = persist ? Pz.tile == DT_PTHREADS * DT_U : Pz.tile % (DT_THREADS * DT_U) == 0;
#define DT_GO(RT, RW)                                                                                              \
  do {                                                                                                             \
    if (persist) {                                                                                                 \
      const unsigned pg = (unsigned)std::min<int64_t>(Pz.ntiles, sm_count);                                        \
      if (full) dtile_persist_kernel<RT, RW, true><<<pg, DT_PTHREADS, psmem, s>>>(P, xp, yp, grid);                \
      else dtile_persist_kernel<RT, RW, false><<<pg, DT_PTHREADS, psmem, s>>>(P, xp, yp, grid);                    \
    } else if (full) dtile_kernel<RT, RW, true><<<grid, DT_THREADS, Pz.smem, s>>>(P, xp, yp);                      \
    else dtile_kernel<RT, RW, false><<<grid, DT_THREADS, Pz.smem, s>>>(P, xp, yp);                                 \
  } while (0)
    if (Pz.real_tables && H.coefs_real) DT_GO(true, true);
    else if (Pz.real_tables) DT_GO(true, false);
    else DT_GO(false, false);
#undef DT_GO
    QOB_LAUNCHED();
    QOB_CUDA(cudaGetLastError());
  }
  return QOB_STATUS_OK;
}
