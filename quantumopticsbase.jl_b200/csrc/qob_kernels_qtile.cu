// qob_kernels_qtile.cu — fused LazySum apply for systems whose subsystems are all 2-dimensional
// (spin-1/2 chains: BASELINE configs 1, 4, 5), the HBM-roofline path.
//
// Reference behaviour replaced: src/operators_lazysum.jl:189-200 runs one full pass over the state
// PER TERM (each through the scalar recursion src/operators_lazytensor.jl:652-685 or the
// permute+zgemm path :333-428).  Here the whole sum is regrouped by "flip mask":
//
//     (H x)[i] = sum_c  w_c(bits of i) * x[i XOR mask_c]
//
// Every product of 2x2 site factors splits into 2^k such components (flip / no-flip per site); all
// components with the same (mask, selector bits) — e.g. XX and YY on one bond, or every diagonal
// term — are merged on the host into one weight table, so the device does ONE gather per distinct
// mask.  A pass loads a tile of 2^T amplitudes (T free index bits: the low L bits for coalescing
// plus a window of high bits) into shared memory with cp.async, and applies every component whose
// mask lies inside the free bits; diagonal factors may sit on any bit (they only select the weight).
// Traffic is 32 B/amplitude for the first pass and 48 B for each further pass (y read-modify-write),
// instead of 48 B per TERM in the reference.
//
// Index bits: bit b of the flat (column-major) index <-> subsystem b+1 (subsystem 1 is fastest,
// src/states.jl:105).  For sharded states the bits >= nbits are the rank (`hi_value`).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>

#include "qob_internal.h"

#define QT_MAXSEG 8
#define QT_MAXSEL 3
#define QT_DIAG_WINDOW 6  // merged diagonal tables span at most this many index bits (64 entries)

// One weight-table lookup + (at the end of a mask group) one gather from the staged tile.
struct QCompDev {   // 16 bytes, read as one broadcast LDS.128
  uint32_t xorB;    // tile-local flip mask * 16 (byte offset XOR inside the staged tile)
  uint32_t sel;     // up to three bit-field runs of the index: run r = bits [9r, 9r+9): shift (6 bits) | width (3 bits)
  uint32_t tabB;    // byte offset of this table inside the pass table
  uint32_t flags;   // bit0: last lookup of its mask group (gather + FMA now); bit1: the group has a single lookup
};

#define QT_MAXPEER 16
#define QT_MAXCOMP 56  // lookup records carried in the kernel parameters (uniform loads, no shared-memory traffic)

struct QPassParams {
  const double2 *tab;     // device: weight tables of this pass (entry 0 is a zero weight)
  int npre;               // diagonal tables evaluated once per thread (selector bits fixed for the thread)
  int ndiag;              // diagonal tables evaluated per amplitude
  int nmulti;             // off-diagonal lookups that share their mask with others (summed before the gather)
  int nsingle;            // off-diagonal lookups with a mask of their own
  int ntab;
  int has_diag;           // some diagonal weight exists: acc starts from (dthread + diag tables) * x[i]
  int nfree_seg, nfixed_seg;
  // segment = (shift in the compact index, length, position in the address)
  unsigned char fs_l[QT_MAXSEG], fs_n[QT_MAXSEG], fs_g[QT_MAXSEG];  // free bits: tile-local index -> address
  unsigned char xs_l[QT_MAXSEG], xs_n[QT_MAXSEG], xs_g[QT_MAXSEG];  // fixed bits: tile id -> address
  unsigned long long hi_or;     // index bits above the local address (rank), already shifted
  unsigned long long eoff[16];  // address part of tile-local index k*THREADS (k < TILE/THREADS), precomputed on the host
  double2 alpha, beta;
  int mode;  // 0: y = a*acc ; 1: y = a*acc + beta*y ; 2: y = a*acc + y      (+ zadd[i] when zadd != nullptr)
  unsigned ntiles;          // tiles this launch processes; the CTAs are persistent and stride over them
  unsigned tile_begin;      // first tile of this launch (chunked launches of one pass)
  int bulk;                 // 1: stage the tile with TMA bulk copies (cp.async.bulk + mbarrier), one per contiguous run
  int run_log2;             // log2 of the amplitudes per contiguous run (= width of the low free block)
  const double2 *zadd;      // optional extra addend in the local layout (contributions received from other ranks)
  // PEER variant (sharded states): the buffer addressed by this pass is the SWAPPED layout of the ranks' slabs.  Index bits
  // [peer_shift, peer_shift+peer_bits) of an address name the rank that holds the element; there it sits at the same
  // address with those bits replaced by this rank's number.  Loads/stores go straight to that rank's memory (NVLink P2P).
  int peer_shift, peer_bits, peer_rank;
  const double2 *xpeer[QT_MAXPEER];
  double2 *ypeer[QT_MAXPEER];
  QCompDev comps[QT_MAXCOMP];  // [npre | ndiag | nmulti | nsingle]
  // per (record, k): byte offsets contributed by the warp-uniform part k*THREADS of the tile-local index:
  // .x -> weight-table offset (selector bits among the thread's varying bits), .y -> gather offset (k*THREADS*16 ^ high mask bits)
  uint2 ck[QT_MAXCOMP][16];
};

__device__ __forceinline__ void qcfma(double2 &acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}

__device__ __forceinline__ unsigned long long qexpand(unsigned v, int nseg, const unsigned char *sl,
                                                      const unsigned char *sn, const unsigned char *sg) {
  unsigned long long a = 0;
#pragma unroll
  for (int s = 0; s < QT_MAXSEG; ++s)
    if (s < nseg) a |= (unsigned long long)((v >> sl[s]) & ((1u << sn[s]) - 1u)) << sg[s];
  return a;
}

// value of the index field described by run r of `sel` (shift 6 bits | width 3 bits), shifted left by `pos`
template <bool IDX64>
__device__ __forceinline__ unsigned qrun(unsigned sel, int r, unsigned lo, unsigned hi, unsigned pos) {
  const unsigned sh = (sel >> (9 * r)) & 63u, w = (sel >> (9 * r + 6)) & 7u;
  unsigned v;
  if (IDX64)
    v = sh < 32 ? __funnelshift_r(lo, hi, sh) : (hi >> (sh & 31));
  else
    v = lo >> sh;
  return (v & ((1u << w) - 1u)) << pos;
}
// table index (in entries) for up to three runs; runs 1 and 2 are rare (selector bits that are not adjacent).
// The field is bitwise in the index: qfield(a | b) == qfield(a) | qfield(b).
template <bool IDX64>
__device__ __forceinline__ unsigned qfield(unsigned sel, unsigned lo, unsigned hi) {
  unsigned idx = qrun<IDX64>(sel, 0, lo, hi, 0);
  if (sel >> 9) {
    const unsigned w0 = (sel >> 6) & 7u, w1 = (sel >> 15) & 7u;
    idx |= qrun<IDX64>(sel, 1, lo, hi, w0);
    if (sel >> 18) idx |= qrun<IDX64>(sel, 2, lo, hi, w0 + w1);
  }
  return idx;
}

// weight load: complex tables hold (re, im); the REALW variant reads only the real part (8 bytes: half the
// shared-memory wavefronts) and multiplies with 2 DFMA instead of 4
template <bool REALW>
struct QW {
  double2 w;
  __device__ __forceinline__ void load(const unsigned char *p) {
    if (REALW) {
      w.x = *reinterpret_cast<const double *>(p);
      w.y = 0.0;
    } else {
      w = *reinterpret_cast<const double2 *>(p);
    }
  }
  __device__ __forceinline__ void fma_into(double2 &acc, const double2 v) const {
    if (REALW) {
      acc.x = fma(w.x, v.x, acc.x);
      acc.y = fma(w.x, v.y, acc.y);
    } else {
      qcfma(acc, w, v);
    }
  }
};

// address of element `a` of the buffer a pass works on: local memory, or (PEER) the owning rank's slab
template <bool PEER, typename PT>
__device__ __forceinline__ PT *qaddr(const QPassParams &P, PT *local, PT *const *peers, unsigned long long a) {
  if (!PEER) return local + a;
  const unsigned long long m = ((1ull << P.peer_bits) - 1ull) << P.peer_shift;
  const unsigned q = (unsigned)((a & m) >> P.peer_shift);
  return peers[q] + ((a & ~m) | ((unsigned long long)P.peer_rank << P.peer_shift));
}

template <int T, int THREADS, int MINB, bool IDX64, bool REALW, bool PEER>
__global__ void __launch_bounds__(THREADS, MINB)
    qtile_kernel(const __grid_constant__ QPassParams P, const double2 *__restrict__ x, double2 *__restrict__ y) {
  constexpr int TILE = 1 << T;
  constexpr int PER = TILE / THREADS;       // amplitudes per thread
  constexpr int U = PER < 8 ? PER : 8;      // amplitudes in flight per thread
  constexpr int ITERS = PER / U;
  constexpr unsigned TMASKB = THREADS * 16u - 1u;  // byte-offset bits that belong to the tid part of the tile-local index
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char *xsB = smem_raw;                                    // staged x tile, byte addressed
  unsigned char *tabB = smem_raw + (size_t)TILE * sizeof(double2);  // weight tables

  const unsigned tid = threadIdx.x;
  const unsigned long long at_tid = qexpand(tid, P.nfree_seg, P.fs_l, P.fs_n, P.fs_g);
  {
    int4 *t4 = reinterpret_cast<int4 *>(tabB);
    const int4 *g4 = reinterpret_cast<const int4 *>(P.tab);
    for (int i = tid; i < P.ntab; i += THREADS) t4[i] = g4[i];
  }
  const unsigned lbt = tid * 16u;
  const int c_diag = P.npre, c_multi = c_diag + P.ndiag, c_single = c_multi + P.nmulti, c_end = c_single + P.nsingle;
  __shared__ __align__(8) unsigned long long tile_bar;  // mbarrier tracking the bulk copies of the current tile
  unsigned bar_phase = 0;
  if (!PEER && P.bulk && tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((unsigned)__cvta_generic_to_shared(&tile_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
  }

  // persistent CTAs: the grid is sized by the host (occupancy x SMs granted to this kernel) and strides over the tiles
#pragma unroll 1
  for (unsigned tile = P.tile_begin + blockIdx.x; tile < P.tile_begin + P.ntiles; tile += gridDim.x) {
  const unsigned long long at = qexpand(tile, P.nfixed_seg, P.xs_l, P.xs_n, P.xs_g) | at_tid;  // per-thread address part
  __syncthreads();  // the previous tile is no longer read (first trip: orders the table stores)

  // ---- stage the x tile.  Local tiles: TMA bulk copies (cp.async.bulk -> SASS UBLKCP), one per contiguous run of the
  // low free block, issued by warp 0 and tracked by an mbarrier, so the fill costs no LSU wavefronts and no per-thread
  // address arithmetic.  Peer tiles (and QOB_QTILE_TMA=0): 16-byte cp.async per amplitude.
  if (!PEER && P.bulk) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&tile_bar);
    if (tid < 32) {
      asm volatile("fence.proxy.async.shared::cta;\n" ::);  // earlier generic-proxy reads of xs precede the async writes
      const unsigned nruns = (unsigned)TILE >> P.run_log2, run_bytes = 16u << P.run_log2;
      if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"((unsigned)TILE * 16u));
      __syncwarp();
      const unsigned long long tbase = at & ~at_tid;  // the tile's base address (fixed bits only)
      for (unsigned r = tid; r < nruns; r += 32) {
        const unsigned long long a = tbase | qexpand(r << P.run_log2, P.nfree_seg, P.fs_l, P.fs_n, P.fs_g);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(xsB + (size_t)r * run_bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                     "l"(x + a), "r"(run_bytes), "r"(bar)
                     : "memory");
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const unsigned l = k * THREADS + tid;
      const unsigned saddr = (unsigned)__cvta_generic_to_shared(xsB + (size_t)l * sizeof(double2));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(qaddr<PEER>(P, x, P.xpeer, at | P.eoff[k])));
    }
    asm volatile("cp.async.commit_group;\n" ::);
  }
  // read-modify-write passes: pull this tile's y (and z) lines into L2 now (no registers held), so that the
  // epilogue's loads find them there instead of paying the DRAM latency after the compute phase.  The low 3
  // tile-local bits are always contiguous address bits (L >= 3): one prefetch per 128-byte line.
  if (!PEER && (tid & 7u) == 0u) {
    if (P.mode != 0) {
#pragma unroll
      for (int k = 0; k < PER; ++k) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(y + (at | P.eoff[k])));
    }
    if (P.zadd != nullptr) {
#pragma unroll
      for (int k = 0; k < PER; ++k) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(P.zadd + (at | P.eoff[k])));
    }
  }
  if (!PEER && P.bulk) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&tile_bar);
    unsigned done = 0;
    while (!done) {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(done)
                   : "r"(bar), "r"(bar_phase)
                   : "memory");
    }
    bar_phase ^= 1u;
  } else {
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
  }

  // thread part of the logical index (tile id, tid and rank bits); the amplitude-dependent part eoff[k] is warp-uniform
  const unsigned g_lo = (unsigned)(at | P.hi_or), g_hi = (unsigned)((at | P.hi_or) >> 32);
  // ---- diagonal weight that does not depend on which of its amplitudes the thread is working on: once per thread
  double2 dthread = make_double2(0.0, 0.0);
  for (int c = 0; c < P.npre; ++c) {
    const QCompDev cd = P.comps[c];
    const double2 w = *reinterpret_cast<const double2 *>(tabB + cd.tabB + qfield<IDX64>(cd.sel, g_lo, g_hi) * 16u);
    dthread.x += w.x;
    dthread.y += w.y;
  }

#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    double2 acc[U];
    if (P.has_diag) {  // acc = (sum of diagonal weights) * x[i]
      double2 d[U];
#pragma unroll
      for (int u = 0; u < U; ++u) d[u] = dthread;
      for (int c = c_diag; c < c_multi; ++c) {
        const QCompDev cd = P.comps[c];
        const unsigned char *wt = tabB + cd.tabB + qfield<IDX64>(cd.sel, g_lo, g_hi) * 16u;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const double2 w = *reinterpret_cast<const double2 *>(wt + P.ck[c][it * U + u].x);
          d[u].x += w.x;
          d[u].y += w.y;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const double2 v = *reinterpret_cast<const double2 *>(xsB + lbt + (unsigned)(it * U + u) * (THREADS * 16u));
        acc[u] = make_double2(d[u].x * v.x - d[u].y * v.y, d[u].x * v.y + d[u].y * v.x);
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = make_double2(0.0, 0.0);
    }
    if (P.nmulti) {  // several lookups per mask (terms that share a flip mask but not their selector bits): rare
      double2 ws[U];
#pragma unroll
      for (int u = 0; u < U; ++u) ws[u] = make_double2(0.0, 0.0);
      for (int c = c_multi; c < c_single; ++c) {
        const QCompDev cd = P.comps[c];
        const unsigned char *wt = tabB + cd.tabB + qfield<IDX64>(cd.sel, g_lo, g_hi) * 16u;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const double2 w = *reinterpret_cast<const double2 *>(wt + P.ck[c][it * U + u].x);
          ws[u].x += w.x;
          ws[u].y += w.y;
        }
        if (cd.flags & 1u) {
          const unsigned xT = lbt ^ (cd.xorB & TMASKB);
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const double2 v = *reinterpret_cast<const double2 *>(xsB + xT + P.ck[c][it * U + u].y);
            qcfma(acc[u], ws[u], v);
            ws[u] = make_double2(0.0, 0.0);
          }
        }
      }
    }
    // hot loop — one record per bond.  Everything that depends only on the record and on (it, u) is warp-uniform
    // (records and eoff live in the kernel parameters), so per (bond, amplitude) the vector pipes see:
    //   LDS (weight; skipped when the weight is the same for all of the thread's amplitudes), LDS.128 (gather), DFMAs.
    for (int c = c_single; c < c_end; ++c) {
      const QCompDev cd = P.comps[c];
      const unsigned char *wt = tabB + cd.tabB + qfield<IDX64>(cd.sel, g_lo, g_hi) * 16u;   // thread part of the lookup
      const unsigned char *xt = xsB + (lbt ^ (cd.xorB & TMASKB));                            // thread part of the gather
      if (cd.flags & 4u) {  // selector bits do not include any bit that varies between this thread's amplitudes
        QW<REALW> w;
        w.load(wt);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (u == 4) asm volatile("" ::: "memory");  // two batches of 4 loads in flight: keeps the kernel at 80 registers
          const double2 v = *reinterpret_cast<const double2 *>(xt + P.ck[c][it * U + u].y);
          w.fma_into(acc[u], v);
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (u == 4) asm volatile("" ::: "memory");
          const uint2 o = P.ck[c][it * U + u];
          QW<REALW> w;
          w.load(wt + o.x);
          const double2 v = *reinterpret_cast<const double2 *>(xt + o.y);
          w.fma_into(acc[u], v);
        }
      }
      asm volatile("" ::: "memory");
    }
    // ---- epilogue: y read-modify-write in batches of 4 independent 16-byte accesses per thread
    constexpr int EB = U < 4 ? U : 4;
#pragma unroll
    for (int u0 = 0; u0 < U; u0 += EB) {
      double2 t[EB];  // beta*y (+ z): everything that is added to alpha*acc
#pragma unroll
      for (int u = 0; u < EB; ++u) t[u] = make_double2(0.0, 0.0);
      if (P.mode != 0) {
#pragma unroll
        for (int u = 0; u < EB; ++u) t[u] = *qaddr<PEER>(P, y, P.ypeer, at | P.eoff[it * U + u0 + u]);
        if (P.mode == 1) {
#pragma unroll
          for (int u = 0; u < EB; ++u)
            t[u] = make_double2(P.beta.x * t[u].x - P.beta.y * t[u].y, P.beta.x * t[u].y + P.beta.y * t[u].x);
        }
      }
      if (!PEER && P.zadd != nullptr) {
#pragma unroll
        for (int u = 0; u < EB; ++u) {
          const double2 z = P.zadd[at | P.eoff[it * U + u0 + u]];
          t[u].x += z.x;
          t[u].y += z.y;
        }
      }
#pragma unroll
      for (int u = 0; u < EB; ++u) {
        double2 o;
        o.x = fma(P.alpha.x, acc[u0 + u].x, fma(-P.alpha.y, acc[u0 + u].y, t[u].x));
        o.y = fma(P.alpha.x, acc[u0 + u].y, fma(P.alpha.y, acc[u0 + u].x, t[u].y));
        *qaddr<PEER>(P, y, P.ypeer, at | P.eoff[it * U + u0 + u]) = o;
      }
      asm volatile("" ::: "memory");
    }
  }
  }  // tile loop
}

// ---------------------------------------------------------------------------------------- host
// A "component" = all terms that share (flip mask, selector bits): weight table of 2^nsel entries.
struct QCompHost {
  uint64_t mask = 0;
  std::vector<int> sel;  // selector bit positions (global), ascending
  struct Contrib {
    int coef_index;
    cplx scalar;
    std::vector<cplx> unit;  // 2^nsel
  };
  std::vector<Contrib> contribs;
  int pass = -1;
};
// A device table: either one off-diagonal component, or several diagonal components merged over a bit window.
struct QTableHost {
  std::vector<int> bits;  // index bits that form the table index, ascending (index bit j of the table <-> bits[j])
  struct Member {
    int comp;
    std::vector<int> pos;  // for each selector bit of the component: its position inside `bits`
  };
  std::vector<Member> members;
  uint32_t tab_off = 0;  // in double2 units inside the pass table
};
struct QPassHost {
  std::vector<int> free_bits;  // ascending, size T
  std::vector<QTableHost> tables;
  QPassParams params;
  DevArray<double2> d_tab;
  std::vector<double2> h_tab;
  bool real_weights = false;  // every off-diagonal weight of this pass has a zero imaginary part (checked per coefficient set)
  std::vector<std::pair<uint32_t, uint32_t>> offdiag_ranges;  // [begin, end) table ranges of the off-diagonal "single" lookups
};
struct QTileProgramHost {
  int nbits = 0, T = 0, L = 0, threads = 256;
  bool idx64 = false;
  uint64_t hi_value = 0;
  std::vector<QCompHost> comps;
  std::vector<std::unique_ptr<QPassHost>> passes;
  size_t smem_bytes(const QPassHost &p) const {
    return ((size_t)1 << T) * sizeof(double2) + p.params.ntab * sizeof(double2);
  }
};

static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

static void make_segments(const std::vector<int> &bits, unsigned char *sl, unsigned char *sn, unsigned char *sg,
                          int &nseg) {
  nseg = 0;
  size_t i = 0;
  while (i < bits.size()) {
    size_t j = i;
    while (j + 1 < bits.size() && bits[j + 1] == bits[j] + 1) ++j;
    if (nseg < QT_MAXSEG) {
      sl[nseg] = (unsigned char)i;
      sn[nseg] = (unsigned char)(j - i + 1);
      sg[nseg] = (unsigned char)bits[i];
    }
    ++nseg;
    i = j + 1;
  }
}

// bit-field runs of an ascending bit list -> packed `sel` word; false when more than 3 runs / run wider than 7
static bool pack_runs(const std::vector<int> &bits, uint32_t &sel) {
  sel = 0;
  int r = 0;
  size_t i = 0;
  while (i < bits.size()) {
    size_t j = i;
    while (j + 1 < bits.size() && bits[j + 1] == bits[j] + 1) ++j;
    const int w = (int)(j - i + 1);
    if (r >= 3 || w > 7 || bits[i] > 63) return false;
    sel |= ((uint32_t)bits[i] | ((uint32_t)w << 6)) << (9 * r);
    ++r;
    i = j + 1;
  }
  return true;
}

static cplx comp_value(const QCompHost &c, const std::vector<cplx> &coefs, size_t r) {
  cplx w = 0.0;
  for (const auto &ct : c.contribs) {
    cplx f = ct.scalar;
    if (ct.coef_index >= 0) f *= coefs[ct.coef_index];
    w += f * ct.unit[r];
  }
  return w;
}

static void fill_tables(QTileProgramHost &h, const std::vector<cplx> &coefs) {
  for (auto &pp : h.passes) {
    QPassHost &p = *pp;
    std::fill(p.h_tab.begin(), p.h_tab.end(), make_double2(0.0, 0.0));
    for (const QTableHost &t : p.tables) {
      const size_t n = (size_t)1 << t.bits.size();
      for (const auto &m : t.members) {
        const QCompHost &c = h.comps[m.comp];
        // component values once, then scattered over the (possibly wider) merged table
        std::vector<cplx> cv((size_t)1 << c.sel.size());
        for (size_t r = 0; r < cv.size(); ++r) cv[r] = comp_value(c, coefs, r);
        for (size_t r = 0; r < n; ++r) {
          size_t ci = 0;
          for (size_t b = 0; b < m.pos.size(); ++b) ci |= ((r >> m.pos[b]) & 1) << b;
          p.h_tab[t.tab_off + r].x += cv[ci].real();
          p.h_tab[t.tab_off + r].y += cv[ci].imag();
        }
      }
    }
    p.real_weights = true;
    for (auto &rg : p.offdiag_ranges)
      for (uint32_t e = rg.first; e < rg.second; ++e) p.real_weights &= (p.h_tab[e].y == 0.0);
  }
}

// merge diagonal components into tables over windows of <= QT_DIAG_WINDOW index bits
static void merge_diag(const QTileProgramHost &h, std::vector<int> ids, std::vector<QTableHost> &out) {
  std::sort(ids.begin(), ids.end(), [&](int a, int b) {
    const auto &sa = h.comps[a].sel, &sb = h.comps[b].sel;
    int la = sa.empty() ? -1 : sa.front(), lb = sb.empty() ? -1 : sb.front();
    if (la != lb) return la < lb;
    return sa < sb;
  });
  std::vector<bool> used(ids.size(), false);
  for (size_t i = 0; i < ids.size(); ++i) {
    if (used[i]) continue;
    QTableHost t;
    std::vector<int> bits = h.comps[ids[i]].sel;
    std::vector<size_t> mem = {i};
    used[i] = true;
    for (size_t j = i + 1; j < ids.size(); ++j) {
      if (used[j]) continue;
      std::vector<int> u = bits;
      for (int b : h.comps[ids[j]].sel)
        if (std::find(u.begin(), u.end(), b) == u.end()) u.push_back(b);
      std::sort(u.begin(), u.end());
      uint32_t dummy;
      if ((int)u.size() <= QT_DIAG_WINDOW && pack_runs(u, dummy)) {
        bits = u;
        mem.push_back(j);
        used[j] = true;
      }
    }
    t.bits = bits;
    for (size_t m : mem) {
      QTableHost::Member mm;
      mm.comp = ids[m];
      for (int b : h.comps[ids[m]].sel)
        mm.pos.push_back((int)(std::find(bits.begin(), bits.end(), b) - bits.begin()));
      t.members.push_back(mm);
    }
    out.push_back(std::move(t));
  }
}

int qtile_build(QTileProgram &prog, int nbits, uint64_t hi_value, const std::vector<QTerm> &terms, int sm_count) {
  auto h = std::make_shared<QTileProgramHost>();
  h->nbits = nbits;
  h->hi_value = hi_value;
  int T = env_int("QOB_QTILE_T", 12);
  if (T > 13) T = 13;
  if (T < 10) T = 10;
  if (T > nbits) T = nbits;
  if (T < 10) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile needs at least 10 index bits (got %d)", nbits);
  int L = env_int("QOB_QTILE_L", 3);
  if (L < 3) L = 3;  // the kernel prefetches y by 128-byte lines: at least 8 contiguous amplitudes per run
  if (L > T - 3) L = T - 3;
  h->T = T;
  h->L = L;
  h->threads = (T == 13) ? 512 : 256;
  const int tid_bits = (T == 13) ? 9 : 8;  // log2(threads): tile-local bits below this are fixed per thread

  // ---- expand every term into flip/no-flip components, merged by (mask, selector bits)
  std::map<std::pair<uint64_t, std::vector<int>>, int> index;
  int max_bit = nbits - 1;
  for (const QTerm &t : terms) {
    const int k = (int)t.bits.size();
    if (k > QT_MAXSEL) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: term on %d sites (max %d)", k, QT_MAXSEL);
    for (int b : t.bits) max_bit = std::max(max_bit, b);
    for (int S = 0; S < (1 << k); ++S) {
      std::vector<cplx> unit((size_t)1 << k);
      bool any = false;
      for (int r = 0; r < (1 << k); ++r) {
        cplx w = 1.0;
        for (int f = 0; f < k; ++f) {
          int i = (r >> f) & 1;
          int j = ((S >> f) & 1) ? 1 - i : i;
          w *= t.m[4 * f + 2 * i + j];
        }
        unit[r] = w;
        any |= (w != cplx(0.0, 0.0));
      }
      if (!any) continue;
      uint64_t mask = 0;
      for (int f = 0; f < k; ++f)
        if ((S >> f) & 1) {
          if (t.bits[f] >= nbits) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: off-diagonal factor on a non-local bit %d", t.bits[f]);
          mask |= 1ull << t.bits[f];
        }
      auto key = std::make_pair(mask, t.bits);
      auto it = index.find(key);
      int id;
      if (it == index.end()) {
        id = (int)h->comps.size();
        index[key] = id;
        QCompHost c;
        c.mask = mask;
        c.sel = t.bits;
        h->comps.push_back(c);
      } else {
        id = it->second;
      }
      h->comps[id].contribs.push_back({t.coef_index, t.scalar, unit});
    }
  }
  h->idx64 = max_bit >= 32;

  // ---- cover the distinct non-zero masks with passes of T free bits
  std::vector<uint64_t> masks;
  for (auto &c : h->comps)
    if (c.mask && std::find(masks.begin(), masks.end(), c.mask) == masks.end()) masks.push_back(c.mask);
  // Greedy cover for a given low-block width L and a cap on the number of free bits whose stride is >= one 2 MiB page
  // (bit 17 and above): every such bit doubles the number of pages a tile touches (measured: 512 pages per operand
  // cost ~25 % of a pass to TLB misses).  A small cost model picks the cheapest (L, cap) combination.
  const int page_bit = 17;
  auto cover = [&](int Lc, int cap, std::vector<uint64_t> &out) -> bool {
    out.clear();
    out.push_back((1ull << T) - 1);  // pass 0: the lowest T bits (fully contiguous tiles)
    std::vector<uint64_t> remaining;
    for (uint64_t m : masks)
      if (m & ~out[0]) remaining.push_back(m);
    const uint64_t lowL = (1ull << Lc) - 1, hiMask = ~((1ull << page_bit) - 1);
    auto min_high = [&](uint64_t m) { uint64_t hgh = m & ~lowL; return hgh ? __builtin_ctzll(hgh) : 64; };
    std::sort(remaining.begin(), remaining.end(), [&](uint64_t a, uint64_t b) {
      int ha = min_high(a), hb = min_high(b);
      if (ha != hb) return ha < hb;
      return a < b;
    });
    while (!remaining.empty()) {
      uint64_t fr = lowL;
      std::vector<uint64_t> rest;
      for (uint64_t m : remaining) {
        const uint64_t u = fr | m;
        if (__builtin_popcountll(u) <= T && __builtin_popcountll(u & hiMask) <= cap) fr = u;
        else rest.push_back(m);
      }
      if (fr == lowL) return false;
      for (int b = 0; b < nbits && __builtin_popcountll(fr) < T; ++b) fr |= 1ull << b;  // spare bits widen the low block
      std::vector<uint64_t> rest2;
      for (uint64_t m : rest)
        if (m & ~fr) rest2.push_back(m);
      remaining.swap(rest2);
      out.push_back(fr);
    }
    return true;
  };
  auto plan_cost = [&](const std::vector<uint64_t> &fs) {
    // pass 0 (contiguous tiles) is bound by the shared-memory pipe (~4.1 TB/s equivalent at 32 B, ~5.5 at 48 B); the most
    // expensive window pass runs first as the write-only pass (32 B/amplitude), all others read-modify-write (48 B)
    double cost = fs.size() > 1 ? 48.0 / 5.5 : 32.0 / 4.1, worst = 0.0;
    for (size_t p = 1; p < fs.size(); ++p) {
      int low = 0;
      while (low < nbits && (fs[p] >> low & 1)) ++low;
      const double bw = low <= 3 ? 5.4 : (low == 4 ? 6.0 : 6.6);  // TB/s measured for 128 / 256 / >= 512-byte runs
      const int hb = __builtin_popcountll(fs[p] & ~((1ull << page_bit) - 1));
      const double tlb = hb <= 7 ? 1.0 : (hb == 8 ? 0.92 : 0.78);
      const double c = 48.0 / (bw * tlb);
      cost += c;
      worst = std::max(worst, c);
    }
    return cost - worst / 3.0;
  };
  std::vector<uint64_t> free_sets;
  {
    double best = 1e300;
    const bool forced = getenv("QOB_QTILE_L") != nullptr;
    for (int Lc = forced ? L : 3; Lc <= (forced ? L : 5) && Lc <= T - 3; ++Lc)
      for (int cap = 7; cap <= 9; ++cap) {
        std::vector<uint64_t> fs;
        if (!cover(Lc, cap, fs)) continue;
        const double c = plan_cost(fs);
        if (c < best - 1e-9) {
          best = c;
          free_sets = fs;
          h->L = Lc;
        }
      }
    if (free_sets.empty()) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: a term does not fit one tile");
    L = h->L;
  }

  // ---- assign components to passes.  Diagonal weights go to pass 0.  An off-diagonal mask that fits several passes
  // (bonds inside the low block, which every pass carries) goes to the least loaded one: pass 0 is bound by the
  // shared-memory pipe (one gather per bond and amplitude) while the window passes have slack under their HBM time.
  {
    std::vector<int> load(free_sets.size(), 0);
    std::map<uint64_t, int> mask_pass;
    // masks with a single candidate first (they fix the loads), then the movable ones
    for (int movable = 0; movable < 2; ++movable)
      for (auto &c : h->comps) {
        if (!c.mask) {
          c.pass = 0;
          continue;
        }
        std::vector<int> cand;
        for (size_t p = 0; p < free_sets.size(); ++p)
          if ((c.mask & ~free_sets[p]) == 0) cand.push_back((int)p);
        if (cand.empty()) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: internal planner error (uncovered mask)");
        if ((cand.size() > 1) != (movable == 1)) continue;
        auto it = mask_pass.find(c.mask);
        if (it != mask_pass.end()) {  // components that share a mask must share the gather
          c.pass = it->second;
          continue;
        }
        int best = cand[0];
        for (int p : cand)
          if (load[p] < load[best] || (load[p] == load[best] && p > best)) best = p;
        c.pass = best;
        mask_pass[c.mask] = best;
        load[best]++;
      }
  }
  for (size_t p = 0; p < free_sets.size(); ++p) {
    std::vector<int> free_bits;
    for (int b = 0; b < nbits; ++b)
      if (free_sets[p] >> b & 1) free_bits.push_back(b);
    // index bits that change between the amplitudes one thread owns: tile-local bits >= tid_bits
    uint64_t varying = 0;
    for (size_t j = tid_bits; j < free_bits.size(); ++j) varying |= 1ull << free_bits[j];
    std::vector<int> diag_pre, diag_amp, flips;
    for (size_t id = 0; id < h->comps.size(); ++id) {
      const QCompHost &c = h->comps[id];
      if (c.pass != (int)p) continue;
      if (c.mask) {
        flips.push_back((int)id);
      } else {
        uint64_t sb = 0;
        for (int b : c.sel) sb |= 1ull << b;
        ((sb & varying) ? diag_amp : diag_pre).push_back((int)id);
      }
    }
    if (p > 0 && flips.empty() && diag_pre.empty() && diag_amp.empty()) continue;
    std::stable_sort(flips.begin(), flips.end(), [&](int a, int b) { return h->comps[a].mask < h->comps[b].mask; });
    std::vector<QTableHost> pre_tabs, amp_tabs;
    merge_diag(*h, diag_pre, pre_tabs);
    merge_diag(*h, diag_amp, amp_tabs);
    // off-diagonal lookups: groups that share a mask ("multi", summed before the gather) and singles
    std::vector<int> multis, singles;
    for (size_t i = 0; i < flips.size(); ++i) {
      const uint64_t m = h->comps[flips[i]].mask;
      const bool first = (i == 0) || h->comps[flips[i - 1]].mask != m;
      const bool last = (i + 1 == flips.size()) || h->comps[flips[i + 1]].mask != m;
      ((first && last) ? singles : multis).push_back(flips[i]);
    }
    if (pre_tabs.size() + amp_tabs.size() + multis.size() > QT_MAXCOMP - 8)
      QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: too many diagonal / shared-mask lookups for one pass");

    // the records live in the kernel parameters: if the singles do not fit, the overflow runs as extra passes over
    // the same tiles (each a read-modify-write of y)
    size_t s_begin = 0;
    bool first_chunk = true;
    do {
      auto ph = std::make_unique<QPassHost>();
      ph->free_bits = free_bits;
      std::vector<QCompDev> dc;
      uint32_t tab_off = 1;  // entry 0 is a zero weight
      auto emit = [&](QTableHost &t, uint32_t xorB, uint32_t flags) -> int {
        QCompDev d;
        d.xorB = xorB;
        if (!pack_runs(t.bits, d.sel)) return QOB_STATUS_UNSUPPORTED;
        d.tabB = tab_off * (uint32_t)sizeof(double2);
        d.flags = flags;
        t.tab_off = tab_off;
        tab_off += 1u << t.bits.size();
        dc.push_back(d);
        ph->tables.push_back(t);
        return QOB_STATUS_OK;
      };
      auto flip_table = [&](int id, QTableHost &t, uint32_t &lmask, bool &invariant) {
        const QCompHost &c = h->comps[id];
        lmask = 0;
        for (size_t j = 0; j < free_bits.size(); ++j)
          if (c.mask >> free_bits[j] & 1) lmask |= 1u << j;
        t.bits = c.sel;
        QTableHost::Member m;
        m.comp = id;
        uint64_t sb = 0;
        for (size_t b = 0; b < c.sel.size(); ++b) {
          m.pos.push_back((int)b);
          sb |= 1ull << c.sel[b];
        }
        t.members.push_back(m);
        invariant = !(sb & varying);
      };
      int npre = 0, ndiag = 0, nmulti = 0, nsingle = 0;
      bool has_diag = false;
      if (first_chunk) {
        for (auto &t : pre_tabs)
          if (emit(t, 0, 0) != QOB_STATUS_OK) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: selector bits too scattered");
        npre = (int)dc.size();
        for (auto &t : amp_tabs)
          if (emit(t, 0, 0) != QOB_STATUS_OK) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: selector bits too scattered");
        ndiag = (int)dc.size() - npre;
        has_diag = npre + ndiag > 0;
        for (size_t i = 0; i < multis.size(); ++i) {
          QTableHost t;
          uint32_t lmask;
          bool inv;
          flip_table(multis[i], t, lmask, inv);
          const bool last = (i + 1 == multis.size()) || h->comps[multis[i + 1]].mask != h->comps[multis[i]].mask;
          if (emit(t, lmask * 16u, last ? 1u : 0u) != QOB_STATUS_OK)
            QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: selector bits too scattered");
          ++nmulti;
        }
      }
      while (s_begin < singles.size() && dc.size() < QT_MAXCOMP) {
        QTableHost t;
        uint32_t lmask;
        bool inv;
        flip_table(singles[s_begin], t, lmask, inv);
        const uint32_t t0 = tab_off;
        if (emit(t, lmask * 16u, 1u | 2u | (inv ? 4u : 0u)) != QOB_STATUS_OK)
          QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: selector bits too scattered");
        ph->offdiag_ranges.push_back({t0, tab_off});
        ++nsingle;
        ++s_begin;
      }
      ph->h_tab.assign(tab_off, make_double2(0.0, 0.0));
      QPassParams &P = ph->params;
      memset(&P, 0, sizeof(P));
      P.npre = npre;
      P.ndiag = ndiag;
      P.nmulti = nmulti;
      P.nsingle = nsingle;
      P.ntab = (int)ph->h_tab.size();
      P.has_diag = has_diag ? 1 : 0;
      for (size_t i = 0; i < dc.size(); ++i) P.comps[i] = dc[i];
      {
        const unsigned tmaskb = (unsigned)h->threads * 16u - 1u;
        for (size_t i = 0; i < dc.size(); ++i)
          for (int k = 0; k < (1 << T) / h->threads && k < 16; ++k) {
            // address part of tile-local index k*threads, then the table field of that part (bitwise in the index)
            unsigned long long e = 0;
            const unsigned v = (unsigned)k * (unsigned)h->threads;
            for (size_t j = 0; j < free_bits.size(); ++j) e |= (unsigned long long)((v >> j) & 1u) << free_bits[j];
            unsigned idx = 0, pos = 0;
            for (int r = 0; r < 3; ++r) {
              const unsigned sh = (dc[i].sel >> (9 * r)) & 63u, w = (dc[i].sel >> (9 * r + 6)) & 7u;
              idx |= (unsigned)((e >> sh) & ((1ull << w) - 1ull)) << pos;
              pos += w;
            }
            P.ck[i][k].x = idx * 16u;
            P.ck[i][k].y = (v * 16u) ^ (dc[i].xorB & ~tmaskb);
          }
      }
      make_segments(free_bits, P.fs_l, P.fs_n, P.fs_g, P.nfree_seg);
      std::vector<int> fixed;
      for (int b = 0; b < nbits; ++b)
        if (!(free_sets[p] >> b & 1)) fixed.push_back(b);
      make_segments(fixed, P.xs_l, P.xs_n, P.xs_g, P.nfixed_seg);
      if (P.nfree_seg > QT_MAXSEG || P.nfixed_seg > QT_MAXSEG) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: too many index segments");
      P.hi_or = (nbits >= 64) ? 0ull : (hi_value << nbits);
      for (int k = 0; k < (1 << T) / h->threads && k < 16; ++k) {
        unsigned long long e = 0;
        const unsigned v = (unsigned)k * (unsigned)h->threads;
        for (size_t j = 0; j < free_bits.size(); ++j) e |= (unsigned long long)((v >> j) & 1u) << free_bits[j];
        P.eoff[k] = e;
      }
      QOB_TRY(ph->d_tab.upload(ph->h_tab));
      P.tab = ph->d_tab.ptr;
      h->passes.push_back(std::move(ph));
      first_chunk = false;
    } while (s_begin < singles.size());
  }
  // ---- order: the first executed pass only WRITES y (32 B/amplitude), the others read-modify-write it (48 B).  Pass 0
  // (contiguous tiles, most bonds) is bound by the shared-memory pipe and barely notices the extra y read, while the
  // window passes are HBM/TLB bound: so the most expensive window pass goes first and saves a third of its traffic.
  if (h->passes.size() > 1 && !(getenv("QOB_QTILE_FIRST") && atoi(getenv("QOB_QTILE_FIRST")) == 0)) {
    size_t best = 0;
    double best_cost = 0.0;
    for (size_t i = 0; i < h->passes.size(); ++i) {
      const QPassHost &ph = *h->passes[i];
      const QPassParams &q = ph.params;
      if (q.npre + q.ndiag + q.nmulti + q.nsingle == 0) continue;
      uint64_t fs = 0;
      for (int b : ph.free_bits) fs |= 1ull << b;
      if (fs == ((1ull << T) - 1)) continue;  // the contiguous pass stays a read-modify-write pass
      int low = 0;
      while (low < nbits && (fs >> low & 1)) ++low;
      const double bw = low <= 3 ? 5.4 : (low == 4 ? 6.0 : 6.6);
      const int hb = __builtin_popcountll(fs & ~((1ull << page_bit) - 1));
      const double tlb = hb <= 7 ? 1.0 : (hb == 8 ? 0.92 : 0.78);
      const double cost = 48.0 / (bw * tlb);
      if (cost > best_cost) {
        best_cost = cost;
        best = i;
      }
    }
    if (best_cost > 0.0 && best != 0) std::rotate(h->passes.begin(), h->passes.begin() + best, h->passes.begin() + best + 1);
  }
  prog.h = h;
  prog.hi_value = hi_value;
  prog.npasses = (int)h->passes.size();
  char buf[256];
  snprintf(buf, sizeof buf, "qtile[bits=%d,T=%d,L=%d,passes=%d,components=%d]", nbits, T, L, prog.npasses,
           (int)h->comps.size());
  prog.describe = buf;
  for (auto &pp : h->passes) {
    prog.describe += " {free:";
    int nseg = pp->params.nfree_seg;
    for (int s = 0; s < nseg; ++s) {
      snprintf(buf, sizeof buf, "%s%d-%d", s ? "," : "", pp->params.fs_g[s], pp->params.fs_g[s] + pp->params.fs_n[s] - 1);
      prog.describe += buf;
    }
    int ninv = 0;
    for (int c = 0; c < pp->params.npre + pp->params.ndiag + pp->params.nmulti + pp->params.nsingle; ++c)
      ninv += (pp->params.comps[c].flags & 4u) ? 1 : 0;
    snprintf(buf, sizeof buf, " lookups:%d diag+%d multi+%d single(%d thread-invariant) per amplitude, %d per thread; tables:%dB}",
             pp->params.ndiag, pp->params.nmulti, pp->params.nsingle, ninv, pp->params.npre,
             (int)(pp->params.ntab * sizeof(double2)));
    prog.describe += buf;
  }
  (void)sm_count;
  return QOB_STATUS_OK;
}

// number of passes that hold work, and the index bits that are fixed (not free) in every one of them
void qtile_info(const QTileProgram &prog, int *npasses, uint64_t *fixed_mask) {
  const QTileProgramHost &h = *prog.h;
  int n = 0;
  uint64_t fixed = h.nbits >= 64 ? ~0ull : ((1ull << h.nbits) - 1);
  for (auto &pp : h.passes) {
    const QPassParams &q = pp->params;
    if (q.npre + q.ndiag + q.nmulti + q.nsingle == 0) continue;
    ++n;
    for (int b : pp->free_bits) fixed &= ~(1ull << b);
  }
  if (npasses) *npasses = n;
  if (fixed_mask) *fixed_mask = n ? fixed : 0;
}

// Reorder the tile numbering of every pass so that the bits of `chunk_mask` (which must be fixed bits) are the MOST
// significant bits of the tile id: a contiguous range of tile ids then corresponds to fixed values of those index bits,
// the same amplitudes in every plan that uses the same chunk bits.
int qtile_set_chunk_bits(QTileProgram &prog, uint64_t chunk_mask) {
  QTileProgramHost &h = *prog.h;
  for (auto &pp : h.passes) {
    const QPassParams &q0 = pp->params;
    if (q0.npre + q0.ndiag + q0.nmulti + q0.nsingle == 0) continue;   // empty passes are never launched
    uint64_t fs = 0;
    for (int b : pp->free_bits) fs |= 1ull << b;
    if (fs & chunk_mask) QOB_FAIL(QOB_STATUS_INVALID_ARG, "chunk bits must not be free bits of a pass");
    std::vector<int> lo, hi;
    for (int b = 0; b < h.nbits; ++b) {
      if (fs >> b & 1) continue;
      ((chunk_mask >> b & 1) ? hi : lo).push_back(b);
    }
    QPassParams &P = pp->params;
    int nlo = 0, nhi = 0;
    unsigned char sl[QT_MAXSEG], sn[QT_MAXSEG], sg[QT_MAXSEG];
    make_segments(lo, P.xs_l, P.xs_n, P.xs_g, nlo);
    make_segments(hi, sl, sn, sg, nhi);
    if (nlo + nhi > QT_MAXSEG) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: too many index segments");
    for (int i = 0; i < nhi; ++i) {
      P.xs_l[nlo + i] = (unsigned char)(sl[i] + lo.size());
      P.xs_n[nlo + i] = sn[i];
      P.xs_g[nlo + i] = sg[i];
    }
    P.nfixed_seg = nlo + nhi;
  }
  return QOB_STATUS_OK;
}

int qtile_set_coefs(QTileProgram &prog, const std::vector<cplx> &coefs, cudaStream_t s) {
  QTileProgramHost &h = *prog.h;
  for (auto &c : h.comps)
    for (auto &ct : c.contribs)
      if (ct.coef_index >= (int)coefs.size()) QOB_FAIL(QOB_STATUS_INVALID_ARG, "coefficient index out of range");
  fill_tables(h, coefs);
  for (auto &pp : h.passes) QOB_TRY(pp->d_tab.upload_async(pp->h_tab, s));
  return QOB_STATUS_OK;
}

template <int T, int THREADS, int MINB, bool IDX64, bool REALW, bool PEER>
static int launch_pass(const QTileProgramHost &h, const QPassHost &p, const QPassParams &P, const void *x, void *y,
                       int max_ctas, cudaStream_t s) {
  size_t smem = h.smem_bytes(p);
  {
    // the opt-in is per device (and per template instantiation): remember the largest size configured on each device
    static std::atomic<size_t> configured[QOB_MAX_DEVICES];
    int dev = 0;
    QOB_CUDA(cudaGetDevice(&dev));
    const int slot = dev >= 0 && dev < QOB_MAX_DEVICES ? dev : 0;
    if (dev != slot || smem > configured[slot].load(std::memory_order_relaxed)) {
      QOB_CUDA(cudaFuncSetAttribute(qtile_kernel<T, THREADS, MINB, IDX64, REALW, PEER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured[slot].store(smem, std::memory_order_relaxed);
    }
  }
  if ((1ull << (h.nbits - T)) > 0x7FFFFFFFull) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: too many tiles");
  const uint64_t ntiles = P.ntiles;
  uint64_t grid = std::min<uint64_t>(ntiles, (uint64_t)std::max(1, max_ctas));
  qtile_kernel<T, THREADS, MINB, IDX64, REALW, PEER><<<(unsigned)grid, THREADS, smem, s>>>(P, (const double2 *)x, (double2 *)y);
  QOB_LAUNCHED();
  QOB_LAUNCHED_FAMILY(PEER ? 4 : 1);
  QOB_CUDA(cudaGetLastError());
  return QOB_STATUS_OK;
}

// ---- optional per-launch event timing (qob_profile_enable / qob_profile_read)
struct ProfEntry {
  cudaEvent_t a, b;
  int pass;
  double bytes;
};
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfEntry> g_prof;
bool qprof_enabled() { return g_prof_on.load(std::memory_order_relaxed); }
void *qprof_begin(cudaStream_t s, int pass, double alg_bytes) {
  if (!qprof_enabled()) return nullptr;
  ProfEntry *pe = new ProfEntry;
  cudaEventCreate(&pe->a);
  cudaEventCreate(&pe->b);
  pe->pass = pass;
  pe->bytes = alg_bytes;
  cudaEventRecord(pe->a, s);
  return pe;
}
void qprof_end(cudaStream_t s, void *token) {
  if (!token) return;
  ProfEntry *pe = (ProfEntry *)token;
  cudaEventRecord(pe->b, s);
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(*pe);
  }
  delete pe;
}
extern "C" int qob_profile_enable(int32_t on) {
  g_prof_on.store(on != 0);
  return QOB_STATUS_OK;
}
extern "C" int qob_profile_read(int32_t max_entries, float *ms, int32_t *pass_index, double *alg_bytes, int32_t *count) {
  std::vector<ProfEntry> taken;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    taken.swap(g_prof);
  }
  int n = 0;
  for (ProfEntry &e : taken) {
    cudaEventSynchronize(e.b);
    float t = 0.f;
    cudaEventElapsedTime(&t, e.a, e.b);
    if (n < max_entries) {
      if (ms) ms[n] = t;
      if (pass_index) pass_index[n] = e.pass;
      if (alg_bytes) alg_bytes[n] = e.bytes;
      ++n;
    }
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  if (count) *count = n;
  return QOB_STATUS_OK;
}

static int g_sm_budget = 0;  // SMs the persistent tile kernels may occupy (0 = all of them)
extern "C" int qob_set_sm_budget(int32_t sms) {
  g_sm_budget = sms < 0 ? 0 : sms;
  return QOB_STATUS_OK;
}

int qtile_launch(const QTileProgram &prog, cplx alpha, const void *x, cplx beta, void *y, cudaStream_t s,
                 const QLaunchOpts *opts) {
  const QTileProgramHost &h = *prog.h;
  const int sm_count = qob_device_sm_count();
  QLaunchOpts none;
  const QLaunchOpts &o = opts ? *opts : none;
  if (o.npeers > QT_MAXPEER) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "more than %d peers", QT_MAXPEER);
  int sms = o.sm_budget > 0 ? o.sm_budget : (g_sm_budget > 0 ? g_sm_budget : sm_count);
  sms = std::min(sms, sm_count);
  // passes that hold no lookup at all only matter for the beta update
  std::vector<const QPassHost *> run;
  for (auto &pp : h.passes) {
    const QPassParams &q = pp->params;
    if (q.npre + q.ndiag + q.nmulti + q.nsingle > 0) run.push_back(pp.get());
  }
  const int64_t n = (int64_t)1 << h.nbits;
  if (run.empty()) {
    if (o.npeers) QOB_FAIL(QOB_STATUS_INVALID_ARG, "peer-addressed apply of a plan without terms");
    if (o.zadd) return launch_axpby(o.zadd, y, n, cplx(1.0, 0.0), beta, s);
    return launch_scale(y, n, beta, s);
  }
  for (size_t pi = 0; pi < run.size(); ++pi) {
    const QPassHost *pp = run[pi];
    const bool first = pi == 0;
    void *prof_token = nullptr;
    if (qprof_enabled())
      prof_token = qprof_begin(s, (int)pi,
                               ((double)(1ull << h.nbits) * ((first && beta == cplx(0.0, 0.0)) ? 32.0 : 48.0) +
                                ((o.zadd && pi + 1 == run.size()) ? 16.0 * (double)(1ull << h.nbits) : 0.0)) /
                                   (double)std::max(1, o.nchunks));
    QPassParams P = pp->params;
    P.alpha = make_double2(alpha.real(), alpha.imag());
    P.beta = make_double2(beta.real(), beta.imag());
    P.mode = first ? (beta == cplx(0.0, 0.0) ? 0 : 1) : 2;
    P.ntiles = (unsigned)(1ull << (h.nbits - h.T));
    P.tile_begin = 0;
    if (o.nchunks > 1) {
      // a chunk = a contiguous range of tile ids = fixed values of the most significant fixed bits (see qtile_set_chunk_bits)
      if (run.size() != 1) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "chunked launches need a plan with exactly one pass");
      if (o.chunk_index < 0 || o.chunk_index >= o.nchunks || (P.ntiles % (unsigned)o.nchunks) != 0)
        QOB_FAIL(QOB_STATUS_INVALID_ARG, "bad chunk %d of %d", o.chunk_index, o.nchunks);
      P.ntiles /= (unsigned)o.nchunks;
      P.tile_begin = P.ntiles * (unsigned)o.chunk_index;
      // beta handling (mode) is per tile, so a chunked pass is just the same pass restricted to some tiles
    }
    {
      // TMA bulk staging pays when the runs are long (measured on N=28/30: the fully contiguous pass gains ~8 %, while
      // 512 x 128-byte or 128 x 512-byte bulk copies per tile are 15-80 % slower than per-thread cp.async).
      // QOB_QTILE_TMA = minimum log2(run length in amplitudes) for the bulk path; 0 disables it.
      static const int tma_min = getenv("QOB_QTILE_TMA") ? atoi(getenv("QOB_QTILE_TMA")) : 9;
      int low = 0;  // contiguous low block of this pass's free bits
      while (low < (int)pp->free_bits.size() && pp->free_bits[low] == low) ++low;
      P.bulk = (tma_min > 0 && low >= tma_min && o.npeers == 0) ? 1 : 0;
      P.run_log2 = low;
    }
    P.zadd = (pi + 1 == run.size()) ? (const double2 *)o.zadd : nullptr;
    P.peer_shift = o.peer_shift;
    P.peer_rank = o.peer_rank;
    P.peer_bits = 0;
    while ((1 << P.peer_bits) < o.npeers) ++P.peer_bits;
    for (int q = 0; q < o.npeers; ++q) {
      P.xpeer[q] = (const double2 *)o.xpeer[q];
      P.ypeer[q] = (double2 *)o.ypeer[q];
    }
    const bool rw = pp->real_weights && !getenv("QOB_QTILE_NO_REALW");
    const bool peer = o.npeers > 0;
    // persistent CTAs (grid = SM budget x occupancy) unless a full grid is requested and no SM budget is in force
    // (measured on N=28/30: one CTA per tile is ~15 % faster than grid-strided persistent CTAs, whose three resident CTAs
    // per SM fall into lockstep; persistence is only used to confine the kernel to an SM budget)
    static const bool env_persist = getenv("QOB_QTILE_PERSIST") && atoi(getenv("QOB_QTILE_PERSIST")) != 0;
    const bool persist = env_persist || sms < sm_count;
#define QT_CASE(TT, TH, MB)                                                                                     \
  case TT: {                                                                                                    \
    const int ctas = persist ? sms * MB : 0x7FFFFFFF;                                                           \
    if (peer) {                                                                                                 \
      if (rw) QOB_TRY((launch_pass<TT, TH, MB, true, true, true>(h, *pp, P, x, y, ctas, s)));                  \
      else QOB_TRY((launch_pass<TT, TH, MB, true, false, true>(h, *pp, P, x, y, ctas, s)));                    \
    } else if (h.idx64 && rw) QOB_TRY((launch_pass<TT, TH, MB, true, true, false>(h, *pp, P, x, y, ctas, s))); \
    else if (h.idx64) QOB_TRY((launch_pass<TT, TH, MB, true, false, false>(h, *pp, P, x, y, ctas, s)));        \
    else if (rw) QOB_TRY((launch_pass<TT, TH, MB, false, true, false>(h, *pp, P, x, y, ctas, s)));             \
    else QOB_TRY((launch_pass<TT, TH, MB, false, false, false>(h, *pp, P, x, y, ctas, s)));                    \
  } break;
    switch (h.T) {
      QT_CASE(10, 256, 3)
      QT_CASE(11, 256, 3)
      QT_CASE(12, 256, 3)
      QT_CASE(13, 512, 1)
      default: QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: unsupported tile size %d", h.T);
    }
#undef QT_CASE
    qprof_end(s, prof_token);
  }
  return QOB_STATUS_OK;
}
