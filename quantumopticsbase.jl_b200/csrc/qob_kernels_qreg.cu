// qob_kernels_qreg.cu — round-2 kernel of the fused LazySum apply on 2-dimensional subsystems (BASELINE configs 4, 5):
// register-blocked tile passes, staged by tensor-map TMA, with pairs of passes chained through L2.
//
// Reference behaviour replaced: one full pass over the state per TERM, src/operators_lazysum.jl:189-200 calling the scalar
// recursion src/operators_lazytensor.jl:652-685.  As in qob_kernels_qtile.cu the sum is regrouped by flip mask,
//     (H x)[i] = sum_c w_c(bits of i) * x[i XOR mask_c],
// and a pass applies every component whose mask fits the 12 free index bits of its 4096-amplitude tile.  What is new:
//
//  * ONE persistent, warp-specialised CTA per SM.  A producer thread walks an ordered tile queue and stages tiles with
//    cp.async.bulk.tensor (SASS UTMALDG): one instruction per 64 KiB tile, whatever its shape — the contiguous low block
//    and the window of high bits are box dimensions of a tensor map over the state.  y of a read-modify-write pass is
//    staged the same way, so neither x nor y costs the LSU any global-load wavefronts (ncu, round 1: the LSU data pipe
//    was the limiter at 88-90 %).
//  * Each of the 256 consumer threads owns 16 amplitudes = 4 index bits ("R bits") and keeps x and the accumulators in
//    registers.  A bond whose mask lies inside the R bits costs no shared-memory access at all; weights that depend on
//    R bits are hoisted into <= 4 registers per (component, tile); the others cost one LDS.128 gather per amplitude.
//  * Two passes whose free bits span <= 20 index bits are CHAINED: the tile queue runs chunk by chunk (a chunk = all
//    amplitudes that share the index bits outside the union, <= 16 MiB of x), pass-2 tiles of chunk c queued one chunk
//    behind pass-1 tiles and released by a per-chunk counter, so pass 2 finds x and y of its chunk in the 126 MB L2.
//    Heisenberg N=28: 3 tile passes but 80 B/amplitude of DRAM traffic instead of 128 (tools/ubench_l2.cu measures the
//    mechanism in isolation: 3.12 -> 2.12 ms for a 2-pass sweep of 2^28 amplitudes).
//
// Index bits: bit b of the flat (column-major) index <-> subsystem b+1 (src/states.jl:105).
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "qob_internal.h"

// Tile size: 2^T amplitudes, T = 12 (64 KiB tiles, one CTA per SM) or T = 11 (32 KiB tiles, TWO CTAs per SM: the same work per
// thread and per amplitude, but two independent pipelines per SM, so one CTA's register-only phases, barriers and waits are
// covered by the other's shared-memory traffic).  The kernel is a template on T; everything else derives from it:
//   consumer threads 2^(T-4) (16 amplitudes each), byte distance between a thread's amplitudes = 16 * 2^(T-4).
#define QR_TMAX 12
#define QR_NSTAGE 3                     // tile buffers: one being loaded, one being worked on, one being stored
#define QR_MAXC 40                      // lookup records per pass (kernel parameter space)
#define QR_MAXSEG 8
#define QR_DIAG_WINDOW 6

// classes of lookup records; records of a pass are sorted by class
enum { QR_PRE_TILE = 0, QR_PRE = 1, QR_DIAGR = 2, QR_INREG = 3, QR_GATHER = 4, QR_GATHER_SLOW = 5 };

// Tile-local index bits: 0..7 = consumer thread, 8..11 = the thread's 16 amplitudes ("R bits").  Amplitude u of thread t sits at
// byte t*16 + u*4096 of the staged tile, so every shared-memory access of the hot loops is [thread base + immediate].
struct QRComp {       // 32 bytes
  uint32_t kind;      // class | SELR << 4 | M << 8   (SELR: which of the 4 R bits select the weight; M: flip mask inside R)
  uint32_t tabE;      // first entry of the weight table inside the pass table
  uint32_t selT;      // selector bits among the 8 thread bits: bit-field runs of the thread id, 9 bits each: shift (6) | width (3)
  uint32_t selG;      // selector bits among the tile-id / rank bits: bit-field runs of the 64-bit flat index
  uint32_t xorB;      // gather: XOR applied to the thread's byte offset inside the tile (thread bits of the mask)
  uint32_t nT;        // number of selector bits among the thread bits (the tile part of the table index is shifted by it)
  // fast path, valid when bit 31 of `fast` is clear: the selector bits outside R are ONE run of thread bits (or none):
  //   entry = tabE + (((tid >> (fast & 31)) & ((fast >> 8) & 255)) << ((fast >> 16) & 31))
  uint32_t fast;
  uint32_t code;      // dense body number for the dispatch switch
};

#define QR_MAXPEER 16
struct QRPass {
  unsigned long long rstride[4];  // element stride in global memory of R bit k
  unsigned char tgbit[8];         // consumer-thread bit b -> flat-index bit
  int nfree_seg;                  // tile-local index -> element offset inside the tile (L2 prefetch of the next tile)
  unsigned char fs_l[QR_MAXSEG], fs_n[QR_MAXSEG], fs_g[QR_MAXSEG];
  int nfixed_seg;                 // compact tile id -> element offset of the tile
  unsigned char xs_l[QR_MAXSEG], xs_n[QR_MAXSEG], xs_g[QR_MAXSEG];
  int rank;                       // tensor-map rank; coordinate d = ((tile >> tshift) & (2^tbits - 1)) << boxlog
  unsigned char dim_tshift[5], dim_tbits[5], dim_boxlog[5];
  int n_pretile, n_pre, n_diagr, n_inreg, n_gather;  // records: [pre (tile) | pre (thread) | diagR | in-register | gather]
  int n_plain;                    // the first n_plain gather records have one weight per thread and no R bit in their mask
  unsigned long long hi_or;       // index bits above the local address (rank of a sharded state), already shifted
  uint32_t tab_smem_off;          // byte offset of this pass's tables inside the shared-memory table area
  uint32_t tab_bytes;
  const unsigned char *tab;       // device copy of the tables
  int mode;                       // 0: y = a*acc (TMA store); 2: y += a*acc (TMA reduce-add); 3: no output, sum conj(x)*acc (expect)
  int signal;                     // 1: count finished tiles per chunk (first pass of a chained pair)
  int wait;                       // 1: tiles wait for their chunk's counter (second pass of a chained pair)
  int stream_out;                 // 1: last use of x and y in this launch: L2 evict_first on loads and stores
  // chained launches: compact tile id = deposit(chunk id, cs_*) | deposit(tile within chunk, js_*)
  int ncs, njs;
  unsigned char cs_l[4], cs_n[4], cs_d[4], js_l[4], js_n[4], js_d[4];
  QRComp comps[QR_MAXC];
  // peer-addressed pass (the exchange of a sharded state): the pass works on the SWAPPED layout of the ranks' slabs.  Index
  // bits [peer_shift, peer_shift + peer_bits) of an element's address name the rank that holds it; there it sits at the same
  // address with those bits replaced by this rank's number.  The tile moves as 2^(T - piece_log2) contiguous pieces.
  // (kept behind the lookup records: the offsets the single-GPU launches read stay where the constant cache had them)
  int peer_bits, peer_shift, peer_rank, piece_log2;
};

struct QRLaunch {
  int npass;             // 1, or 2 chained passes
  unsigned tpc_log2;     // log2(tiles per chunk and pass)
  unsigned nchunks, lag; // chunks; pass 2 runs `lag` chunks behind pass 1
  unsigned total_items;
  int prefetch;          // (unused)
  int nstage;            // tile buffers in use: 3, or 2 when a kernel of another stream has to fit beside this one on every SM
  unsigned *queue;       // work counter
  unsigned *done;        // per chunk: finished pass-1 tiles
  int static_queue;      // 1: tiles are dealt round-robin (item = blockIdx.x + k*gridDim.x): reproducible partial sums
  double2 *partials;     // mode 3: one partial sum per consumer warp, [blockIdx.x * 8 + warp]
  long long *stats;      // debug (QOB_QREG_STATS=1): per-CTA cycle counts of the pipeline phases, 16 per CTA
  double2 alpha, beta;
  QRPass pass[2];
  int remap;             // single pass over one tile range: compact tile id = deposit(chunk_fixed, cs_*) | deposit(item, js_*)
  unsigned chunk_fixed;
  const double2 *xpeer[QR_MAXPEER];   // peer-addressed pass: every rank's slab of x and of the buffer that receives the results
  double2 *ypeer[QR_MAXPEER];
};

// ------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ unsigned qr_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void qr_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void qr_mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void qr_mbar_expect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void qr_mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    // the hint lets the hardware park the warp until the phase completes instead of spinning through the issue slots
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(bar), "r"(parity), "r"(0x989680u)
                 : "memory");
  }
}
__device__ __forceinline__ void qr_tma_load(unsigned dst, const CUtensorMap *map, unsigned bar, int rank, const int *c,
                                            unsigned long long pol) {
  const unsigned long long m = (unsigned long long)map;
  switch (rank) {
    case 1:
      asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3}], [%2], %4;\n" ::"r"(dst),
                   "l"(m), "r"(bar), "r"(c[0]), "l"(pol)
                   : "memory");
      break;
    case 2:
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(dst),
                   "l"(m), "r"(bar), "r"(c[0]), "r"(c[1]), "l"(pol)
                   : "memory");
      break;
    case 3:
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;\n" ::"r"(dst),
                   "l"(m), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "l"(pol)
                   : "memory");
      break;
    case 4:
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;\n" ::"r"(dst),
                   "l"(m), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "l"(pol)
                   : "memory");
      break;
    default:
      asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;\n" ::"r"(dst),
                   "l"(m), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "l"(pol)
                   : "memory");
      break;
  }
}
// result tile: shared memory -> global through the tensor map, either a plain store (first pass, beta == 0: y is never read)
// or an f64 add performed by the L2 (every other pass: y += tile) — no y load, no per-thread global stores
__device__ __forceinline__ void qr_tma_store(const CUtensorMap *map, unsigned src, int rank, const int *c, bool add) {
  const unsigned long long m = (unsigned long long)map;
#define QR_TS(N, COORDS, ...)                                                                                                          \
  if (add)                                                                                                                             \
    asm volatile("cp.reduce.async.bulk.tensor." N ".global.shared::cta.add.tile.bulk_group [%0, " COORDS "], [%1];\n" ::"l"(m), "r"(src), \
                 __VA_ARGS__                                                                                                           \
                 : "memory");                                                                                                          \
  else                                                                                                                                 \
    asm volatile("cp.async.bulk.tensor." N ".global.shared::cta.tile.bulk_group [%0, " COORDS "], [%1];\n" ::"l"(m), "r"(src), __VA_ARGS__ \
                 : "memory");
  switch (rank) {
    case 1: QR_TS("1d", "{%2}", "r"(c[0])) break;
    case 2: QR_TS("2d", "{%2, %3}", "r"(c[0]), "r"(c[1])) break;
    case 3: QR_TS("3d", "{%2, %3, %4}", "r"(c[0]), "r"(c[1]), "r"(c[2])) break;
    case 4: QR_TS("4d", "{%2, %3, %4, %5}", "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3])) break;
    default: QR_TS("5d", "{%2, %3, %4, %5, %6}", "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])) break;
  }
#undef QR_TS
}

// peer-addressed passes: one contiguous piece of a tile, global (any rank's slab, over NVLink) <-> shared memory
__device__ __forceinline__ void qr_bulk_load(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
__device__ __forceinline__ void qr_bulk_store(void *dst, unsigned src, unsigned bytes, bool add) {
  if (add)
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;\n" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
  else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
// element `a` of the swapped layout -> (owning rank, element offset in that rank's slab)
__device__ __forceinline__ unsigned qr_peer_of(const QRPass &P, unsigned long long a, unsigned long long &phys) {
  const unsigned long long m = ((1ull << P.peer_bits) - 1ull) << P.peer_shift;
  phys = (a & ~m) | ((unsigned long long)P.peer_rank << P.peer_shift);
  return (unsigned)((a & m) >> P.peer_shift);
}

// pull a tile into L2 without occupying shared memory (the DRAM latency is paid here, several tiles ahead)
__device__ __forceinline__ void qr_tma_prefetch(const CUtensorMap *map, int rank, const int *c) {
  const unsigned long long m = (unsigned long long)map;
  switch (rank) {
    case 1: asm volatile("cp.async.bulk.prefetch.tensor.1d.L2.global.tile [%0, {%1}];\n" ::"l"(m), "r"(c[0]) : "memory"); break;
    case 2: asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n" ::"l"(m), "r"(c[0]), "r"(c[1]) : "memory"); break;
    case 3:
      asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];\n" ::"l"(m), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory");
      break;
    case 4:
      asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];\n" ::"l"(m), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3])
                   : "memory");
      break;
    default:
      asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];\n" ::"l"(m), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]),
                   "r"(c[4])
                   : "memory");
      break;
  }
}

__host__ __device__ constexpr int qr_popc(int v) { return (v & 1) + ((v >> 1) & 1) + ((v >> 2) & 1) + ((v >> 3) & 1); }
__host__ __device__ constexpr int qr_pext(int u, int sel) {
  int r = 0, k = 0;
  for (int b = 0; b < 4; ++b)
    if ((sel >> b) & 1) {
      r |= ((u >> b) & 1) << k;
      ++k;
    }
  return r;
}

__device__ __forceinline__ unsigned long long qr_expand(unsigned v, int nseg, const unsigned char *sl, const unsigned char *sn,
                                                        const unsigned char *sg) {
  unsigned long long a = 0;
#pragma unroll
  for (int s = 0; s < QR_MAXSEG; ++s)
    if (s < nseg) a |= (unsigned long long)((v >> sl[s]) & ((1u << sn[s]) - 1u)) << sg[s];
  return a;
}
__device__ __forceinline__ unsigned qr_deposit(unsigned v, int nseg, const unsigned char *sl, const unsigned char *sn,
                                               const unsigned char *sd) {
  unsigned a = 0;
#pragma unroll
  for (int s = 0; s < 4; ++s)
    if (s < nseg) a |= ((v >> sl[s]) & ((1u << sn[s]) - 1u)) << sd[s];
  return a;
}

// table index (in entries) of selector bits given as up to three bit-field runs of a 64-bit (32-bit) value
__device__ __forceinline__ unsigned qr_run(unsigned sel, int r, unsigned lo, unsigned hi, unsigned pos) {
  const unsigned sh = (sel >> (9 * r)) & 63u, w = (sel >> (9 * r + 6)) & 7u;
  const unsigned v = sh < 32 ? __funnelshift_r(lo, hi, sh) : (hi >> (sh & 31));
  return (v & ((1u << w) - 1u)) << pos;
}
__device__ __forceinline__ unsigned qr_field(unsigned sel, unsigned lo, unsigned hi) {
  unsigned idx = qr_run(sel, 0, lo, hi, 0);
  if (sel >> 9) {
    const unsigned w0 = (sel >> 6) & 7u, w1 = (sel >> 15) & 7u;
    idx |= qr_run(sel, 1, lo, hi, w0);
    if (sel >> 18) idx |= qr_run(sel, 2, lo, hi, w0 + w1);
  }
  return idx;
}
__device__ __forceinline__ unsigned qr_field32(unsigned sel, unsigned v) {
  unsigned idx = (v >> (sel & 63u)) & ((1u << ((sel >> 6) & 7u)) - 1u);
  if (sel >> 9) {
    const unsigned w0 = (sel >> 6) & 7u, w1 = (sel >> 15) & 7u;
    idx |= ((v >> ((sel >> 9) & 63u)) & ((1u << w1) - 1u)) << w0;
    if (sel >> 18) idx |= ((v >> ((sel >> 18) & 63u)) & ((1u << ((sel >> 24) & 7u)) - 1u)) << (w0 + w1);
  }
  return idx;
}
// general index of the selector bits outside R: thread part | tile part << nT (kept out of line: rare)
__device__ __noinline__ unsigned qr_index_slow(unsigned selT, unsigned selG, unsigned nT, unsigned tid, unsigned g_lo, unsigned g_hi) {
  unsigned idx = 0;
  if (selT) idx = qr_field32(selT, tid);
  if (selG) idx |= qr_field(selG, g_lo, g_hi) << nT;
  return idx;
}
// first table entry this thread uses: `shift` = number of selector bits inside R (the low index bits of the table).
// hot = {kind | code << 16, tabE, fast, xorB} from shared memory; the cold fields are only touched on the slow path.
__device__ __forceinline__ unsigned qr_entry(const uint4 hot, const QRComp &cold, unsigned shift, unsigned tid, unsigned g_lo, unsigned g_hi) {
  const unsigned f = hot.z;
  if (!(f >> 31)) return hot.y + (((tid >> (f & 31u)) & ((f >> 8) & 255u)) << ((f >> 16) & 31u));
  return hot.y + (qr_index_slow(cold.selT, cold.selG, cold.nT, tid, g_lo, g_hi) << shift);
}

// weights: REALW tables hold 8-byte reals (2 DFMA per product), otherwise (re, im) pairs (4 DFMA)
template <bool REALW>
struct QRW {
  double re, im;
  __device__ __forceinline__ void load(const unsigned char *p) {
    if (REALW) {
      re = *reinterpret_cast<const double *>(p);
      im = 0.0;
    } else {
      const double2 t = *reinterpret_cast<const double2 *>(p);
      re = t.x;
      im = t.y;
    }
  }
  __device__ __forceinline__ void fma_into(double2 &acc, const double2 v) const {
    if (REALW) {
      acc.x = fma(re, v.x, acc.x);
      acc.y = fma(re, v.y, acc.y);
    } else {
      acc.x = fma(re, v.x, acc.x);
      acc.x = fma(-im, v.y, acc.x);
      acc.y = fma(re, v.y, acc.y);
      acc.y = fma(im, v.x, acc.y);
    }
  }
};

// bond inside the thread's 16 amplitudes: no shared-memory access at all
template <bool REALW, int SELR, int M>
__device__ __forceinline__ void qr_inreg(double2 (&acc)[16], const double2 (&xr)[16], const unsigned char *wt) {
  constexpr int NW = 1 << qr_popc(SELR);
  QRW<REALW> w[NW];
#pragma unroll
  for (int v = 0; v < NW; ++v) w[v].load(wt + v * (REALW ? 8 : 16));
#pragma unroll
  for (int u = 0; u < 16; ++u) w[qr_pext(u, SELR)].fma_into(acc[u], xr[u ^ M]);
}
// bond that leaves the thread: one LDS.128 per amplitude, weights hoisted.  The partner of amplitude u sits at
// [xt + (u ^ M)*4096]: an immediate offset when the mask has no R bit (MR == false), one XOR per amplitude otherwise.
template <bool REALW, int SELR, bool MR, unsigned US>
__device__ __forceinline__ void qr_gather(double2 (&acc)[16], const unsigned char *xt, const unsigned char *wt, unsigned mxor) {
  constexpr int NW = 1 << qr_popc(SELR);
  QRW<REALW> w[NW];
#pragma unroll
  for (int v = 0; v < NW; ++v) w[v].load(wt + v * (REALW ? 8 : 16));
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      v[u] = *reinterpret_cast<const double2 *>(xt + (MR ? (((8 * h + u) * US) ^ mxor) : (8 * h + u) * US));
#pragma unroll
    for (int u = 0; u < 8; ++u) w[qr_pext(8 * h + u, SELR)].fma_into(acc[8 * h + u], v[u]);
  }
}
// any selector pattern: weight looked up per amplitude
template <bool REALW, unsigned US>
__device__ __forceinline__ void qr_gather_slow(double2 (&acc)[16], const unsigned char *xt, const unsigned char *wt, unsigned selr,
                                               unsigned mxor) {
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    unsigned r = 0, k = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if ((selr >> b) & 1u) {
        r |= ((u >> b) & 1u) << k;
        ++k;
      }
    QRW<REALW> w;
    w.load(wt + r * (REALW ? 8 : 16));
    const double2 v = *reinterpret_cast<const double2 *>(xt + ((u * US) ^ mxor));
    w.fma_into(acc[u], v);
  }
}

struct QRItem {     // 64 bytes, written by the producer for the consumers
  int pass;         // -1: no more work
  unsigned tile;    // compact tile id
  unsigned chunk;
  unsigned pad;
  double dre, dim;  // diagonal weight that is the same for every amplitude of the tile
  int co[5];        // tensor-map coordinates of the tile (the result goes out through the same box)
  int pad2;
  unsigned long long tbase;   // element offset of the tile (peer-addressed passes move the tile piece by piece)
};

// work item -> (pass, chunk, tile within chunk); false when the queue is exhausted
__device__ __forceinline__ bool qr_decode(const QRLaunch &L, unsigned item, int &p, unsigned &c, unsigned &j) {
  if (item >= L.total_items) return false;
  if (L.npass == 1) {
    p = 0;
    c = 0;
    j = item;
    return true;
  }
  const unsigned b = item >> L.tpc_log2, C = L.nchunks, lag = L.lag;
  j = item & ((1u << L.tpc_log2) - 1u);
  if (C <= lag) {
    p = b < C ? 0 : 1;
    c = b < C ? b : b - C;
  } else if (b < lag) {
    p = 0;
    c = b;
  } else {
    const unsigned b1 = b - lag, mid = 2u * (C - lag);
    if (b1 < mid) {
      p = (int)(b1 & 1u);
      c = p ? (b1 >> 1) : lag + (b1 >> 1);
    } else {
      p = 1;
      c = C - lag + (b1 - mid);
    }
  }
  return true;
}

// PEER: the peer-addressed variant (exchange of a sharded state); the single-GPU instantiations carry none of its code
template <bool REALW, int T, bool PEER = false>
__global__ void __launch_bounds__((1 << (T - 4)) + 64, T == 12 ? 1 : 2)
    qreg_kernel(const __grid_constant__ QRLaunch L, const __grid_constant__ CUtensorMap mx0, const __grid_constant__ CUtensorMap my0,
                const __grid_constant__ CUtensorMap mx1, const __grid_constant__ CUtensorMap my1) {
  constexpr unsigned QR_CTHREADS = 1u << (T - 4);        // consumer threads (8 or 4 warps), 16 amplitudes each
  constexpr unsigned QR_THREADS = QR_CTHREADS + 64;      // + one loader warp and one storer warp
  constexpr unsigned QR_TILE_BYTES = 16u << T;
  constexpr unsigned QR_USTRIDE = QR_CTHREADS * 16u;     // byte distance in the staged tile between consecutive amplitudes u of a thread
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((128u - (qr_smem(smem_raw) & 127u)) & 127u);  // TMA destinations: 128-byte aligned
  unsigned char *xs0 = smem;                              // QR_NSTAGE tile buffers: x comes in, the result tile leaves from the same one
  const unsigned nst = (unsigned)L.nstage;
  unsigned char *tabs = smem + nst * QR_TILE_BYTES;       // weight tables of the pass(es)
  const unsigned tab_total = L.pass[0].tab_bytes + (L.npass > 1 ? L.pass[1].tab_bytes : 0u);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(tabs + ((tab_total + 15u) & ~15u));
  // bars[s]: x of stage s has landed; [3+s]: stage s is free; [6+s]: the result tile of stage s is staged
  QRItem *slots = reinterpret_cast<QRItem *>(bars + 12);
  // hot half of every lookup record {kind | code << 16, tabE, fast, xorB}: read with one broadcast LDS.128 instead of
  // indexed constant-bank loads (7 KB of parameters do not stay in the immediate-constant cache)
  uint4 *recs = reinterpret_cast<uint4 *>(slots + QR_NSTAGE);
  const unsigned tid = threadIdx.x;
  constexpr int WB = REALW ? 8 : 16;

  if (tid == 0) {
    for (int st = 0; st < QR_NSTAGE; ++st) {
      qr_mbar_init(qr_smem(bars + st), 1);
      qr_mbar_init(qr_smem(bars + 3 + st), 1);
      qr_mbar_init(qr_smem(bars + 6 + st), QR_CTHREADS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
  }
  for (int p = 0; p < L.npass; ++p) {
    const int4 *g4 = reinterpret_cast<const int4 *>(L.pass[p].tab);
    int4 *t4 = reinterpret_cast<int4 *>(tabs + L.pass[p].tab_smem_off);
    for (unsigned i = tid; i < L.pass[p].tab_bytes / 16u; i += QR_THREADS) t4[i] = g4[i];
    for (unsigned i = tid; i < QR_MAXC; i += QR_THREADS) {
      const QRComp &cd = L.pass[p].comps[i];
      recs[p * QR_MAXC + i] = make_uint4(cd.kind | (cd.code << 16), cd.tabE, cd.fast, cd.xorB);
    }
  }
  __syncthreads();

  if (tid >= QR_CTHREADS + 32) {
    // ------------------------------------------------------------------ storer (one thread)
    // Sends every staged result tile to global memory with one TMA instruction — a plain store for the pass that defines y,
    // an f64 add performed by the L2 for every other pass — frees the staging buffer as soon as the copy engine has read it,
    // and announces finished pass-1 tiles of a chained pair once their writes are complete.
    if (tid == QR_CTHREADS + 32) {
      unsigned stage = 0, ophase = 0;   // bit s of ophase: parity of the next wait on staged[s]
      while (true) {
        qr_mbar_wait(qr_smem(bars + 6 + stage), (ophase >> stage) & 1u);
        ophase ^= 1u << stage;
        const QRItem it = slots[stage];
        if (it.pass < 0) break;
        const QRPass &P = L.pass[it.pass];
        if (P.mode != 3) {
          if (PEER) {
            const unsigned pbytes = 16u << P.piece_log2;
            for (unsigned jp = 0; jp < (1u << (T - P.piece_log2)); ++jp) {
              unsigned long long phys;
              const unsigned q = qr_peer_of(P, it.tbase | qr_expand(jp << P.piece_log2, P.nfree_seg, P.fs_l, P.fs_n, P.fs_g), phys);
              qr_bulk_store(L.ypeer[q] + phys, qr_smem(xs0 + stage * QR_TILE_BYTES + jp * pbytes), pbytes, P.mode != 0);
            }
          } else {
            qr_tma_store(it.pass ? &my1 : &my0, qr_smem(xs0 + stage * QR_TILE_BYTES), P.rank, it.co, P.mode != 0);
          }
          asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
        }
        qr_mbar_arrive(qr_smem(bars + 3 + stage));   // the buffer may be loaded again
        if (P.signal) {
          asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
          __threadfence();
          atomicAdd(L.done + it.chunk, 1u);
        }
        stage = stage + 1 == nst ? 0 : stage + 1;
      }
      asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    }
    return;
  }
  if (tid >= QR_CTHREADS) {
    // ------------------------------------------------------------------ producer warp
    // Lane 0 claims work items from the ordered queue (the next one while the current loads are in flight, so the latency
    // of the atomic is hidden) and issues the TMA loads; all lanes evaluate the diagonal tables whose selector bits are
    // tile-id bits (one table per lane).
    const unsigned lane = tid & 31u;
    unsigned long long pol_keep, pol_stream;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    unsigned stage = 0, xphase = 0;  // bit s of xphase: parity of the next wait on x-empty[s]
    long long st_ex = 0, st_dep = 0, st_ey = 0, st_n = 0;
    const long long st_t0 = clock64();
    unsigned item = 0;
    if (lane == 0) item = L.static_queue ? blockIdx.x : atomicAdd(L.queue, 1u);
    while (true) {
      item = __shfl_sync(0xffffffffu, item, 0);
      int p = 0;
      unsigned t = 0, c = 0, j = 0;
      const bool more = qr_decode(L, item, p, c, j);   // every lane decodes
      if (lane == 0) {
        const long long q0 = clock64();
        qr_mbar_wait(qr_smem(bars + 3 + stage), ((xphase >> stage) & 1u) ^ 1u);
        st_ex += clock64() - q0;
        xphase ^= 1u << stage;
      }
      if (!more) {
        if (lane == 0) {
          QRItem it;
          it.pass = -1;
          it.tile = it.chunk = it.pad = 0;
          it.dre = it.dim = 0.0;
          for (int d = 0; d < 5; ++d) it.co[d] = 0;
          slots[stage] = it;
          qr_mbar_arrive(qr_smem(bars + stage));
          if (L.stats) {
            long long *o = L.stats + 16 * blockIdx.x;
            o[0] = st_ex, o[1] = st_dep, o[2] = st_ey, o[3] = clock64() - st_t0, o[4] = st_n;
          }
        }
        break;
      }
      ++st_n;
      const QRPass &P = L.pass[p];
      t = (L.npass == 1 && !L.remap) ? j
                                     : (qr_deposit(L.remap ? L.chunk_fixed : c, P.ncs, P.cs_l, P.cs_n, P.cs_d) |
                                        qr_deposit(j, P.njs, P.js_l, P.js_n, P.js_d));
      // tile-constant diagonal weight: lane k evaluates table k
      double dre = 0.0, dim = 0.0;
      if (P.n_pretile > 0) {
        const unsigned long long gidx = qr_expand(t, P.nfixed_seg, P.xs_l, P.xs_n, P.xs_g) | P.hi_or;
        const unsigned char *tb = tabs + P.tab_smem_off;
        for (int k = (int)lane; k < P.n_pretile; k += 32) {
          const QRComp &cd = P.comps[k];
          QRW<REALW> w;
          w.load(tb + (size_t)(cd.tabE + qr_field(cd.selG, (unsigned)gidx, (unsigned)(gidx >> 32))) * WB);
          dre += w.re;
          dim += w.im;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          dre += __shfl_xor_sync(0xffffffffu, dre, o);
          dim += __shfl_xor_sync(0xffffffffu, dim, o);
        }
      }
      if (lane == 0) {
        int co[5];
#pragma unroll
        for (int d = 0; d < 5; ++d) co[d] = (int)(((t >> P.dim_tshift[d]) & ((1u << P.dim_tbits[d]) - 1u)) << P.dim_boxlog[d]);
        QRItem it;
        it.pass = p;
        it.tile = t;
        it.chunk = c;
        it.pad = 0;
        it.dre = dre;
        it.dim = dim;
#pragma unroll
        for (int d = 0; d < 5; ++d) it.co[d] = co[d];
        it.pad2 = 0;
        it.tbase = PEER ? qr_expand(t, P.nfixed_seg, P.xs_l, P.xs_n, P.xs_g) : 0ull;
        slots[stage] = it;
        if (P.wait) {
          const unsigned need = 1u << L.tpc_log2;   // every pass-1 tile of the chunk
          unsigned have;
          const long long q0 = clock64();
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(have) : "l"(L.done + c) : "memory");
            if (have < need) __nanosleep(64);
          } while (have < need);
          st_dep += clock64() - q0;
          asm volatile("fence.proxy.async;\n" ::: "memory");  // the tiles written by other CTAs are read by the async proxy
        }
        const unsigned long long pol = P.stream_out ? pol_stream : pol_keep;
        qr_mbar_expect(qr_smem(bars + stage), QR_TILE_BYTES);
        if (!PEER) qr_tma_load(qr_smem(xs0 + stage * QR_TILE_BYTES), p ? &mx1 : &mx0, qr_smem(bars + stage), P.rank, co, pol);
        item = L.static_queue ? item + gridDim.x : atomicAdd(L.queue, 1u);   // the next item: the atomic's latency overlaps the consumers' work
      }
      __syncwarp();
      if (PEER) {
        // every lane fetches its pieces of the tile from the ranks that hold them (NVLink loads straight into shared memory)
        const unsigned long long tb = qr_expand(t, P.nfixed_seg, P.xs_l, P.xs_n, P.xs_g);
        const unsigned pbytes = 16u << P.piece_log2;
        for (unsigned jp = lane; jp < (1u << (T - P.piece_log2)); jp += 32) {
          unsigned long long phys;
          const unsigned q = qr_peer_of(P, tb | qr_expand(jp << P.piece_log2, P.nfree_seg, P.fs_l, P.fs_n, P.fs_g), phys);
          qr_bulk_load(qr_smem(xs0 + stage * QR_TILE_BYTES + jp * pbytes), L.xpeer[q] + phys, pbytes, qr_smem(bars + stage));
        }
      }
      stage = stage + 1 == nst ? 0 : stage + 1;
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const unsigned so = tid * 16u;
  unsigned stage = 0, xphase = 0;
  double2 esum = make_double2(0.0, 0.0);   // mode 3: this thread's share of <x| op x>
  long long sc_wx = 0, sc_cmp = 0, sc_wy = 0, sc_epi = 0;
#pragma unroll 1
  while (true) {
    long long q0 = clock64();
    qr_mbar_wait(qr_smem(bars + stage), (xphase >> stage) & 1u);
    long long q1 = clock64();
    sc_wx += q1 - q0;
    xphase ^= 1u << stage;
    const QRItem it = slots[stage];
    if (it.pass < 0) break;
    const QRPass &P = L.pass[it.pass];
    const unsigned char *xs = xs0 + stage * QR_TILE_BYTES;
    const unsigned char *tb = tabs + P.tab_smem_off;
    const uint4 *hr = recs + it.pass * QR_MAXC;
    const unsigned long long tbase = qr_expand(it.tile, P.nfixed_seg, P.xs_l, P.xs_n, P.xs_g);
    const unsigned long long gidx = tbase | P.hi_or;   // the thread bits enter the lookups through `tid` (selT)
    const unsigned g_lo = (unsigned)gidx, g_hi = (unsigned)(gidx >> 32);

    double2 acc[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) acc[u] = make_double2(0.0, 0.0);
    {
      {
        // ---- own amplitudes in registers: diagonal weights and the bonds inside the R bits
        double2 xr[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) xr[u] = *reinterpret_cast<const double2 *>(xs + so + u * QR_USTRIDE);
        int c = P.n_pretile;
        const int c_diag_end = P.n_pretile + P.n_pre + P.n_diagr;
        if (c_diag_end > 0) {
          // diagonal weight: tile constant (from the producer) + tables on thread bits + tables indexed by the R bits
          double dre = it.dre, dim = it.dim;
          for (; c < P.n_pretile + P.n_pre; ++c) {
            QRW<REALW> w;
            w.load(tb + qr_entry(hr[c], P.comps[c], 0u, tid, g_lo, g_hi) * WB);
            dre += w.re;
            dim += w.im;
          }
          if (REALW) {
            double d[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) d[u] = dre;
            for (; c < c_diag_end; ++c) {
              const unsigned char *wt = tb + qr_entry(hr[c], P.comps[c], 4u, tid, g_lo, g_hi) * WB;
#pragma unroll
              for (int u = 0; u < 16; ++u) d[u] += *reinterpret_cast<const double *>(wt + u * 8);
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) acc[u] = make_double2(d[u] * xr[u].x, d[u] * xr[u].y);
          } else {
            QRW<REALW> d[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              d[u].re = dre;
              d[u].im = dim;
            }
            for (; c < c_diag_end; ++c) {
              const unsigned char *wt = tb + qr_entry(hr[c], P.comps[c], 4u, tid, g_lo, g_hi) * WB;
#pragma unroll
              for (int u = 0; u < 16; ++u) {
                QRW<REALW> w;
                w.load(wt + u * WB);
                d[u].re += w.re;
                d[u].im += w.im;
              }
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) d[u].fma_into(acc[u], xr[u]);
          }
        }
        const int c_in_end = c_diag_end + P.n_inreg;
        for (c = c_diag_end; c < c_in_end; ++c) {
          const uint4 cd = hr[c];
          const unsigned char *wt = tb + qr_entry(cd, P.comps[c], __popc((cd.x >> 4) & 15u), tid, g_lo, g_hi) * WB;
          switch (cd.x >> 16) {  // SELR == M
#define QR_IN(CODE, MM) \
  case CODE: qr_inreg<REALW, MM, MM>(acc, xr, wt); break;
            QR_IN(0, 1) QR_IN(1, 2) QR_IN(2, 4) QR_IN(3, 8) QR_IN(4, 3) QR_IN(5, 5) QR_IN(6, 6) QR_IN(7, 9) QR_IN(8, 10) QR_IN(9, 12)
#undef QR_IN
            default: break;
          }
        }
      }
      {
        const int c0 = P.n_pretile + P.n_pre + P.n_diagr + P.n_inreg, c1 = c0 + P.n_gather;
        int c = c0;
        // ---- plain gathers first (one weight per thread, mask without R bits: the bulk of a pass).  No dispatch, and the
        // loads of the next half-record are in flight while the DFMAs of the current one run.
        const int cp = c0 + P.n_plain;
        if (c < cp) {
          double2 va[8], vb[8];
          QRW<REALW> w;
          const unsigned char *xt;
          {
            const uint4 cd = hr[c];
            w.load(tb + qr_entry(cd, P.comps[c], 0u, tid, g_lo, g_hi) * WB);
            xt = xs + (so ^ cd.w);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) va[u] = *reinterpret_cast<const double2 *>(xt + u * QR_USTRIDE);
#pragma unroll 1
          for (; c < cp; ++c) {
#pragma unroll
            for (int u = 0; u < 8; ++u) vb[u] = *reinterpret_cast<const double2 *>(xt + (8 + u) * QR_USTRIDE);
            QRW<REALW> wn = w;
            const unsigned char *xn = xt;
            const bool more = c + 1 < cp;
            if (more) {
              const uint4 nd = hr[c + 1];
              wn.load(tb + qr_entry(nd, P.comps[c + 1], 0u, tid, g_lo, g_hi) * WB);
              xn = xs + (so ^ nd.w);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) w.fma_into(acc[u], va[u]);
            if (more) {
#pragma unroll
              for (int u = 0; u < 8; ++u) va[u] = *reinterpret_cast<const double2 *>(xn + u * QR_USTRIDE);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) w.fma_into(acc[8 + u], vb[u]);
            w = wn;
            xt = xn;
          }
        }
        for (; c < c1; ++c) {
          const uint4 cd = hr[c];
          const unsigned selr = (cd.x >> 4) & 15u;
          const unsigned char *wt = tb + qr_entry(cd, P.comps[c], __popc(selr), tid, g_lo, g_hi) * WB;
          const unsigned char *xt = xs + (so ^ cd.w);          // xorB: thread bits of the mask
          const unsigned mxor = ((cd.x >> 8) & 15u) * QR_USTRIDE;  // R bits of the mask
          switch (cd.x >> 16) {  // 0..10: mask without R bits; 11..21: with R bits; 22: any selector pattern
#define QR_GA(CODE, SS)                                              \
  case CODE: qr_gather<REALW, SS, false, QR_USTRIDE>(acc, xt, wt, 0u); break;    \
  case CODE + 11: qr_gather<REALW, SS, true, QR_USTRIDE>(acc, xt, wt, mxor); break;
            QR_GA(0, 0) QR_GA(1, 1) QR_GA(2, 2) QR_GA(3, 4) QR_GA(4, 8) QR_GA(5, 3) QR_GA(6, 5) QR_GA(7, 6) QR_GA(8, 9) QR_GA(9, 10) QR_GA(10, 12)
#undef QR_GA
            default: qr_gather_slow<REALW, QR_USTRIDE>(acc, xt, wt, selr, mxor); break;
          }
        }
      }
    }
    if (P.mode == 3) {
      // expect(op, psi): nothing is written — every pass contributes sum_i conj(x_i) * acc_i, and the passes add up
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const double2 xo = *reinterpret_cast<const double2 *>(xs + so + u * QR_USTRIDE);
        esum.x = fma(xo.x, acc[u].x, fma(xo.y, acc[u].y, esum.x));
        esum.y = fma(xo.x, acc[u].y, fma(-xo.y, acc[u].x, esum.y));
      }
      __syncwarp();
      if ((tid & 31u) == 0) qr_mbar_arrive(qr_smem(bars + 6 + stage));
      stage = stage + 1 == nst ? 0 : stage + 1;
      continue;
    }
    // all reads of this tile are done by every consumer warp: its buffer now takes the result tile, alpha * acc, which the
    // storer thread sends off through TMA; the buffer returns to the loader when the copy engine has read it
    asm volatile("bar.sync 1, %0;\n" ::"n"(1 << (T - 4)) : "memory");
    q0 = clock64();
    sc_cmp += q0 - q1;
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const double2 o = make_double2(L.alpha.x * acc[u].x - L.alpha.y * acc[u].y, L.alpha.x * acc[u].y + L.alpha.y * acc[u].x);
      *reinterpret_cast<double2 *>(const_cast<unsigned char *>(xs) + so + u * QR_USTRIDE) = o;
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> visible to the bulk copy engine
    __syncwarp();
    if ((tid & 31u) == 0) qr_mbar_arrive(qr_smem(bars + 6 + stage));
    sc_epi += clock64() - q0;
    stage = stage + 1 == nst ? 0 : stage + 1;
  }
  if (L.partials) {  // fixed-order reduction: lanes of a warp, then one slot per warp; the host adds the slots in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      esum.x += __shfl_xor_sync(0xffffffffu, esum.x, o);
      esum.y += __shfl_xor_sync(0xffffffffu, esum.y, o);
    }
    if ((tid & 31u) == 0) L.partials[blockIdx.x * (QR_CTHREADS / 32) + (tid >> 5)] = esum;
  }
  // the loader's end marker sits in slots[stage]: pass it on to the storer
  __syncwarp();
  if ((tid & 31u) == 0) qr_mbar_arrive(qr_smem(bars + 6 + stage));
  if (L.stats && (tid & 31u) == 0) {
    long long *o = L.stats + 16 * blockIdx.x + 8;
    if (tid == 0) o[0] = sc_wx, o[1] = sc_cmp, o[2] = sc_wy, o[3] = sc_epi;
    if (tid == QR_CTHREADS - 32) o[4] = sc_wx, o[5] = sc_cmp, o[6] = sc_wy, o[7] = sc_epi;
  }
}

// ------------------------------------------------------------------------------------------ host
struct QRCompHost {
  uint64_t mask = 0;
  std::vector<int> sel;  // selector bit positions (flat index), ascending
  struct Contrib {
    int coef_index;
    cplx scalar;
    std::vector<cplx> unit;  // 2^nsel
  };
  std::vector<Contrib> contribs;
  int pass = -1;
};
struct QRTableHost {
  std::vector<int> bits;  // flat-index bit behind each bit of the table index (lowest first)
  struct Member {
    int comp;
    std::vector<int> pos;  // for each selector bit of the component: its position inside `bits`
  };
  std::vector<Member> members;
  uint32_t tab_off = 0;  // entries
};
struct QRPassHost {
  std::vector<int> free_bits;  // ascending, T of them
  std::vector<int> rpos;       // the 4 tile-local positions that form R (ascending)
  std::vector<QRTableHost> tables;
  QRPass params;
  std::vector<cplx> h_tab;
  DevArray<unsigned char> d_tab;
  size_t tab_entries = 0;
  // tensor map geometry (elements are doubles)
  int rank = 0;
  uint64_t gdim[5], gstride[5];
  uint32_t box[5];
  // cached tensor maps
  const void *map_x_ptr = nullptr, *map_y_ptr = nullptr;
  alignas(64) CUtensorMap map_x, map_y;
  int n_work = 0;
};
struct QRegProgramHost {
  int nbits = 0;
  int T = 12;                 // log2 of the tile size (11: two CTAs per SM, 12: one)
  uint64_t hi_value = 0;
  std::vector<QRCompHost> comps;
  std::vector<std::unique_ptr<QRPassHost>> passes;   // execution order
  struct Group {
    int first, count;          // passes [first, first+count) run in one launch
    unsigned tpc_log2 = 0, nchunks = 1, lag = 0;
  };
  std::vector<Group> groups;
  bool real_tables = false;
  std::mutex mu;
  std::map<cudaStream_t, unsigned *> sync_bufs;      // per stream: work counter + per-chunk counters
  std::map<cudaStream_t, double2 *> part_bufs;       // per stream: partial sums of the fused expect
  size_t sync_words = 0;
  ~QRegProgramHost() {
    for (auto &kv : sync_bufs)
      if (kv.second) cudaFree(kv.second);
    for (auto &kv : part_bufs)
      if (kv.second) cudaFree(kv.second);
  }
};

static int qr_env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

static void qr_segments(const std::vector<int> &bits, unsigned char *sl, unsigned char *sn, unsigned char *sg, int &nseg) {
  nseg = 0;
  size_t i = 0;
  while (i < bits.size()) {
    size_t j = i;
    while (j + 1 < bits.size() && bits[j + 1] == bits[j] + 1) ++j;
    if (nseg < QR_MAXSEG) {
      sl[nseg] = (unsigned char)i;
      sn[nseg] = (unsigned char)(j - i + 1);
      sg[nseg] = (unsigned char)bits[i];
    }
    ++nseg;
    i = j + 1;
  }
}
static bool qr_pack_runs(const std::vector<int> &bits, uint32_t &sel) {
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

static cplx qr_comp_value(const QRCompHost &c, const std::vector<cplx> &coefs, size_t r) {
  cplx w = 0.0;
  for (const auto &ct : c.contribs) {
    cplx f = ct.scalar;
    if (ct.coef_index >= 0) f *= coefs[ct.coef_index];
    w += f * ct.unit[r];
  }
  return w;
}

static void qr_fill_tables(QRegProgramHost &h, const std::vector<cplx> &coefs) {
  bool real = true;
  for (auto &pp : h.passes) {
    QRPassHost &p = *pp;
    std::fill(p.h_tab.begin(), p.h_tab.end(), cplx(0.0, 0.0));
    for (const QRTableHost &t : p.tables) {
      const size_t n = (size_t)1 << t.bits.size();
      for (const auto &m : t.members) {
        const QRCompHost &c = h.comps[m.comp];
        std::vector<cplx> cv((size_t)1 << c.sel.size());
        for (size_t r = 0; r < cv.size(); ++r) cv[r] = qr_comp_value(c, coefs, r);
        for (size_t r = 0; r < n; ++r) {
          size_t ci = 0;
          for (size_t b = 0; b < m.pos.size(); ++b) ci |= ((r >> m.pos[b]) & 1) << b;
          p.h_tab[t.tab_off + r] += cv[ci];
        }
      }
    }
    for (const cplx &z : p.h_tab) real &= (z.imag() == 0.0);
  }
  h.real_tables = real && !getenv("QOB_QTILE_NO_REALW");
}

// tile-local bit positions 0..T-5 belong to the thread id, the top four are the R bits (the thread's 16 amplitudes)
static void qr_layout(QRPassHost &ph, int nbits, uint64_t hi_value) {
  QRPass &P = ph.params;
  const std::vector<int> &fb = ph.free_bits;
  const int tb = (int)fb.size() - 4;   // thread bits: tile-local positions below the 4 R bits
  for (int k = 0; k < 4; ++k) P.rstride[k] = 1ull << fb[tb + k];
  for (int b = 0; b < tb; ++b) P.tgbit[b] = (unsigned char)fb[b];
  std::vector<int> fixed;
  for (int i = 0; i < nbits; ++i)
    if (std::find(fb.begin(), fb.end(), i) == fb.end()) fixed.push_back(i);
  qr_segments(fixed, P.xs_l, P.xs_n, P.xs_g, P.nfixed_seg);
  qr_segments(fb, P.fs_l, P.fs_n, P.fs_g, P.nfree_seg);
  P.hi_or = (nbits >= 64) ? 0ull : (hi_value << nbits);
}

// Tensor map of a pass: every run of free bits, merged with the run of fixed bits above it, is one dimension whose box
// extent covers only the free part; runs of more than 8 bits (256 elements) are split.  Elements are doubles (2 per amplitude).
static bool qr_tensor_geometry(QRPassHost &ph, int nbits) {
  const std::vector<int> &fb = ph.free_bits;
  std::vector<int> fixed;
  for (int i = 0; i < nbits; ++i)
    if (std::find(fb.begin(), fb.end(), i) == fb.end()) fixed.push_back(i);
  auto fixed_rank = [&](int bit) {  // compact position of a fixed bit
    return (int)(std::find(fixed.begin(), fixed.end(), bit) - fixed.begin());
  };
  QRPass &P = ph.params;
  int rank = 0;
  size_t i = 0;
  if (fb.empty() || fb[0] != 0) return false;
  while (i < fb.size()) {
    size_t j = i;
    while (j + 1 < fb.size() && fb[j + 1] == fb[j] + 1) ++j;
    int lo = fb[i], n = (int)(j - i + 1);            // run of free bits [lo, lo+n)
    int top = lo + n;                                // fixed bits [top, next free run or nbits)
    int nfix = 0;
    while (top + nfix < nbits && std::find(fb.begin(), fb.end(), top + nfix) == fb.end()) ++nfix;
    // split the free run into pieces of <= 8 bits (7 for the first run: 2 doubles per amplitude)
    int done = 0;
    while (done < n) {
      const int cap = (lo == 0 && done == 0) ? 7 : 8;
      const int piece = std::min(cap, n - done);
      const bool last = done + piece == n;
      if (rank >= 5) return false;
      const int start = lo + done;
      const int el = (start == 0) ? 1 : 0;           // the first dimension counts doubles
      ph.box[rank] = 1u << (piece + el);
      ph.gdim[rank] = 1ull << (piece + el + (last ? nfix : 0));
      ph.gstride[rank] = 16ull << start;             // bytes (unused for dimension 0)
      P.dim_boxlog[rank] = (unsigned char)(piece + el);
      P.dim_tshift[rank] = (unsigned char)((last && nfix) ? fixed_rank(top) : 0);
      P.dim_tbits[rank] = (unsigned char)(last ? nfix : 0);
      if (ph.gstride[rank] >= (1ull << 40)) return false;
      ++rank;
      done += piece;
    }
    i = j + 1;
  }
  ph.rank = rank;
  P.rank = rank;
  for (int d = rank; d < 5; ++d) P.dim_boxlog[d] = P.dim_tshift[d] = P.dim_tbits[d] = 0;
  return true;
}

typedef CUresult (*qr_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static qr_encode_fn qr_get_encode() {
  static qr_encode_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (qr_encode_fn)p;
  });
  return fn;
}
static int qr_encode_map(const QRPassHost &ph, const void *base, CUtensorMap *out) {
  qr_encode_fn enc = qr_get_encode();
  if (!enc) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t box[5], estr[5];
  for (int d = 0; d < ph.rank; ++d) {
    gdim[d] = ph.gdim[d];
    box[d] = ph.box[d];
    estr[d] = 1;
    if (d > 0) gstr[d - 1] = ph.gstride[d];
  }
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)ph.rank, const_cast<void *>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return QOB_STATUS_OK;
}

// merge diagonal components into tables over <= QR_DIAG_WINDOW selector bits; `ok(bits)` says whether a bit set can be
// addressed by one record (bit-field runs)
template <class OK>
static void qr_merge_diag(const QRegProgramHost &h, std::vector<int> ids, std::vector<QRTableHost> &out, OK ok) {
  std::sort(ids.begin(), ids.end(), [&](int a, int b) {
    const auto &sa = h.comps[a].sel, &sb = h.comps[b].sel;
    int la = sa.empty() ? -1 : sa.front(), lb = sb.empty() ? -1 : sb.front();
    if (la != lb) return la < lb;
    return sa < sb;
  });
  std::vector<bool> used(ids.size(), false);
  for (size_t i = 0; i < ids.size(); ++i) {
    if (used[i]) continue;
    QRTableHost t;
    std::vector<int> bits = h.comps[ids[i]].sel;
    std::vector<size_t> mem = {i};
    used[i] = true;
    for (size_t j = i + 1; j < ids.size(); ++j) {
      if (used[j]) continue;
      std::vector<int> u = bits;
      for (int b : h.comps[ids[j]].sel)
        if (std::find(u.begin(), u.end(), b) == u.end()) u.push_back(b);
      std::sort(u.begin(), u.end());
      if ((int)u.size() <= QR_DIAG_WINDOW && ok(u)) {
        bits = u;
        mem.push_back(j);
        used[j] = true;
      }
    }
    t.bits = bits;
    for (size_t m : mem) {
      QRTableHost::Member mm;
      mm.comp = ids[m];
      mm.pos.clear();
      t.members.push_back(mm);
    }
    out.push_back(std::move(t));
  }
}

int qreg_build(QRegProgram &prog, int nbits, uint64_t hi_value, const std::vector<QTerm> &terms, int sm_count) {
  (void)sm_count;
  auto h = std::make_shared<QRegProgramHost>();
  h->nbits = nbits;
  h->hi_value = hi_value;
  int T = qr_env_int("QOB_QREG_T", 12);
  if (T != 11 && T != 12) T = 12;
  h->T = T;
  if (nbits < T) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg needs at least %d index bits (got %d)", T, nbits);
  if (nbits - T > 31) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: too many tiles");
  const int L = 3;

  // ---- expand every term into flip / no-flip components, merged by (mask, selector bits)
  std::map<std::pair<uint64_t, std::vector<int>>, int> index;
  for (const QTerm &t : terms) {
    const int k = (int)t.bits.size();
    if (k > 3) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: term on %d sites (max 3)", k);
    for (int S = 0; S < (1 << k); ++S) {
      std::vector<cplx> unit((size_t)1 << k);
      bool any = false;
      for (int r = 0; r < (1 << k); ++r) {
        cplx w = 1.0;
        for (int f = 0; f < k; ++f) {
          const int i = (r >> f) & 1, j = ((S >> f) & 1) ? 1 - i : i;
          w *= t.m[4 * f + 2 * i + j];
        }
        unit[r] = w;
        any |= (w != cplx(0.0, 0.0));
      }
      if (!any) continue;
      uint64_t mask = 0;
      for (int f = 0; f < k; ++f)
        if ((S >> f) & 1) {
          if (t.bits[f] >= nbits) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: off-diagonal factor on a non-local bit %d", t.bits[f]);
          mask |= 1ull << t.bits[f];
        }
      auto key = std::make_pair(mask, t.bits);
      auto it = index.find(key);
      int id;
      if (it == index.end()) {
        id = (int)h->comps.size();
        index[key] = id;
        QRCompHost c;
        c.mask = mask;
        c.sel = t.bits;
        h->comps.push_back(c);
      } else {
        id = it->second;
      }
      h->comps[id].contribs.push_back({t.coef_index, t.scalar, unit});
    }
  }

  // ---- cover the distinct masks with passes of T free bits: pass 0 = the lowest T bits (contiguous tiles), every other
  // pass = the low block of L bits (128-byte runs) plus a window of high bits, chosen greedily from the lowest uncovered bit
  std::vector<uint64_t> masks;
  for (auto &c : h->comps)
    if (c.mask && std::find(masks.begin(), masks.end(), c.mask) == masks.end()) masks.push_back(c.mask);
  std::vector<uint64_t> free_sets;
  {
    free_sets.push_back((1ull << T) - 1);
    std::vector<uint64_t> remaining;
    for (uint64_t m : masks)
      if (m & ~free_sets[0]) remaining.push_back(m);
    const uint64_t lowL = (1ull << L) - 1;
    auto min_high = [&](uint64_t m) { uint64_t hgh = m & ~lowL; return hgh ? __builtin_ctzll(hgh) : 64; };
    std::sort(remaining.begin(), remaining.end(), [&](uint64_t a, uint64_t b) {
      const int ha = min_high(a), hb = min_high(b);
      if (ha != hb) return ha < hb;
      return a < b;
    });
    while (!remaining.empty()) {
      uint64_t fr = lowL;
      std::vector<uint64_t> rest;
      for (uint64_t m : remaining) {
        const uint64_t u = fr | m;
        if (__builtin_popcountll(u) <= T) fr = u;
        else rest.push_back(m);
      }
      if (fr == lowL) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: a term does not fit one tile");
      for (int b = 0; b < nbits && __builtin_popcountll(fr) < T; ++b) fr |= 1ull << b;  // spare bits widen the low block
      std::vector<uint64_t> rest2;
      for (uint64_t m : rest)
        if (m & ~fr) rest2.push_back(m);
      remaining.swap(rest2);
      free_sets.push_back(fr);
    }
  }
  const int np = (int)free_sets.size();
  std::vector<std::vector<int>> fbits(np);
  for (int p = 0; p < np; ++p)
    for (int b = 0; b < nbits; ++b)
      if (free_sets[p] >> b & 1) fbits[p].push_back(b);

  // ---- the R bits of a pass are its 4 highest free bits (tile-local positions 8..11): a mask inside them, on <= 2 bits and
  // with no other selector bit inside R, is applied from registers
  auto in_reg_ok = [&](const QRCompHost &c, uint64_t rbits) {
    if (!c.mask || (c.mask & ~rbits) || __builtin_popcountll(c.mask) > 2) return false;
    uint64_t sb = 0;
    for (int b : c.sel) sb |= 1ull << b;
    return (sb & rbits) == c.mask;
  };
  std::vector<std::vector<int>> rpos(np);
  std::vector<uint64_t> rbits(np, 0);
  for (int p = 0; p < np; ++p) {
    rpos[p] = {T - 4, T - 3, T - 2, T - 1};
    for (int k : rpos[p]) rbits[p] |= 1ull << fbits[p][k];
  }

  // ---- assign components to passes.  Diagonal weights go to pass 0.  A mask goes where it is register-resident if that
  // is possible, otherwise to the candidate with the fewest gathers so far (masks with one candidate are placed first).
  {
    std::vector<int> load(np, 0);
    std::map<uint64_t, int> mask_pass;
    for (int movable = 0; movable < 2; ++movable)
      for (auto &c : h->comps) {
        if (!c.mask) {
          c.pass = 0;
          continue;
        }
        std::vector<int> cand;
        for (int p = 0; p < np; ++p)
          if ((c.mask & ~free_sets[p]) == 0) cand.push_back(p);
        if (cand.empty()) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: internal planner error (uncovered mask)");
        if ((cand.size() > 1) != (movable == 1)) continue;
        auto it = mask_pass.find(c.mask);
        if (it != mask_pass.end()) {
          c.pass = it->second;
          continue;
        }
        int best = -1;
        for (int p : cand)
          if (in_reg_ok(c, rbits[p])) {
            best = p;
            break;
          }
        if (best < 0) {
          // the pass that runs alone (the last of an odd count) streams x through DRAM with shared-memory time to spare,
          // the chained ones are bound by the shared-memory pipe: the lone pass may carry two gathers more than the others
          auto eff = [&](int p) { return load[p] - ((np % 2 == 1 && np >= 3 && p == np - 1) ? 2 : 0); };
          best = cand[0];
          for (int p : cand)
            if (eff(p) < eff(best) || (eff(p) == eff(best) && p > best)) best = p;
          load[best]++;
        }
        c.pass = best;
        mask_pass[c.mask] = best;
      }
  }

  // ---- records and tables of every pass
  for (int p = 0; p < np; ++p) {
    auto ph = std::make_unique<QRPassHost>();
    ph->free_bits = fbits[p];
    ph->rpos = rpos[p];
    QRPass &P = ph->params;
    memset(&P, 0, sizeof(P));
    qr_layout(*ph, nbits, hi_value);
    if (P.nfree_seg > QR_MAXSEG || P.nfixed_seg > QR_MAXSEG) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: too many index segments");
    if (!qr_tensor_geometry(*ph, nbits)) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: tile shape needs a tensor map of rank > 5");
    std::vector<int> rg;  // flat-index bits of R, ascending
    for (int k : ph->rpos) rg.push_back(ph->free_bits[k]);
    // a selector bit is an R bit, a thread bit (tile-local position < 8, i.e. a bit of the thread id) or a tile / rank bit
    auto tpos_of = [&](int bit) {  // position in the thread id, -1 if none
      for (int q = 0; q < T - 4; ++q)
        if (ph->free_bits[q] == bit) return q;
      return -1;
    };
    auto is_r = [&](int bit) { return std::find(rg.begin(), rg.end(), bit) != rg.end(); };
    // split selector bits outside R into thread positions and tile bits and pack both as bit-field runs
    auto pack_outside = [&](const std::vector<int> &bits, std::vector<int> &tbits, std::vector<int> &gbits, uint32_t &selT, uint32_t &selG) {
      tbits.clear();
      gbits.clear();
      std::vector<int> tp;
      for (int b : bits) {
        if (is_r(b)) continue;
        const int q = tpos_of(b);
        if (q >= 0) {
          tbits.push_back(b);
          tp.push_back(q);
        } else {
          gbits.push_back(b);
        }
      }
      return qr_pack_runs(tp, selT) && qr_pack_runs(gbits, selG);
    };
    std::vector<int> pretile_ids, pre_ids, diagr_ids, inreg_ids, gather_ids;
    for (size_t id = 0; id < h->comps.size(); ++id) {
      const QRCompHost &c = h->comps[id];
      if (c.pass != p) continue;
      uint64_t sb = 0;
      bool any_thread = false;
      for (int b : c.sel) {
        sb |= 1ull << b;
        any_thread |= tpos_of(b) >= 0;
      }
      if (!c.mask) {
        if (sb & rbits[p]) diagr_ids.push_back((int)id);
        else (any_thread ? pre_ids : pretile_ids).push_back((int)id);
      } else if (in_reg_ok(c, rbits[p])) {
        inreg_ids.push_back((int)id);
      } else {
        gather_ids.push_back((int)id);
      }
    }
    ph->n_work = (int)(pretile_ids.size() + pre_ids.size() + diagr_ids.size() + inreg_ids.size() + gather_ids.size());
    std::vector<QRComp> recs;
    uint32_t tab_off = 0;
    // finish a table whose index bits are [inside R (given) | thread bits | tile bits] and emit its record
    auto emit = [&](QRTableHost &t, const std::vector<int> &inside, const std::vector<int> &all_bits, uint32_t kind, uint32_t xorB) -> bool {
      QRComp r;
      memset(&r, 0, sizeof r);
      std::vector<int> tb_, gb_;
      if (!pack_outside(all_bits, tb_, gb_, r.selT, r.selG)) return false;
      t.bits = inside;
      for (int b : tb_) t.bits.push_back(b);
      for (int b : gb_) t.bits.push_back(b);
      for (auto &m : t.members) {
        m.pos.clear();
        for (int b : h->comps[m.comp].sel) m.pos.push_back((int)(std::find(t.bits.begin(), t.bits.end(), b) - t.bits.begin()));
      }
      r.kind = kind;
      r.nT = (uint32_t)tb_.size();
      r.xorB = xorB;
      {
        // fast path: no tile-part selectors and at most one run of thread bits; the table index of the run is shifted past
        // the selector bits inside R (all 4 of them for a diagR table)
        const unsigned cls = kind & 15u;
        const unsigned shift = cls == QR_DIAGR ? 4u : (cls == QR_PRE || cls == QR_PRE_TILE ? 0u : (unsigned)__builtin_popcount((kind >> 4) & 15u));
        if (r.selG == 0 && (r.selT >> 9) == 0) {
          const unsigned sh = r.selT & 63u, w = (r.selT >> 6) & 7u;
          r.fast = (sh & 31u) | (((1u << w) - 1u) << 8) | (shift << 16);
        } else {
          r.fast = 1u << 31;
        }
        // dense body numbers (see the dispatch switches of the kernel)
        static const int order[11] = {0, 1, 2, 4, 8, 3, 5, 6, 9, 10, 12};
        const unsigned selr = (kind >> 4) & 15u, mR = (kind >> 8) & 15u;
        r.code = 255;
        if (cls == QR_INREG) {
          for (int q = 1; q < 11; ++q)
            if ((unsigned)order[q] == mR) r.code = (uint32_t)(q - 1);
        } else if (cls == QR_GATHER) {
          for (int q = 0; q < 11; ++q)
            if ((unsigned)order[q] == selr) r.code = (uint32_t)(q + (mR ? 11 : 0));
        }
      }
      r.tabE = tab_off;
      t.tab_off = tab_off;
      tab_off += 1u << t.bits.size();
      recs.push_back(r);
      ph->tables.push_back(t);
      return true;
    };
    auto runs_ok = [&](const std::vector<int> &bits) {
      std::vector<int> tb_, gb_;
      uint32_t a_, b_;
      return pack_outside(bits, tb_, gb_, a_, b_);
    };
    // (a) diagonal tables on tile / rank bits only: evaluated once per tile by the producer warp
    {
      std::vector<QRTableHost> tabs;
      qr_merge_diag(*h, pretile_ids, tabs, runs_ok);
      for (auto &t : tabs)
        if (!emit(t, {}, t.bits, QR_PRE_TILE, 0)) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: selector bits too scattered");
      P.n_pretile = (int)tabs.size();
    }
    // (b) diagonal tables that involve thread bits (but no R bit): once per thread and tile
    {
      std::vector<QRTableHost> tabs;
      qr_merge_diag(*h, pre_ids, tabs, runs_ok);
      for (auto &t : tabs)
        if (!emit(t, {}, t.bits, QR_PRE, 0)) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: selector bits too scattered");
      P.n_pre = (int)tabs.size();
    }
    // (c) diagonal tables indexed by all 4 R bits (low index bits) and <= 4 selector bits outside R
    {
      std::vector<bool> used(diagr_ids.size(), false);
      int n = 0;
      for (size_t i = 0; i < diagr_ids.size(); ++i) {
        if (used[i]) continue;
        std::vector<int> outside;
        QRTableHost t;
        for (size_t j = i; j < diagr_ids.size(); ++j) {
          if (used[j]) continue;
          std::vector<int> u = outside;
          for (int b : h->comps[diagr_ids[j]].sel)
            if (!is_r(b) && std::find(u.begin(), u.end(), b) == u.end()) u.push_back(b);
          std::sort(u.begin(), u.end());
          if (u.size() <= 4 && runs_ok(u)) {
            outside = u;
            QRTableHost::Member mm;
            mm.comp = diagr_ids[j];
            t.members.push_back(mm);
            used[j] = true;
          }
        }
        if (t.members.empty()) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: selector bits too scattered");
        if (!emit(t, rg, outside, QR_DIAGR | (15u << 4), 0)) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: selector bits too scattered");
        ++n;
      }
      P.n_diagr = n;
    }
    // (d, e) off-diagonal records: table index = [selector bits inside R | on thread bits | on tile bits], each ascending
    auto flip_record = [&](int id, int cls) -> bool {
      const QRCompHost &c = h->comps[id];
      std::vector<int> inside;
      for (int b : c.sel)
        if (is_r(b)) inside.push_back(b);
      unsigned selr = 0, mR = 0, xorB = 0;
      for (int k = 0; k < 4; ++k) {
        if (std::find(c.sel.begin(), c.sel.end(), rg[k]) != c.sel.end()) selr |= 1u << k;
        if (c.mask >> rg[k] & 1) mR |= 1u << k;
      }
      for (int pos = 0; pos < T - 4; ++pos)   // thread bits of the mask; its R bits travel in `kind` (mR)
        if (c.mask >> ph->free_bits[pos] & 1) xorB |= 16u << pos;
      if (cls == QR_GATHER && __builtin_popcount(selr) > 2) cls = QR_GATHER_SLOW;
      QRTableHost t;
      QRTableHost::Member mm;
      mm.comp = id;
      t.members.push_back(mm);
      return emit(t, inside, c.sel, (uint32_t)cls | (selr << 4) | (mR << 8), xorB);
    };
    for (int id : inreg_ids)
      if (!flip_record(id, QR_INREG)) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: selector bits too scattered");
    P.n_inreg = (int)inreg_ids.size();
    {
      // plain gathers (no selector bit and no mask bit inside R) first: the kernel runs them in its pipelined loop
      auto is_plain = [&](int id) {
        const QRCompHost &c = h->comps[id];
        uint64_t sb = 0;
        for (int b : c.sel) sb |= 1ull << b;
        return !(sb & rbits[p]) && !(c.mask & rbits[p]);
      };
      std::stable_partition(gather_ids.begin(), gather_ids.end(), is_plain);
      int nplain = 0;
      for (int id : gather_ids) {
        if (!flip_record(id, QR_GATHER)) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: selector bits too scattered");
        if (is_plain(id) && recs.back().code == 0) ++nplain;
      }
      P.n_plain = nplain;
    }
    P.n_gather = (int)gather_ids.size();
    if (recs.size() > QR_MAXC) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: %d lookup records in one pass (max %d)", (int)recs.size(), QR_MAXC);
    for (size_t i = 0; i < recs.size(); ++i) P.comps[i] = recs[i];
    ph->tab_entries = std::max<uint32_t>(tab_off, 1u);
    ph->h_tab.assign(ph->tab_entries, cplx(0.0, 0.0));
    if (ph->tab_entries * 16 > 12 * 1024) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: weight tables too large");
    h->passes.push_back(std::move(ph));
  }
  // drop passes without work (keep pass 0 when nothing at all is left: it carries the beta update)
  {
    std::vector<std::unique_ptr<QRPassHost>> keep;
    for (auto &pp : h->passes)
      if (pp->n_work > 0) keep.push_back(std::move(pp));
    h->passes.swap(keep);
  }
  if (h->passes.empty()) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: no work");

  // ---- group consecutive passes into chained launches: the union of their free bits must span <= chain_bits index bits
  const int chain_bits = qr_env_int("QOB_QREG_CHAIN_BITS", 20);
  // With an odd number of passes one of them runs alone, bound by DRAM (it streams x and y once) with shared-memory time to
  // spare, while a chained pair is bound by the shared-memory pipe with DRAM time to spare.  The lone pass goes FIRST: the
  // first launch only writes y (32 B/amplitude instead of 48), so the DRAM-bound launch is the one that saves a third of its
  // traffic; the pair after it accumulates into y through the L2.
  if (chain_bits >= T && h->passes.size() >= 3 && h->passes.size() % 2 == 1)
    std::rotate(h->passes.begin(), h->passes.end() - 1, h->passes.end());
  for (size_t i = 0; i < h->passes.size();) {
    QRegProgramHost::Group g;
    g.first = (int)i;
    g.count = 1;
    if (i + 1 < h->passes.size() && chain_bits >= T) {
      uint64_t u = 0;
      for (int b : h->passes[i]->free_bits) u |= 1ull << b;
      for (int b : h->passes[i + 1]->free_bits) u |= 1ull << b;
      const int ub = __builtin_popcountll(u);
      if (ub <= chain_bits) {
        g.count = 2;
        g.tpc_log2 = (unsigned)(ub - T);
        g.nchunks = 1u << (nbits - ub);
        const unsigned tpc = 1u << g.tpc_log2;
        g.lag = std::max(1u, (unsigned)((qr_env_int("QOB_QREG_LAG_TILES", 192) + tpc - 1) / tpc));
        // chunk / in-chunk split of the compact tile id of both passes
        for (int q = 0; q < 2; ++q) {
          QRPassHost &ph = *h->passes[i + q];
          QRPass &P = ph.params;
          std::vector<int> fixed;
          for (int b = 0; b < nbits; ++b)
            if (std::find(ph.free_bits.begin(), ph.free_bits.end(), b) == ph.free_bits.end()) fixed.push_back(b);
          // walk the fixed bits in compact order; bits outside the union number the chunk, bits inside it the tile in the chunk
          int ci = 0, ji = 0;
          P.ncs = P.njs = 0;
          int prev_kind = -1;
          for (size_t k = 0; k < fixed.size(); ++k) {
            const int kind = (u >> fixed[k] & 1) ? 1 : 0;  // 1: inside the union
            unsigned char *sl = kind ? P.js_l : P.cs_l, *sn = kind ? P.js_n : P.cs_n, *sd = kind ? P.js_d : P.cs_d;
            int &ns = kind ? P.njs : P.ncs;
            int &src = kind ? ji : ci;
            if (kind == prev_kind) {
              sn[ns - 1]++;
            } else {
              if (ns >= 4) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: chunk structure too scattered");
              sl[ns] = (unsigned char)src;
              sn[ns] = 1;
              sd[ns] = (unsigned char)k;
              ++ns;
            }
            ++src;
            prev_kind = kind;
          }
        }
      }
    }
    h->groups.push_back(g);
    i += g.count;
  }
  size_t maxchunks = 1;
  for (auto &g : h->groups) maxchunks = std::max<size_t>(maxchunks, g.nchunks);
  h->sync_words = 32 + maxchunks;

  prog.h = h;
  prog.npasses = (int)h->passes.size();
  char buf[320];
  snprintf(buf, sizeof buf, "qreg[bits=%d,T=%d,passes=%d,launches=%d,components=%d]", nbits, T, prog.npasses, (int)h->groups.size(),
           (int)h->comps.size());
  prog.describe = buf;
  for (auto &g : h->groups) {
    prog.describe += g.count == 2 ? " {chained" : " {single";
    if (g.count == 2) {
      snprintf(buf, sizeof buf, " chunks=%u x %u tiles, lag=%u:", g.nchunks, 1u << g.tpc_log2, g.lag);
      prog.describe += buf;
    }
    for (int q = 0; q < g.count; ++q) {
      const QRPassHost &ph = *h->passes[g.first + q];
      prog.describe += " [free:";
      unsigned char sl[QR_MAXSEG], sn[QR_MAXSEG], sg[QR_MAXSEG];
      int ns = 0;
      qr_segments(ph.free_bits, sl, sn, sg, ns);
      for (int s = 0; s < ns && s < QR_MAXSEG; ++s) {
        snprintf(buf, sizeof buf, "%s%d-%d", s ? "," : "", sg[s], sg[s] + sn[s] - 1);
        prog.describe += buf;
      }
      snprintf(buf, sizeof buf, " R:%d,%d,%d,%d diag:%d+%d+%d in-register:%d gathers:%d rank:%d]", ph.free_bits[ph.rpos[0]],
               ph.free_bits[ph.rpos[1]], ph.free_bits[ph.rpos[2]], ph.free_bits[ph.rpos[3]], ph.params.n_pretile, ph.params.n_pre,
               ph.params.n_diagr, ph.params.n_inreg, ph.params.n_gather, ph.rank);
      prog.describe += buf;
    }
    prog.describe += "}";
  }
  return QOB_STATUS_OK;
}

int qreg_set_coefs(QRegProgram &prog, const std::vector<cplx> &coefs, cudaStream_t s) {
  QRegProgramHost &h = *prog.h;
  for (auto &c : h.comps)
    for (auto &ct : c.contribs)
      if (ct.coef_index >= (int)coefs.size()) QOB_FAIL(QOB_STATUS_INVALID_ARG, "coefficient index out of range");
  qr_fill_tables(h, coefs);
  for (auto &pp : h.passes) {
    // device format: 8-byte reals when every weight of every pass is real, (re, im) pairs otherwise; padded to 16 bytes
    const size_t wb = h.real_tables ? 8 : 16;
    std::vector<unsigned char> raw(((pp->tab_entries * wb) + 15) & ~(size_t)15, 0);
    for (size_t i = 0; i < pp->tab_entries; ++i) {
      if (h.real_tables) {
        const double v = pp->h_tab[i].real();
        memcpy(raw.data() + i * 8, &v, 8);
      } else {
        const double v[2] = {pp->h_tab[i].real(), pp->h_tab[i].imag()};
        memcpy(raw.data() + i * 16, v, 16);
      }
    }
    pp->params.tab_bytes = (uint32_t)raw.size();
    if (pp->d_tab.n < ((pp->tab_entries * 16 + 15) & ~(size_t)15)) {
      std::vector<unsigned char> full((pp->tab_entries * 16 + 15) & ~(size_t)15, 0);
      memcpy(full.data(), raw.data(), raw.size());
      QOB_TRY(pp->d_tab.upload(full));
    } else {
      QOB_TRY(pp->d_tab.upload_async(raw, s));
    }
    pp->params.tab = pp->d_tab.ptr;
  }
  return QOB_STATUS_OK;
}


static int qreg_launch_impl(const QRegProgram &prog, cplx alpha, const void *x, cplx beta, void *y, cudaStream_t s, const QRegOpts &o,
                            cplx *expect_out);
int qreg_launch(const QRegProgram &prog, cplx alpha, const void *x, cplx beta, void *y, cudaStream_t s, int max_ctas) {
  QRegOpts o;
  o.max_ctas = max_ctas;
  return qreg_launch_impl(prog, alpha, x, beta, y, s, o, nullptr);
}
int qreg_launch_ex(const QRegProgram &prog, cplx alpha, const void *x, cplx beta, void *y, cudaStream_t s, const QRegOpts &o) {
  if (!qreg_supports(prog, o)) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: peer-addressed / chunked launch not supported by this program");
  return qreg_launch_impl(prog, alpha, x, beta, y, s, o, nullptr);
}
// lowest run of free bits of a pass = the contiguous pieces a peer-addressed launch moves
static int qr_low_run(const QRPassHost &ph) {
  int r = 0;
  while (r < (int)ph.free_bits.size() && ph.free_bits[r] == r) ++r;
  return r;
}
bool qreg_supports(const QRegProgram &prog, const QRegOpts &o) {
  if (!prog.h) return false;
  const QRegProgramHost &h = *prog.h;
  if (o.npeers == 0 && o.nchunks <= 1) return true;
  if (h.passes.size() != 1 || h.T != 12) return false;
  const QRPassHost &ph = *h.passes[0];
  if (o.npeers > 0) {
    if (o.npeers > QR_MAXPEER || (o.npeers & (o.npeers - 1)) || !o.xpeer || !o.ypeer) return false;
    const int r = qr_low_run(ph);
    if (r < 6 || o.peer_shift < r) return false;       // pieces of >= 1 KiB that never straddle two ranks
  }
  if (o.nchunks > 1) {
    if (o.nchunks & (o.nchunks - 1)) return false;
    if (__builtin_popcountll(o.chunk_mask) != __builtin_ctz((unsigned)o.nchunks)) return false;
    for (int b : ph.free_bits)
      if (o.chunk_mask >> b & 1) return false;         // range bits must be fixed bits of the pass
    if (o.chunk_mask >> h.nbits) return false;
  }
  return true;
}
// <x| op |x> without writing op*x: every launch runs in mode 3 and leaves per-warp partial sums; tiles are dealt statically
// and the partials are added in a fixed order, so the result is reproducible bit for bit
int qreg_expect(const QRegProgram &prog, const void *x, cplx *out, cudaStream_t s) {
  return qreg_launch_impl(prog, cplx(1.0, 0.0), x, cplx(0.0, 0.0), nullptr, s, QRegOpts(), out);
}
static int qreg_launch_impl(const QRegProgram &prog, cplx alpha, const void *x, cplx beta, void *y, cudaStream_t s, const QRegOpts &o,
                            cplx *expect_out) {
  QRegProgramHost &h = *prog.h;
  const int max_ctas = o.max_ctas;
  const bool peer = o.npeers > 0;
  int dev = 0, sms = 148;
  QOB_CUDA(cudaGetDevice(&dev));
  QOB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  unsigned *sync = nullptr;
  {
    std::lock_guard<std::mutex> lk(h.mu);
    auto it = h.sync_bufs.find(s);
    if (it == h.sync_bufs.end()) {
      QOB_CUDA(cudaMalloc(&sync, h.sync_words * sizeof(unsigned)));
      h.sync_bufs[s] = sync;
    } else {
      sync = it->second;
    }
  }
  // y = alpha*H x + beta*y: beta == 0 -> the first pass stores (y is never read); beta == 1 -> every pass adds; any other
  // beta -> y is scaled first (one extra light pass), then every pass adds
  if (peer && beta != cplx(0.0, 0.0) && beta != cplx(1.0, 0.0))
    QOB_FAIL(QOB_STATUS_INVALID_ARG, "peer-addressed launch: beta must be 0 (store) or 1 (add into the owners' buffers)");
  if (!expect_out && beta != cplx(0.0, 0.0) && beta != cplx(1.0, 0.0)) QOB_TRY(launch_scale(y, (int64_t)1 << h.nbits, beta, s));
  double2 *partials = nullptr;
  const size_t part_per_launch = (size_t)sms * 8;   // consumer warps per SM: 1 CTA x 8 (T = 12) or 2 CTAs x 4 (T = 11)
  if (expect_out) {
    std::lock_guard<std::mutex> lk(h.mu);
    auto it = h.part_bufs.find(s);
    if (it == h.part_bufs.end()) {
      QOB_CUDA(cudaMalloc(&partials, 8 * part_per_launch * sizeof(double2)));
      h.part_bufs[s] = partials;
    } else {
      partials = it->second;
    }
    if (h.groups.size() > 8) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: too many launches for the fused expect");
    QOB_CUDA(cudaMemsetAsync(partials, 0, 8 * part_per_launch * sizeof(double2), s));
  }
  bool first = true;
  int gi = 0;
  for (const auto &g : h.groups) {
    // algorithmic bytes of this launch at DRAM level: x read + y written once (+ y read when it accumulates), however many
    // tile passes are chained inside it
    void *prof_token = qprof_enabled() ? qprof_begin(s, gi, (double)(1ull << h.nbits) * ((first && beta == cplx(0.0, 0.0)) ? 32.0 : 48.0))
                                       : nullptr;
    ++gi;
    QRLaunch Lp;
    memset(&Lp, 0, sizeof Lp);
    Lp.npass = g.count;
    Lp.alpha = make_double2(alpha.real(), alpha.imag());
    Lp.beta = make_double2(beta.real(), beta.imag());
    Lp.queue = sync;
    Lp.done = sync + 32;
    const unsigned ntiles = 1u << (h.nbits - h.T);
    Lp.tpc_log2 = g.tpc_log2;
    Lp.nchunks = g.nchunks;
    Lp.lag = g.lag;
    Lp.total_items = ntiles * (unsigned)g.count;
    Lp.prefetch = 0;
    static long long *stats_buf = nullptr;
    const bool want_stats = qr_env_int("QOB_QREG_STATS", 0) != 0;
    if (want_stats && !stats_buf) QOB_CUDA(cudaMalloc(&stats_buf, 16 * 256 * sizeof(long long)));
    Lp.stats = want_stats ? stats_buf : nullptr;
    alignas(64) CUtensorMap maps[4];
    uint32_t tab_off = 0;
    for (int q = 0; q < g.count; ++q) {
      QRPassHost &ph = *h.passes[g.first + q];
      {
        std::lock_guard<std::mutex> lk(h.mu);
        if (peer) {
          // no tensor maps: the tile moves as contiguous pieces addressed per rank (any valid map fills the parameter slots)
          if (!ph.map_x_ptr) QOB_TRY(qr_encode_map(ph, o.xpeer[0], &ph.map_x));
          if (!ph.map_y_ptr) ph.map_y = ph.map_x;
        } else if (ph.map_x_ptr != x) {
          QOB_TRY(qr_encode_map(ph, x, &ph.map_x));
          ph.map_x_ptr = x;
        }
        if (!peer && !expect_out && ph.map_y_ptr != y) {
          QOB_TRY(qr_encode_map(ph, y, &ph.map_y));
          ph.map_y_ptr = y;
        }
        maps[2 * q] = ph.map_x;
        maps[2 * q + 1] = ph.map_y;
      }
      Lp.pass[q] = ph.params;
      QRPass &P = Lp.pass[q];
      P.mode = expect_out ? 3 : ((first && beta == cplx(0.0, 0.0)) ? 0 : 2);   // 0: stored; 2: added to y by the L2; 3: reduced
      first = false;
      P.signal = (g.count == 2 && q == 0 && !expect_out) ? 1 : 0;   // mode 3 writes nothing: the chunk order alone keeps x in L2
      P.wait = (g.count == 2 && q == 1 && !expect_out) ? 1 : 0;
      P.stream_out = (q == g.count - 1) ? 1 : 0;
      P.tab_smem_off = tab_off;
      tab_off += (P.tab_bytes + 15u) & ~15u;
    }
    if (g.count == 1) {
      maps[2] = maps[0];
      maps[3] = maps[1];
    }
    Lp.nstage = QR_NSTAGE;
    unsigned items = ntiles * (unsigned)g.count;
    if (peer) {
      QRPass &P = Lp.pass[0];
      P.peer_bits = __builtin_ctz((unsigned)o.npeers);
      P.peer_shift = o.peer_shift;
      P.peer_rank = o.peer_rank;
      P.piece_log2 = std::min(qr_low_run(*h.passes[g.first]), 10);   // pieces of <= 16 KiB
      for (int q = 0; q < o.npeers; ++q) {
        Lp.xpeer[q] = (const double2 *)o.xpeer[q];
        Lp.ypeer[q] = (double2 *)o.ypeer[q];
      }
    }
    if (o.nchunks > 1) {
      // tile range `chunk_index`: the chunk bits (fixed bits of the pass) hold chunk_index, the other fixed bits count the tiles
      QRPassHost &ph = *h.passes[g.first];
      QRPass &P = Lp.pass[0];
      std::vector<int> fixed;
      for (int b = 0; b < h.nbits; ++b)
        if (std::find(ph.free_bits.begin(), ph.free_bits.end(), b) == ph.free_bits.end()) fixed.push_back(b);
      int ci = 0, ji = 0, prev_kind = -1;
      P.ncs = P.njs = 0;
      for (size_t k = 0; k < fixed.size(); ++k) {
        const int kind = (o.chunk_mask >> fixed[k] & 1) ? 0 : 1;   // 0: numbers the range, 1: counts inside it
        unsigned char *sl = kind ? P.js_l : P.cs_l, *sn = kind ? P.js_n : P.cs_n, *sd = kind ? P.js_d : P.cs_d;
        int &ns = kind ? P.njs : P.ncs;
        int &src = kind ? ji : ci;
        if (kind == prev_kind) {
          sn[ns - 1]++;
        } else {
          if (ns >= 4) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qreg: tile-range structure too scattered");
          sl[ns] = (unsigned char)src;
          sn[ns] = 1;
          sd[ns] = (unsigned char)k;
          ++ns;
        }
        ++src;
        prev_kind = kind;
      }
      Lp.remap = 1;
      Lp.chunk_fixed = (unsigned)o.chunk_index;
      items = ntiles / (unsigned)o.nchunks;
      Lp.total_items = items;
    }
    Lp.static_queue = expect_out ? 1 : 0;
    Lp.partials = expect_out ? partials + (size_t)(gi - 1) * part_per_launch : nullptr;
    const size_t smem = (size_t)Lp.nstage * ((size_t)16 << h.T) + tab_off + 16 + 96 + QR_NSTAGE * sizeof(QRItem) + 2 * QR_MAXC * 16 + 128;
    QOB_CUDA(cudaMemsetAsync(sync, 0, (32 + (size_t)g.nchunks) * sizeof(unsigned), s));
    // one persistent CTA per SM; max_ctas > 0 leaves the other SMs to a kernel of another stream (a CTA of this kernel fills
    // an SM's shared memory, so the two kernels never share an SM: the exchange of a sharded apply gets its own SMs and NVLink
    // rate, the local passes the rest)
    unsigned grid = std::min<unsigned>((unsigned)sms * (h.T == 11 ? 2u : 1u), items);
    if (max_ctas > 0) grid = std::min<unsigned>(grid, (unsigned)max_ctas);
    auto launch = [&](auto kern) -> int {
      QOB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, (1u << (h.T - 4)) + 64u, smem, s>>>(Lp, maps[0], maps[1], maps[2], maps[3]);
      QOB_LAUNCHED();
      QOB_LAUNCHED_FAMILY(peer ? 3 : 2);
      QOB_CUDA(cudaGetLastError());
      return QOB_STATUS_OK;
    };
    if (h.T == 11) {
      if (h.real_tables) QOB_TRY(launch(qreg_kernel<true, 11>));
      else QOB_TRY(launch(qreg_kernel<false, 11>));
    } else {
      if (peer) {
        if (h.real_tables) QOB_TRY(launch(qreg_kernel<true, 12, true>));
        else QOB_TRY(launch(qreg_kernel<false, 12, true>));
      } else if (h.real_tables) {
        QOB_TRY(launch(qreg_kernel<true, 12>));
      } else {
        QOB_TRY(launch(qreg_kernel<false, 12>));
      }
    }
    qprof_end(s, prof_token);
    if (want_stats && !expect_out) {
      std::vector<long long> hs(16 * 256);
      cudaStreamSynchronize(s);
      cudaMemcpy(hs.data(), stats_buf, hs.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      for (int b : {0, 1, 73, 147}) {
        const long long *o = hs.data() + 16 * b;
        fprintf(stderr, "[qreg stats] launch %d cta %3d: tiles %lld total %lld | producer wait x-empty %lld dep %lld y-empty %lld | consumer w0 wait-x %lld compute %lld wait-y %lld epilogue %lld | w7 %lld %lld %lld %lld\n",
                gi - 1, b, o[4], o[3], o[0], o[1], o[2], o[8], o[9], o[10], o[11], o[12], o[13], o[14], o[15]);
      }
    }
  }
  if (expect_out) {
    std::vector<double2> hp(h.groups.size() * part_per_launch);
    QOB_CUDA(cudaMemcpyAsync(hp.data(), partials, hp.size() * sizeof(double2), cudaMemcpyDeviceToHost, s));
    QOB_CUDA(cudaStreamSynchronize(s));
    double re = 0.0, im = 0.0;
    for (const double2 &v : hp) {
      re += v.x;
      im += v.y;
    }
    *expect_out = cplx(re, im);
  }
  return QOB_STATUS_OK;
}
