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

struct QCompDev {       // 16 bytes, read as one int4 broadcast from shared memory
  uint32_t lmask;       // tile-local flip mask
  uint32_t sel;         // selector bit positions: byte j = global bit of table-index bit j; 0xFF = unused
  uint32_t tab_off;     // offset of this component's 2^nsel weights in the pass table
  uint32_t flags;       // bit0: last component of its mask group (do the gather+FMA now)
};

struct QPassParams {
  const QCompDev *comps;  // device
  const double2 *tab;     // device
  int ncomp, ntab;
  int nfree_seg, nfixed_seg;
  // segment = (shift in the compact index, length, position in the address)
  unsigned char fs_l[QT_MAXSEG], fs_n[QT_MAXSEG], fs_g[QT_MAXSEG];  // free bits: tile-local index -> address
  unsigned char xs_l[QT_MAXSEG], xs_n[QT_MAXSEG], xs_g[QT_MAXSEG];  // fixed bits: tile id -> address
  unsigned long long hi_or;  // index bits above the local address (rank), already shifted
  double2 alpha, beta;
  int mode;  // 0: y = a*acc ; 1: y = a*acc + beta*y ; 2: y = a*acc + y
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

template <int T, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
    qtile_kernel(const __grid_constant__ QPassParams P, const double2 *__restrict__ x, double2 *__restrict__ y) {
  constexpr int TILE = 1 << T;
  constexpr int U = 4;                    // amplitudes in flight per thread
  constexpr int ITERS = TILE / (THREADS * U);
  static_assert(ITERS >= 1, "tile too small for this block size");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *xs = reinterpret_cast<double2 *>(smem_raw);
  double2 *tab = xs + TILE;
  QCompDev *comps = reinterpret_cast<QCompDev *>(tab + P.ntab);

  const unsigned tid = threadIdx.x;
  const unsigned long long base = qexpand(blockIdx.x, P.nfixed_seg, P.xs_l, P.xs_n, P.xs_g);
  const unsigned long long a_tid = qexpand(tid, P.nfree_seg, P.fs_l, P.fs_n, P.fs_g);

  // ---- stage the x tile: 16-byte cp.async per amplitude, lanes walk the contiguous low block
#pragma unroll 4
  for (int it = 0; it < TILE / THREADS; ++it) {
    const unsigned l = it * THREADS + tid;
    const unsigned long long a = base | a_tid | qexpand(it * THREADS, P.nfree_seg, P.fs_l, P.fs_n, P.fs_g);
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(xs + l);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(x + a));
  }
  asm volatile("cp.async.commit_group;\n" ::);
  for (int i = tid; i < P.ntab; i += THREADS) tab[i] = P.tab[i];
  for (int i = tid; i < P.ncomp; i += THREADS) comps[i] = P.comps[i];
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();

  const int ncomp = P.ncomp;
  const unsigned long long at = base | a_tid;  // per-thread part of the address
  for (int it = 0; it < ITERS; ++it) {
    unsigned glo[U], ghi[U];
    double2 acc[U], ws[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // (it*U+u)*THREADS is warp-uniform: its expansion lives in uniform registers
      const unsigned long long gg = at | qexpand((it * U + u) * THREADS, P.nfree_seg, P.fs_l, P.fs_n, P.fs_g) | P.hi_or;
      glo[u] = (unsigned)gg;
      ghi[u] = (unsigned)(gg >> 32);
      acc[u] = make_double2(0.0, 0.0);
      ws[u] = make_double2(0.0, 0.0);
    }
    for (int c = 0; c < ncomp; ++c) {
      const QCompDev cd = comps[c];
      const unsigned s0 = cd.sel & 0xFF, s1 = (cd.sel >> 8) & 0xFF, s2 = (cd.sel >> 16) & 0xFF;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        unsigned idx = 0;
        if (s0 != 0xFF) idx |= (((s0 < 32) ? glo[u] : ghi[u]) >> (s0 & 31)) & 1u;
        if (s1 != 0xFF) idx |= ((((s1 < 32) ? glo[u] : ghi[u]) >> (s1 & 31)) & 1u) << 1;
        if (s2 != 0xFF) idx |= ((((s2 < 32) ? glo[u] : ghi[u]) >> (s2 & 31)) & 1u) << 2;
        const double2 w = tab[cd.tab_off + idx];
        ws[u].x += w.x;
        ws[u].y += w.y;
      }
      if (cd.flags & 1u) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned l = (it * U + u) * THREADS + tid;
          const double2 v = xs[l ^ cd.lmask];
          qcfma(acc[u], ws[u], v);
          ws[u] = make_double2(0.0, 0.0);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long a = at | qexpand((it * U + u) * THREADS, P.nfree_seg, P.fs_l, P.fs_n, P.fs_g);
      double2 o;
      o.x = P.alpha.x * acc[u].x - P.alpha.y * acc[u].y;
      o.y = P.alpha.x * acc[u].y + P.alpha.y * acc[u].x;
      if (P.mode == 1) {
        qcfma(o, P.beta, y[a]);
      } else if (P.mode == 2) {
        const double2 yo = y[a];
        o.x += yo.x;
        o.y += yo.y;
      }
      y[a] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------- host
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
  uint32_t tab_off = 0;
};
struct QPassHost {
  std::vector<int> free_bits;  // ascending, size T
  std::vector<int> comp_ids;   // sorted by mask
  QPassParams params;
  DevArray<QCompDev> d_comps;
  DevArray<double2> d_tab;
  std::vector<double2> h_tab;
};
struct QTileProgramHost {
  int nbits = 0, T = 0, L = 0, threads = 256;
  uint64_t hi_value = 0;
  std::vector<QCompHost> comps;
  std::vector<std::unique_ptr<QPassHost>> passes;
  size_t smem_bytes(const QPassHost &p) const {
    return ((size_t)1 << T) * sizeof(double2) + p.params.ntab * sizeof(double2) + p.params.ncomp * sizeof(QCompDev);
  }
};

bool qtile_supported_term(const QTerm &t) { return t.bits.size() <= QT_MAXSEL; }

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
    sl[nseg] = (unsigned char)i;
    sn[nseg] = (unsigned char)(j - i + 1);
    sg[nseg] = (unsigned char)bits[i];
    ++nseg;
    i = j + 1;
  }
}

static void fill_tables(QTileProgramHost &h, const std::vector<cplx> &coefs) {
  for (auto &pp : h.passes) {
    QPassHost &p = *pp;
    for (int id : p.comp_ids) {
      const QCompHost &c = h.comps[id];
      const size_t n = (size_t)1 << c.sel.size();
      for (size_t r = 0; r < n; ++r) {
        cplx w = 0.0;
        for (const auto &ct : c.contribs) {
          cplx f = ct.scalar;
          if (ct.coef_index >= 0) f *= coefs[ct.coef_index];
          w += f * ct.unit[r];
        }
        p.h_tab[c.tab_off + r] = make_double2(w.real(), w.imag());
      }
    }
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
  if (L < 1) L = 1;
  if (L > T - 3) L = T - 3;
  h->T = T;
  h->L = L;
  h->threads = (T == 13) ? 512 : 256;

  // ---- expand every term into flip/no-flip components, merged by (mask, selector bits)
  std::map<std::pair<uint64_t, std::vector<int>>, int> index;
  for (const QTerm &t : terms) {
    const int k = (int)t.bits.size();
    if (k > QT_MAXSEL) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: term on %d sites (max %d)", k, QT_MAXSEL);
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

  // ---- cover the distinct non-zero masks with passes of T free bits
  std::vector<uint64_t> masks;
  for (auto &c : h->comps)
    if (c.mask && std::find(masks.begin(), masks.end(), c.mask) == masks.end()) masks.push_back(c.mask);
  std::vector<uint64_t> free_sets;
  free_sets.push_back((T >= 64) ? ~0ull : ((1ull << T) - 1));  // pass 0: the lowest T bits (fully contiguous tiles)
  std::vector<uint64_t> remaining;
  for (uint64_t m : masks)
    if (m & ~free_sets[0]) remaining.push_back(m);
  const uint64_t lowL = (1ull << L) - 1;
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
      if (__builtin_popcountll(fr | m) <= T) fr |= m;
      else rest.push_back(m);
    }
    if (fr == lowL) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: a term does not fit one tile");
    // a second sweep: masks that became coverable because their bits were added meanwhile
    for (int b = 0; b < nbits && __builtin_popcountll(fr) < T; ++b) fr |= 1ull << b;  // spend spare bits on the low block
    std::vector<uint64_t> rest2;
    for (uint64_t m : rest)
      if (m & ~fr) rest2.push_back(m);
    remaining.swap(rest2);
    free_sets.push_back(fr);
  }

  // ---- assign components to passes (diagonal -> pass 0; others -> first pass containing the mask)
  for (auto &c : h->comps) {
    c.pass = -1;
    for (size_t p = 0; p < free_sets.size(); ++p)
      if ((c.mask & ~free_sets[p]) == 0) {
        c.pass = (int)p;
        break;
      }
    if (c.pass < 0) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: internal planner error (uncovered mask)");
  }
  for (size_t p = 0; p < free_sets.size(); ++p) {
    auto ph = std::make_unique<QPassHost>();
    for (int b = 0; b < nbits; ++b)
      if (free_sets[p] >> b & 1) ph->free_bits.push_back(b);
    for (size_t id = 0; id < h->comps.size(); ++id)
      if (h->comps[id].pass == (int)p) ph->comp_ids.push_back((int)id);
    if (p > 0 && ph->comp_ids.empty()) continue;
    std::stable_sort(ph->comp_ids.begin(), ph->comp_ids.end(),
                     [&](int a, int b) { return h->comps[a].mask < h->comps[b].mask; });
    // device records
    std::vector<QCompDev> dc;
    uint32_t tab_off = 0;
    for (size_t i = 0; i < ph->comp_ids.size(); ++i) {
      QCompHost &c = h->comps[ph->comp_ids[i]];
      QCompDev d;
      uint32_t lmask = 0;
      for (size_t j = 0; j < ph->free_bits.size(); ++j)
        if (c.mask >> ph->free_bits[j] & 1) lmask |= 1u << j;
      d.lmask = lmask;
      d.sel = 0xFFFFFFFFu;
      for (size_t j = 0; j < c.sel.size(); ++j) d.sel = (d.sel & ~(0xFFu << (8 * j))) | ((uint32_t)c.sel[j] << (8 * j));
      d.tab_off = tab_off;
      c.tab_off = tab_off;
      tab_off += 1u << c.sel.size();
      bool last = (i + 1 == ph->comp_ids.size()) || h->comps[ph->comp_ids[i + 1]].mask != c.mask;
      d.flags = last ? 1u : 0u;
      dc.push_back(d);
    }
    if (dc.empty()) {  // a sum with no term at all in pass 0: keep one zero-weight component so y is still written
      QCompDev d = {0u, 0xFFFFFFFFu, 0u, 1u};
      dc.push_back(d);
      tab_off = 1;
    }
    ph->h_tab.assign(std::max<uint32_t>(tab_off, 1), make_double2(0.0, 0.0));
    QPassParams &P = ph->params;
    memset(&P, 0, sizeof(P));
    P.ncomp = (int)dc.size();
    P.ntab = (int)ph->h_tab.size();
    make_segments(ph->free_bits, P.fs_l, P.fs_n, P.fs_g, P.nfree_seg);
    std::vector<int> fixed;
    for (int b = 0; b < nbits; ++b)
      if (!(free_sets[p] >> b & 1)) fixed.push_back(b);
    make_segments(fixed, P.xs_l, P.xs_n, P.xs_g, P.nfixed_seg);
    if (P.nfree_seg > QT_MAXSEG || P.nfixed_seg > QT_MAXSEG) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: too many index segments");
    P.hi_or = (nbits >= 64) ? 0ull : (hi_value << nbits);
    QOB_TRY(ph->d_comps.upload(dc));
    QOB_TRY(ph->d_tab.upload(ph->h_tab));
    P.comps = ph->d_comps.ptr;
    P.tab = ph->d_tab.ptr;
    h->passes.push_back(std::move(ph));
  }
  prog.h = h;
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
    snprintf(buf, sizeof buf, " comps:%d}", pp->params.ncomp);
    prog.describe += buf;
  }
  (void)sm_count;
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

template <int T, int THREADS, int MINB>
static int launch_pass(const QTileProgramHost &h, const QPassHost &p, const QPassParams &P, const void *x, void *y,
                       cudaStream_t s) {
  size_t smem = h.smem_bytes(p);
  static size_t configured = 0;
  if (smem > configured) {
    QOB_CUDA(cudaFuncSetAttribute(qtile_kernel<T, THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const uint64_t ntiles = 1ull << (h.nbits - T);
  if (ntiles > 0x7FFFFFFFull) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: too many tiles");
  qtile_kernel<T, THREADS, MINB><<<(unsigned)ntiles, THREADS, smem, s>>>(P, (const double2 *)x, (double2 *)y);
  QOB_LAUNCHED();
  QOB_CUDA(cudaGetLastError());
  return QOB_STATUS_OK;
}

// ---- optional per-launch event timing (qob_profile_enable / qob_profile_read)
struct ProfEntry {
  cudaEvent_t a, b;
  int pass;
  double bytes;
};
static bool g_prof_on = false;
static std::vector<ProfEntry> g_prof;
extern "C" int qob_profile_enable(int32_t on) {
  g_prof_on = on != 0;
  return QOB_STATUS_OK;
}
extern "C" int qob_profile_read(int32_t max_entries, float *ms, int32_t *pass_index, double *alg_bytes, int32_t *count) {
  int n = 0;
  for (ProfEntry &e : g_prof) {
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
  g_prof.clear();
  if (count) *count = n;
  return QOB_STATUS_OK;
}

int qtile_launch(const QTileProgram &prog, cplx alpha, const void *x, cplx beta, void *y, cudaStream_t s) {
  const QTileProgramHost &h = *prog.h;
  bool first = true;
  int pass_no = 0;
  for (auto &pp : h.passes) {
    ProfEntry pe;
    if (g_prof_on) {
      cudaEventCreate(&pe.a);
      cudaEventCreate(&pe.b);
      pe.pass = pass_no;
      pe.bytes = (double)(1ull << h.nbits) * ((first && beta == cplx(0.0, 0.0)) ? 32.0 : 48.0);
      cudaEventRecord(pe.a, s);
    }
    ++pass_no;
    QPassParams P = pp->params;
    P.alpha = make_double2(alpha.real(), alpha.imag());
    P.beta = make_double2(beta.real(), beta.imag());
    P.mode = first ? (beta == cplx(0.0, 0.0) ? 0 : 1) : 2;
    first = false;
    switch (h.T) {
      case 10: QOB_TRY((launch_pass<10, 256, 3>(h, *pp, P, x, y, s))); break;
      case 11: QOB_TRY((launch_pass<11, 256, 3>(h, *pp, P, x, y, s))); break;
      case 12: QOB_TRY((launch_pass<12, 256, 3>(h, *pp, P, x, y, s))); break;
      case 13: QOB_TRY((launch_pass<13, 512, 1>(h, *pp, P, x, y, s))); break;
      default: QOB_FAIL(QOB_STATUS_UNSUPPORTED, "qtile: unsupported tile size %d", h.T);
    }
    if (g_prof_on) {
      cudaEventRecord(pe.b, s);
      g_prof.push_back(pe);
    }
  }
  return QOB_STATUS_OK;
}
