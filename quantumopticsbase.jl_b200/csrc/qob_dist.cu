// qob_dist.cu — the sharded (multi-GPU) LazySum apply behind the C ABI (include/qob200.h, `qob_dist_*`): planning, CUDA-IPC
// mapping of the ranks' slabs, device-side cross-rank barriers and the stream / event choreography of the fused exchange.
// The host program (Python, Julia, C) only has to move 64-byte IPC handles between its processes once; no torch, no NCCL and
// no host round trip sits in the data path.
//
// No reference counterpart: QuantumOpticsBase has no distributed path.  The state is sharded on its
// highest-stride axes (rank = top p index bits of the reference's column-major array); see quantumopticsbase.jl_b200/dist.py for the
// original Python orchestration that this file restates (it stays as the torch.distributed / NCCL fallback).
//
// One apply:
//   side stream : barrier | remote-term pass, chunk 0 (PEER tile kernel: loads x tiles from the owners' slabs over NVLink,
//                 stores results into the owners' contribution slabs) | barrier | chunk 1 | barrier | ...
//   main stream : communication-free terms, group A (beside the exchange)      | group B chunk 0 (+ contributions) | chunk 1 ...
// Direct mode (qob_dist_bind_result: the result slab is mapped by the peers too; no contribution slab, 2 slabs per rank):
//   main stream : y = beta*y | ------------ every communication-free term, on all but k SMs, adding into y ------------ | wait
//   side stream :             barrier | remote-term pass on k SMs: x pieces come from the owners' slabs by bulk copy, results
//                                       are ADDED into the owners' y slabs (cp.reduce.async.bulk, f64 add in the owner's L2) | barrier
// The barrier is a tiny kernel: every rank writes its epoch into every peer's signal pad (system-scope release) and spins until
// every peer's epoch has arrived in its own pad (system-scope acquire).
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "qob_internal.h"

struct qob_dist {
  qob_op *sum = nullptr;
  qob_ctx *ctx = nullptr;
  int rank = 0, world = 1, p = 0, n = 0, nloc = 0;
  int n_local = 0, n_remote = 0;
  int plan_a = -1, plan_b = -1, plan_r = -1;
  bool direct_ok = false;   // the exchange pass can ADD into the owners' result slabs (round-2 kernel, peer-addressed)
  std::vector<void *> y_peers;
  bool timing = false;      // qob_dist_exchange_timing: CUDA events around the exchange of every apply
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;
  int swap_lo = -1;
  int nchunks = 1;
  uint64_t chunk_mask = 0;
  int swap_sms = 32;
  // bound buffers
  bool bound = false;
  std::vector<void *> x_peers, z_peers, flag_peers;
  unsigned long long **d_flag_table = nullptr;   // device copy of flag_peers
  unsigned long long epoch = 0;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_main = nullptr, ev_done = nullptr;
  std::vector<cudaEvent_t> ev_chunk;
  std::string text;
};

__global__ void dist_barrier_kernel(unsigned long long *const *flags, int rank, int world, unsigned long long epoch) {
  const int q = threadIdx.x;
  if (q >= world) return;
  // everything this GPU wrote before (peer stores of the exchange pass included) is ordered before the signal
  __threadfence_system();
  unsigned long long *theirs = flags[q] + rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
  const unsigned long long *mine = flags[rank] + q;
  unsigned long long seen;
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
  } while (seen < epoch);
}

static int dist_barrier(qob_dist *d, cudaStream_t s) {
  ++d->epoch;
  dist_barrier_kernel<<<1, 32, 0, s>>>(d->d_flag_table, d->rank, d->world, d->epoch);
  QOB_LAUNCHED();
  QOB_CUDA(cudaGetLastError());
  return QOB_STATUS_OK;
}

static int env_dist(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

extern "C" {

// ---- memory that other processes can map: plain cudaMalloc (CUDA IPC cannot export sub-allocations of a caching allocator)
int qob_dist_alloc(qob_ctx *ctx, int64_t bytes, void **ptr) {
  if (!ctx || !ptr || bytes <= 0) QOB_FAIL(QOB_STATUS_INVALID_ARG, "bad allocation request");
  if (ctx->device < 0) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "planning-only context (no CUDA device): libqob200 has no CPU fallback");
  int prev = -1;
  cudaGetDevice(&prev);
  QOB_CUDA(cudaSetDevice(ctx->device));
  const cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (prev >= 0) cudaSetDevice(prev);
  if (e != cudaSuccess) QOB_FAIL(QOB_STATUS_ALLOC, "cudaMalloc(%lld bytes) failed: %s", (long long)bytes, cudaGetErrorString(e));
  return QOB_STATUS_OK;
}
int qob_dist_free(qob_ctx *ctx, void *ptr) {
  (void)ctx;
  if (ptr) QOB_CUDA(cudaFree(ptr));
  return QOB_STATUS_OK;
}
int qob_ipc_export(void *ptr, uint8_t *handle64) {
  if (!ptr || !handle64) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) QOB_FAIL(QOB_STATUS_NCCL_ERROR, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  memcpy(handle64, &h, 64);
  return QOB_STATUS_OK;
}
int qob_ipc_open(qob_ctx *ctx, const uint8_t *handle64, void **ptr) {
  if (!ctx || !handle64 || !ptr) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  if (ctx->device < 0) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "planning-only context (no CUDA device): libqob200 has no CPU fallback");
  int prev = -1;
  cudaGetDevice(&prev);
  QOB_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  const cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (prev >= 0) cudaSetDevice(prev);
  if (e != cudaSuccess) QOB_FAIL(QOB_STATUS_NCCL_ERROR, "cudaIpcOpenMemHandle failed (no peer access between the devices?): %s", cudaGetErrorString(e));
  return QOB_STATUS_OK;
}
int qob_ipc_close(qob_ctx *ctx, void *ptr) {
  (void)ctx;
  if (ptr) {
    const cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) QOB_FAIL(QOB_STATUS_NCCL_ERROR, "cudaIpcCloseMemHandle failed: %s", cudaGetErrorString(e));
  }
  return QOB_STATUS_OK;
}

// ---- planning: which terms need the exchange, the swap window, the chunking of exchange and fold-in
int qob_dist_create(qob_op *sum, int32_t rank, int32_t world, qob_dist **out) {
  if (!sum || !out) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  if (world < 1 || (world & (world - 1)) || rank < 0 || rank >= world) QOB_FAIL(QOB_STATUS_INVALID_ARG, "world must be a power of two, 0 <= rank < world");
  int64_t dl = 0, dr = 0;
  QOB_TRY(qob_op_dims(sum, &dl, &dr));
  uint64_t od = 0, al = 0;
  QOB_TRY(qob_lazysum_term_masks(sum, 0, &od, &al));   // fails unless `sum` is a LazySum of LazyTensors on spin-1/2 subsystems
  auto d = std::make_unique<qob_dist>();
  d->sum = sum;
  d->ctx = qob_op_context(sum);
  d->rank = rank;
  d->world = world;
  int n = 0;
  while (((int64_t)1 << n) < dl) ++n;
  if (((int64_t)1 << n) != dl || dl != dr) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "sharded apply needs a square operator on 2-dimensional subsystems");
  d->n = n;
  while ((1 << d->p) < world) ++d->p;
  d->nloc = n - d->p;
  if (d->nloc < 10) QOB_FAIL(QOB_STATUS_INVALID_ARG, "at least 2^10 amplitudes per rank");
  const uint64_t lowmask = (1ull << d->nloc) - 1;
  const int split_bit = env_dist("QOB_DIST_SPLIT_BIT", d->nloc - 6);
  // classify the terms: R = off-diagonal on a sharded bit; A = local, below split_bit (runs beside the exchange); B = the rest
  std::vector<uint8_t> selA, selB, selR;
  uint64_t touched = 0;
  int cA = 0, cB = 0, cR = 0;
  for (int i = 0;; ++i) {
    const int st = qob_lazysum_term_masks(sum, i, &od, &al);
    if (st != QOB_STATUS_OK) break;   // past the last term
    selA.push_back(0);
    selB.push_back(0);
    selR.push_back(0);
    if (od & ~lowmask) {
      selR[i] = 1;
      ++cR;
      touched |= al;
    } else if (od == 0 || (od >> split_bit) == 0) {
      selA[i] = 1;
      ++cA;
    } else {
      selB[i] = 1;
      ++cB;
    }
  }
  qob_set_error("");
  d->n_remote = cR;
  d->n_local = cA + cB;
  if (!cR) {  // nothing to exchange: one local plan
    for (size_t i = 0; i < selA.size(); ++i)
      if (selB[i]) selA[i] = 1, selB[i] = 0;
    cA += cB;
    cB = 0;
  }
  std::vector<int32_t> ident(n), swapped(n);
  for (int k = 0; k < n; ++k) ident[k] = swapped[k] = k;
  QOB_TRY(qob_layout_plan_create(sum, d->nloc, ident.data(), (uint64_t)rank, selA.data(), &d->plan_a));
  if (cB) QOB_TRY(qob_layout_plan_create(sum, d->nloc, ident.data(), (uint64_t)rank, selB.data(), &d->plan_b));
  if (cR) {
    // highest window of p local bits that no exchanged term touches: the sharded bits trade places with it
    int s = -1;
    for (int c = d->nloc - d->p; c >= 0 && s < 0; --c) {
      const uint64_t win = ((1ull << d->p) - 1) << c;
      if (!(touched & lowmask & win)) s = c;
    }
    if (s < 0) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "no free window of %d local bits for the axis swap", d->p);
    d->swap_lo = s;
    for (int k = d->nloc; k < n; ++k) swapped[k] = s + (k - d->nloc);
    for (int k = s; k < s + d->p; ++k) swapped[k] = d->nloc + (k - s);
    QOB_TRY(qob_layout_plan_create(sum, d->nloc, swapped.data(), (uint64_t)rank, selR.data(), &d->plan_r));
  }
  // pipeline the fold-in behind the exchange: chunk both single-pass plans on their top common fixed bits
  d->nchunks = 1;
  const int want = env_dist("QOB_DIST_CHUNKS", 4);
  if (d->plan_r >= 0 && d->plan_b >= 0 && want > 1) {
    int32_t ns = 0, nb = 0;
    uint64_t fs = 0, fb = 0;
    QOB_TRY(qob_layout_plan_info(sum, d->plan_r, &ns, &fs));
    QOB_TRY(qob_layout_plan_info(sum, d->plan_b, &nb, &fb));
    if (ns == 1 && nb == 1) {
      const uint64_t window = ((1ull << d->p) - 1) << d->swap_lo;
      const uint64_t common = fs & fb & lowmask & ~window;
      int nbits = 0, target = 0;
      while ((1 << (target + 1)) <= want) ++target;
      uint64_t mask = 0;
      for (int b = d->nloc - 1; b >= 0 && nbits < std::max(1, target); --b)
        if (common >> b & 1) {
          mask |= 1ull << b;
          ++nbits;
        }
      if (nbits >= 1) {
        QOB_TRY(qob_layout_plan_set_chunk_bits(sum, d->plan_r, mask));
        QOB_TRY(qob_layout_plan_set_chunk_bits(sum, d->plan_b, mask));
        d->nchunks = 1 << nbits;
        d->chunk_mask = mask;
      }
    }
  }
  d->swap_sms = env_dist("QOB_DIST_SWAP_SMS", 32);
  if (cR) {
    d->direct_ok = layout_plan_peer_qreg(sum, d->plan_r, world, d->swap_lo);
  }
  char buf[8192];
  std::string t = "dist[rank " + std::to_string(rank) + "/" + std::to_string(world) + ", 2^" + std::to_string(d->nloc) + " amplitudes per rank, " +
                  std::to_string(d->n_local) + " local + " + std::to_string(d->n_remote) + " exchanged terms, chunks=" + std::to_string(d->nchunks) + "]";
  if (qob_layout_plan_describe(sum, d->plan_a, buf, sizeof buf) == QOB_STATUS_OK) t += std::string(" A: ") + buf;
  if (d->plan_b >= 0 && qob_layout_plan_describe(sum, d->plan_b, buf, sizeof buf) == QOB_STATUS_OK) t += std::string(" | B: ") + buf;
  if (d->plan_r >= 0 && qob_layout_plan_describe(sum, d->plan_r, buf, sizeof buf) == QOB_STATUS_OK)
    t += " | exchanged (window bit " + std::to_string(d->swap_lo) + "): " + buf;
  if (d->direct_ok) t += " | direct mode available (the exchange adds into the owners' result slabs)";
  d->text = t;
  *out = d.release();
  return QOB_STATUS_OK;
}

int qob_dist_info(qob_dist *d, int32_t *nloc, int32_t *n_remote, int32_t *nchunks, int64_t *slab_bytes, int64_t *flag_bytes) {
  if (!d) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null handle");
  if (nloc) *nloc = d->nloc;
  if (n_remote) *n_remote = d->n_remote;
  if (nchunks) *nchunks = d->nchunks;
  if (slab_bytes) *slab_bytes = (int64_t)16 << d->nloc;
  if (flag_bytes) *flag_bytes = (int64_t)8 * std::max(d->world, 32);
  return QOB_STATUS_OK;
}

int qob_dist_describe(qob_dist *d, char *buf, int64_t buflen) {
  if (!d || !buf) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  snprintf(buf, (size_t)buflen, "%s", d->text.c_str());
  return QOB_STATUS_OK;
}

// x_peers / z_peers / flag_peers: `world` device pointers each, entry q = rank q's buffer as mapped into THIS process
// (entry `rank` = this rank's own allocation).  x: the state slab (2^nloc ComplexF64), z: the contribution slab (same size),
// flags: `flag_bytes` bytes, zero-initialised before the first apply on every rank.
int qob_dist_bind(qob_dist *d, void *const *x_peers, void *const *z_peers, void *const *flag_peers) {
  if (!d || !x_peers) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  d->x_peers.assign(x_peers, x_peers + d->world);
  if (d->n_remote) {
    if (!flag_peers) QOB_FAIL(QOB_STATUS_INVALID_ARG, "exchanged terms need signal pads");
    if (!z_peers && !d->direct_ok)
      QOB_FAIL(QOB_STATUS_INVALID_ARG, "exchanged terms need contribution slabs (this plan cannot add into the result slabs directly)");
    if (z_peers) d->z_peers.assign(z_peers, z_peers + d->world);
    else d->z_peers.clear();
    d->flag_peers.assign(flag_peers, flag_peers + d->world);
    for (int q = 0; q < d->world; ++q)
      if (!d->x_peers[q] || (z_peers && !d->z_peers[q]) || !d->flag_peers[q]) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null peer pointer %d", q);
    if (!d->ctx || d->ctx->device < 0) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "planning-only context (no CUDA device): libqob200 has no CPU fallback");
    int prev = -1;
    cudaGetDevice(&prev);
    QOB_CUDA(cudaSetDevice(d->ctx->device));
    struct Restore {
      int p;
      ~Restore() {
        if (p >= 0) cudaSetDevice(p);
      }
    } restore{prev};
    if (!d->d_flag_table) QOB_CUDA(cudaMalloc(&d->d_flag_table, sizeof(void *) * d->world));
    QOB_CUDA(cudaMemcpy(d->d_flag_table, d->flag_peers.data(), sizeof(void *) * d->world, cudaMemcpyHostToDevice));
    if (!d->side) {
      int lo = 0, hi = 0;
      QOB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      QOB_CUDA(cudaStreamCreateWithPriority(&d->side, cudaStreamNonBlocking, hi));   // the exchange kernel's CTAs get free SM slots first
      QOB_CUDA(cudaEventCreateWithFlags(&d->ev_main, cudaEventDisableTiming));
      QOB_CUDA(cudaEventCreateWithFlags(&d->ev_done, cudaEventDisableTiming));
      d->ev_chunk.resize(d->nchunks);
      for (auto &e : d->ev_chunk) QOB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
  }
  d->bound = true;
  return QOB_STATUS_OK;
}

// Direct mode: the result slab of every rank as mapped into this process.  An apply whose `y` is this rank's entry then needs no
// contribution slab: the exchange pass adds into the owners' results.  yes = 0 from qob_dist_direct_capable: not available for
// this plan (slabs below 2^20 amplitudes, scattered pieces); bind contribution slabs instead.
int qob_dist_direct_capable(qob_dist *d, int32_t *yes) {
  if (!d || !yes) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  *yes = (d->direct_ok && env_dist("QOB_DIST_DIRECT", 1) != 0) ? 1 : 0;
  return QOB_STATUS_OK;
}
int qob_dist_bind_result(qob_dist *d, void *const *y_peers) {
  if (!d || !y_peers) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  if (!d->direct_ok) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "this plan cannot add into the result slabs directly (see qob_dist_direct_capable)");
  for (int q = 0; q < d->world; ++q)
    if (!y_peers[q]) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null peer pointer %d", q);
  d->y_peers.assign(y_peers, y_peers + d->world);
  return QOB_STATUS_OK;
}

// y_local = alpha * (H x)_local + beta * y_local with x = the bound slab of this rank.  Collective: every rank calls it, in the
// same order; the ranks meet in device-side barriers, the host never blocks.
int qob_dist_apply(qob_dist *d, qob_c64 alpha, qob_c64 beta, void *y, void *stream) {
  if (!d || !y) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  if (!d->bound) QOB_FAIL(QOB_STATUS_INVALID_ARG, "qob_dist_bind has not been called");
  cudaStream_t main = (cudaStream_t)stream;
  if (!d->ctx || d->ctx->device < 0) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "planning-only context (no CUDA device): libqob200 has no CPU fallback");
  struct Dev {   // the barrier kernels are launched from here: switch to the context's device, restore the caller's afterwards
    int prev = -1;
    explicit Dev(int dev) {
      cudaGetDevice(&prev);
      if (prev != dev) cudaSetDevice(dev);
      else prev = -1;
    }
    ~Dev() {
      if (prev >= 0) cudaSetDevice(prev);
    }
  } dev_guard(d->ctx->device);
  const void *x = d->x_peers[d->rank];
  const qob_c64 one = {1.0, 0.0}, zero = {0.0, 0.0};
  const bool alpha_zero = alpha.re == 0.0 && alpha.im == 0.0;
  if (d->plan_r < 0 || alpha_zero) {
    QOB_TRY(qob_layout_plan_apply(d->sum, d->plan_a, alpha, x, beta, y, stream));
    if (d->plan_b >= 0 && !alpha_zero) QOB_TRY(qob_layout_plan_apply(d->sum, d->plan_b, alpha, x, one, y, stream));
    return QOB_STATUS_OK;
  }
  int sms = qob_device_sm_count();
  const int k = std::max(4, std::min(d->swap_sms, sms / 2));
  auto time_begin = [&]() -> int {
    if (!d->timing) return QOB_STATUS_OK;
    cudaEvent_t a, b;
    QOB_CUDA(cudaEventCreate(&a));
    QOB_CUDA(cudaEventCreate(&b));
    d->timed.emplace_back(a, b);
    QOB_CUDA(cudaEventRecord(a, d->side));
    return QOB_STATUS_OK;
  };
  auto time_end = [&]() -> int {
    if (d->timing && !d->timed.empty()) QOB_CUDA(cudaEventRecord(d->timed.back().second, d->side));
    return QOB_STATUS_OK;
  };
  if (d->direct_ok && !d->y_peers.empty() && y == d->y_peers[d->rank] && env_dist("QOB_DIST_DIRECT", 1) != 0) {
    // direct mode: prepare y, then every pass (local and exchanged, on any rank) adds into it
    const int64_t namp = (int64_t)1 << d->nloc;
    // QOB_DIST_TRACE=1: phase times of this apply on stderr (synchronises at the end of the call; debugging only)
    const bool trace = env_dist("QOB_DIST_TRACE", 0) != 0;
    cudaEvent_t te[7] = {};
    auto mark = [&](int i, cudaStream_t s) {
      if (!trace) return;
      cudaEventCreate(&te[i]);
      cudaEventRecord(te[i], s);
    };
    mark(0, main);
    if (beta.re == 0.0 && beta.im == 0.0) QOB_CUDA(cudaMemsetAsync(y, 0, (size_t)namp * 16, main));
    else if (!(beta.re == 1.0 && beta.im == 0.0)) QOB_TRY(launch_scale(y, namp, cplx(beta.re, beta.im), main));
    QOB_CUDA(cudaEventRecord(d->ev_main, main));
    mark(1, main);
    QOB_CUDA(cudaStreamWaitEvent(d->side, d->ev_main, 0));
    QOB_TRY(dist_barrier(d, d->side));                               // every rank: x ready, y prepared
    QOB_TRY(time_begin());
    mark(3, d->side);
    QOB_TRY(qob_layout_plan_apply_ex(d->sum, d->plan_r, alpha, nullptr, one, nullptr, nullptr, d->world, (const void *const *)d->x_peers.data(),
                                     d->y_peers.data(), d->swap_lo, k, 0, 1, d->side));
    mark(4, d->side);
    QOB_TRY(dist_barrier(d, d->side));                               // every contribution has landed; nobody reads this x any more
    QOB_TRY(time_end());
    mark(5, d->side);
    QOB_CUDA(cudaEventRecord(d->ev_done, d->side));
    // both groups of communication-free terms, on the SMs the exchange leaves free, adding into y
    QOB_TRY(qob_layout_plan_apply_ex(d->sum, d->plan_a, alpha, x, one, y, nullptr, 0, nullptr, nullptr, 0, -k, 0, 1, main));
    if (d->plan_b >= 0) QOB_TRY(qob_layout_plan_apply_ex(d->sum, d->plan_b, alpha, x, one, y, nullptr, 0, nullptr, nullptr, 0, -k, 0, 1, main));
    mark(2, main);
    QOB_CUDA(cudaStreamWaitEvent(main, d->ev_done, 0));
    mark(6, main);
    if (trace) {
      cudaStreamSynchronize(main);
      cudaStreamSynchronize(d->side);
      float t[7] = {};
      for (int i = 1; i < 7; ++i) cudaEventElapsedTime(&t[i], te[0], te[i]);
      fprintf(stderr, "[qob_dist trace] rank %d: y prepared %.2f | local passes done %.2f | exchange: barrier passed %.2f, kernel done %.2f, "
                      "all ranks done %.2f | apply done %.2f ms\n", d->rank, t[1], t[2], t[3], t[4], t[5], t[6]);
      for (auto &e : te) cudaEventDestroy(e);
    }
    return QOB_STATUS_OK;
  }
  if (d->z_peers.empty()) QOB_FAIL(QOB_STATUS_INVALID_ARG, "no contribution slabs bound and y is not the bound result slab");
  void *z = d->z_peers[d->rank];
  if (d->plan_b >= 0) {
    const int nc = d->nchunks;
    QOB_CUDA(cudaEventRecord(d->ev_main, main));
    QOB_CUDA(cudaStreamWaitEvent(d->side, d->ev_main, 0));            // x is ready on this rank
    QOB_TRY(dist_barrier(d, d->side));                               // ... and on every rank; last apply's contributions are consumed
    QOB_TRY(time_begin());
    for (int c = 0; c < nc; ++c) {
      QOB_TRY(qob_layout_plan_apply_ex(d->sum, d->plan_r, alpha, nullptr, zero, nullptr, nullptr, d->world, (const void *const *)d->x_peers.data(),
                                       d->z_peers.data(), d->swap_lo, k, c, nc, d->side));
      QOB_TRY(dist_barrier(d, d->side));                             // chunk c of every rank's contributions has landed
      QOB_CUDA(cudaEventRecord(d->ev_chunk[c], d->side));
    }
    QOB_TRY(time_end());
    // beside the exchange: the communication-free group A on the SMs the exchange leaves free (sm_budget -k)
    QOB_TRY(qob_layout_plan_apply_ex(d->sum, d->plan_a, alpha, x, beta, y, nullptr, 0, nullptr, nullptr, 0, -k, 0, 1, main));
    for (int c = 0; c < nc; ++c) {                                   // fold the contributions in, chunk by chunk, behind the exchange
      QOB_CUDA(cudaStreamWaitEvent(main, d->ev_chunk[c], 0));
      QOB_TRY(qob_layout_plan_apply_ex(d->sum, d->plan_b, alpha, x, one, y, z, 0, nullptr, nullptr, 0, 0, c, nc, main));
    }
  } else {
    QOB_TRY(dist_barrier(d, main));
    QOB_TRY(qob_layout_plan_apply_ex(d->sum, d->plan_r, alpha, nullptr, zero, nullptr, nullptr, d->world, (const void *const *)d->x_peers.data(),
                                     d->z_peers.data(), d->swap_lo, 0, 0, 1, main));
    QOB_TRY(dist_barrier(d, main));
    QOB_TRY(qob_layout_plan_apply_ex(d->sum, d->plan_a, alpha, x, beta, y, z, 0, nullptr, nullptr, 0, 0, 0, 1, main));
  }
  return QOB_STATUS_OK;
}

// CUDA-event timing of the exchange (first barrier passed -> last contributions landed on every rank) of the applies issued
// while it is enabled; qob_dist_exchange_ms synchronises the exchange stream and returns the mean over those applies
int qob_dist_exchange_timing(qob_dist *d, int32_t enable) {
  if (!d) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null handle");
  d->timing = enable != 0;
  return QOB_STATUS_OK;
}
int qob_dist_exchange_ms(qob_dist *d, double *mean_ms, int32_t *count, int64_t *bytes_per_direction) {
  if (!d) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null handle");
  double sum = 0.0;
  int n = 0;
  if (d->side) QOB_CUDA(cudaStreamSynchronize(d->side));
  for (auto &ev : d->timed) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) sum += ms, ++n;
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  cudaGetLastError();
  d->timed.clear();
  if (mean_ms) *mean_ms = n ? sum / n : 0.0;
  if (count) *count = n;
  // per apply and direction on this GPU's links: the (1 - 1/P) share of its swapped-layout x pieces comes in from the peers and the
  // same share of the peers' results comes in too (and the mirror image goes out)
  if (bytes_per_direction) *bytes_per_direction = (int64_t)(2.0 * (1.0 - 1.0 / d->world) * 16.0 * (double)((int64_t)1 << d->nloc));
  return QOB_STATUS_OK;
}

int qob_dist_destroy(qob_dist *d) {
  if (!d) return QOB_STATUS_OK;
  if (d->side) {
    cudaStreamSynchronize(d->side);
    cudaStreamDestroy(d->side);
  }
  if (d->ev_main) cudaEventDestroy(d->ev_main);
  if (d->ev_done) cudaEventDestroy(d->ev_done);
  for (auto e : d->ev_chunk)
    if (e) cudaEventDestroy(e);
  if (d->d_flag_table) cudaFree(d->d_flag_table);
  delete d;
  return QOB_STATUS_OK;
}

}  // extern "C"
