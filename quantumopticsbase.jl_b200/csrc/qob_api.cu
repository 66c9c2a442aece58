// qob_api.cu — the C ABI of libqob200.so (include/qob200.h) and the host-side engine behind it:
// operator handles that mirror the reference's lazy-operator types, the planner that maps a
// LazySum of LazyTensors onto fused device programs, and scratch management.
//
// Reference call stacks replaced: mul! for LazySum (src/operators_lazysum.jl:189-238),
// LazyTensor (src/operators_lazytensor.jl:539-609), LazyProduct (src/operators_lazyproduct.jl:103-163),
// SparseOperator (src/operators_sparse.jl:199-202) and dense Operator (src/operators_dense.jl:394-396).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "qob_internal.h"

std::atomic<int64_t> g_launch_count{0};
std::atomic<int64_t> g_family_count[8];
#define QT_MAXPEER_API 16
thread_local bool t_planning_only = false;
struct PlanningScope {
  bool prev;
  explicit PlanningScope(bool on) : prev(t_planning_only) { t_planning_only = on; }
  ~PlanningScope() { t_planning_only = prev; }
};

int qob_device_sm_count() {
  static std::atomic<int> cache[QOB_MAX_DEVICES];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  const bool cached = dev >= 0 && dev < QOB_MAX_DEVICES;
  int n = cached ? cache[dev].load(std::memory_order_relaxed) : 0;
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (cached) cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

static thread_local char t_err[1024] = "";
void qob_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof t_err, fmt, ap);
  va_end(ap);
}

static inline cplx C(qob_c64 z) { return cplx(z.re, z.im); }
static const cplx ZERO(0.0, 0.0), ONE(1.0, 0.0);

// ============================================================================ HostMat
cplx HostMat::at(int64_t i, int64_t j) const {
  if (kind == QOB_FACTOR_DENSE) return dense[i + j * rows];
  if (kind == QOB_FACTOR_EYE) return i == j ? ONE : ZERO;
  cplx v = 0.0;  // duplicates (legal in CSC) add up, as every reference loop treats them
  for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p)
    if (rowidx[p] == i) v += vals[p];
  return v;
}

HostMat HostMat::transposed() const {
  HostMat t;
  t.kind = kind;
  t.rows = cols;
  t.cols = rows;
  if (kind == QOB_FACTOR_DENSE) {
    t.dense.resize(dense.size());
    for (int64_t i = 0; i < rows; ++i)
      for (int64_t j = 0; j < cols; ++j) t.dense[j + i * cols] = dense[i + j * rows];
  } else if (kind == QOB_FACTOR_CSC) {
    t.colptr.assign(rows + 1, 0);
    for (int64_t r : rowidx) t.colptr[r + 1]++;
    for (int64_t r = 0; r < rows; ++r) t.colptr[r + 1] += t.colptr[r];
    t.rowidx.resize(rowidx.size());
    t.vals.resize(vals.size());
    std::vector<int64_t> fill(t.colptr.begin(), t.colptr.end() - 1);
    for (int64_t j = 0; j < cols; ++j)
      for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
        int64_t q = fill[rowidx[p]]++;
        t.rowidx[q] = j;
        t.vals[q] = vals[p];
      }
  }
  return t;
}

void HostMat::to_csr(std::vector<int32_t> &rp, std::vector<int32_t> &ci, std::vector<cplx> &v) const {
  rp.assign(rows + 1, 0);
  ci.clear();
  v.clear();
  if (kind == QOB_FACTOR_DENSE) {
    for (int64_t i = 0; i < rows; ++i) {
      for (int64_t j = 0; j < cols; ++j) {
        ci.push_back((int32_t)j);
        v.push_back(dense[i + j * rows]);
      }
      rp[i + 1] = (int32_t)ci.size();
    }
  } else if (kind == QOB_FACTOR_EYE) {
    for (int64_t i = 0; i < rows; ++i) {
      if (i < cols) {
        ci.push_back((int32_t)i);
        v.push_back(ONE);
      }
      rp[i + 1] = (int32_t)ci.size();
    }
  } else {
    HostMat t = transposed();  // CSC of the transpose == CSR of this
    for (int64_t i = 0; i < rows; ++i) rp[i + 1] = (int32_t)t.colptr[i + 1];
    for (size_t p = 0; p < t.rowidx.size(); ++p) {
      ci.push_back((int32_t)t.rowidx[p]);
      v.push_back(t.vals[p]);
    }
  }
}

int64_t HostMat::max_row_nnz() const {
  if (kind == QOB_FACTOR_DENSE) return cols;
  if (kind == QOB_FACTOR_EYE) return 1;
  std::vector<int64_t> cnt(rows, 0);
  for (int64_t r : rowidx) cnt[r]++;
  int64_t m = 0;
  for (int64_t c : cnt) m = std::max(m, c);
  return m;
}

int hostmat_from_factor(const qob_factor *f, HostMat &out) {
  if (!f) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null factor");
  if (f->nrows < 0 || f->ncols < 0) QOB_FAIL(QOB_STATUS_INVALID_ARG, "negative factor shape");
  if (f->trans < QOB_OP_N || f->trans > QOB_OP_C) QOB_FAIL(QOB_STATUS_INVALID_ARG, "bad trans flag %d", f->trans);
  HostMat m;
  m.kind = f->kind;
  m.rows = f->nrows;
  m.cols = f->ncols;
  if (f->kind == QOB_FACTOR_DENSE) {
    if (!f->dense && f->nrows * f->ncols > 0) QOB_FAIL(QOB_STATUS_INVALID_ARG, "dense factor without data");
    m.dense.resize((size_t)(f->nrows * f->ncols));
    for (size_t i = 0; i < m.dense.size(); ++i) m.dense[i] = C(f->dense[i]);
  } else if (f->kind == QOB_FACTOR_CSC) {
    if (!f->colptr) QOB_FAIL(QOB_STATUS_INVALID_ARG, "CSC factor without colptr");
    m.colptr.resize(f->ncols + 1);
    for (int64_t j = 0; j <= f->ncols; ++j) m.colptr[j] = f->colptr[j] - 1;
    int64_t nnz = m.colptr[f->ncols];
    if (m.colptr[0] != 0 || nnz < 0) QOB_FAIL(QOB_STATUS_INVALID_ARG, "CSC colptr must be 1-based and monotone");
    if (nnz > 0 && (!f->rowval || !f->nzval)) QOB_FAIL(QOB_STATUS_INVALID_ARG, "CSC factor with %lld stored entries but no rowval / nzval", (long long)nnz);
    m.rowidx.resize(nnz);
    m.vals.resize(nnz);
    for (int64_t j = 0; j < f->ncols; ++j)
      if (m.colptr[j + 1] < m.colptr[j]) QOB_FAIL(QOB_STATUS_INVALID_ARG, "CSC colptr not monotone");
    for (int64_t p = 0; p < nnz; ++p) {
      m.rowidx[p] = f->rowval[p] - 1;
      if (m.rowidx[p] < 0 || m.rowidx[p] >= f->nrows) QOB_FAIL(QOB_STATUS_INVALID_ARG, "CSC rowval out of range");
      m.vals[p] = C(f->nzval[p]);
    }
  } else if (f->kind != QOB_FACTOR_EYE) {
    // the reference throws MethodError / ArgumentError for factor types it has no kernel for
    // (src/operators_lazytensor.jl:639-641, test/test_operators_lazytensor.jl:409-415)
    QOB_FAIL(QOB_STATUS_UNSUPPORTED, "unsupported factor kind %d", f->kind);
  }
  if (f->trans != QOB_OP_N) {
    m = m.transposed();
    if (f->trans == QOB_OP_C) {
      for (cplx &z : m.dense) z = std::conj(z);
      for (cplx &z : m.vals) z = std::conj(z);
    }
  }
  out = std::move(m);
  return QOB_STATUS_OK;
}

// ============================================================================ context
int qob_ctx::get_scratch(cudaStream_t s, int slot, size_t bytes, void **out) {
  std::lock_guard<std::mutex> lk(mu);
  DevBuf &b = scratch[std::make_pair(s, slot)];
  if (b.bytes < bytes) {
    if (b.ptr) {
      cudaStreamSynchronize(s);
      cudaFree(b.ptr);
      b.ptr = nullptr;
      b.bytes = 0;
    }
    QOB_CUDA(cudaMalloc(&b.ptr, bytes));
    b.bytes = bytes;
  }
  *out = b.ptr;
  return QOB_STATUS_OK;
}
int64_t qob_ctx::scratch_bytes() {
  std::lock_guard<std::mutex> lk(mu);
  int64_t t = 0;
  for (auto &kv : scratch) t += (int64_t)kv.second.bytes;
  return t;
}
void qob_ctx::clear_scratch() {
  std::lock_guard<std::mutex> lk(mu);
  if (device >= 0) cudaDeviceSynchronize();
  for (auto &kv : scratch)
    if (kv.second.ptr) cudaFree(kv.second.ptr);
  scratch.clear();
}

static std::atomic<int> g_slot_counter{1000};

// ============================================================================ operator handles
enum OpKind { OP_LAZYTENSOR, OP_SPARSE, OP_DENSE, OP_LAZYSUM, OP_LAZYPRODUCT, OP_LINDBLAD, OP_DIRECTSUM };

struct qob_op {
  qob_ctx *ctx;
  OpKind kind;
  int64_t dl = 0, dr = 0;
  std::atomic<int> refs{1};
  int slot_base;  // scratch slots [slot_base, slot_base+8) belong to this handle
  qob_op(qob_ctx *c, OpKind k) : ctx(c), kind(k), slot_base(g_slot_counter.fetch_add(8)) {
    if (ctx) ctx->live_ops.fetch_add(1);
  }
  virtual ~qob_op() {
    if (!ctx) return;
    ctx->live_ops.fetch_sub(1);
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (auto it = ctx->scratch.begin(); it != ctx->scratch.end();) {
      if (it->first.second >= slot_base && it->first.second < slot_base + 8) {
        if (it->second.ptr) {
          cudaStreamSynchronize(it->first.first);
          cudaFree(it->second.ptr);
        }
        it = ctx->scratch.erase(it);
      } else {
        ++it;
      }
    }
  }
  virtual int apply(int side, cplx alpha, const void *x, cplx beta, void *y, int64_t batch, cudaStream_t s) = 0;
  virtual std::string describe(int side, int64_t batch) = 0;
};
// Switch to the context's device for the duration of an entry point and restore the caller's current device afterwards.
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) {
      err = cudaSetDevice(device);
      changed = err == cudaSuccess;
    }
  }
  ~DeviceGuard() {
    if (changed && prev >= 0) cudaSetDevice(prev);
  }
};
#define QOB_DEVICE(dev)                                                                                       \
  DeviceGuard device_guard__(dev);                                                                            \
  if (device_guard__.err != cudaSuccess)                                                                      \
  QOB_FAIL(QOB_STATUS_CUDA_ERROR, "cudaSetDevice(%d) failed: %s", (int)(dev), cudaGetErrorString(device_guard__.err))

qob_ctx *qob_op_context(const qob_op *op) { return op ? op->ctx : nullptr; }
static void op_retain(qob_op *o) { o->refs.fetch_add(1); }
static void op_release(qob_op *o) {
  if (o->refs.fetch_sub(1) == 1) delete o;
}

static int env_i(const char *n, int d) {
  const char *v = getenv(n);
  return v && *v ? atoi(v) : d;
}

static bool is_pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }
static int ilog2(int64_t v) {
  int l = 0;
  while ((1ll << l) < v) ++l;
  return l;
}

// ---------------------------------------------------------------------------- tensor groups
// A set of LazyTensor terms on one pair of composite bases, compiled per (side, batch) into pieces.
struct GroupTerm {
  int coef_index;  // -1: no external coefficient
  cplx scalar;     // LazyTensor.factor
  std::vector<int> sites;
  std::vector<std::shared_ptr<HostMat>> mats;  // left orientation (rows: basis_l, cols: basis_r)
};

struct SeqStep {
  int axis;
  bool dense_axis;
  AxisMatrixDev amat;                 // dense DMMA step
  std::unique_ptr<GatherProgram> gp;  // small / sparse step
  std::vector<int64_t> dims_after;
};
struct SeqTerm {
  int coef_index;
  cplx scalar;
  std::vector<std::unique_ptr<SeqStep>> steps;
  int64_t max_tmp = 0;  // elements (without batch) of the largest intermediate
};

struct Compiled {
  int side;
  int64_t batch;
  std::vector<int64_t> dims_out, dims_in;
  int64_t d_out = 1, d_in = 1;
  bool has_qtile = false, has_gather = false, has_dtile = false, has_qreg = false;
  QTileProgram qtile;
  QRegProgram qreg;
  DTileProgram dtile;
  GatherProgram gather;
  std::vector<std::unique_ptr<SeqTerm>> seq;
  std::string text;
};

struct TensorGroup {
  qob_ctx *ctx;
  int slot_base;
  std::vector<int64_t> dims_l, dims_r;
  std::vector<GroupTerm> terms;
  std::vector<std::unique_ptr<Compiled>> cache;
  std::vector<cplx> last_coefs;  // coefficients the cached programs were last loaded with
  bool coefs_dirty = true;
  std::recursive_mutex mu;       // guards the compiled-program cache, the coefficient upload AND the launches that read the
                                 // tables: handles may be applied from several host threads and streams (the reference is
                                 // reentrant per task, operators_lazytensor.jl:233).  The weight tables are one device buffer
                                 // per program, so an upload has to be ordered against kernels on OTHER streams too:
  cudaEvent_t ev_upload = nullptr;                 // recorded after the last table upload, on `upload_stream`
  cudaStream_t upload_stream = nullptr;
  bool uploaded_once = false;
  std::map<cudaStream_t, cudaEvent_t> ev_use;      // per stream: recorded after its last launch that read the tables
  ~TensorGroup() {
    if (ev_upload) cudaEventDestroy(ev_upload);
    for (auto &kv : ev_use)
      if (kv.second) cudaEventDestroy(kv.second);
  }

  int compile(int side, int64_t batch, Compiled **out);
  // expect_out != nullptr: compute <x| group |x> instead (fused reduction, nothing written); QOB_STATUS_UNSUPPORTED when the
  // compiled program is not a single round-2 tile program (the caller then applies into scratch and reduces)
  int apply(int side, cplx alpha, const void *x, cplx beta, void *y, int64_t batch, const std::vector<cplx> &coefs,
            cudaStream_t s, cplx *expect_out = nullptr);
};

static const int GATHER_MAXF = 4;

int TensorGroup::compile(int side, int64_t batch, Compiled **out) {
  std::lock_guard<std::recursive_mutex> lk(mu);
  for (auto &c : cache)
    if (c->side == side && c->batch == batch) {
      *out = c.get();
      return QOB_STATUS_OK;
    }
  auto cp = std::make_unique<Compiled>();
  Compiled &c = *cp;
  c.side = side;
  c.batch = batch;
  const int n = (int)dims_l.size();
  c.dims_out = side == QOB_SIDE_LEFT ? dims_l : dims_r;
  c.dims_in = side == QOB_SIDE_LEFT ? dims_r : dims_l;
  for (int k = 0; k < n; ++k) {
    c.d_out *= c.dims_out[k];
    c.d_in *= c.dims_in[k];
  }
  bool all2 = true, all_same = true;
  for (int k = 0; k < n; ++k) {
    all2 &= (c.dims_out[k] == 2 && c.dims_in[k] == 2);
    all_same &= c.dims_out[k] == c.dims_in[k];
  }
  const int dense_min = env_i("QOB_AXIS_MIN_DIM", 16);
  const int qt_min_bits = std::min(env_i("QOB_QTILE_MIN_BITS", 18), env_i("QOB_QREG_MIN_BITS", 20));
  const bool qt_batch_ok = is_pow2(batch);
  const int nbits = n + (qt_batch_ok ? ilog2(batch) : 0);
  const bool qt_ok = all2 && qt_batch_ok && nbits >= qt_min_bits && nbits <= 62 && !env_i("QOB_DISABLE_QTILE", 0);

  std::vector<QTerm> qterms;
  std::vector<OrientedTerm> gterms, q_as_g;
  for (const GroupTerm &t : terms) {
    OrientedTerm o;
    o.coef_index = t.coef_index;
    o.scalar = t.scalar;
    int n_iso = 0;
    if (!all_same)
      for (int k = 0; k < n; ++k)
        if (c.dims_out[k] != c.dims_in[k] && std::find(t.sites.begin(), t.sites.end(), k) == t.sites.end()) ++n_iso;
    bool heavy = false;
    double rowprod = 1.0;
    for (size_t f = 0; f < t.sites.size(); ++f) {
      HostMat m = side == QOB_SIDE_LEFT ? *t.mats[f] : t.mats[f]->transposed();
      if (m.kind == QOB_FACTOR_DENSE && std::max(m.rows, m.cols) >= dense_min) heavy = true;
      rowprod *= (double)std::max<int64_t>(1, m.max_row_nnz());
      o.axes.push_back(t.sites[f]);
      o.mats.push_back(std::move(m));
    }
    if ((int)t.sites.size() + n_iso > GATHER_MAXF || rowprod > 512.0) heavy = true;
    if (heavy) {
      // sequential factor-by-factor application with temporaries (the reference's _tp_sum_matmul!, :443-488)
      auto st = std::make_unique<SeqTerm>();
      st->coef_index = t.coef_index;
      st->scalar = t.scalar;
      std::vector<int64_t> cur = c.dims_in;
      std::vector<std::pair<int, HostMat>> steps;
      for (size_t f = 0; f < o.axes.size(); ++f) steps.push_back({o.axes[f], o.mats[f]});
      if (!all_same)
        for (int k = 0; k < n; ++k)
          if (c.dims_out[k] != c.dims_in[k] && std::find(t.sites.begin(), t.sites.end(), k) == t.sites.end()) {
            HostMat e;
            e.kind = QOB_FACTOR_EYE;
            e.rows = c.dims_out[k];
            e.cols = c.dims_in[k];
            steps.push_back({k, e});
          }
      for (auto &sp : steps) {
        auto step = std::make_unique<SeqStep>();
        step->axis = sp.first;
        const HostMat &m = sp.second;
        if (m.rows != c.dims_out[sp.first] || m.cols != cur[sp.first])
          QOB_FAIL(QOB_STATUS_DIM_MISMATCH, "factor on subsystem %d does not match the bases", sp.first + 1);
        std::vector<int64_t> nxt = cur;
        nxt[sp.first] = m.rows;
        step->dense_axis = (m.kind == QOB_FACTOR_DENSE && std::max(m.rows, m.cols) >= dense_min);
        if (step->dense_axis) {
          QOB_TRY(prepare_axis_matrix(m, step->amat));
        } else {
          step->gp = std::make_unique<GatherProgram>();
          OrientedTerm one;
          one.coef_index = -1;
          one.scalar = ONE;
          one.axes = {sp.first};
          one.mats = {m};
          // dims for this single step: only axis sp.first changes
          std::vector<OrientedTerm> v1;
          v1.push_back(std::move(one));
          QOB_TRY(gather_program_build(*step->gp, nxt, cur, v1));
          QOB_TRY(gather_program_set_coefs(*step->gp, {}, 0));
        }
        step->dims_after = nxt;
        int64_t sz = 1;
        for (int64_t d : nxt) sz *= d;
        st->max_tmp = std::max(st->max_tmp, sz);
        cur = nxt;
        st->steps.push_back(std::move(step));
      }
      c.seq.push_back(std::move(st));
      continue;
    }
    bool q_ok = qt_ok && o.axes.size() <= 3;
    if (q_ok) {
      QTerm q;
      q.coef_index = t.coef_index;
      q.scalar = t.scalar;
      const int shift = side == QOB_SIDE_LEFT ? 0 : ilog2(batch);
      for (size_t f = 0; f < o.axes.size(); ++f) {
        q.bits.push_back(o.axes[f] + shift);
        for (int i = 0; i < 2; ++i)
          for (int j = 0; j < 2; ++j) q.m.push_back(o.mats[f].at(i, j));
      }
      qterms.push_back(std::move(q));
      q_as_g.push_back(std::move(o));
    } else {
      gterms.push_back(std::move(o));
    }
  }
  if (!qterms.empty() && nbits >= env_i("QOB_QREG_MIN_BITS", 20) && !env_i("QOB_DISABLE_QREG", 0)) {
    // large states: register-blocked tile passes staged by TMA, pairs of passes chained through L2 (qob_kernels_qreg.cu)
    const int st = qreg_build(c.qreg, nbits, 0, qterms, ctx->sm_count);
    if (st == QOB_STATUS_OK) c.has_qreg = true;
    else if (st != QOB_STATUS_UNSUPPORTED) return st;
  }
  if (!qterms.empty() && !c.has_qreg) {
    const int st = qtile_build(c.qtile, nbits, 0, qterms, ctx->sm_count);
    if (st == QOB_STATUS_UNSUPPORTED) {
      // the tile planner declined (selector bits too scattered, too many shared-mask lookups, ...): the generic fused
      // kernel computes the same map
      for (auto &o : q_as_g) gterms.push_back(std::move(o));
    } else {
      QOB_TRY(st);
      c.has_qtile = true;
    }
  }
  if (!gterms.empty() && all_same && !env_i("QOB_DISABLE_DTILE", 0) &&
      (double)c.d_out * (double)batch >= (double)env_i("QOB_DTILE_MIN_ELEMS", 1 << 18)) {
    // large state on subsystems of any dimension: mixed-radix tile passes.  The batch is one more tensor axis: the
    // slowest one for op*X, the fastest one for X*op (split in two when it is long, so that a short run of it can be
    // the coalesced low block of every tile).
    std::vector<int64_t> dd;
    int shift = 0;
    bool ok = true;
    if (side == QOB_SIDE_RIGHT && batch > 1) {
      int64_t lo = batch;
      if (batch > 64) {
        lo = 0;
        for (int64_t v = 64; v >= 8 && !lo; --v)
          if (batch % v == 0) lo = v;
        if (!lo) ok = batch <= 256, lo = batch;
      }
      dd.push_back(lo);
      shift = 1;
      if (lo != batch) {
        dd.push_back(batch / lo);
        shift = 2;
      }
    }
    for (int k = 0; k < n; ++k) dd.push_back(c.dims_out[k]);
    if (side == QOB_SIDE_LEFT && batch > 1) {
      // a small system on many columns: a few whole columns per tile (the batch axis carries no factor and may be split)
      int64_t lo = 1;
      for (int64_t v = std::min<int64_t>(batch, 4096 / std::max<int64_t>(1, c.d_out)); v > 1 && lo == 1; --v)
        if (batch % v == 0) lo = v;
      if (lo > 1) dd.push_back(lo);
      if (batch / lo > 1) dd.push_back(batch / lo);
    }
    if (ok) {
      std::vector<OrientedTerm> shifted = gterms;
      for (auto &o : shifted)
        for (int &a : o.axes) a += shift;
      std::vector<int> declined;
      const int st = dtile_build(c.dtile, dd, shifted, &declined);
      if (st == QOB_STATUS_OK) {
        c.has_dtile = true;
        std::vector<OrientedTerm> rest;   // the terms the tile planner left out go through the gather kernel
        for (int t : declined) rest.push_back(std::move(gterms[t]));
        gterms = std::move(rest);
      } else if (st != QOB_STATUS_UNSUPPORTED) {
        return st;
      }
    }
  }
  if (!gterms.empty()) {
    QOB_TRY(gather_program_build(c.gather, c.dims_out, c.dims_in, gterms));
    c.has_gather = true;
  }
  c.text = std::string(side == QOB_SIDE_LEFT ? "left" : "right") + " batch=" + std::to_string(batch) + ":";
  if (c.has_qreg) c.text += " " + c.qreg.describe;
  if (c.has_qtile) c.text += " " + c.qtile.describe;
  if (c.has_dtile) c.text += " " + c.dtile.describe;
  if (c.has_gather) c.text += " " + c.gather.describe;
  if (!c.seq.empty()) {
    c.text += " seq[terms=" + std::to_string(c.seq.size()) + ":";
    for (auto &st : c.seq)
      for (auto &sp : st->steps) c.text += sp->dense_axis ? " dmma@" + std::to_string(sp->axis + 1) : " gather@" + std::to_string(sp->axis + 1);
    c.text += "]";
  }
  coefs_dirty = true;
  *out = cp.get();
  cache.push_back(std::move(cp));
  return QOB_STATUS_OK;
}

int TensorGroup::apply(int side, cplx alpha, const void *x, cplx beta, void *y, int64_t batch,
                       const std::vector<cplx> &coefs, cudaStream_t s, cplx *expect_out) {
  Compiled *cp = nullptr;
  QOB_TRY(compile(side, batch, &cp));
  Compiled &c = *cp;
  if (expect_out && !(c.has_qreg && !c.has_qtile && !c.has_dtile && !c.has_gather && c.seq.empty())) return QOB_STATUS_UNSUPPORTED;
  // The lock is held until every launch of this apply is enqueued: the kernels' variant flags (real weight tables) are
  // read consistently with the tables they were computed from, and the event bookkeeping below is atomic per apply.
  std::unique_lock<std::recursive_mutex> lk(mu);
  bool on_device = !t_planning_only && ctx && ctx->device >= 0;
  if (on_device) {
    // while the stream is being captured into a CUDA graph the cross-stream bookkeeping is off: events recorded outside the
    // capture cannot be waited on inside it, and a replayed graph owns its ordering (tables must be loaded before capture)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) on_device = false;
  }
  if (coefs_dirty || coefs != last_coefs) {
    if (on_device) {
      // kernels still reading the old tables on other streams must finish before the tables are overwritten
      for (auto &kv : ev_use)
        if (kv.first != s && kv.second) QOB_CUDA(cudaStreamWaitEvent(s, kv.second, 0));
    }
    for (auto &cc : cache) {
      if (cc->has_qreg) QOB_TRY(qreg_set_coefs(cc->qreg, coefs, s));
      if (cc->has_qtile) QOB_TRY(qtile_set_coefs(cc->qtile, coefs, s));
      if (cc->has_dtile) QOB_TRY(dtile_set_coefs(cc->dtile, coefs, s));
      if (cc->has_gather) QOB_TRY(gather_program_set_coefs(cc->gather, coefs, s));
    }
    last_coefs = coefs;
    coefs_dirty = false;
    if (on_device) {
      if (!ev_upload) QOB_CUDA(cudaEventCreateWithFlags(&ev_upload, cudaEventDisableTiming));
      QOB_CUDA(cudaEventRecord(ev_upload, s));
      upload_stream = s;
      uploaded_once = true;
    }
  }
  // a launch on another stream than the one that uploaded the tables waits for that upload
  if (on_device && uploaded_once && s != upload_stream) QOB_CUDA(cudaStreamWaitEvent(s, ev_upload, 0));
  struct UseMark {  // record "tables in use" on this stream when the apply is done enqueueing
    TensorGroup *g;
    cudaStream_t s;
    bool on;
    ~UseMark() {
      if (!on) return;
      cudaEvent_t &e = g->ev_use[s];
      if (!e) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      if (e) cudaEventRecord(e, s);
    }
  } use_mark{this, s, on_device};
  const int64_t pre = side == QOB_SIDE_LEFT ? 1 : batch, post = side == QOB_SIDE_LEFT ? batch : 1;
  bool first = true;
  auto beta_now = [&]() {
    cplx b = first ? beta : ONE;
    first = false;
    return b;
  };
  if (expect_out) return qreg_expect(c.qreg, x, expect_out, s);
  if (c.has_qreg) QOB_TRY(qreg_launch(c.qreg, alpha, x, beta_now(), y, s));
  if (c.has_qtile) QOB_TRY(qtile_launch(c.qtile, alpha, x, beta_now(), y, s));
  if (c.has_dtile) QOB_TRY(dtile_launch(c.dtile, alpha, x, beta_now(), y, s));
  if (c.has_gather) QOB_TRY(gather_program_launch(c.gather, pre, post, alpha, x, beta_now(), y, s));
  for (auto &stp : c.seq) {
    SeqTerm &st = *stp;
    cplx a = alpha * st.scalar;
    if (st.coef_index >= 0) a *= coefs[st.coef_index];
    const size_t nsteps = st.steps.size();
    cplx b = beta_now();
    if (nsteps == 0) {  // factor-less LazyTensor with equal bases: scaled identity (:449-454)
      QOB_TRY(launch_axpby(x, y, c.d_out * batch, a, b, s));
      continue;
    }
    void *tmp[2] = {nullptr, nullptr};
    if (nsteps >= 2) QOB_TRY(ctx->get_scratch(s, slot_base + 0, (size_t)st.max_tmp * batch * sizeof(double2), &tmp[0]));
    if (nsteps >= 3) QOB_TRY(ctx->get_scratch(s, slot_base + 1, (size_t)st.max_tmp * batch * sizeof(double2), &tmp[1]));
    const void *src = x;
    std::vector<int64_t> cur = c.dims_in;
    for (size_t i = 0; i < nsteps; ++i) {
      SeqStep &sp = *st.steps[i];
      void *dst = (i + 1 == nsteps) ? y : tmp[i & 1];
      cplx sa = (i == 0) ? a : ONE;
      cplx sb = (i + 1 == nsteps) ? b : ZERO;
      if (sp.dense_axis) {
        int64_t L = pre, R = post;
        for (int k = 0; k < sp.axis; ++k) L *= cur[k];
        for (size_t k = sp.axis + 1; k < cur.size(); ++k) R *= cur[k];
        QOB_TRY(launch_axis_dense(sp.amat, L, R, sa, src, sb, dst, s));
      } else {
        QOB_TRY(gather_program_launch(*sp.gp, pre, post, sa, src, sb, dst, s));
      }
      src = dst;
      cur = sp.dims_after;
    }
  }
  if (first) QOB_TRY(launch_scale(y, c.d_out * batch, beta, s));  // group without any term
  return QOB_STATUS_OK;
}

// ---------------------------------------------------------------------------- concrete ops
static int check_alias(const void *x, int64_t nx, const void *y, int64_t ny) {
  const char *a = (const char *)x, *b = (const char *)y;
  if (a < b + ny * 16 && b < a + nx * 16 && nx > 0 && ny > 0)
    QOB_FAIL(QOB_STATUS_ALIASING, "output matrix must not be aliased with input matrix");
  return QOB_STATUS_OK;
}

struct LazyTensorOp : qob_op {
  TensorGroup group;
  LazyTensorOp(qob_ctx *c) : qob_op(c, OP_LAZYTENSOR) {
    group.ctx = c;
    group.slot_base = slot_base;
  }
  int apply(int side, cplx alpha, const void *x, cplx beta, void *y, int64_t batch, cudaStream_t s) override {
    // alpha == 0 shortcut: only the beta update (src/operators_lazytensor.jl:540)
    if (alpha == ZERO) return launch_scale(y, (side == QOB_SIDE_LEFT ? dl : dr) * batch, beta, s);
    return group.apply(side, alpha, x, beta, y, batch, {}, s);
  }
  std::string describe(int side, int64_t batch) override {
    Compiled *c = nullptr;
    if (group.compile(side, batch, &c) != QOB_STATUS_OK) return std::string("lazytensor: ") + t_err;
    return "lazytensor " + c->text;
  }
};

struct SparseOp : qob_op {
  HostMat m;           // after trans
  SparseDev csr, csc;  // csr: rows of m (left apply); csc: columns of m (right apply)
  SparseOp(qob_ctx *c) : qob_op(c, OP_SPARSE) {}
  int apply(int side, cplx alpha, const void *x, cplx beta, void *y, int64_t batch, cudaStream_t s) override {
    // gemm!/gemv!: dimension checks happen in qob_op_apply; alpha == 0 still leaves only the beta update
    if (side == QOB_SIDE_LEFT) return launch_spmm_left(csr, dl, dr, batch, alpha, x, beta, y, s);
    return launch_spmm_right(csc, batch, dl, dr, alpha, x, beta, y, s);
  }
  std::string describe(int, int64_t) override {
    return "sparse " + std::to_string(dl) + "x" + std::to_string(dr) + " nnz=" + std::to_string(m.vals.size()) +
           " spmm(CSR/CSC gather, 1-4 outputs per thread" + (csc.order.n ? ", right side in Cuthill-McKee column order)" : ")");
  }
};

struct LazySumOp : qob_op {
  std::vector<cplx> coefs;
  std::mutex coef_mu;                  // qob_lazysum_set_coefs may race with an apply on another host thread
  std::vector<qob_op *> terms;
  std::unique_ptr<TensorGroup> group;  // fused LazyTensor children
  std::vector<int> group_coef_index;   // child index of every group term
  std::vector<int> others;             // children applied one by one
  // sharded apply (qob_dist_*): programs on the rank-local index bits
  struct LayoutPlan {
    int nloc = 0;
    int nterms = 0;
    QTileProgram prog;
    // the round-2 kernel for plain launches of this plan (no peer addressing, no extra addend, no tile ranges, no SM budget):
    // the communication-free group of a sharded apply is the bulk of its HBM traffic
    bool has_qreg = false;
    QRegProgram qreg;
    uint64_t chunk_mask = 0;   // qob_layout_plan_set_chunk_bits
  };
  std::vector<std::unique_ptr<LayoutPlan>> layouts;
  LazySumOp(qob_ctx *c) : qob_op(c, OP_LAZYSUM) {}
  ~LazySumOp() override {
    for (qob_op *t : terms) op_release(t);
  }
  int apply(int side, cplx alpha, const void *x, cplx beta, void *y, int64_t batch, cudaStream_t s) override {
    const int64_t n_out = (side == QOB_SIDE_LEFT ? dl : dr) * batch;
    // empty sum or alpha == 0: only _zero_op_mul! (src/operators_lazysum.jl:190-192)
    if (terms.empty() || alpha == ZERO) return launch_scale(y, n_out, beta, s);
    std::vector<cplx> cf;
    {
      std::lock_guard<std::mutex> lk(coef_mu);
      cf = coefs;   // snapshot: one consistent coefficient set per apply
    }
    bool first = true;
    if (group) {
      QOB_TRY(group->apply(side, alpha, x, beta, y, batch, cf, s));
      first = false;
    }
    for (int i : others) {
      QOB_TRY(terms[i]->apply(side, alpha * cf[i], x, first ? beta : ONE, y, batch, s));
      first = false;
    }
    return QOB_STATUS_OK;
  }
  std::string describe(int side, int64_t batch) override {
    std::string t = "lazysum[terms=" + std::to_string(terms.size()) + "]";
    if (group) {
      Compiled *c = nullptr;
      if (group->compile(side, batch, &c) == QOB_STATUS_OK) t += " fused{" + c->text + "}";
      else t += std::string(" fused{error: ") + t_err + "}";
    }
    for (int i : others) t += " + " + terms[i]->describe(side, batch);
    return t;
  }
};

struct LazyProductOp : qob_op {
  cplx factor;
  std::vector<qob_op *> ops;
  LazyProductOp(qob_ctx *c) : qob_op(c, OP_LAZYPRODUCT) {}
  ~LazyProductOp() override {
    for (qob_op *t : ops) op_release(t);
  }
  int apply(int side, cplx alpha, const void *x, cplx beta, void *y, int64_t batch, cudaStream_t s) override {
    if (alpha == ZERO) return launch_scale(y, (side == QOB_SIDE_LEFT ? dl : dr) * batch, beta, s);
    const int m = (int)ops.size();
    if (m == 1) return ops[0]->apply(side, factor * alpha, x, beta, y, batch, s);
    // left:  t = factor*O_m x, t = O_k t ..., y = alpha*O_1 t + beta*y   (operators_lazyproduct.jl:103-115)
    // right: t = factor*x O_1, t = t O_k ..., y = alpha*t O_m + beta*y   (operators_lazyproduct.jl:117-129)
    int64_t max_dim = 0;
    for (qob_op *o : ops) max_dim = std::max(max_dim, std::max(o->dl, o->dr));
    void *tmp[2];
    QOB_TRY(ctx->get_scratch(s, slot_base + 0, (size_t)max_dim * batch * sizeof(double2), &tmp[0]));
    QOB_TRY(ctx->get_scratch(s, slot_base + 1, (size_t)max_dim * batch * sizeof(double2), &tmp[1]));
    const void *src = x;
    for (int step = 0; step < m; ++step) {
      int k = side == QOB_SIDE_LEFT ? m - 1 - step : step;
      void *dst = (step == m - 1) ? y : tmp[step & 1];
      cplx a = step == 0 ? factor : ONE, b = ZERO;
      if (step == m - 1) {
        a = alpha;
        b = beta;
      }
      QOB_TRY(ops[k]->apply(side, a, src, b, dst, batch, s));
      src = dst;
    }
    return QOB_STATUS_OK;
  }
  std::string describe(int side, int64_t batch) override {
    std::string t = "lazyproduct[";
    for (qob_op *o : ops) t += o->describe(side, batch) + "; ";
    return t + "]";
  }
};

// LazyDirectSum (src/spinors.jl:158-163, mul! :221-247): block-diagonal operator on a SumBasis.  Block i maps the slice
// [index[i], index[i+1]) of the state to the same slice of the result; like the reference the slices are cut by the RIGHT
// basis lengths for both vectors, so every block has to be square (the reference's `Ket(bases_l[i], result.data[...])`
// throws DimensionMismatch otherwise).  Only Ket / Bra methods exist in the reference: batch must be 1.
struct LazyDirectSumOp : qob_op {
  std::vector<qob_op *> ops;
  LazyDirectSumOp(qob_ctx *c) : qob_op(c, OP_DIRECTSUM) {}
  ~LazyDirectSumOp() override {
    for (qob_op *t : ops) op_release(t);
  }
  int apply(int side, cplx alpha, const void *x, cplx beta, void *y, int64_t batch, cudaStream_t s) override {
    if (batch != 1) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "LazyDirectSum: mul! is only defined for Ket and Bra states (src/spinors.jl:221-247)");
    int64_t off = 0;
    for (qob_op *o : ops) {
      if (o->dl != o->dr)
        QOB_FAIL(QOB_STATUS_DIM_MISMATCH, "LazyDirectSum block of size %lldx%lld: the reference slices state and result by the same lengths",
                 (long long)o->dl, (long long)o->dr);
      QOB_TRY(o->apply(side, alpha, (const char *)x + off * 16, beta, (char *)y + off * 16, 1, s));
      off += o->dr;
    }
    return QOB_STATUS_OK;
  }
  std::string describe(int side, int64_t batch) override {
    std::string t = "lazydirectsum[";
    for (qob_op *o : ops) t += o->describe(side, batch) + "; ";
    return t + "]";
  }
};

// ============================================================================ C ABI
extern "C" {

int qob_version(void) { return QOB200_VERSION; }
const char *qob_last_error(void) { return t_err; }
const char *qob_status_string(int st) {
  switch (st) {
    case QOB_STATUS_OK: return "ok";
    case QOB_STATUS_DIM_MISMATCH: return "DimensionMismatch";
    case QOB_STATUS_ALIASING: return "ArgumentError(aliasing)";
    case QOB_STATUS_INVALID_ARG: return "ArgumentError";
    case QOB_STATUS_UNSUPPORTED: return "MethodError(unsupported)";
    case QOB_STATUS_CUDA_ERROR: return "CUDA error";
    case QOB_STATUS_NCCL_ERROR: return "NCCL error";
    case QOB_STATUS_ALLOC: return "allocation failure";
  }
  return "unknown";
}

int qob_ctx_create(int device, qob_ctx **out) {
  if (!out) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null out pointer");
  if (device == -1) {  // planning-only context: constructors, validation and qob_op_describe work, compute does not
    qob_ctx *c = new qob_ctx();
    c->device = -1;
    *out = c;
    return QOB_STATUS_OK;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    QOB_FAIL(QOB_STATUS_CUDA_ERROR, "no CUDA device available (%s): libqob200 has no CPU fallback",
             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= count) QOB_FAIL(QOB_STATUS_INVALID_ARG, "device %d out of range (0..%d)", device, count - 1);
  QOB_CUDA(cudaSetDevice(device));
  qob_ctx *c = new qob_ctx();
  c->device = device;
  cudaDeviceProp prop;
  QOB_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  *out = c;
  return QOB_STATUS_OK;
}
int qob_ctx_destroy(qob_ctx *ctx) {
  if (!ctx) return QOB_STATUS_OK;
  // handles keep a pointer to their context (scratch pool, streams): destroying it under them would be a use-after-free
  if (ctx->live_ops.load() > 0)
    QOB_FAIL(QOB_STATUS_INVALID_ARG, "context still has %d live operator handle(s): destroy them first", ctx->live_ops.load());
  ctx->clear_scratch();
  for (cudaStream_t st : ctx->pipe_streams)
    if (st) cudaStreamDestroy(st);
  delete ctx;
  return QOB_STATUS_OK;
}
int qob_ctx_scratch_bytes(qob_ctx *ctx, int64_t *bytes) {
  if (!ctx || !bytes) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  *bytes = ctx->scratch_bytes();
  return QOB_STATUS_OK;
}
int qob_ctx_clear_scratch(qob_ctx *ctx) {
  if (!ctx) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null context");
  ctx->clear_scratch();
  return QOB_STATUS_OK;
}

static int make_group_term(const std::vector<int64_t> &dims_l, const std::vector<int64_t> &dims_r, int32_t nfac,
                           const int32_t *sites, const qob_factor *factors, cplx factor, GroupTerm &t) {
  const int n = (int)dims_l.size();
  t.coef_index = -1;
  t.scalar = factor;
  int prev = 0;
  for (int f = 0; f < nfac; ++f) {
    int sidx = sites[f];
    // check_indices + issorted (src/operators_lazytensor.jl:23-27)
    if (sidx < 1 || sidx > n) QOB_FAIL(QOB_STATUS_INVALID_ARG, "site index %d out of range 1..%d", sidx, n);
    if (sidx <= prev) QOB_FAIL(QOB_STATUS_INVALID_ARG, "LazyTensor indices must be sorted and unique");
    prev = sidx;
    auto m = std::make_shared<HostMat>();
    QOB_TRY(hostmat_from_factor(&factors[f], *m));
    // ops[n].basis_l == bl.bases[indices[n]] etc. (:30-31)
    if (m->rows != dims_l[sidx - 1] || m->cols != dims_r[sidx - 1])
      QOB_FAIL(QOB_STATUS_INVALID_ARG, "operator on subsystem %d is %lldx%lld but the bases have dimensions %lldx%lld", sidx,
               (long long)m->rows, (long long)m->cols, (long long)dims_l[sidx - 1], (long long)dims_r[sidx - 1]);
    if (m->is_square_eye()) continue;  // square identities are dropped (_tpops_tuple, :523-537)
    t.sites.push_back(sidx - 1);
    t.mats.push_back(m);
  }
  return QOB_STATUS_OK;
}

int qob_lazytensor_create(qob_ctx *ctx, int32_t nsub, const int64_t *dims_l, const int64_t *dims_r, int32_t nfac,
                          const int32_t *sites, const qob_factor *factors, qob_c64 factor, qob_op **out) {
  if (!ctx || !out || nsub < 1 || !dims_l || !dims_r || nfac < 0) QOB_FAIL(QOB_STATUS_INVALID_ARG, "bad LazyTensor arguments");
  auto op = std::make_unique<LazyTensorOp>(ctx);
  op->group.dims_l.assign(dims_l, dims_l + nsub);
  op->group.dims_r.assign(dims_r, dims_r + nsub);
  op->dl = op->dr = 1;
  for (int k = 0; k < nsub; ++k) {
    if (dims_l[k] < 1 || dims_r[k] < 1) QOB_FAIL(QOB_STATUS_INVALID_ARG, "subsystem dimensions must be >= 1");
    op->dl *= dims_l[k];
    op->dr *= dims_r[k];
  }
  GroupTerm t;
  QOB_TRY(make_group_term(op->group.dims_l, op->group.dims_r, nfac, sites, factors, C(factor), t));
  op->group.terms.push_back(std::move(t));
  *out = op.release();
  return QOB_STATUS_OK;
}

static int sparse_upload(const std::vector<int32_t> &ptr, const std::vector<int32_t> &idx, const std::vector<cplx> &v,
                         SparseDev &d) {
  std::vector<double2> vv(std::max<size_t>(1, v.size()), make_double2(0, 0));
  for (size_t i = 0; i < v.size(); ++i) vv[i] = make_double2(v[i].real(), v[i].imag());
  std::vector<int32_t> ii(idx);
  if (ii.empty()) ii.push_back(0);
  QOB_TRY(d.ptr.upload(ptr));
  QOB_TRY(d.idx.upload(ii));
  QOB_TRY(d.val.upload(vv));
  d.nptr = (int64_t)ptr.size();
  return QOB_STATUS_OK;
}

// Cuthill-McKee ordering of the graph of a square sparse operator (pattern of M + M^T): breadth-first from a
// minimum-degree vertex of every component, neighbours by increasing degree.  The right-side SpMM walks its output
// columns in this order so that the input columns it gathers from stay close in time (L2 reuse).
static std::vector<int32_t> cuthill_mckee_order(const HostMat &m) {
  const int64_t n = m.cols;
  std::vector<std::vector<int32_t>> adj((size_t)n);
  for (int64_t c = 0; c < n; ++c)
    for (int64_t p = m.colptr[c]; p < m.colptr[c + 1]; ++p) {
      const int64_t r = m.rowidx[p];
      if (r == c) continue;
      adj[(size_t)c].push_back((int32_t)r);
      adj[(size_t)r].push_back((int32_t)c);
    }
  for (auto &a : adj) {
    std::sort(a.begin(), a.end());
    a.erase(std::unique(a.begin(), a.end()), a.end());
  }
  std::vector<int32_t> by_degree((size_t)n), order;
  for (int64_t i = 0; i < n; ++i) by_degree[(size_t)i] = (int32_t)i;
  std::stable_sort(by_degree.begin(), by_degree.end(), [&](int32_t a, int32_t b) { return adj[a].size() < adj[b].size(); });
  std::vector<char> seen((size_t)n, 0);
  order.reserve((size_t)n);
  for (int32_t start : by_degree) {
    if (seen[start]) continue;
    seen[start] = 1;
    size_t head = order.size();
    order.push_back(start);
    while (head < order.size()) {
      const int32_t v = order[head++];
      std::vector<int32_t> nb;
      for (int32_t w : adj[v])
        if (!seen[w]) {
          seen[w] = 1;
          nb.push_back(w);
        }
      std::stable_sort(nb.begin(), nb.end(), [&](int32_t a, int32_t b) { return adj[a].size() < adj[b].size(); });
      for (int32_t w : nb) order.push_back(w);
    }
  }
  return order;
}

int qob_sparse_create(qob_ctx *ctx, const qob_factor *f, qob_op **out) {
  if (!ctx || !out || !f) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  PlanningScope ps(ctx->device < 0);
  if (f->kind != QOB_FACTOR_CSC) QOB_FAIL(QOB_STATUS_INVALID_ARG, "qob_sparse_create needs a CSC factor");
  auto op = std::make_unique<SparseOp>(ctx);
  QOB_TRY(hostmat_from_factor(f, op->m));
  op->dl = op->m.rows;
  op->dr = op->m.cols;
  if (op->m.vals.size() > 0x7FFFFFF0ull) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "more than 2^31 nonzeros");
  std::vector<int32_t> rp, ci;
  std::vector<cplx> v;
  op->m.to_csr(rp, ci, v);
  QOB_TRY(sparse_upload(rp, ci, v, op->csr));
  std::vector<int32_t> cp(op->m.colptr.begin(), op->m.colptr.end()), ri(op->m.rowidx.begin(), op->m.rowidx.end());
  QOB_TRY(sparse_upload(cp, ri, op->m.vals, op->csc));
  if (op->m.rows == op->m.cols && op->m.cols >= 1024) {
    std::vector<int32_t> order = cuthill_mckee_order(op->m);
    bool identity = true;
    for (size_t i = 0; i < order.size() && identity; ++i) identity = order[i] == (int32_t)i;
    if (!identity) QOB_TRY(op->csc.order.upload(order));
  }
  *out = op.release();
  return QOB_STATUS_OK;
}

int qob_dense_create(qob_ctx *ctx, const qob_factor *f, qob_op **out) {
  if (!ctx || !out || !f) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  if (f->kind != QOB_FACTOR_DENSE) QOB_FAIL(QOB_STATUS_INVALID_ARG, "qob_dense_create needs a dense factor");
  // a dense Operator is a one-subsystem tensor with one dense factor: BLAS gemm/gemv in the reference
  // (src/operators_dense.jl:394-396), the DMMA axis kernel (d >= 16) or the gather kernel here.
  HostMat m;
  QOB_TRY(hostmat_from_factor(f, m));
  int64_t dl = m.rows, dr = m.cols;
  int32_t site = 1;
  qob_c64 one = {1.0, 0.0};
  qob_op *o = nullptr;
  QOB_TRY(qob_lazytensor_create(ctx, 1, &dl, &dr, 1, &site, f, one, &o));
  o->kind = OP_DENSE;
  *out = o;
  return QOB_STATUS_OK;
}

// ---------------------------------------------------------------------------- fused master-equation right-hand side
// (kernel in qob_kernels_lindblad.cu).  Host side: Heff = H - i/2 sum_k r_k J_k^+ J_k by rows,
// G = H + i/2 sum_k r_k J_k^+ J_k by columns, sqrt(r_k) J_k by rows.
typedef std::map<std::pair<int64_t, int64_t>, cplx> Coo;   // ordered by (first, second)
static void coo_add(Coo &c, const HostMat &m, cplx scale, bool key_by_column) {
  auto put = [&](int64_t r, int64_t col, cplx v) { c[key_by_column ? std::make_pair(col, r) : std::make_pair(r, col)] += scale * v; };
  if (m.kind == QOB_FACTOR_CSC) {
    for (int64_t col = 0; col < m.cols; ++col)
      for (int64_t p = m.colptr[col]; p < m.colptr[col + 1]; ++p) put(m.rowidx[p], col, m.vals[p]);
  } else if (m.kind == QOB_FACTOR_DENSE) {
    for (int64_t col = 0; col < m.cols; ++col)
      for (int64_t r = 0; r < m.rows; ++r) {
        const cplx v = m.dense[(size_t)(r + col * m.rows)];
        if (v != ZERO) put(r, col, v);
      }
  } else {
    for (int64_t i = 0; i < std::min(m.rows, m.cols); ++i) put(i, i, ONE);
  }
}
static void coo_compress(const Coo &c, int64_t n, std::vector<int32_t> &ptr, std::vector<int32_t> &idx, std::vector<cplx> &val) {
  ptr.assign((size_t)n + 1, 0);
  idx.clear();
  val.clear();
  for (const auto &kv : c) {
    ++ptr[(size_t)kv.first.first + 1];
    idx.push_back((int32_t)kv.first.second);
    val.push_back(kv.second);
  }
  for (int64_t i = 0; i < n; ++i) ptr[(size_t)i + 1] += ptr[(size_t)i];
}

struct LindbladOp : qob_op {
  LindbladDev dev;
  // host copies (introspection / tests)
  std::vector<int32_t> h_ptr, h_col, g_ptr, g_row, j_ptr, j_col;
  std::vector<cplx> h_val, g_val, j_val;
  int nJ = 0;
  LindbladOp(qob_ctx *c) : qob_op(c, OP_LINDBLAD) {}
  int apply(int, cplx, const void *, cplx, void *, int64_t, cudaStream_t) override {
    QOB_FAIL(QOB_STATUS_UNSUPPORTED, "a Lindblad right-hand side is applied with qob_lindblad_apply");
  }
  std::string describe(int, int64_t) override {
    return "lindblad " + std::to_string(dl) + "x" + std::to_string(dl) + " Heff nnz=" + std::to_string(h_val.size()) + " jumps=" +
           std::to_string(nJ) + " (nnz=" + std::to_string(j_val.size()) + ") fused{one kernel, one output element per thread}";
  }
};

int qob_lindblad_create(qob_ctx *ctx, const qob_factor *H, int32_t nJ, const qob_factor *J, const double *rates, qob_op **out) {
  if (!ctx || !H || !out || nJ < 0 || (nJ > 0 && !J)) QOB_FAIL(QOB_STATUS_INVALID_ARG, "bad Lindblad arguments");
  PlanningScope ps(ctx->device < 0);
  auto op = std::make_unique<LindbladOp>(ctx);
  HostMat h;
  QOB_TRY(hostmat_from_factor(H, h));
  if (h.rows != h.cols) QOB_FAIL(QOB_STATUS_DIM_MISMATCH, "the Hamiltonian must be square, got %lldx%lld", (long long)h.rows, (long long)h.cols);
  const int64_t D = h.rows;
  if (D > 0x7FFFFFF0ll) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "dimension too large");
  op->dl = op->dr = D;
  op->nJ = nJ;
  Coo jdj;   // sum_k r_k J_k^+ J_k, keyed (row, col)
  op->j_ptr.assign((size_t)nJ * (size_t)(D + 1), 0);
  for (int k = 0; k < nJ; ++k) {
    HostMat jm;
    QOB_TRY(hostmat_from_factor(&J[k], jm));
    if (jm.rows != D || jm.cols != D)
      QOB_FAIL(QOB_STATUS_DIM_MISMATCH, "jump operator %d is %lldx%lld, expected %lldx%lld", k + 1, (long long)jm.rows, (long long)jm.cols,
               (long long)D, (long long)D);
    const double r = rates ? rates[k] : 1.0;
    if (!(r >= 0.0)) QOB_FAIL(QOB_STATUS_INVALID_ARG, "rate %d must be >= 0", k + 1);
    Coo jk;
    coo_add(jk, jm, cplx(std::sqrt(r), 0.0), false);
    std::vector<int32_t> ptr, col;
    std::vector<cplx> val;
    coo_compress(jk, D, ptr, col, val);
    const int32_t base = (int32_t)op->j_col.size();
    if ((size_t)base + col.size() > 0x7FFFFFF0ull) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "too many nonzeros in the jump operators");
    for (int64_t i = 0; i <= D; ++i) op->j_ptr[(size_t)k * (size_t)(D + 1) + (size_t)i] = base + ptr[(size_t)i];
    op->j_col.insert(op->j_col.end(), col.begin(), col.end());
    op->j_val.insert(op->j_val.end(), val.begin(), val.end());
    // (J^+ J)[a, b] = sum_r conj(J[r, a]) J[r, b]: pairs of entries of every row (the sqrt(r) factors multiply to r)
    for (int64_t row = 0; row < D; ++row)
      for (int32_t p = ptr[(size_t)row]; p < ptr[(size_t)row + 1]; ++p)
        for (int32_t q = ptr[(size_t)row]; q < ptr[(size_t)row + 1]; ++q)
          jdj[std::make_pair((int64_t)col[(size_t)p], (int64_t)col[(size_t)q])] += std::conj(val[(size_t)p]) * val[(size_t)q];
  }
  Coo heff, g;   // Heff by rows: key (row, col); G by columns: key (col, row)
  coo_add(heff, h, ONE, false);
  coo_add(g, h, ONE, true);
  for (const auto &kv : jdj) {
    heff[kv.first] += cplx(0.0, -0.5) * kv.second;
    g[std::make_pair(kv.first.second, kv.first.first)] += cplx(0.0, 0.5) * kv.second;
  }
  coo_compress(heff, D, op->h_ptr, op->h_col, op->h_val);
  coo_compress(g, D, op->g_ptr, op->g_row, op->g_val);
  auto up_i = [](DevArray<int32_t> &d, std::vector<int32_t> v) {
    if (v.empty()) v.push_back(0);
    return d.upload(v);
  };
  auto up_v = [](DevArray<double2> &d, const std::vector<cplx> &v) {
    std::vector<double2> t(std::max<size_t>(1, v.size()), make_double2(0.0, 0.0));
    for (size_t i = 0; i < v.size(); ++i) t[i] = make_double2(v[i].real(), v[i].imag());
    return d.upload(t);
  };
  op->dev.D = D;
  op->dev.nJ = nJ;
  QOB_TRY(up_i(op->dev.h_ptr, op->h_ptr));
  QOB_TRY(up_i(op->dev.h_col, op->h_col));
  QOB_TRY(up_v(op->dev.h_val, op->h_val));
  QOB_TRY(up_i(op->dev.g_ptr, op->g_ptr));
  QOB_TRY(up_i(op->dev.g_row, op->g_row));
  QOB_TRY(up_v(op->dev.g_val, op->g_val));
  QOB_TRY(up_i(op->dev.j_ptr, op->j_ptr));
  QOB_TRY(up_i(op->dev.j_col, op->j_col));
  QOB_TRY(up_v(op->dev.j_val, op->j_val));
  *out = op.release();
  return QOB_STATUS_OK;
}

int qob_lindblad_apply(qob_op *L, qob_c64 alpha, const void *rho, qob_c64 beta, void *drho, void *stream) {
  if (!L || L->kind != OP_LINDBLAD) QOB_FAIL(QOB_STATUS_INVALID_ARG, "not a Lindblad handle");
  LindbladOp *op = static_cast<LindbladOp *>(L);
  const int64_t n = op->dl * op->dl;
  if (n == 0) return QOB_STATUS_OK;
  if (!rho || !drho) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null data pointer");
  QOB_TRY(check_alias(rho, n, drho, n));
  if (op->ctx->device < 0) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "planning-only context (no CUDA device): libqob200 has no CPU fallback");
  QOB_DEVICE(op->ctx->device);
  if (C(alpha) == ZERO) return launch_scale(drho, n, C(beta), (cudaStream_t)stream);   // only the beta update, like mul!
  return launch_lindblad(op->dev, C(alpha), rho, C(beta), drho, (cudaStream_t)stream);
}

int qob_lindblad_dense(qob_op *L, int32_t which, qob_c64 *out) {
  if (!L || L->kind != OP_LINDBLAD || !out) QOB_FAIL(QOB_STATUS_INVALID_ARG, "not a Lindblad handle");
  LindbladOp *op = static_cast<LindbladOp *>(L);
  const int64_t D = op->dl;
  if (which < 0 || which > 1 + op->nJ) QOB_FAIL(QOB_STATUS_INVALID_ARG, "matrix index out of range");
  for (int64_t i = 0; i < D * D; ++i) out[i].re = out[i].im = 0.0;
  auto put = [&](int64_t r, int64_t c, cplx v) {
    out[r + c * D].re += v.real();
    out[r + c * D].im += v.imag();
  };
  if (which == 0) {
    for (int64_t r = 0; r < D; ++r)
      for (int32_t p = op->h_ptr[(size_t)r]; p < op->h_ptr[(size_t)r + 1]; ++p) put(r, op->h_col[(size_t)p], op->h_val[(size_t)p]);
  } else if (which == 1) {
    for (int64_t c = 0; c < D; ++c)
      for (int32_t p = op->g_ptr[(size_t)c]; p < op->g_ptr[(size_t)c + 1]; ++p) put(op->g_row[(size_t)p], c, op->g_val[(size_t)p]);
  } else {
    const size_t k = (size_t)(which - 2);
    for (int64_t r = 0; r < D; ++r)
      for (int32_t p = op->j_ptr[k * (size_t)(D + 1) + (size_t)r]; p < op->j_ptr[k * (size_t)(D + 1) + (size_t)r + 1]; ++p)
        put(r, op->j_col[(size_t)p], op->j_val[(size_t)p]);
  }
  return QOB_STATUS_OK;
}

int qob_lazysum_create(qob_ctx *ctx, int64_t dim_l, int64_t dim_r, int32_t nterms, const qob_c64 *coefs,
                       qob_op *const *terms, qob_op **out) {
  if (!ctx || !out || nterms < 0 || (nterms > 0 && (!coefs || !terms))) QOB_FAIL(QOB_STATUS_INVALID_ARG, "bad LazySum arguments");
  auto op = std::make_unique<LazySumOp>(ctx);
  op->dl = dim_l;
  op->dr = dim_r;
  for (int i = 0; i < nterms; ++i) {
    if (!terms[i]) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null term %d", i);
    if (terms[i]->ctx != ctx) QOB_FAIL(QOB_STATUS_INVALID_ARG, "LazySum term %d belongs to another context (device)", i + 1);
    // _check_bases (src/operators_lazysum.jl:6-11): IncompatibleBases -> dimension mismatch here
    if (terms[i]->dl != dim_l || terms[i]->dr != dim_r)
      QOB_FAIL(QOB_STATUS_DIM_MISMATCH, "LazySum term %d has dimensions %lldx%lld, expected %lldx%lld", i + 1,
               (long long)terms[i]->dl, (long long)terms[i]->dr, (long long)dim_l, (long long)dim_r);
  }
  for (int i = 0; i < nterms; ++i) {
    op_retain(terms[i]);
    op->terms.push_back(terms[i]);
    op->coefs.push_back(C(coefs[i]));
  }
  // fuse every LazyTensor child that lives on the same composite bases as the first one
  const LazyTensorOp *ref = nullptr;
  for (int i = 0; i < nterms; ++i) {
    LazyTensorOp *lt = (terms[i]->kind == OP_LAZYTENSOR) ? static_cast<LazyTensorOp *>(terms[i]) : nullptr;
    if (lt && !ref) ref = lt;
    if (lt && lt->group.dims_l == ref->group.dims_l && lt->group.dims_r == ref->group.dims_r) {
      if (!op->group) {
        op->group = std::make_unique<TensorGroup>();
        op->group->ctx = ctx;
        op->group->slot_base = op->slot_base + 2;
        op->group->dims_l = lt->group.dims_l;
        op->group->dims_r = lt->group.dims_r;
      }
      GroupTerm t = lt->group.terms[0];
      t.coef_index = i;
      op->group->terms.push_back(std::move(t));
    } else {
      op->others.push_back(i);
    }
  }
  *out = op.release();
  return QOB_STATUS_OK;
}

int qob_lazysum_set_coefs(qob_op *sum, int32_t nterms, const qob_c64 *coefs) {
  if (!sum || sum->kind != OP_LAZYSUM) QOB_FAIL(QOB_STATUS_INVALID_ARG, "not a LazySum handle");
  LazySumOp *s = static_cast<LazySumOp *>(sum);
  if ((size_t)nterms != s->coefs.size()) QOB_FAIL(QOB_STATUS_INVALID_ARG, "LazySum has %d terms, got %d coefficients", (int)s->coefs.size(), nterms);
  if (!coefs && nterms > 0) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null coefficient array");
  std::lock_guard<std::mutex> lk(s->coef_mu);
  for (int i = 0; i < nterms; ++i) s->coefs[i] = C(coefs[i]);
  return QOB_STATUS_OK;
}

int qob_lazyproduct_create(qob_ctx *ctx, int32_t nops, qob_op *const *ops, qob_c64 factor, qob_op **out) {
  if (!ctx || !out) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  if (nops < 1 || !ops) QOB_FAIL(QOB_STATUS_INVALID_ARG, "LazyProduct needs at least one operator!");
  for (int i = 0; i < nops; ++i)
    if (!ops[i]) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null operator %d", i);
  for (int i = 0; i < nops; ++i)
    if (ops[i]->ctx != ctx) QOB_FAIL(QOB_STATUS_INVALID_ARG, "LazyProduct operator %d belongs to another context (device)", i + 1);
  for (int i = 1; i < nops; ++i)  // check_multiplicable (src/operators_lazyproduct.jl:4-10)
    if (ops[i - 1]->dr != ops[i]->dl) QOB_FAIL(QOB_STATUS_DIM_MISMATCH, "LazyProduct operators %d and %d are not multiplicable", i, i + 1);
  auto op = std::make_unique<LazyProductOp>(ctx);
  op->factor = C(factor);
  op->dl = ops[0]->dl;
  op->dr = ops[nops - 1]->dr;
  for (int i = 0; i < nops; ++i) {
    op_retain(ops[i]);
    op->ops.push_back(ops[i]);
  }
  *out = op.release();
  return QOB_STATUS_OK;
}

int qob_op_destroy(qob_op *op) {
  if (op) op_release(op);
  return QOB_STATUS_OK;
}

int qob_op_dims(const qob_op *op, int64_t *dim_l, int64_t *dim_r) {
  if (!op || !dim_l || !dim_r) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  *dim_l = op->dl;
  *dim_r = op->dr;
  return QOB_STATUS_OK;
}

int qob_op_apply(qob_op *op, int32_t side, qob_c64 alpha, const void *x, qob_c64 beta, void *y, int64_t batch,
                 void *stream) {
  if (!op) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null operator");
  if (side != QOB_SIDE_LEFT && side != QOB_SIDE_RIGHT) QOB_FAIL(QOB_STATUS_INVALID_ARG, "bad side %d", side);
  if (batch < 0) QOB_FAIL(QOB_STATUS_INVALID_ARG, "negative batch");
  const int64_t n_in = (side == QOB_SIDE_LEFT ? op->dr : op->dl) * batch;
  const int64_t n_out = (side == QOB_SIDE_LEFT ? op->dl : op->dr) * batch;
  if (n_out == 0) return QOB_STATUS_OK;
  if (!y || (!x && n_in > 0)) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null data pointer");
  QOB_TRY(check_alias(x, n_in, y, n_out));
  if (op->ctx->device < 0) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "planning-only context (no CUDA device): libqob200 has no CPU fallback");
  QOB_DEVICE(op->ctx->device);
  return op->apply(side, C(alpha), x, C(beta), y, batch, (cudaStream_t)stream);
}

static int64_t env_int64(const char *name, int64_t dflt) {
  const char *v = getenv(name);
  return v && *v ? atoll(v) : dflt;
}

// Host-buffer apply of `batch` kets in groups of `g` columns over two lanes (device buffer pairs) and three streams:
// up (H2D), one compute stream per lane, down (D2H).  Events only order what must be ordered:
//   up:   x_j after lane's previous apply has consumed its x buffer (and y_j staged after the previous download, beta != 0)
//   lane: apply_j after x_j (and y_j) are up and the lane's previous result has gone down
//   down: y_j after apply_j
static int apply_host_pipelined(qob_op *op, cplx alpha, const qob_c64 *x, cplx beta, qob_c64 *y, int64_t batch, int64_t g,
                                int64_t d_in, int64_t d_out) {
  qob_ctx *ctx = op->ctx;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (int i = 0; i < 4; ++i)
      if (!ctx->pipe_streams[i]) QOB_CUDA(cudaStreamCreateWithFlags(&ctx->pipe_streams[i], cudaStreamNonBlocking));
  }
  cudaStream_t up = ctx->pipe_streams[0], down = ctx->pipe_streams[1], lane[2] = {ctx->pipe_streams[2], ctx->pipe_streams[3]};
  void *dx[2] = {nullptr, nullptr}, *dy[2] = {nullptr, nullptr};
  for (int l = 0; l < 2; ++l) {
    QOB_TRY(ctx->get_scratch(lane[l], op->slot_base + 6, (size_t)std::max<int64_t>(1, d_in * g) * 16, &dx[l]));
    QOB_TRY(ctx->get_scratch(lane[l], op->slot_base + 7, (size_t)std::max<int64_t>(1, d_out * g) * 16, &dy[l]));
  }
  struct Events {  // destroyed on every exit path
    cudaEvent_t e[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    ~Events() {
      for (cudaEvent_t v : e)
        if (v) cudaEventDestroy(v);
    }
  } evs;
  for (int i = 0; i < 6; ++i) QOB_CUDA(cudaEventCreateWithFlags(&evs.e[i], cudaEventDisableTiming));
  cudaEvent_t *e_up = evs.e, *e_done = evs.e + 2, *e_down = evs.e + 4;
  int rc = QOB_STATUS_OK;
  const int64_t ngroups = (batch + g - 1) / g;
  for (int64_t j = 0; j < ngroups && rc == QOB_STATUS_OK; ++j) {
    const int l = (int)(j & 1);
    const int64_t c0 = j * g, nc = std::min(g, batch - c0);
    cudaError_t ce = cudaSuccess;
    auto ok = [&](cudaError_t e) { if (ce == cudaSuccess) ce = e; };
    if (j >= 2) ok(cudaStreamWaitEvent(up, e_done[l], 0));
    ok(cudaMemcpyAsync(dx[l], x + c0 * d_in, (size_t)(nc * d_in) * 16, cudaMemcpyHostToDevice, up));
    if (beta != ZERO) {
      if (j >= 2) ok(cudaStreamWaitEvent(up, e_down[l], 0));
      ok(cudaMemcpyAsync(dy[l], y + c0 * d_out, (size_t)(nc * d_out) * 16, cudaMemcpyHostToDevice, up));
    }
    ok(cudaEventRecord(e_up[l], up));
    ok(cudaStreamWaitEvent(lane[l], e_up[l], 0));
    if (j >= 2) ok(cudaStreamWaitEvent(lane[l], e_down[l], 0));
    if (ce != cudaSuccess) {
      qob_set_error("CUDA error in pipelined host apply: %s", cudaGetErrorString(ce));
      rc = QOB_STATUS_CUDA_ERROR;
      break;
    }
    rc = op->apply(QOB_SIDE_LEFT, alpha, dx[l], beta, dy[l], nc, lane[l]);
    if (rc != QOB_STATUS_OK) break;
    ok(cudaEventRecord(e_done[l], lane[l]));
    ok(cudaStreamWaitEvent(down, e_done[l], 0));
    ok(cudaMemcpyAsync(y + c0 * d_out, dy[l], (size_t)(nc * d_out) * 16, cudaMemcpyDeviceToHost, down));
    ok(cudaEventRecord(e_down[l], down));
    if (ce != cudaSuccess) {
      qob_set_error("CUDA error in pipelined host apply: %s", cudaGetErrorString(ce));
      rc = QOB_STATUS_CUDA_ERROR;
    }
  }
  cudaError_t e1 = cudaStreamSynchronize(up), e2 = cudaStreamSynchronize(lane[0]), e3 = cudaStreamSynchronize(lane[1]),
              e4 = cudaStreamSynchronize(down);
  if (rc != QOB_STATUS_OK) return rc;
  for (cudaError_t e : {e1, e2, e3, e4})
    if (e != cudaSuccess) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "CUDA error in pipelined host apply: %s", cudaGetErrorString(e));
  return QOB_STATUS_OK;
}

int qob_op_apply_host(qob_op *op, int32_t side, qob_c64 alpha, const qob_c64 *x, qob_c64 beta, qob_c64 *y,
                      int64_t batch) {
  if (!op) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null operator");
  if (side != QOB_SIDE_LEFT && side != QOB_SIDE_RIGHT) QOB_FAIL(QOB_STATUS_INVALID_ARG, "bad side %d", side);
  const int64_t n_in = (side == QOB_SIDE_LEFT ? op->dr : op->dl) * batch;
  const int64_t n_out = (side == QOB_SIDE_LEFT ? op->dl : op->dr) * batch;
  if (n_out == 0) return QOB_STATUS_OK;
  if (!x || !y) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null data pointer");
  if ((const void *)x == (const void *)y) QOB_FAIL(QOB_STATUS_ALIASING, "output matrix must not be aliased with input matrix");
  if (op->ctx->device < 0) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "planning-only context (no CUDA device): libqob200 has no CPU fallback");
  QOB_DEVICE(op->ctx->device);
  const int64_t d_in = side == QOB_SIDE_LEFT ? op->dr : op->dl, d_out = side == QOB_SIDE_LEFT ? op->dl : op->dr;
  // A batch of kets (LEFT side: the columns are contiguous) that is large enough is streamed through the device in column
  // groups: group j+1 goes up on one copy engine while group j is applied and group j-1 comes down on the other, so the
  // call is bound by ONE direction of the host link instead of the sum of both directions plus the kernels.
  const int64_t pipe_min = env_int64("QOB_HOST_PIPE_MIN_BYTES", 32ll << 20);
  const int64_t col_bytes = 16 * std::max(d_in, d_out);
  if (side == QOB_SIDE_LEFT && batch >= 2 && pipe_min > 0 && col_bytes * batch >= 2 * pipe_min) {
    const int64_t g = std::max<int64_t>(1, std::min<int64_t>(batch / 2, pipe_min / col_bytes));
    return apply_host_pipelined(op, C(alpha), x, C(beta), y, batch, g, d_in, d_out);
  }
  void *dx = nullptr, *dy = nullptr;
  cudaStream_t s = 0;
  QOB_TRY(op->ctx->get_scratch(s, op->slot_base + 6, (size_t)std::max<int64_t>(1, n_in) * 16, &dx));
  QOB_TRY(op->ctx->get_scratch(s, op->slot_base + 7, (size_t)n_out * 16, &dy));
  QOB_CUDA(cudaMemcpyAsync(dx, x, (size_t)n_in * 16, cudaMemcpyHostToDevice, s));
  if (C(beta) != ZERO) QOB_CUDA(cudaMemcpyAsync(dy, y, (size_t)n_out * 16, cudaMemcpyHostToDevice, s));
  QOB_TRY(op->apply(side, C(alpha), dx, C(beta), dy, batch, s));
  QOB_CUDA(cudaMemcpyAsync(y, dy, (size_t)n_out * 16, cudaMemcpyDeviceToHost, s));
  QOB_CUDA(cudaStreamSynchronize(s));
  return QOB_STATUS_OK;
}

int64_t qob_launch_count(void) { return g_launch_count.load(); }
int64_t qob_launch_count_of(int32_t family) { return family >= 0 && family < 8 ? g_family_count[family].load() : -1; }

int qob_op_describe(qob_op *op, int32_t side, int64_t batch, char *buf, int64_t buflen) {
  if (!op || !buf || buflen < 1) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  PlanningScope ps(op->ctx->device < 0);
  std::string t = op->describe(side, batch);
  snprintf(buf, (size_t)buflen, "%s", t.c_str());
  return QOB_STATUS_OK;
}

int qob_fill_state(void *x, int64_t offset, int64_t n, uint64_t seed, double scale, void *stream) {
  if (!x && n > 0) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null pointer");
  return launch_fill_state(x, offset, n, seed, scale, (cudaStream_t)stream);
}
int qob_norm2(const void *x, int64_t n, double *out, void *stream) {
  if (!out) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null pointer");
  return launch_norm2(x, n, out, (cudaStream_t)stream);
}
int qob_dot(const void *x, const void *y, int64_t n, qob_c64 *out, void *stream) {
  if (!out) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null pointer");
  cplx r;
  QOB_TRY(launch_dot(x, y, n, &r, (cudaStream_t)stream));
  out->re = r.real();
  out->im = r.imag();
  return QOB_STATUS_OK;
}

// ---- LazyDirectSum, expect / variance, partial traces -----------------------------------------
int qob_lazydirectsum_create(qob_ctx *ctx, int32_t nops, qob_op *const *ops, qob_op **out) {
  if (!ctx || !out) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  if (nops < 1 || !ops) QOB_FAIL(QOB_STATUS_INVALID_ARG, "LazyDirectSum needs at least one operator");
  auto op = std::make_unique<LazyDirectSumOp>(ctx);
  for (int i = 0; i < nops; ++i) {
    if (!ops[i]) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null operator %d", i);
    if (ops[i]->ctx != ctx) QOB_FAIL(QOB_STATUS_INVALID_ARG, "LazyDirectSum operator %d belongs to another context (device)", i + 1);
  }
  for (int i = 0; i < nops; ++i) {
    op_retain(ops[i]);
    op->ops.push_back(ops[i]);
    op->dl += ops[i]->dl;
    op->dr += ops[i]->dr;
  }
  *out = op.release();
  return QOB_STATUS_OK;
}

// <x| op |x> = dot(x, op*x)  (src/operators.jl:119): mul! into the handle's scratch, then one deterministic device reduction;
// only the scalar crosses to the host.
int qob_expect(qob_op *op, const void *x, qob_c64 *out, void *stream) {
  if (!op || !out) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  if (op->dl != op->dr) QOB_FAIL(QOB_STATUS_DIM_MISMATCH, "expect needs an operator with equal left and right bases");
  if (!x && op->dr > 0) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null data pointer");
  if (op->ctx->device < 0) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "planning-only context (no CUDA device): libqob200 has no CPU fallback");
  QOB_DEVICE(op->ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  cplx r;
  // a LazySum of LazyTensors on a large spin-1/2 state: every tile pass reduces conj(x) * (its share of op x) on the fly —
  // one sweep over x per launch, no result vector written, no second pass for the dot product
  if (op->kind == OP_LAZYSUM) {
    LazySumOp *S = static_cast<LazySumOp *>(op);
    if (S->group && S->others.empty() && !S->terms.empty()) {
      std::vector<cplx> cf;
      {
        std::lock_guard<std::mutex> lk(S->coef_mu);
        cf = S->coefs;
      }
      const int st = S->group->apply(QOB_SIDE_LEFT, ONE, x, ZERO, nullptr, 1, cf, s, &r);
      if (st == QOB_STATUS_OK) {
        out->re = r.real();
        out->im = r.imag();
        return QOB_STATUS_OK;
      }
      if (st != QOB_STATUS_UNSUPPORTED) return st;
    }
  }
  void *tmp = nullptr;
  QOB_TRY(op->ctx->get_scratch(s, op->slot_base + 4, (size_t)std::max<int64_t>(1, op->dl) * 16, &tmp));
  QOB_TRY(op->apply(QOB_SIDE_LEFT, ONE, x, ZERO, tmp, 1, s));
  QOB_TRY(launch_dot(x, tmp, op->dl, &r, s));
  out->re = r.real();
  out->im = r.imag();
  return QOB_STATUS_OK;
}

// variance(op, psi) = psi' (op (op psi)) - (psi' (op psi))^2  (src/operators.jl:139-142), two applications like the reference
int qob_variance(qob_op *op, const void *x, qob_c64 *out, void *stream) {
  if (!op || !out) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  if (op->dl != op->dr) QOB_FAIL(QOB_STATUS_DIM_MISMATCH, "variance needs an operator with equal left and right bases");
  if (!x && op->dr > 0) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null data pointer");
  if (op->ctx->device < 0) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "planning-only context (no CUDA device): libqob200 has no CPU fallback");
  QOB_DEVICE(op->ctx->device);
  cudaStream_t s = (cudaStream_t)stream;
  void *t1 = nullptr, *t2 = nullptr;
  QOB_TRY(op->ctx->get_scratch(s, op->slot_base + 4, (size_t)std::max<int64_t>(1, op->dl) * 16, &t1));
  QOB_TRY(op->ctx->get_scratch(s, op->slot_base + 5, (size_t)std::max<int64_t>(1, op->dl) * 16, &t2));
  QOB_TRY(op->apply(QOB_SIDE_LEFT, ONE, x, ZERO, t1, 1, s));
  QOB_TRY(op->apply(QOB_SIDE_LEFT, ONE, t1, ZERO, t2, 1, s));
  cplx e1, e2;
  QOB_TRY(launch_dot(x, t1, op->dl, &e1, s));
  QOB_TRY(launch_dot(x, t2, op->dl, &e2, s));
  const cplx v = e2 - e1 * e1;
  out->re = v.real();
  out->im = v.imag();
  return QOB_STATUS_OK;
}

static const int PTRACE_SLOT = 900;  // context-level scratch slots (operator handles start at 1000)
int qob_ptrace_op(qob_ctx *ctx, int32_t nsub, const int64_t *dims_l, const int64_t *dims_r, int32_t ntraced, const int32_t *traced,
                  const void *a, void *result, void *stream) {
  if (!ctx) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null context");
  PlanningScope ps(ctx->device < 0);
  if (ctx->device >= 0) {
    if (!a || !result) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null data pointer");
    QOB_DEVICE(ctx->device);
    return launch_ptrace_op(ctx, PTRACE_SLOT, nsub, dims_l, dims_r, ntraced, traced, a, result, (cudaStream_t)stream);
  }
  return launch_ptrace_op(ctx, PTRACE_SLOT, nsub, dims_l, dims_r, ntraced, traced, a, result, (cudaStream_t)stream);
}
int qob_ptrace_state(qob_ctx *ctx, int32_t nsub, const int64_t *dims, int32_t ntraced, const int32_t *traced, int32_t is_bra,
                     const void *psi, void *result, void *stream) {
  if (!ctx) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null context");
  PlanningScope ps(ctx->device < 0);
  if (ctx->device >= 0) {
    if (!psi || !result) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null data pointer");
    QOB_DEVICE(ctx->device);
    return launch_ptrace_state(ctx, PTRACE_SLOT + 1, nsub, dims, ntraced, traced, is_bra != 0, psi, result, (cudaStream_t)stream);
  }
  return launch_ptrace_state(ctx, PTRACE_SLOT + 1, nsub, dims, ntraced, traced, is_bra != 0, psi, result, (cudaStream_t)stream);
}

// ---- sharded apply: per-rank compute in an arbitrary index layout ----------------
static int qubit_sum(qob_op *sum, LazySumOp **out) {
  if (!sum || sum->kind != OP_LAZYSUM) QOB_FAIL(QOB_STATUS_INVALID_ARG, "not a LazySum handle");
  LazySumOp *S = static_cast<LazySumOp *>(sum);
  if (!S->group || !S->others.empty()) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "layout plans need a LazySum made only of LazyTensor terms");
  const TensorGroup &G = *S->group;
  if (G.dims_l.size() > 62) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "more than 62 subsystems");
  for (size_t k = 0; k < G.dims_l.size(); ++k)
    if (G.dims_l[k] != 2 || G.dims_r[k] != 2) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "layout plans support 2-dimensional subsystems only");
  *out = S;
  return QOB_STATUS_OK;
}

int qob_lazysum_term_masks(qob_op *sum, int32_t term, uint64_t *offdiag_mask, uint64_t *site_mask) {
  LazySumOp *S = nullptr;
  QOB_TRY(qubit_sum(sum, &S));
  if (term < 0 || term >= (int)S->group->terms.size()) QOB_FAIL(QOB_STATUS_INVALID_ARG, "term index out of range");
  const GroupTerm &t = S->group->terms[term];
  uint64_t od = 0, all = 0;
  for (size_t f = 0; f < t.sites.size(); ++f) {
    all |= 1ull << t.sites[f];
    if (t.mats[f]->at(0, 1) != ZERO || t.mats[f]->at(1, 0) != ZERO) od |= 1ull << t.sites[f];
  }
  if (offdiag_mask) *offdiag_mask = od;
  if (site_mask) *site_mask = all;
  return QOB_STATUS_OK;
}

int qob_layout_plan_create(qob_op *sum, int32_t nbits_local, const int32_t *bitpos, uint64_t hi_value,
                           const uint8_t *term_select, int32_t *plan_id) {
  LazySumOp *S = nullptr;
  QOB_TRY(qubit_sum(sum, &S));
  PlanningScope ps(S->ctx->device < 0);
  if (!bitpos || !plan_id) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null argument");
  const TensorGroup &G = *S->group;
  const int n = (int)G.dims_l.size();
  uint64_t seen = 0;
  for (int k = 0; k < n; ++k) {
    if (bitpos[k] < 0 || bitpos[k] >= 62 || (seen >> bitpos[k] & 1)) QOB_FAIL(QOB_STATUS_INVALID_ARG, "bitpos must be a one-to-one map into [0, 62)");
    seen |= 1ull << bitpos[k];
  }
  if (nbits_local < 10 || nbits_local > n) QOB_FAIL(QOB_STATUS_INVALID_ARG, "nbits_local must be in [10, nsub]");
  std::vector<QTerm> qt;
  for (size_t ti = 0; ti < G.terms.size(); ++ti) {
    if (term_select && !term_select[ti]) continue;
    const GroupTerm &t = G.terms[ti];
    if (t.sites.size() > 3) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "layout plans support terms on at most 3 sites");
    std::vector<int> order(t.sites.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return bitpos[t.sites[a]] < bitpos[t.sites[b]]; });
    QTerm q;
    q.coef_index = t.coef_index;
    q.scalar = t.scalar;
    for (int oi : order) {
      q.bits.push_back(bitpos[t.sites[oi]]);
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) q.m.push_back(t.mats[oi]->at(i, j));
    }
    qt.push_back(std::move(q));
  }
  auto lp = std::make_unique<LazySumOp::LayoutPlan>();
  lp->nloc = nbits_local;
  lp->nterms = (int)qt.size();
  QOB_TRY(qtile_build(lp->prog, nbits_local, hi_value, qt, S->ctx->sm_count));
  if (!qt.empty() && nbits_local >= env_i("QOB_QREG_MIN_BITS", 20) && !env_i("QOB_DISABLE_QREG", 0)) {
    const int st = qreg_build(lp->qreg, nbits_local, hi_value, qt, S->ctx->sm_count);
    if (st == QOB_STATUS_OK) lp->has_qreg = true;
    else if (st != QOB_STATUS_UNSUPPORTED) return st;
  }
  S->layouts.push_back(std::move(lp));
  *plan_id = (int)S->layouts.size() - 1;
  return QOB_STATUS_OK;
}

int qob_layout_plan_apply(qob_op *sum, int32_t plan_id, qob_c64 alpha, const void *x, qob_c64 beta, void *y,
                          void *stream) {
  LazySumOp *S = nullptr;
  QOB_TRY(qubit_sum(sum, &S));
  if (plan_id < 0 || plan_id >= (int)S->layouts.size()) QOB_FAIL(QOB_STATUS_INVALID_ARG, "bad layout plan id");
  if (S->ctx->device < 0) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "planning-only context (no CUDA device): libqob200 has no CPU fallback");
  LazySumOp::LayoutPlan &lp = *S->layouts[plan_id];
  const int64_t n = 1ll << lp.nloc;
  cudaStream_t s = (cudaStream_t)stream;
  if (!x || !y) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null data pointer");
  QOB_TRY(check_alias(x, n, y, n));
  QOB_DEVICE(S->ctx->device);
  if (C(alpha) == ZERO) return launch_scale(y, n, C(beta), s);
  std::vector<cplx> cf;
  {
    std::lock_guard<std::mutex> lk(S->coef_mu);
    cf = S->coefs;
  }
  if (lp.has_qreg) {
    QOB_TRY(qreg_set_coefs(lp.qreg, cf, s));
    return qreg_launch(lp.qreg, C(alpha), x, C(beta), y, s);
  }
  QOB_TRY(qtile_set_coefs(lp.prog, cf, s));
  return qtile_launch(lp.prog, C(alpha), x, C(beta), y, s);
}

int qob_layout_plan_apply_ex(qob_op *sum, int32_t plan_id, qob_c64 alpha, const void *x, qob_c64 beta, void *y,
                             const void *zadd, int32_t npeers, const void *const *x_peers, void *const *y_peers,
                             int32_t peer_shift, int32_t sm_budget, int32_t chunk_index, int32_t nchunks, void *stream) {
  LazySumOp *S = nullptr;
  QOB_TRY(qubit_sum(sum, &S));
  if (plan_id < 0 || plan_id >= (int)S->layouts.size()) QOB_FAIL(QOB_STATUS_INVALID_ARG, "bad layout plan id");
  if (S->ctx->device < 0) QOB_FAIL(QOB_STATUS_CUDA_ERROR, "planning-only context (no CUDA device): libqob200 has no CPU fallback");
  LazySumOp::LayoutPlan &lp = *S->layouts[plan_id];
  const int64_t n = 1ll << lp.nloc;
  cudaStream_t s = (cudaStream_t)stream;
  QLaunchOpts o;
  o.sm_budget = sm_budget;
  o.zadd = zadd;
  o.chunk_index = chunk_index;
  o.nchunks = nchunks < 1 ? 1 : nchunks;
  if (npeers > 0) {
    if (!is_pow2(npeers) || !x_peers || !y_peers) QOB_FAIL(QOB_STATUS_INVALID_ARG, "npeers must be a power of two with pointer tables");
    const int pb = ilog2(npeers);
    if (peer_shift < 0 || peer_shift + pb > lp.nloc) QOB_FAIL(QOB_STATUS_INVALID_ARG, "peer window outside the local index bits");
    if (zadd) QOB_FAIL(QOB_STATUS_INVALID_ARG, "zadd is not available in the peer-addressed form");
    for (int q = 0; q < npeers; ++q)
      if (!x_peers[q] || !y_peers[q]) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null peer pointer");
    o.npeers = npeers;
    o.peer_shift = peer_shift;
    o.peer_rank = (int)(lp.prog.h_hi_value() & (uint64_t)(npeers - 1));
    o.xpeer = x_peers;
    o.ypeer = y_peers;
  } else {
    if (!x || !y) QOB_FAIL(QOB_STATUS_INVALID_ARG, "null data pointer");
    QOB_TRY(check_alias(x, n, y, n));
    if (zadd) QOB_TRY(check_alias(zadd, n, y, n));
  }
  QOB_DEVICE(S->ctx->device);
  if (C(alpha) == ZERO) {
    if (npeers > 0) QOB_FAIL(QOB_STATUS_INVALID_ARG, "alpha == 0 in the peer-addressed form");
    if (zadd) return launch_axpby(zadd, y, n, ONE, C(beta), s);
    return launch_scale(y, n, C(beta), s);
  }
  std::vector<cplx> cf;
  {
    std::lock_guard<std::mutex> lk(S->coef_mu);
    cf = S->coefs;
  }
  // Which kernel.  The round-2 kernel (one persistent CTA fills an SM) runs every launch it supports: plain ones, peer-addressed
  // ones (the exchange: contiguous pieces over NVLink) and tile ranges.  Two kernels that run beside each other split the SMs:
  //   sm_budget  > 0 : at most that many SMs (the exchange);      sm_budget < -1 : all but |sm_budget| SMs (the local passes)
  //   sm_budget == -1: the round-1 kernel, whose small CTAs fill whatever room the other kernel leaves on any SM
  // (a full-grid round-2 launch beside a round-1 exchange serialised: 8 GPUs, N=33: 80 ms instead of 53 ms).
  static const bool qreg_peer_off = getenv("QOB_DIST_QREG") && atoi(getenv("QOB_DIST_QREG")) == 0;
  if (lp.has_qreg && !zadd && sm_budget != -1 && !(qreg_peer_off && (npeers > 0 || o.nchunks > 1 || sm_budget != 0))) {
    QRegOpts ro;
    ro.max_ctas = sm_budget > 0 ? sm_budget : (sm_budget < -1 ? std::max(1, S->ctx->sm_count + sm_budget) : 0);
    ro.npeers = o.npeers;
    ro.xpeer = o.xpeer;
    ro.ypeer = o.ypeer;
    ro.peer_shift = o.peer_shift;
    ro.peer_rank = o.peer_rank;
    ro.chunk_mask = lp.chunk_mask;
    ro.chunk_index = o.chunk_index;
    ro.nchunks = o.nchunks;
    if (qreg_supports(lp.qreg, ro)) {
      QOB_TRY(qreg_set_coefs(lp.qreg, cf, s));
      return qreg_launch_ex(lp.qreg, C(alpha), x, C(beta), y, s, ro);
    }
  }
  if (npeers > 0 && C(beta) != ZERO)
    QOB_FAIL(QOB_STATUS_UNSUPPORTED, "peer-addressed launch with beta != 0 needs the round-2 tile kernel (adds performed by the owner's L2); "
                                     "this plan runs the round-1 kernel, which can only store");
  if (o.sm_budget < 0) o.sm_budget = 0;
  QOB_TRY(qtile_set_coefs(lp.prog, cf, s));
  return qtile_launch(lp.prog, C(alpha), x, C(beta), y, s, &o);
}

}  // extern "C"
bool layout_plan_peer_qreg(qob_op *sum, int plan_id, int npeers, int peer_shift) {
  LazySumOp *S = nullptr;
  if (qubit_sum(sum, &S) != QOB_STATUS_OK || plan_id < 0 || plan_id >= (int)S->layouts.size()) return false;
  const LazySumOp::LayoutPlan &lp = *S->layouts[plan_id];
  if (!lp.has_qreg || (getenv("QOB_DIST_QREG") && atoi(getenv("QOB_DIST_QREG")) == 0)) return false;
  static const void *dummy[QT_MAXPEER_API] = {};
  QRegOpts ro;
  ro.npeers = npeers;
  ro.xpeer = dummy;
  ro.ypeer = (void *const *)dummy;
  ro.peer_shift = peer_shift;
  return qreg_supports(lp.qreg, ro);
}
extern "C" {

int qob_layout_plan_info(qob_op *sum, int32_t plan_id, int32_t *npasses, uint64_t *fixed_mask) {
  LazySumOp *S = nullptr;
  QOB_TRY(qubit_sum(sum, &S));
  if (plan_id < 0 || plan_id >= (int)S->layouts.size()) QOB_FAIL(QOB_STATUS_INVALID_ARG, "bad layout plan id");
  int n = 0;
  uint64_t m = 0;
  qtile_info(S->layouts[plan_id]->prog, &n, &m);
  if (npasses) *npasses = n;
  if (fixed_mask) *fixed_mask = m;
  return QOB_STATUS_OK;
}

int qob_layout_plan_set_chunk_bits(qob_op *sum, int32_t plan_id, uint64_t chunk_mask) {
  LazySumOp *S = nullptr;
  QOB_TRY(qubit_sum(sum, &S));
  if (plan_id < 0 || plan_id >= (int)S->layouts.size()) QOB_FAIL(QOB_STATUS_INVALID_ARG, "bad layout plan id");
  QOB_TRY(qtile_set_chunk_bits(S->layouts[plan_id]->prog, chunk_mask));
  S->layouts[plan_id]->chunk_mask = chunk_mask;
  return QOB_STATUS_OK;
}

int qob_layout_plan_describe(qob_op *sum, int32_t plan_id, char *buf, int64_t buflen) {
  LazySumOp *S = nullptr;
  QOB_TRY(qubit_sum(sum, &S));
  if (plan_id < 0 || plan_id >= (int)S->layouts.size() || !buf) QOB_FAIL(QOB_STATUS_INVALID_ARG, "bad layout plan id");
  const LazySumOp::LayoutPlan &lp = *S->layouts[plan_id];
  if (lp.has_qreg) snprintf(buf, (size_t)buflen, "%s (plain launches) / %s (peer-addressed, chunked or budgeted launches)", lp.qreg.describe.c_str(), lp.prog.describe.c_str());
  else snprintf(buf, (size_t)buflen, "%s", lp.prog.describe.c_str());
  return QOB_STATUS_OK;
}

}  // extern "C"
