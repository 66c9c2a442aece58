// qob_internal.h — shared declarations of libqob200.so (host side of the engine).
// Not part of the public ABI; see include/qob200.h for that.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <complex>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/qob200.h"

typedef std::complex<double> cplx;

// ---------------------------------------------------------------- errors
void qob_set_error(const char *fmt, ...);
#define QOB_FAIL(code, ...)     \
  do {                          \
    qob_set_error(__VA_ARGS__); \
    return (code);              \
  } while (0)
#define QOB_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      QOB_FAIL(QOB_STATUS_CUDA_ERROR, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
               __FILE__, __LINE__);                                                           \
  } while (0)
#define QOB_TRY(expr)                 \
  do {                                \
    int s__ = (expr);                 \
    if (s__ != QOB_STATUS_OK) return s__; \
  } while (0)

#define QOB_MAX_DEVICES 64
// SM count of the CURRENT device (cached per device; a process may drive several devices, one qob_ctx each)
int qob_device_sm_count();

extern std::atomic<int64_t> g_launch_count;
// set by the ABI entry points while they work for a planning-only context (device < 0): device uploads
// become no-ops so that validation and planning can be exercised (and tested) on a box without a GPU.
extern thread_local bool t_planning_only;
#define QOB_LAUNCHED() (g_launch_count.fetch_add(1, std::memory_order_relaxed))
// launches per kernel family (qob_launch_count_of): 0 gather, 1 round-1 tile kernel, 2 round-2 tile kernel, 3 round-2 tile kernel
// peer-addressed (exchange over NVLink), 4 round-1 tile kernel peer-addressed
extern std::atomic<int64_t> g_family_count[8];
#define QOB_LAUNCHED_FAMILY(f) (g_family_count[f].fetch_add(1, std::memory_order_relaxed))

// ---------------------------------------------------------------- host matrices
// A site factor after `trans` has been applied, in one of three normal forms.
struct HostMat {
  int kind = QOB_FACTOR_DENSE;  // qob_factor_kind
  int64_t rows = 0, cols = 0;
  std::vector<cplx> dense;             // column-major rows x cols   (kind DENSE)
  std::vector<int64_t> colptr, rowidx; // 0-based CSC                (kind CSC)
  std::vector<cplx> vals;
  cplx at(int64_t i, int64_t j) const; // slow accessor (tests / tiny factors)
  bool is_square_eye() const { return kind == QOB_FACTOR_EYE && rows == cols; }
  HostMat transposed() const;          // plain transpose (no conjugate)
  // CSR rows of this matrix: for dense every entry (explicit zeros included, like BLAS),
  // for CSC the stored entries, for Eye the min(rows, cols) ones.
  void to_csr(std::vector<int32_t> &rowptr, std::vector<int32_t> &colidx, std::vector<cplx> &v) const;
  int64_t max_row_nnz() const;
};
int hostmat_from_factor(const qob_factor *f, HostMat &out);

// ---------------------------------------------------------------- context / scratch
struct DevBuf {
  void *ptr = nullptr;
  size_t bytes = 0;
};
struct qob_ctx {
  int device = 0;
  int sm_count = 148;
  size_t smem_optin = 0;
  std::atomic<int> live_ops{0};   // operator handles created on this context and not yet destroyed
  std::mutex mu;
  // scratch slots keyed by (stream, slot): the analogue of the reference's LRU temp cache keyed by
  // (stage symbol, task id) — operators_lazytensor.jl:303-315
  std::map<std::pair<cudaStream_t, int>, DevBuf> scratch;
  cudaStream_t pipe_streams[4] = {nullptr, nullptr, nullptr, nullptr};   // up, down, two compute lanes (qob_op_apply_host)
  int get_scratch(cudaStream_t s, int slot, size_t bytes, void **out);
  int64_t scratch_bytes();
  void clear_scratch();
};

// ---------------------------------------------------------------- device-program pieces
template <class T>
struct DevArray {  // owning device array uploaded from a host vector
  T *ptr = nullptr;
  size_t n = 0;
  DevArray() {}
  DevArray(const DevArray &) = delete;
  DevArray &operator=(const DevArray &) = delete;
  ~DevArray() {
    if (ptr) cudaFree(ptr);
  }
  int upload(const std::vector<T> &h) {
    if (t_planning_only) {
      n = h.size();
      return QOB_STATUS_OK;
    }
    if (ptr && n < h.size()) {
      cudaFree(ptr);
      ptr = nullptr;
    }
    n = h.size();
    if (!ptr && n) QOB_CUDA(cudaMalloc(&ptr, n * sizeof(T)));
    if (n) QOB_CUDA(cudaMemcpy(ptr, h.data(), n * sizeof(T), cudaMemcpyHostToDevice));
    return QOB_STATUS_OK;
  }
  int upload_async(const std::vector<T> &h, cudaStream_t s) {
    if (t_planning_only) return QOB_STATUS_OK;
    if (h.size() > n || !ptr) return upload(h);
    QOB_CUDA(cudaMemcpyAsync(ptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return QOB_STATUS_OK;
  }
};

// ---------------------------------------------------------------- kernels (qob_kernels_*.cu)

// y[i] = beta*y[i] (beta == 0 -> zero fill without reading)
int launch_scale(void *y, int64_t n, cplx beta, cudaStream_t s);
// y = alpha*x + beta*y
int launch_axpby(const void *x, void *y, int64_t n, cplx alpha, cplx beta, cudaStream_t s);

// Dense single-axis contraction (DMMA): y[l,i,r] = alpha*sum_j A[i,j] x[l,j,r] + beta*y[l,i,r]
// A: device, row-major (i, j) split planes prepared by prepare_axis_matrix.
struct AxisMatrixDev {
  DevArray<double> planes;  // [2][dl_pad][dr_pad] row-major, re plane then im plane
  int dl = 0, dr = 0, dl_pad = 0, dr_pad = 0;
};
int prepare_axis_matrix(const HostMat &m, AxisMatrixDev &out);
int launch_axis_dense(const AxisMatrixDev &A, int64_t L, int64_t R, cplx alpha, const void *x, cplx beta,
                      void *y, cudaStream_t s);

// Sparse x dense.  left: R(m x n) = beta R + alpha S(m x k) B(k x n), S given as CSR (rowptr over m rows).
// right: R(q x n) = beta R + alpha B(q x m) S(m x n), S given as CSC (colptr over n columns).
struct SparseDev {
  DevArray<int32_t> ptr, idx;
  DevArray<double2> val;
  DevArray<int32_t> order;   // CSC only, optional: column processing order of the right-side SpMM (Cuthill-McKee)
  int64_t nptr = 0;
};
int launch_spmm_left(const SparseDev &csr, int64_t m, int64_t k, int64_t n, cplx alpha, const void *B, cplx beta,
                     void *R, cudaStream_t s);
int launch_spmm_right(const SparseDev &csc, int64_t q, int64_t m, int64_t n, cplx alpha, const void *B, cplx beta,
                      void *R, cudaStream_t s);

// Fused master-equation right-hand side (qob_kernels_lindblad.cu): Heff = H - i/2 sum_k J_k^+ J_k as CSR rows,
// G = H + i/2 sum_k J_k^+ J_k by columns (CSC), the jump operators (times sqrt(rate)) as CSR rows sharing one index / value
// array, row pointers of operator k at k*(D+1).
struct LindbladDev {
  int64_t D = 0;
  int nJ = 0;
  DevArray<int32_t> h_ptr, h_col, g_ptr, g_row, j_ptr, j_col;
  DevArray<double2> h_val, g_val, j_val;
};
int launch_lindblad(const LindbladDev &L, cplx alpha, const void *rho, cplx beta, void *drho, cudaStream_t s);

// Generic fused gather program --------------------------------------------------------------
struct GatherProgram {
  // host description
  int64_t d_out = 0, d_in = 0;  // tensor sizes (without batch)
  int nterms = 0;
  int max_fac = 0;
  std::vector<int> coef_of_term;  // coef index per compiled term (for set_coefs)
  std::vector<cplx> scalars;      // per compiled term
  // device
  DevArray<int32_t> d_i32;  // packed int tables
  DevArray<int64_t> d_i64;
  DevArray<double2> d_vals;
  DevArray<double2> d_coef;  // per term: coef*scalar (alpha applied in kernel)
  size_t off_fac_i = 0, off_rowptr = 0, off_colidx = 0, off_segs = 0;  // section offsets inside d_i32 / d_i64
  std::string describe;
};
// dims_out/dims_in per axis; mats already oriented (rows = out index, cols = in index).
struct OrientedTerm {
  int coef_index;
  cplx scalar;
  std::vector<int> axes;
  std::vector<HostMat> mats;
};
int gather_program_build(GatherProgram &p, const std::vector<int64_t> &dims_out, const std::vector<int64_t> &dims_in,
                         const std::vector<OrientedTerm> &terms);
int gather_program_set_coefs(GatherProgram &p, const std::vector<cplx> &coefs, cudaStream_t s);
int gather_program_launch(const GatherProgram &p, int64_t pre, int64_t post, cplx alpha, const void *x, cplx beta,
                          void *y, cudaStream_t s);

// Qubit tile program ------------------------------------------------------------------------
struct QTerm {  // a term on bit positions of the flat index (all dims 2)
  int coef_index;
  cplx scalar;
  std::vector<int> bits;          // physical bit positions, ascending
  std::vector<cplx> m;            // per factor 4 entries row-major: [a00 a01 a10 a11] (out row, in col)
};
struct QTileProgramHost;
struct QTileProgram {
  std::shared_ptr<QTileProgramHost> h;
  std::string describe;
  int npasses = 0;
  uint64_t hi_value = 0;
  uint64_t h_hi_value() const { return hi_value; }
};
// nbits: bits of the local flat index; hi_value: value of index bits >= nbits (rank offset for sharded states)
int qtile_build(QTileProgram &p, int nbits, uint64_t hi_value, const std::vector<QTerm> &terms, int sm_count);
int qtile_set_coefs(QTileProgram &p, const std::vector<cplx> &coefs, cudaStream_t s);
struct QLaunchOpts {
  int sm_budget = 0;              // SMs this launch may occupy (0: library default, see qob_set_sm_budget)
  const void *zadd = nullptr;     // extra addend (local layout), applied by the last pass
  int npeers = 0;                 // > 0: x and y are addressed across ranks (swapped layout over peer memory)
  int peer_shift = 0, peer_rank = 0;
  const void *const *xpeer = nullptr;
  void *const *ypeer = nullptr;
  int chunk_index = 0, nchunks = 1;  // run only chunk `chunk_index` of `nchunks` equal tile ranges (single-pass plans)
};
void qtile_info(const QTileProgram &p, int *npasses, uint64_t *fixed_mask);
int qtile_set_chunk_bits(QTileProgram &p, uint64_t chunk_mask);
int qtile_launch(const QTileProgram &p, cplx alpha, const void *x, cplx beta, void *y, cudaStream_t s,
                 const QLaunchOpts *opts = nullptr);

// Register-blocked, TMA-staged, L2-chained tile program (qob_kernels_qreg.cu): the single-GPU path for >= 2^20 amplitudes
struct QRegProgramHost;
struct QRegProgram {
  std::shared_ptr<QRegProgramHost> h;
  std::string describe;
  int npasses = 0;
};
// Returns QOB_STATUS_UNSUPPORTED when the terms do not fit the scheme (the caller then uses the qtile kernel).
int qreg_build(QRegProgram &p, int nbits, uint64_t hi_value, const std::vector<QTerm> &terms, int sm_count);
int qreg_set_coefs(QRegProgram &p, const std::vector<cplx> &coefs, cudaStream_t s);
// max_ctas > 0: the launch occupies at most that many SMs (one persistent CTA each); the others stay free for a kernel of
// another stream (the fused exchange of a sharded apply and the local passes share the GPU this way)
int qreg_launch(const QRegProgram &p, cplx alpha, const void *x, cplx beta, void *y, cudaStream_t s, int max_ctas = 0);
// Peer-addressed launch / launch over one tile range (the exchange step of a sharded apply; qob_layout_plan_apply_ex).
struct QRegOpts {
  int max_ctas = 0;
  int npeers = 0;                       // > 0: x and y are the SWAPPED layout spread over the ranks' slabs
  const void *const *xpeer = nullptr;
  void *const *ypeer = nullptr;
  int peer_shift = 0, peer_rank = 0;
  uint64_t chunk_mask = 0;              // nchunks > 1: these fixed index bits number the tile ranges
  int chunk_index = 0, nchunks = 1;
};
// can this program run peer-addressed / chunked with these options? (single pass, contiguous pieces of >= 1 KiB, ...)
bool qreg_supports(const QRegProgram &p, const QRegOpts &o);
// can layout plan `plan_id` of a LazySum run its peer-addressed launches on the round-2 kernel (bulk-copy pieces, adds into the
// owners' buffers)?  (qob_api.cu; used by the sharded apply to choose its schedule)
struct qob_op;
bool layout_plan_peer_qreg(qob_op *sum, int plan_id, int npeers, int peer_shift);
int qreg_launch_ex(const QRegProgram &p, cplx alpha, const void *x, cplx beta, void *y, cudaStream_t s, const QRegOpts &o);

// <x| op |x> in one sweep over x (no result vector is written): every pass reduces conj(x)*acc; deterministic
int qreg_expect(const QRegProgram &p, const void *x, cplx *out, cudaStream_t s);

// per-launch event timing of the tile kernels (qob_profile_enable / qob_profile_read); thread safe
bool qprof_enabled();
void *qprof_begin(cudaStream_t s, int pass, double alg_bytes);
void qprof_end(cudaStream_t s, void *token);

// Mixed-radix tile program (any subsystem dimensions, large states) --------------------------
struct DTileProgramHost;
struct DTileProgram {
  std::shared_ptr<DTileProgramHost> h;
  std::string describe;
  int npasses = 0;
};
// dims: every tensor axis of the state incl. the batch axis (fastest first); terms: square oriented factors on those axes.
// Returns QOB_STATUS_UNSUPPORTED when the terms do not fit the scheme (the caller then uses the gather kernel).
// `declined` (optional) receives the indices of terms the scheme cannot take (too many components, larger than a
// tile): the program then covers the others and the caller applies the declined ones another way.
int dtile_build(DTileProgram &p, const std::vector<int64_t> &dims, const std::vector<OrientedTerm> &terms,
                std::vector<int> *declined = nullptr);
int dtile_set_coefs(DTileProgram &p, const std::vector<cplx> &coefs, cudaStream_t s);
int dtile_launch(const DTileProgram &p, cplx alpha, const void *x, cplx beta, void *y, cudaStream_t s);

// Partial traces (qob_kernels_ptrace.cu); `slot` names the context scratch slot used for offset tables / partial sums
int launch_ptrace_op(qob_ctx *ctx, int slot, int nsub, const int64_t *dims_l, const int64_t *dims_r, int ntraced, const int32_t *traced,
                     const void *a, void *result, cudaStream_t s);
int launch_ptrace_state(qob_ctx *ctx, int slot, int nsub, const int64_t *dims, int ntraced, const int32_t *traced, bool bra, const void *psi,
                        void *result, cudaStream_t s);

// the context an operator handle was created on (qob_dist.cu)
qob_ctx *qob_op_context(const qob_op *op);

// misc device helpers
int launch_fill_state(void *x, int64_t offset, int64_t n, uint64_t seed, double scale, cudaStream_t s);
int launch_norm2(const void *x, int64_t n, double *host_out, cudaStream_t s);
int launch_dot(const void *x, const void *y, int64_t n, cplx *host_out, cudaStream_t s);
