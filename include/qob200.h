/*
 * qob200.h — C ABI of libqob200.so: B200-native (sm_100a) operator application
 * `mul!(result, op, state, alpha, beta)` for QuantumOpticsBase.jl's LazyTensor,
 * LazySum, LazyProduct, SparseOperator and dense Operator acting on Ket / Bra /
 * dense Operator data in ComplexF64.
 *
 * The reference has NO FFI for this path: the boundary is Julia multiple dispatch
 * on `LinearAlgebra.mul!` (reference src/QuantumOpticsBase.jl:4).  Every entry
 * point below cites the reference method(s) it replaces; `INTEGRATION.md` shows
 * the `ccall` stubs a maintainer adds (julia/QOB200.jl holds the full glue).
 *
 * Conventions (all taken from the reference):
 *   - ComplexF64 = two IEEE doubles (re, im), interleaved.
 *   - all dense data column-major; composite index has subsystem 1 fastest
 *     (src/states.jl:105, src/operators_dense.jl:134,296-308).
 *   - site indices are 1-based and sorted (src/operators_lazytensor.jl:26).
 *   - CSC arrays are Julia SparseMatrixCSC{ComplexF64,Int64}: colptr/rowval 1-based.
 *   - Y = alpha*A*B + beta*Y; alpha==0 -> only the beta update; beta==0 -> Y is
 *     never read (NaNs die) (src/operators_lazysum.jl:179-192,
 *     src/operators_lazytensor.jl:540, src/sparsematrix.jl:103-107).
 *   - state/result pointers are DEVICE pointers owned by the caller (CuPtr);
 *     operator definitions (factors, CSC arrays, coefficients) are HOST pointers
 *     copied at creation time.
 *   - every function returns a qob_status; qob_last_error() gives the message
 *     of the calling thread's last failure.
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     returns QOB_STATUS_CUDA_ERROR.
 */
#ifndef QOB200_H
#define QOB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QOB200_VERSION 100

typedef struct { double re, im; } qob_c64;

typedef enum {
  QOB_STATUS_OK = 0,
  QOB_STATUS_DIM_MISMATCH = 1,   /* Julia DimensionMismatch (sparsematrix.jl:100-102, operators_lazytensor.jl:695-703) */
  QOB_STATUS_ALIASING = 2,       /* Julia ArgumentError: result aliases an input (operators_lazytensor.jl:704-708) */
  QOB_STATUS_INVALID_ARG = 3,    /* Julia ArgumentError / AssertionError in constructors (operators_lazytensor.jl:23-32) */
  QOB_STATUS_UNSUPPORTED = 4,    /* Julia MethodError / "not implemented" (sparsematrix.jl:177-188) */
  QOB_STATUS_CUDA_ERROR = 5,
  QOB_STATUS_NCCL_ERROR = 6,
  QOB_STATUS_ALLOC = 7
} qob_status;

typedef enum { QOB_SIDE_LEFT = 0, QOB_SIDE_RIGHT = 1 } qob_side;

/* how a stored matrix enters: as is, transposed, or adjoint (Julia `Adjoint` wrapper,
 * operators_dense.jl:128, operators_sparse.jl:6) */
typedef enum { QOB_OP_N = 0, QOB_OP_T = 1, QOB_OP_C = 2 } qob_trans;

typedef enum {
  QOB_FACTOR_DENSE = 0,  /* Matrix{ComplexF64}, column-major nrows x ncols                 */
  QOB_FACTOR_CSC = 1,    /* SparseMatrixCSC{ComplexF64,Int64}                              */
  QOB_FACTOR_EYE = 2     /* FillArrays.Eye(nrows, ncols), possibly non-square (isometry)   */
} qob_factor_kind;

/* One site operator (`.data` of an Operator) as the reference stores it. HOST pointers. */
typedef struct {
  int32_t kind;           /* qob_factor_kind */
  int32_t trans;          /* qob_trans applied to the STORED matrix */
  int64_t nrows, ncols;   /* shape of the STORED matrix (before trans) */
  const qob_c64 *dense;   /* kind DENSE: nrows*ncols values, column-major */
  const int64_t *colptr;  /* kind CSC: ncols+1 entries, 1-based */
  const int64_t *rowval;  /* kind CSC: nnz entries, 1-based */
  const qob_c64 *nzval;   /* kind CSC: nnz entries */
} qob_factor;

typedef struct qob_ctx qob_ctx;  /* one per (process, device) */
typedef struct qob_op qob_op;    /* immutable operator handle (refcounted by sums/products that hold it) */

int qob_version(void);
const char *qob_last_error(void);
const char *qob_status_string(int status);

/* Device context: binds to CUDA device `device`, owns scratch (the analogue of the
 * reference's LRU temp cache, operators_lazytensor.jl:222-279, keyed per stream). */
int qob_ctx_create(int device, qob_ctx **out);
int qob_ctx_destroy(qob_ctx *ctx);
/* lazytensor_cachesize()/lazytensor_clear_cache() analogues (operators_lazytensor.jl:241-279) */
int qob_ctx_scratch_bytes(qob_ctx *ctx, int64_t *bytes);
int qob_ctx_clear_scratch(qob_ctx *ctx);

/* LazyTensor(bl, br, indices, operators, factor) — operators_lazytensor.jl:15-37.
 * dims_l/dims_r: per-subsystem dimensions (`_comp_size`, :517-518); sites: 1-based sorted. */
int qob_lazytensor_create(qob_ctx *ctx, int32_t nsub, const int64_t *dims_l, const int64_t *dims_r,
                          int32_t nfac, const int32_t *sites, const qob_factor *factors,
                          qob_c64 factor, qob_op **out);

/* SparseOperator (Operator{…,SparseMatrixCSC} or its lazy Adjoint) — operators_sparse.jl:5-23,199-202 */
int qob_sparse_create(qob_ctx *ctx, const qob_factor *m, qob_op **out);

/* DenseOperator used as an operator (BLAS path, operators_dense.jl:394-396). */
int qob_dense_create(qob_ctx *ctx, const qob_factor *m, qob_op **out);

/* LazySum(basis_l, basis_r, factors, operators) — operators_lazysum.jl:41-51.
 * dim_l/dim_r are the total dimensions (needed for the empty sum, :118-121,190-192). */
int qob_lazysum_create(qob_ctx *ctx, int64_t dim_l, int64_t dim_r, int32_t nterms,
                       const qob_c64 *coefs, qob_op *const *terms, qob_op **out);
/* TimeDependentSum set_time! rewrites LazySum.factors (time_dependent_operator.jl:279-290): cheap update. */
int qob_lazysum_set_coefs(qob_op *sum, int32_t nterms, const qob_c64 *coefs);

/* LazyProduct(operators, factor) — operators_lazyproduct.jl:32-52; temporaries are plan-owned
 * device buffers (the reference pre-allocates ket_l/bra_r, :12-21). */
int qob_lazyproduct_create(qob_ctx *ctx, int32_t nops, qob_op *const *ops, qob_c64 factor, qob_op **out);

/* LazyDirectSum(op1, op2, ...) — src/spinors.jl:158-169; mul! for Ket / Bra :221-247 (block i acts on slice i of the
 * SumBasis state; blocks must be square, batch must be 1 — the reference defines no other method). */
int qob_lazydirectsum_create(qob_ctx *ctx, int32_t nops, qob_op *const *ops, qob_op **out);

int qob_op_destroy(qob_op *op);
int qob_op_dims(const qob_op *op, int64_t *dim_l, int64_t *dim_r);

/* mul!(result, op, b, alpha, beta)  [side LEFT ]: y(dim_l x batch) = alpha*op*x(dim_r x batch) + beta*y
 * mul!(result, a, op, alpha, beta)  [side RIGHT]: y(batch x dim_r) = alpha*x(batch x dim_l)*op + beta*y
 * batch = 1 is the Ket (LEFT) / Bra (RIGHT) case.
 * Replaces: operators_lazytensor.jl:539-609, operators_lazysum.jl:189-238,
 * operators_lazyproduct.jl:103-163, operators_sparse.jl:199-202, operators_dense.jl:394-396.
 * x, y: device pointers to ComplexF64, column-major. stream: a cudaStream_t (0 = legacy default). */
int qob_op_apply(qob_op *op, int32_t side, qob_c64 alpha, const void *x, qob_c64 beta, void *y,
                 int64_t batch, void *stream);

/* Same call with HOST buffers: stages x (and y when beta != 0) to the device, applies, copies y back.
 * This is the end-to-end path bench.py times as `e2e`.  A LEFT-side batch of kets of at least
 * 2 x QOB_HOST_PIPE_MIN_BYTES (env, default 32 MiB; 0 disables) is streamed through the device in column
 * groups: upload of group j+1, kernels of group j and download of group j-1 run on separate streams, so the
 * call is bound by one direction of the host link.  Pinned (page-locked) host buffers are needed for the
 * overlap; pageable ones still give correct results. */
int qob_op_apply_host(qob_op *op, int32_t side, qob_c64 alpha, const qob_c64 *x, qob_c64 beta,
                      qob_c64 *y, int64_t batch);

/* Introspection used by tests/bench: number of kernel launches issued by this library in this
 * process, and a text description of the plan chosen for `op` (passes, tiles, kernels). */
int64_t qob_launch_count(void);
/* launches per kernel family: 1 round-1 tile kernel, 2 round-2 tile kernel (qreg), 3 round-2 tile kernel peer-addressed (the
 * exchange of a sharded apply over NVLink), 4 round-1 tile kernel peer-addressed; -1 for an unknown family */
int64_t qob_launch_count_of(int32_t family);
int qob_op_describe(qob_op *op, int32_t side, int64_t batch, char *buf, int64_t buflen);

/* Per-kernel timing for the roofline report: while enabled, every tile-pass launch is bracketed by
 * CUDA events on the launching stream.  qob_profile_read synchronises those events and returns, in
 * launch order, the duration (ms), the pass index and the algorithmic bytes (32 B/amplitude for a
 * pass that only writes y, 48 B/amplitude for a read-modify-write pass) of each launch, then clears. */
int qob_profile_enable(int32_t on);
int qob_profile_read(int32_t max_entries, float *ms, int32_t *pass_index, double *alg_bytes, int32_t *count);

/* Fused master-equation right-hand side (no single reference function — the reference builds
 * it from six mul! calls per step, test/test_sciml_broadcast_interfaces.jl:36-43):
 *   drho = alpha * ( -i (H rho - rho H) + sum_k r_k ( J_k rho J_k^+ - 1/2 (J_k^+ J_k rho + rho J_k^+ J_k) ) ) + beta * drho
 * H, J_k: square D x D factors (CSC as SparseOperator holds them, or dense); rates r_k >= 0 (NULL: all 1).
 * rho, drho: device pointers to D x D ComplexF64, column-major; they must not alias.  alpha == 0 leaves only the beta update,
 * beta == 0 never reads drho.  qob_lindblad_dense writes one of the assembled host matrices (0: H - i/2 sum r_k J_k^+ J_k,
 * 1: H + i/2 sum r_k J_k^+ J_k, 2+k: sqrt(r_k) J_k) as a dense column-major D x D array (tests, introspection). */
int qob_lindblad_create(qob_ctx *ctx, const qob_factor *H, int32_t nJ, const qob_factor *J, const double *rates, qob_op **out);
int qob_lindblad_apply(qob_op *L, qob_c64 alpha, const void *rho, qob_c64 beta, void *drho, void *stream);
int qob_lindblad_dense(qob_op *L, int32_t which, qob_c64 *out);

/* Counter-based synthetic input generator shared with the oracle (oracle/qob_oracle.c:orc_fill_state):
 * x[i] = scale * (u(seed, 2i), u(seed, 2i+1)), u uniform in [-1, 1) from splitmix64. */
int qob_fill_state(void *x, int64_t offset, int64_t n, uint64_t seed, double scale, void *stream);
/* sum |x|^2 and <x|y> reductions for size-independent parity properties (device result -> host). */
int qob_norm2(const void *x, int64_t n, double *out, void *stream);
int qob_dot(const void *x, const void *y, int64_t n, qob_c64 *out, void *stream);

/* expect(op, psi) = dot(psi, op*psi) and variance(op, psi) = psi'(op(op psi)) - (psi'(op psi))^2 for a Ket
 * (src/operators.jl:119,139-142).  x: device pointer to dim_r ComplexF64; the scalar result is written to HOST memory
 * (the call synchronises `stream`).  The intermediate op*psi lives in the handle's scratch (per stream). */
int qob_expect(qob_op *op, const void *x, qob_c64 *out, void *stream);
int qob_variance(qob_op *op, const void *x, qob_c64 *out, void *stream);

/* ptrace(a::DataOperator, indices) / ptrace(psi::Ket, indices) / ptrace(psi::Bra, indices)
 * (src/operators_dense.jl:191-215, generated loop nests :311-383).  `traced`: ntraced distinct 1-based subsystem indices.
 * qob_ptrace_op:    a = dense prod(dims_l) x prod(dims_r) device matrix -> result (kept dims_l) x (kept dims_r), column-major;
 *                   traced subsystems need dims_l == dims_r.
 * qob_ptrace_state: psi = device vector over dims -> result M x M with M = prod(kept dims): psi psi^+ summed over the traced
 *                   subsystems (is_bra != 0: conj(psi[Il]) * psi[Ir], the reference's Bra method).
 * Errors follow check_ptrace_arguments (src/operators.jl:153-176): QOB_STATUS_INVALID_ARG when all subsystems are traced, an
 * index is out of range or repeated, or a traced subsystem is not square. */
int qob_ptrace_op(qob_ctx *ctx, int32_t nsub, const int64_t *dims_l, const int64_t *dims_r, int32_t ntraced,
                  const int32_t *traced, const void *a, void *result, void *stream);
int qob_ptrace_state(qob_ctx *ctx, int32_t nsub, const int64_t *dims, int32_t ntraced, const int32_t *traced,
                     int32_t is_bra, const void *psi, void *result, void *stream);

/* ---- sharded (multi-GPU) LazySum apply: one process per GPU -----------------------------------
 * The state is sharded on its highest-stride axes: rank r of P=2^p owns the contiguous slab
 * [r*D/P, (r+1)*D/P) of the reference's linear array, i.e. the top p index bits are the rank.
 * The library provides the per-rank compute for ANY index layout; the host (python/dist.py, or the
 * Julia glue) owns the exchange step (NCCL all-to-all axis swap) between layouts:
 *   - qob_lazysum_term_masks: which subsystems a LazyTensor term touches / touches off-diagonally
 *     (a diagonal factor on a sharded axis needs no communication, only a rank-dependent weight);
 *   - qob_layout_plan_create: tile program for the selected terms of a spin-1/2 LazySum when
 *     subsystem k's index bit sits at position bitpos[k] of a virtual index whose positions
 *     < nbits_local address the local buffer and whose positions >= nbits_local are constant on this
 *     rank (bit j of hi_value = position nbits_local + j).  Off-diagonal factors must be local.
 *   - qob_layout_plan_apply: y = alpha * (selected terms) x + beta * y on the local buffers. */
int qob_lazysum_term_masks(qob_op *sum, int32_t term, uint64_t *offdiag_mask, uint64_t *site_mask);
int qob_layout_plan_create(qob_op *sum, int32_t nbits_local, const int32_t *bitpos, uint64_t hi_value,
                           const uint8_t *term_select, int32_t *plan_id);
int qob_layout_plan_apply(qob_op *sum, int32_t plan_id, qob_c64 alpha, const void *x, qob_c64 beta, void *y,
                          void *stream);
int qob_layout_plan_describe(qob_op *sum, int32_t plan_id, char *buf, int64_t buflen);
/* Extended form used by the fused exchange (compute + NVLink peer memory in ONE kernel, no staging copy, no NCCL):
 *   zadd     (optional) extra addend in the local layout: y = alpha*(terms)x + beta*y + zadd, folded into the last pass;
 *   npeers>0 the plan's buffer is the SWAPPED layout of the ranks' slabs: index bits [peer_shift, peer_shift+log2 npeers)
 *            of an address select the owning rank; the tile kernel loads x straight from x_peers[owner] and stores its
 *            result straight into y_peers[owner] (device pointers into every rank's symmetric / IPC-mapped memory).
 *            x and y are ignored in that case.
 *            beta = 0: the results are stored (every element of the owners' buffers is written exactly once);
 *            beta = 1: the results are ADDED into the owners' buffers (f64 add performed by the owner's L2; needs the
 *            round-2 tile kernel: slabs of >= 2^20 amplitudes, contiguous pieces of >= 1 KiB); other values: INVALID_ARG.
 *   sm_budget>0 limits the launch to that many SMs so that another kernel runs beside it (the exchange pass);
 *   sm_budget<-1 leaves |sm_budget| SMs free: the launch runs on all the others (the local passes beside the exchange).  One
 *            persistent CTA of the round-2 tile kernel fills an SM, so two launches split this way never share an SM;
 *   sm_budget=-1 says that this launch runs beside another kernel and should SHARE SMs with it: the round-1 tile kernel,
 *            whose small CTAs fit next to another kernel's (the pre-round-2 schedule);
 *   chunk_index/nchunks: run one of nchunks equal tile ranges (nchunks <= 1: the whole pass). */
int qob_layout_plan_apply_ex(qob_op *sum, int32_t plan_id, qob_c64 alpha, const void *x, qob_c64 beta, void *y,
                             const void *zadd, int32_t npeers, const void *const *x_peers, void *const *y_peers,
                             int32_t peer_shift, int32_t sm_budget, int32_t chunk_index, int32_t nchunks, void *stream);
/* Chunked launches (to pipeline the fold-in of received contributions behind the exchange): a plan with exactly one pass
 * can be run on chunk `chunk_index` of `nchunks` equal ranges of its tiles.  qob_layout_plan_info reports how many passes
 * hold work and which index bits are fixed (not free) in all of them; qob_layout_plan_set_chunk_bits makes the given
 * fixed bits the most significant bits of the tile numbering, so that chunk c covers the same amplitudes (those bits = c)
 * in every plan configured with the same chunk bits. */
int qob_layout_plan_info(qob_op *sum, int32_t plan_id, int32_t *npasses, uint64_t *fixed_mask);
int qob_layout_plan_set_chunk_bits(qob_op *sum, int32_t plan_id, uint64_t chunk_mask);
/* ---- the whole sharded apply behind the ABI: planning, CUDA-IPC mapping, device-side barriers, stream choreography ------
 * For hosts without torch.distributed (the Julia glue, a C program): each process
 *   1. builds the same LazySum and calls qob_dist_create(sum, rank, world);
 *   2. allocates its state slab, contribution slab and signal pad with qob_dist_alloc (sizes from qob_dist_info; the pad must be
 *      zeroed), exports them with qob_ipc_export (64-byte cudaIpcMemHandle_t), moves the handles to the other processes by any
 *      means it has (MPI, sockets, files), maps the peers' buffers with qob_ipc_open, and hands all pointers to qob_dist_bind;
 *   3. calls qob_dist_apply(dist, alpha, beta, y, stream) collectively: y_local = alpha*(H x)_local + beta*y_local with x = its
 *      bound state slab.  The exchange of the terms that act on sharded axes is fused into the tile kernel (peer loads / stores
 *      over NVLink); the ranks meet in device-side barriers (a kernel that signals every peer's pad and waits on its own); the
 *      host never blocks and nothing goes through NCCL.
 * IPC / peer-mapping failures return QOB_STATUS_NCCL_ERROR (the "communication layer" status). */
typedef struct qob_dist qob_dist;
int qob_dist_alloc(qob_ctx *ctx, int64_t bytes, void **ptr);
int qob_dist_free(qob_ctx *ctx, void *ptr);
int qob_ipc_export(void *ptr, uint8_t *handle64);
int qob_ipc_open(qob_ctx *ctx, const uint8_t *handle64, void **ptr);
int qob_ipc_close(qob_ctx *ctx, void *ptr);
int qob_dist_create(qob_op *sum, int32_t rank, int32_t world, qob_dist **out);
int qob_dist_info(qob_dist *d, int32_t *nbits_local, int32_t *n_exchanged_terms, int32_t *nchunks, int64_t *slab_bytes,
                  int64_t *flag_bytes);
int qob_dist_bind(qob_dist *d, void *const *x_peers, void *const *z_peers, void *const *flag_peers);
/* Direct mode (when qob_dist_direct_capable says yes): also allocate the RESULT slab with qob_dist_alloc, exchange its handle
 * like the state's and pass the table to qob_dist_bind_result.  qob_dist_apply with y = this rank's entry then needs no
 * contribution slab (z_peers may be NULL in qob_dist_bind): the exchange pass adds its results into the owners' result slabs
 * (cp.reduce.async.bulk, an f64 add performed by the owner's L2), so a rank holds 2 slabs instead of 3 — N=33 fits on 2 GPUs.
 * The order in which the contributions are added is not fixed: results are reproducible to rounding, not bit for bit. */
int qob_dist_direct_capable(qob_dist *d, int32_t *yes);
int qob_dist_bind_result(qob_dist *d, void *const *y_peers);
int qob_dist_apply(qob_dist *d, qob_c64 alpha, qob_c64 beta, void *y, void *stream);
int qob_dist_describe(qob_dist *d, char *buf, int64_t buflen);
/* CUDA-event timing of the exchange step (first barrier passed -> every rank's contributions landed) of the applies issued while
 * enabled; qob_dist_exchange_ms returns the mean over them and the bytes that cross this GPU's NVLink per direction per apply. */
int qob_dist_exchange_timing(qob_dist *d, int32_t enable);
int qob_dist_exchange_ms(qob_dist *d, double *mean_ms, int32_t *count, int64_t *bytes_per_direction);
int qob_dist_destroy(qob_dist *d);

/* SMs the persistent tile kernels may occupy by default (0 = all). */
int qob_set_sm_budget(int32_t sms);

#ifdef __cplusplus
}
#endif
#endif /* QOB200_H */
