// qob_kernels_lindblad.cu — fused master-equation right-hand side for sparse H and sparse jump operators on a dense rho
// (SURVEY.md §8f row 3).  The reference builds it from mul! calls (test/test_sciml_broadcast_interfaces.jl:36-43; two
// sparse gemm! for the commutator, four more per jump operator), each a full pass over rho / drho.  Here
//     drho = alpha * ( -i (Heff rho - rho G) + sum_k J_k rho J_k^+ ) + beta * drho,
//     Heff = H - i/2 sum_k J_k^+ J_k,  G = H + i/2 sum_k J_k^+ J_k  (= Heff^+ for Hermitian H; H is NOT assumed Hermitian:
//     the reference's call pattern is -i (H rho - rho H) for any H)
// is ONE kernel: every thread owns one element (i, j) of drho (i fastest: the coalesced direction of Julia's column-major
// matrices) and gathers
//     sum_{p in row i of Heff} Heff[p] rho[col p, j]           (times -i)
//   + sum_{q in column j of G} rho[i, row q] G[q]                (times +i; G is stored by columns = CSR of G^T)
//   + sum_k sum_{p in row i of J_k} sum_{q in row j of J_k} J_k[p] conj(J_k[q]) rho[col p, col q]
// from rho through L1/L2: rho is read from DRAM about once per distinct gather pattern, drho is written once.
// The rates are folded into J_k (sqrt) on the host; Heff and the CSR rows are built on the host (qob_api.cu).
#include "qob_internal.h"

__device__ __forceinline__ void lfma(double2 &acc, double2 a, double2 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}

__global__ void __launch_bounds__(256)
    lindblad_kernel(long long D, const int *__restrict__ h_ptr, const int *__restrict__ h_col, const double2 *__restrict__ h_val,
                    const int *__restrict__ g_ptr, const int *__restrict__ g_row, const double2 *__restrict__ g_val, int nJ, const int *__restrict__ j_ptr, const int *__restrict__ j_col, const double2 *__restrict__ j_val,
                    double2 alpha, double2 beta, int beta_zero, const double2 *__restrict__ rho, double2 *__restrict__ drho,
                    unsigned row_blocks) {
  const unsigned rb = blockIdx.x % row_blocks;
  const long long j = blockIdx.x / row_blocks;
  const long long i = (long long)rb * 256 + threadIdx.x;
  if (i >= D) return;
  const double2 *rcol = rho + j * D;   // column j of rho
  const double2 *rrow = rho + i;       // row i of rho (stride D)
  double2 t1 = make_double2(0.0, 0.0), t2 = make_double2(0.0, 0.0), acc = make_double2(0.0, 0.0);
  for (int p = h_ptr[i]; p < h_ptr[i + 1]; ++p) lfma(t1, h_val[p], rcol[h_col[p]]);
  for (int q = g_ptr[j]; q < g_ptr[j + 1]; ++q) lfma(t2, g_val[q], rrow[(long long)g_row[q] * D]);
  // -i*t1 + i*t2
  acc.x = t1.y - t2.y;
  acc.y = t2.x - t1.x;
  for (int k = 0; k < nJ; ++k) {
    const int *ptr = j_ptr + (long long)k * (D + 1);
    const int p0 = ptr[i], p1 = ptr[i + 1], q0 = ptr[j], q1 = ptr[j + 1];
    for (int q = q0; q < q1; ++q) {
      const double2 vq = j_val[q];
      const double2 *rc = rho + (long long)j_col[q] * D;
      double2 inner = make_double2(0.0, 0.0);
      for (int p = p0; p < p1; ++p) lfma(inner, j_val[p], rc[j_col[p]]);
      lfma(acc, make_double2(vq.x, -vq.y), inner);
    }
  }
  double2 o = make_double2(alpha.x * acc.x - alpha.y * acc.y, alpha.x * acc.y + alpha.y * acc.x);
  double2 *out = drho + j * D + i;
  if (!beta_zero) lfma(o, beta, *out);
  *out = o;
}

int launch_lindblad(const LindbladDev &L, cplx alpha, const void *rho, cplx beta, void *drho, cudaStream_t s) {
  if (L.D == 0) return QOB_STATUS_OK;
  const int64_t row_blocks = (L.D + 255) / 256, blocks = row_blocks * L.D;
  if (blocks >= ((int64_t)1 << 31)) QOB_FAIL(QOB_STATUS_UNSUPPORTED, "Lindblad operator too large for one launch");
  lindblad_kernel<<<(unsigned)blocks, 256, 0, s>>>(L.D, L.h_ptr.ptr, L.h_col.ptr, L.h_val.ptr, L.g_ptr.ptr, L.g_row.ptr,
                                                  L.g_val.ptr, L.nJ, L.j_ptr.ptr, L.j_col.ptr,
                                                  L.j_val.ptr, make_double2(alpha.real(), alpha.imag()),
                                                  make_double2(beta.real(), beta.imag()), beta == cplx(0.0, 0.0),
                                                  (const double2 *)rho, (double2 *)drho, (unsigned)row_blocks);
  QOB_LAUNCHED();
  QOB_CUDA(cudaGetLastError());
  return QOB_STATUS_OK;
}
