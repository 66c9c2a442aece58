"""
qob_oracle.py — CPU restatement (numpy/OpenBLAS + oracle/qob_oracle.c) of the reference's
`mul!(result, op, state, alpha, beta)` hot path, object model included.

THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product (qob200 / libqob200.so)
never does.  It never reads /root/reference at run time.

Pinning status: the reference is Julia, which is absent from this image and from the GPU box,
and its own tests store no golden vectors (inputs come from Julia's RNG).  The reference pins
this path by IDENTITIES (lazy/sparse result == explicit dense-kron result, known-answer site
operators).  tests/test_oracle_identities.py re-runs those identities on this file; see
DESIGN.md "Oracle".

Citations are relative to /root/reference.  Layout facts: subsystem 1 is the fastest index
(src/states.jl:105, src/operators_dense.jl:134), everything column-major.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
C128 = np.complex128


class DimensionMismatch(Exception):
    """Julia DimensionMismatch."""


class IncompatibleBases(Exception):
    """QuantumInterface.IncompatibleBases."""


class ArgumentError(Exception):
    """Julia ArgumentError."""


class MethodError(Exception):
    """Julia MethodError (no method for these operand types)."""


def build(force: bool = False) -> str:
    """Compile oracle/qob_oracle.c -> oracle/libqob_oracle.so (gcc, -O2, scalar like the reference)."""
    src = os.path.join(_HERE, "qob_oracle.c")
    out = os.path.join(_HERE, "libqob_oracle.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-march=native", "-fPIC", "-shared", "-std=c11", "-o", out, src, "-lm"])
    return out


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _sc(z):
    return np.array([complex(z).real, complex(z).imag], dtype=np.float64)


# --------------------------------------------------------------------------------------
# data wrappers
# --------------------------------------------------------------------------------------
class Eye:
    """FillArrays.Eye(m, n) — possibly non-square (src/operators_lazytensor.jl:436-440)."""

    def __init__(self, m, n=None):
        self.shape = (int(m), int(m if n is None else n))


class Adj:
    """LinearAlgebra.Adjoint wrapper (lazy `dagger`, src/operators_dense.jl:128)."""

    def __init__(self, parent):
        self.parent = parent
        self.shape = (parent.shape[1], parent.shape[0])


def is_sparse(d):
    return sp.issparse(d)


def materialize(d):
    """Dense ndarray (F-order) of any data wrapper."""
    if isinstance(d, Adj):
        return np.asfortranarray(materialize(d.parent).conj().T)
    if isinstance(d, Eye):
        return np.asfortranarray(np.eye(d.shape[0], d.shape[1], dtype=C128))
    if sp.issparse(d):
        return np.asfortranarray(d.toarray().astype(C128))
    return np.asfortranarray(np.asarray(d, dtype=C128))


def _transpose_data(d):
    """`transpose(op.data)` as used by op_transform=transpose (src/operators_lazytensor.jl:569,604)."""
    if isinstance(d, Adj):
        p = d.parent
        if isinstance(p, Eye):
            return Eye(p.shape[0], p.shape[1])  # transpose(adjoint(Eye)) = Eye
        return p.conj() if not sp.issparse(p) else p.conj().tocsc()
    if isinstance(d, Eye):
        return Eye(d.shape[1], d.shape[0])
    if sp.issparse(d):
        return d.T.tocsc()
    return d.T


class Op:
    """Operator{BL,BR,T} (src/operators_dense.jl:12-20): bases (as per-subsystem dims) + data."""

    def __init__(self, dims_l, dims_r, data):
        self.dims_l = tuple(int(x) for x in dims_l)
        self.dims_r = tuple(int(x) for x in dims_r)
        if not isinstance(data, (Eye, Adj)) and not sp.issparse(data):
            data = np.asfortranarray(np.asarray(data, dtype=C128))
        elif sp.issparse(data):
            data = sp.csc_matrix(data, dtype=C128)
        self.data = data
        if data.shape != (int(np.prod(self.dims_l)), int(np.prod(self.dims_r))):
            raise DimensionMismatch("Tried to assign data of size %s to bases of length %d and %d" % (
                data.shape, np.prod(self.dims_l), np.prod(self.dims_r)))

    @property
    def shape(self):
        return self.data.shape

    def dagger(self):
        d = self.data
        return Op(self.dims_r, self.dims_l, d.parent if isinstance(d, Adj) else Adj(d))

    def copy(self):
        d = self.data
        return Op(self.dims_l, self.dims_r, d.copy() if hasattr(d, "copy") else d)


class Ket:
    """Ket{B,T} (src/states.jl:11-32)."""

    def __init__(self, dims, data):
        self.dims = tuple(int(x) for x in dims)
        self.data = np.ascontiguousarray(np.asarray(data, dtype=C128)).reshape(-1)
        if self.data.shape[0] != int(np.prod(self.dims)):
            raise DimensionMismatch("Ket data length")


class Bra(Ket):
    """Bra{B,T} (src/states.jl:11-32); data are the plain (unconjugated) components."""


class LazyTensor:
    """LazyTensor(bl, br, indices, operators, factor) (src/operators_lazytensor.jl:15-37)."""

    def __init__(self, dims_l, dims_r, indices, operators, factor=1.0):
        self.dims_l = tuple(int(x) for x in dims_l)
        self.dims_r = tuple(int(x) for x in dims_r)
        if isinstance(indices, (int, np.integer)):
            indices, operators = [indices], [operators]
        self.indices = [int(i) for i in indices]
        self.operators = list(operators)
        self.factor = complex(factor)
        n = len(self.dims_l)
        assert n == len(self.dims_r)
        assert all(1 <= i <= n for i in self.indices), "check_indices"
        assert len(self.indices) == len(self.operators)
        assert self.indices == sorted(self.indices), "indices must be sorted"
        assert len(set(self.indices)) == len(self.indices)
        for i, o in zip(self.indices, self.operators):
            assert o.shape == (self.dims_l[i - 1], self.dims_r[i - 1]), "site operator basis mismatch"

    @property
    def shape(self):
        return (int(np.prod(self.dims_l)), int(np.prod(self.dims_r)))

    def suboperator(self, i):
        return self.operators[self.indices.index(i)]


class LazySum:
    """LazySum(basis_l, basis_r, factors, operators) (src/operators_lazysum.jl:41-51)."""

    def __init__(self, dims_l, dims_r, factors, operators):
        self.dims_l = tuple(int(x) for x in dims_l)
        self.dims_r = tuple(int(x) for x in dims_r)
        if len(factors) != len(operators):
            raise ArgumentError("LazySum `operators` and `factors` have different lengths.")
        self.factors = [complex(f) for f in factors]
        self.operators = list(operators)
        for o in self.operators:  # _check_bases, :6-11
            if tuple(o.dims_l) != self.dims_l or tuple(o.dims_r) != self.dims_r:
                raise IncompatibleBases()

    @property
    def shape(self):
        return (int(np.prod(self.dims_l)), int(np.prod(self.dims_r)))


class LazyProduct:
    """LazyProduct(operators, factor) (src/operators_lazyproduct.jl:32-52)."""

    def __init__(self, operators, factor=1.0):
        self.operators = list(operators)
        if not self.operators:
            raise ArgumentError("LazyProduct needs at least one operator!")
        for a, b in zip(self.operators[:-1], self.operators[1:]):  # check_multiplicable
            if tuple(a.dims_r) != tuple(b.dims_l):
                raise IncompatibleBases()
        self.factor = complex(factor)
        self.dims_l = tuple(self.operators[0].dims_l)
        self.dims_r = tuple(self.operators[-1].dims_r)

    @property
    def shape(self):
        return (int(np.prod(self.dims_l)), int(np.prod(self.dims_r)))


# --------------------------------------------------------------------------------------
# explicit dense twins (what the reference's tests compare against)
# --------------------------------------------------------------------------------------
def dense(op) -> np.ndarray:
    """dense(op).data: explicit matrix of any operator (tensor = kron(b, a), src/operators_dense.jl:134)."""
    if isinstance(op, Op):
        return materialize(op.data)
    if isinstance(op, LazyTensor):
        out = np.ones((1, 1), dtype=C128)
        for k in range(len(op.dims_l)):
            if (k + 1) in op.indices:
                m = materialize(op.suboperator(k + 1).data)
            else:
                m = np.eye(op.dims_l[k], op.dims_r[k], dtype=C128)
            out = np.kron(m, out)
        return np.asfortranarray(op.factor * out)
    if isinstance(op, LazySum):
        out = np.zeros(op.shape, dtype=C128)
        for f, o in zip(op.factors, op.operators):
            out = out + f * dense(o)
        return np.asfortranarray(out)
    if isinstance(op, LazyProduct):
        out = dense(op.operators[0])
        for o in op.operators[1:]:
            out = out @ dense(o)
        return np.asfortranarray(op.factor * out)
    raise MethodError(type(op))


# --------------------------------------------------------------------------------------
# scalar kernels (C) — src/sparsematrix.jl
# --------------------------------------------------------------------------------------
def _csc_arrays(m):
    m = sp.csc_matrix(m, dtype=C128)
    colptr = (m.indptr.astype(np.int64) + 1)
    rowval = (m.indices.astype(np.int64) + 1)
    nzval = np.ascontiguousarray(m.data.astype(C128))
    return colptr, rowval, nzval


def _zero_op_mul(data, beta):
    """src/operators_lazysum.jl:179-186"""
    flat = data.reshape(-1, order="F") if data.ndim > 1 else data
    lib().orc_zero_op_mul(_p(flat), ctypes.c_int64(flat.size), _p(_sc(beta)))
    return data


def _F(a):
    """flat column-major view that writes through."""
    if a.ndim == 1:
        assert a.flags.c_contiguous
        return a
    assert a.flags.f_contiguous, "dense operator data must be column-major"
    return a.reshape(-1, order="F")


def gemm(alpha, A, B, beta, R):
    """gemm!(alpha, A, B, beta, result) where exactly one of A, B is sparse (possibly Adj) — src/sparsematrix.jl:99-188."""
    L = lib()
    a_sp = sp.issparse(A) or (isinstance(A, Adj) and sp.issparse(A.parent))
    b_sp = sp.issparse(B) or (isinstance(B, Adj) and sp.issparse(B.parent))
    if a_sp and b_sp:
        raise MethodError("sparse*sparse gemm! is not implemented (src/sparsematrix.jl:177-188)")
    Rf = _F(R)
    i64 = ctypes.c_int64
    if a_sp:
        Bd = np.asfortranarray(B)
        if isinstance(A, Adj):
            M = A.parent
            cp, rv, nz = _csc_arrays(M)
            rc = L.orc_gemm_adjsp_dense(_p(_sc(alpha)), i64(M.shape[0]), i64(M.shape[1]), _p(cp), _p(rv), _p(nz),
                                        _p(_F(Bd)), i64(Bd.shape[0]), i64(Bd.shape[1]), _p(_sc(beta)), _p(Rf),
                                        i64(R.shape[0]), i64(R.shape[1]))
        else:
            cp, rv, nz = _csc_arrays(A)
            rc = L.orc_gemm_sp_dense(_p(_sc(alpha)), i64(A.shape[0]), i64(A.shape[1]), _p(cp), _p(rv), _p(nz),
                                     _p(_F(Bd)), i64(Bd.shape[0]), i64(Bd.shape[1]), _p(_sc(beta)), _p(Rf),
                                     i64(R.shape[0]), i64(R.shape[1]))
    else:
        Ad = np.asfortranarray(A)
        if isinstance(B, Adj):
            M = B.parent
            cp, rv, nz = _csc_arrays(M)
            rc = L.orc_gemm_dense_adjsp(_p(_sc(alpha)), _p(_F(Ad)), i64(Ad.shape[0]), i64(Ad.shape[1]),
                                        i64(M.shape[0]), i64(M.shape[1]), _p(cp), _p(rv), _p(nz), _p(_sc(beta)),
                                        _p(Rf), i64(R.shape[0]), i64(R.shape[1]))
        else:
            cp, rv, nz = _csc_arrays(B)
            rc = L.orc_gemm_dense_sp(_p(_sc(alpha)), _p(_F(Ad)), i64(Ad.shape[0]), i64(Ad.shape[1]),
                                     i64(B.shape[0]), i64(B.shape[1]), _p(cp), _p(rv), _p(nz), _p(_sc(beta)),
                                     _p(Rf), i64(R.shape[0]), i64(R.shape[1]))
    if rc != 0:
        raise DimensionMismatch()
    return R


def gemv(alpha, A, B, beta, r):
    """gemv!(alpha, M, v, beta, result) / gemv!(alpha, v, M, beta, result) — src/sparsematrix.jl:190-238."""
    L = lib()
    i64 = ctypes.c_int64
    if sp.issparse(A):
        cp, rv, nz = _csc_arrays(A)
        rc = L.orc_gemv_sp(_p(_sc(alpha)), i64(A.shape[0]), i64(A.shape[1]), _p(cp), _p(rv), _p(nz), _p(B),
                           i64(B.shape[0]), _p(_sc(beta)), _p(r), i64(r.shape[0]))
    else:
        cp, rv, nz = _csc_arrays(B)
        rc = L.orc_gemv_vsp(_p(_sc(alpha)), _p(A), i64(A.shape[0]), i64(B.shape[0]), i64(B.shape[1]), _p(cp), _p(rv),
                            _p(nz), _p(_sc(beta)), _p(r), i64(r.shape[0]))
    if rc != 0:
        raise DimensionMismatch()
    return r


# --------------------------------------------------------------------------------------
# LazyTensor dense-factor path — src/operators_lazytensor.jl:281-488
# --------------------------------------------------------------------------------------
def _as_matrix(a):
    """matrix usable by numpy/scipy matmul (BLAS zgemm for dense; SparseArrays-like product for CSC)."""
    if isinstance(a, Adj):
        p = a.parent
        if isinstance(p, Eye):
            return np.eye(p.shape[1], p.shape[0], dtype=C128)
        return p.conj().T if not sp.issparse(p) else p.conj().T.tocsc()
    if isinstance(a, Eye):
        return np.eye(a.shape[0], a.shape[1], dtype=C128)
    return a


def _gemm_blas(result_r, A, B, alpha, beta):
    """LinearAlgebra.mul!(C, A, B, alpha, beta): BLAS semantics — beta == 0 never reads C."""
    prod = A @ B
    if sp.issparse(prod):
        prod = prod.toarray()
    if beta == 0:
        result_r[...] = alpha * prod
    else:
        result_r[...] = alpha * prod + beta * result_r


def _tp_matmul_first(result, a, b, alpha, beta):
    """:281-290 — result_r(d_out x rest) = alpha*a*b_r(d_first x rest) + beta*result_r"""
    a = _as_matrix(a)
    d_first = a.shape[1]
    br = b.reshape((d_first, b.size // d_first), order="F")
    rr = result.reshape((a.shape[0], b.size // d_first), order="F")
    _gemm_blas(rr, a, br, alpha, beta)
    return result


def _tp_matmul_last(result, a, b, alpha, beta):
    """:292-301 — result_r(rest x d_out) = alpha*b_r(rest x d_last)*transpose(a) + beta*result_r"""
    a = _as_matrix(a)
    d_last = a.shape[1]
    br = b.reshape((b.size // d_last, d_last), order="F")
    rr = result.reshape((b.size // d_last, a.shape[0]), order="F")
    at = a.T
    if sp.issparse(at):
        prod = (at.T @ br.T).T  # keep the dense operand on the right for scipy
        if beta == 0:
            rr[...] = alpha * prod
        else:
            rr[...] = alpha * prod + beta * rr
    else:
        _gemm_blas(rr, br, at, alpha, beta)
    return result


def _tp_matmul_mid(result, a, loc, b, alpha, beta):
    """:333-404 — b, result are N-d F-order views; loc is 1-based."""
    sz1 = int(np.prod(b.shape[: loc - 1]))
    sz3 = int(np.prod(b.shape[loc:]))
    a_shape = a.shape
    br = b.reshape((sz1, b.shape[loc - 1], sz3), order="F")
    rr = result.reshape((sz1, a_shape[0], sz3), order="F")
    if isinstance(a, Eye):  # non-square Eye: slice copy, :346-371
        if not (b.shape[loc - 1] == a_shape[1] and result.shape[loc - 1] == a_shape[0]):
            raise DimensionMismatch("Dimensions of Eye matrix do not match subspace dimensions.")
        d = min(a_shape)
        if beta == 0:
            result[...] = 0
            rr[:, :d, :] = alpha * br[:, :d, :]
        else:
            result *= beta
            rr[:, :d, :] += alpha * br[:, :d, :]
        return result
    move_left = sz1 < sz3
    perm = (1, 0, 2) if move_left else (0, 2, 1)
    br_p = np.asfortranarray(np.transpose(br, perm))  # @strided permutedims!(br_p, br, perm)
    rshape_p = tuple(rr.shape[i] for i in perm)
    rr_p = np.empty(rshape_p, dtype=C128, order="F")
    if beta != 0:
        rr_p[...] = np.transpose(rr, perm)
    if move_left:
        _tp_matmul_first(rr_p, a, br_p, alpha, beta)
    else:
        _tp_matmul_last(rr_p, a, br_p, alpha, beta)
    rr[...] = np.transpose(rr_p, perm)  # perm is an involution
    return result


def _tp_matmul(result, a, loc, b, alpha, beta):
    """:406-428"""
    if loc == 1:
        return _tp_matmul_first(result, a, b, alpha, beta)
    if loc == b.ndim:
        return _tp_matmul_last(result, a, b, alpha, beta)
    return _tp_matmul_mid(result, a, loc, b, alpha, beta)


def _is_square_eye(d):
    """:436-440"""
    if isinstance(d, Adj):
        return _is_square_eye(d.parent)
    return isinstance(d, Eye) and d.shape[0] == d.shape[1]


def _tpops_tuple(operators, indices, shift=0, op_transform=None):
    """:523-537 — (matrix, axis) pairs, square Eyes filtered out."""
    pairs = []
    for o, i in zip(operators, indices):
        d = o.data if op_transform is None else op_transform(o.data)
        if not _is_square_eye(d):
            pairs.append((d, i + shift))
    return pairs


def _explicit_isometries(used_indices, shp_l, shp_r, shift=0):
    """:491-514 — Eye(sl, sr) on untouched axes whose left/right sizes differ."""
    if tuple(shp_l) == tuple(shp_r):
        return []
    out = []
    for i, (sl, sr) in enumerate(zip(shp_l, shp_r), start=1):
        if sl != sr and (i + shift) not in used_indices:
            out.append((Eye(sl, sr), i + shift))
    return out


def _tp_sum_matmul(result_data, tp_ops, iso_ops, b_data, alpha, beta):
    """:443-488 — sequential application with 0/1/2 ping-pong temporaries."""
    ops = list(tp_ops) + list(iso_ops)
    n = len(ops)

    def tmp_for(op, loc, arr):
        shp = tuple(op.shape[0] if (i + 1) == loc else arr.shape[i] for i in range(arr.ndim))
        return np.empty(shp, dtype=C128, order="F")

    if n == 0:
        if beta == 0:
            result_data[...] = alpha * b_data
        else:
            result_data[...] = alpha * b_data + beta * result_data
    elif n == 1:
        _tp_matmul(result_data, ops[0][0], ops[0][1], b_data, alpha, beta)
    elif n == 2:
        tmp = tmp_for(ops[0][0], ops[0][1], b_data)
        _tp_matmul(tmp, ops[0][0], ops[0][1], b_data, alpha, 0.0)
        _tp_matmul(result_data, ops[1][0], ops[1][1], tmp, 1.0, beta)
    else:
        tmp1 = tmp_for(ops[0][0], ops[0][1], b_data)
        _tp_matmul(tmp1, ops[0][0], ops[0][1], b_data, alpha, 0.0)
        for i in range(1, n - 1):
            tmp2 = tmp_for(ops[i][0], ops[i][1], tmp1)
            _tp_matmul(tmp2, ops[i][0], ops[i][1], tmp1, 1.0, 0.0)
            tmp1 = tmp2
        _tp_matmul(result_data, ops[n - 1][0], ops[n - 1][1], tmp1, 1.0, beta)
    return result_data


# --------------------------------------------------------------------------------------
# LazyTensor pure-sparse path — src/operators_lazytensor.jl:612-751
# --------------------------------------------------------------------------------------
def _is_pure_sparse(operators):
    """:520 — all factors SparseOpPureType (plain CSC) or EyeOpType (Eye or Adjoint Eye)."""
    def ok(d):
        if sp.issparse(d):
            return True
        if isinstance(d, Eye):
            return True
        return isinstance(d, Adj) and isinstance(d.parent, Eye)
    return all(ok(o.data) for o in operators)


def _strides(shape):
    """src/operators_dense.jl:296-308"""
    s, out = 1, []
    for d in shape:
        out.append(s)
        s *= d
    return out


def _gemm_puresparse(alpha, h: LazyTensor, op: np.ndarray, beta, result: np.ndarray, right: bool):
    """`_gemm_puresparse` both orders (:711-739) with `check_mul!_compatibility` (:691-708)."""
    if result is op or (np.shares_memory(result, op) and result.size and op.size):
        raise ArgumentError("output matrix must not be aliased with input matrix")
    hs = h.shape
    if not right:
        size_b = op.shape
        if hs[1] != size_b[0]:
            raise DimensionMismatch("A and B dimensions do not match. Can't do `A*B`")
        if tuple(result.shape) != (hs[0],) + tuple(size_b[1:]):
            raise DimensionMismatch("Output dimensions do not match A*B. Can't do `R.=A*B`")
    else:
        if op.ndim == 1:
            if hs[0] != op.shape[0] or result.shape != (hs[1],):
                raise DimensionMismatch("A and B dimensions do not match. Can't do `A*B`")
        else:
            if op.shape[1] != hs[0]:
                raise DimensionMismatch("A and B dimensions do not match. Can't do `A*B`")
            if tuple(result.shape) != (op.shape[0], hs[1]):
                raise DimensionMismatch("Output dimensions do not match A*B. Can't do `R.=A*B`")
    n = len(h.dims_l)
    shape = np.array([min(a, b) for a, b in zip(h.dims_l, h.dims_r)], dtype=np.int64)
    strides_j = np.array(_strides(h.dims_l), dtype=np.int64)
    strides_k = np.array(_strides(h.dims_r), dtype=np.int64)
    kind = np.zeros(n, dtype=np.int32)
    ncols = np.zeros(n, dtype=np.int64)
    PP = ctypes.c_void_p * n
    cps, rvs, nzs, keep = PP(), PP(), PP(), []
    for i, o in zip(h.indices, h.operators):
        d = o.data
        if sp.issparse(d):
            cp, rv, nz = _csc_arrays(d)
            keep.append((cp, rv, nz))
            kind[i - 1] = 1
            ncols[i - 1] = d.shape[1]
            cps[i - 1], rvs[i - 1], nzs[i - 1] = cp.ctypes.data, rv.ctypes.data, nz.ctypes.data
        elif not (isinstance(d, Eye) or (isinstance(d, Adj) and isinstance(d.parent, Eye))):
            raise ArgumentError("gemm! of LazyTensor is not implemented for %s" % type(d))
    opf, resf = _F(op), _F(result)
    if not right:
        op_ld, res_ld = op.shape[0], result.shape[0]
        n_free = 1 if op.ndim == 1 else op.shape[1]
    else:
        # result[I, K] += val*op[I, J]; vectors are treated as a Bra (:618-620)
        if op.ndim == 1:
            op_ld, res_ld, n_free = 1, 1, 1
        else:
            op_ld, res_ld, n_free = op.shape[0], result.shape[0], op.shape[0]
    lib().orc_lazytensor_puresparse(
        ctypes.c_int32(1 if right else 0), ctypes.c_int32(n), _p(shape), _p(strides_k), _p(strides_j), _p(kind),
        _p(ncols), cps, rvs, nzs, _p(_sc(alpha * h.factor)), _p(opf), ctypes.c_int64(op_ld), _p(_sc(beta)),
        _p(resf), ctypes.c_int64(res_ld), ctypes.c_int64(resf.size), ctypes.c_int64(n_free))
    return result


# --------------------------------------------------------------------------------------
# mul! dispatch
# --------------------------------------------------------------------------------------
def _is_denseop(x):
    return isinstance(x, Op) and not sp.issparse(x.data) and not isinstance(x.data, (Eye, Adj)) or \
        (isinstance(x, Op) and isinstance(x.data, Adj) and isinstance(x.data.parent, np.ndarray))


def _is_sparseop(x):
    return isinstance(x, Op) and (sp.issparse(x.data) or (isinstance(x.data, Adj) and sp.issparse(x.data.parent)))


def _dense_data(x):
    """dense `.data` of a DenseOpType (plain or Adjoint wrapped) as an F-order array; plain data alias through."""
    d = x.data
    if isinstance(d, Adj):
        return np.asfortranarray(d.parent.conj().T)
    return d


def _check_basis(cond):
    if not cond:
        # bases are type parameters in the reference: a mismatch is "no method" / IncompatibleBases
        raise IncompatibleBases()


def mul(result, a, b, alpha=1.0, beta=0.0):
    """mul!(result, a, b, alpha, beta) -> result.  Dispatch mirrors the reference's method table."""
    alpha, beta = complex(alpha), complex(beta)
    # ---- state on the right: result::Ket = a * b::Ket ; operator on the right: result = a::(Bra|Op) * b
    if isinstance(b, Ket) and not isinstance(b, Bra):
        return _mul_left(result, a, b, alpha, beta, ket=True)
    if isinstance(a, Bra):
        return _mul_right(result, a, b, alpha, beta, bra=True)
    if _is_denseop(b) and not _is_denseop(a):
        return _mul_left(result, a, b, alpha, beta, ket=False)
    if _is_denseop(a) and not _is_denseop(b):
        return _mul_right(result, a, b, alpha, beta, bra=False)
    if _is_denseop(a) and _is_denseop(b):  # src/operators_dense.jl:394
        _check_basis(a.dims_r == b.dims_l and result.dims_l == a.dims_l and result.dims_r == b.dims_r)
        _gemm_blas(result.data, _dense_data(a), _dense_data(b), alpha, beta)
        return result
    raise MethodError((type(result), type(a), type(b)))


def _state_dims_l(x, ket):
    return x.dims if ket else x.dims_l


def _mul_left(result, a, b, alpha, beta, ket):
    """result = alpha * a * b + beta * result with b a Ket (ket=True) or a dense Operator."""
    b_dl = b.dims if ket else b.dims_l
    r_dl = result.dims if ket else result.dims_l
    if not ket:
        _check_basis(tuple(result.dims_r) == tuple(b.dims_r))
    if isinstance(a, LazySum):  # src/operators_lazysum.jl:189-200, 215-226
        _check_basis(tuple(a.dims_l) == tuple(r_dl) and tuple(a.dims_r) == tuple(b_dl))
        if len(a.operators) == 0 or alpha == 0:
            _zero_op_mul(result.data, beta)
        else:
            mul(result, a.operators[0], b, alpha * a.factors[0], beta)
            for f, o in zip(a.factors[1:], a.operators[1:]):
                mul(result, o, b, alpha * f, 1.0)
        return result
    if isinstance(a, LazyProduct):  # src/operators_lazyproduct.jl:103-115, 131-146
        _check_basis(tuple(a.dims_l) == tuple(r_dl) and tuple(a.dims_r) == tuple(b_dl))
        if alpha == 0:
            _zero_op_mul(result.data, beta)
            return result
        ops = a.operators
        if len(ops) == 1:
            return mul(result, ops[0], b, a.factor * alpha, beta)

        def tmp_like(o):
            n = int(np.prod(o.dims_l))
            if ket:
                return Ket(o.dims_l, np.zeros(n, dtype=C128))
            return Op(o.dims_l, b.dims_r, np.zeros((n, b.data.shape[1]), dtype=C128, order="F"))
        t = tmp_like(ops[-1])
        mul(t, ops[-1], b, a.factor, 0.0)
        for o in reversed(ops[1:-1]):
            t2 = tmp_like(o)
            mul(t2, o, t, 1.0, 0.0)
            t = t2
        return mul(result, ops[0], t, alpha, beta)
    if isinstance(a, LazyTensor):  # src/operators_lazytensor.jl:539-557, 576-591
        _check_basis(tuple(a.dims_l) == tuple(r_dl) and tuple(a.dims_r) == tuple(b_dl))
        if alpha == 0:
            _zero_op_mul(result.data, beta)
            return result
        plain_b = ket or not isinstance(b.data, Adj)
        if len(a.operators) > 0 and _is_pure_sparse(a.operators) and plain_b:
            _gemm_puresparse(alpha, a, b.data, beta, result.data, right=False)
            return result
        bd = b.data if ket else _dense_data(b)
        extra_b = () if ket else tuple(b.dims_r)
        extra_r = () if ket else tuple(result.dims_r)
        b_nd = _F(bd).reshape(tuple(b_dl) + extra_b, order="F")
        r_nd = _F(result.data).reshape(tuple(r_dl) + extra_r, order="F")
        tp = _tpops_tuple(a.operators, a.indices)
        iso = _explicit_isometries(a.indices, a.dims_l, a.dims_r)
        _tp_sum_matmul(r_nd, tp, iso, b_nd, alpha * a.factor, beta)
        return result
    if _is_sparseop(a):  # src/operators_sparse.jl:199,201
        _check_basis(tuple(a.dims_l) == tuple(r_dl) and tuple(a.dims_r) == tuple(b_dl))
        if ket:
            if isinstance(a.data, Adj):
                raise MethodError("mul!(Ket, SparseOpAdjType, Ket) has no sparse method (operators_sparse.jl:201)")
            gemv(alpha, a.data, b.data, beta, result.data)
        else:
            gemm(alpha, a.data, _dense_data(b), beta, result.data)
        return result
    if _is_denseop(a):  # src/operators_dense.jl:394-395
        _check_basis(tuple(a.dims_l) == tuple(r_dl) and tuple(a.dims_r) == tuple(b_dl))
        _gemm_blas(result.data, _dense_data(a), b.data if ket else _dense_data(b), alpha, beta)
        return result
    raise MethodError((type(result), type(a), type(b)))


def _mul_right(result, a, b, alpha, beta, bra):
    """result = alpha * a * b + beta * result with a a Bra (bra=True) or a dense Operator, b the operator."""
    a_dr = a.dims if bra else a.dims_r
    r_dr = result.dims if bra else result.dims_r
    if not bra:
        _check_basis(tuple(result.dims_l) == tuple(a.dims_l))
    if isinstance(b, LazySum):  # src/operators_lazysum.jl:202-213, 227-238
        _check_basis(tuple(b.dims_l) == tuple(a_dr) and tuple(b.dims_r) == tuple(r_dr))
        if len(b.operators) == 0 or alpha == 0:
            _zero_op_mul(result.data, beta)
        else:
            mul(result, a, b.operators[0], alpha * b.factors[0], beta)
            for f, o in zip(b.factors[1:], b.operators[1:]):
                mul(result, a, o, alpha * f, 1.0)
        return result
    if isinstance(b, LazyProduct):  # src/operators_lazyproduct.jl:117-129, 148-163
        _check_basis(tuple(b.dims_l) == tuple(a_dr) and tuple(b.dims_r) == tuple(r_dr))
        if alpha == 0:
            _zero_op_mul(result.data, beta)
            return result
        ops = b.operators
        if len(ops) == 1:
            return mul(result, a, ops[0], b.factor * alpha, beta)

        def tmp_like(o):
            n = int(np.prod(o.dims_r))
            if bra:
                return Bra(o.dims_r, np.zeros(n, dtype=C128))
            return Op(a.dims_l, o.dims_r, np.zeros((a.data.shape[0], n), dtype=C128, order="F"))
        t = tmp_like(ops[0])
        mul(t, a, ops[0], b.factor, 0.0)
        for o in ops[1:-1]:
            t2 = tmp_like(o)
            mul(t2, t, o, 1.0, 0.0)
            t = t2
        return mul(result, t, ops[-1], alpha, beta)
    if isinstance(b, LazyTensor):  # src/operators_lazytensor.jl:559-574, 593-609
        _check_basis(tuple(b.dims_l) == tuple(a_dr) and tuple(b.dims_r) == tuple(r_dr))
        if alpha == 0:
            _zero_op_mul(result.data, beta)
            return result
        plain_a = bra or not isinstance(a.data, Adj)
        if len(b.operators) > 0 and _is_pure_sparse(b.operators) and plain_a:
            _gemm_puresparse(alpha, b, a.data, beta, result.data, right=True)
            return result
        ad = a.data if bra else _dense_data(a)
        if bra:
            a_nd = ad.reshape(tuple(a_dr), order="F")
            r_nd = result.data.reshape(tuple(r_dr), order="F")
            shift = 0
        else:
            a_nd = _F(ad).reshape(tuple(a.dims_l) + tuple(a.dims_r), order="F")
            r_nd = _F(result.data).reshape(tuple(result.dims_l) + tuple(result.dims_r), order="F")
            shift = len(a.dims_l)
        tp = _tpops_tuple(b.operators, b.indices, shift=shift, op_transform=_transpose_data)
        iso = _explicit_isometries([i + shift for i in b.indices], b.dims_r, b.dims_l, shift)
        _tp_sum_matmul(r_nd, tp, iso, a_nd, alpha * b.factor, beta)
        return result
    if _is_sparseop(b):  # src/operators_sparse.jl:200,202
        _check_basis(tuple(b.dims_l) == tuple(a_dr) and tuple(b.dims_r) == tuple(r_dr))
        if bra:
            if isinstance(b.data, Adj):
                raise MethodError("mul!(Bra, Bra, SparseOpAdjType) has no sparse method (operators_sparse.jl:202)")
            gemv(alpha, a.data, b.data, beta, result.data)
        else:
            gemm(alpha, _dense_data(a), b.data, beta, result.data)
        return result
    if _is_denseop(b):  # src/operators_dense.jl:394,396
        _check_basis(tuple(b.dims_l) == tuple(a_dr) and tuple(b.dims_r) == tuple(r_dr))
        if bra:  # mul!(result.data, transpose(b.data), a.data, alpha, beta)
            _gemm_blas(result.data, _dense_data(b).T, a.data, alpha, beta)
        else:
            _gemm_blas(result.data, _dense_data(a), _dense_data(b), alpha, beta)
        return result
    raise MethodError((type(result), type(a), type(b)))


# --------------------------------------------------------------------------------------
# site operators that define the benchmark inputs (all sparse, as the reference builds them)
# --------------------------------------------------------------------------------------
def _spdiagm(n, offsets):
    m = sp.lil_matrix((n, n), dtype=C128)
    for off, vals in offsets.items():
        for t, v in enumerate(vals):
            i, j = (t, t + off) if off >= 0 else (t - off, t)
            m[i, j] = v
    return sp.csc_matrix(m)


def sigmax(spin=0.5):
    """src/spin.jl:14-20"""
    n = int(round(2 * spin + 1))
    d = [np.sqrt((spin + 1) * 2 * a - a * (a + 1)) for a in range(1, n)]
    return Op((n,), (n,), _spdiagm(n, {1: d, -1: d}))


def sigmay(spin=0.5):
    """src/spin.jl:34-40"""
    n = int(round(2 * spin + 1))
    d = [1j * np.sqrt((spin + 1) * 2 * a - a * (a + 1)) for a in range(1, n)]
    return Op((n,), (n,), _spdiagm(n, {-1: d, 1: [-x for x in d]}))


def sigmaz(spin=0.5):
    """src/spin.jl:54-60"""
    n = int(round(2 * spin + 1))
    d = [2 * (spin - t) for t in range(n)]
    return Op((n,), (n,), _spdiagm(n, {0: d}))


def sigmap(spin=0.5):
    """src/spin.jl:68-75"""
    n = int(round(2 * spin + 1))
    S = (spin + 1) * spin
    ms = [spin - 1 - t for t in range(n - 1)]
    return Op((n,), (n,), _spdiagm(n, {1: [np.sqrt(S - m * (m + 1)) for m in ms]}))


def sigmam(spin=0.5):
    """src/spin.jl:83-90"""
    n = int(round(2 * spin + 1))
    S = (spin + 1) * spin
    ms = [spin - t for t in range(n - 1)]
    return Op((n,), (n,), _spdiagm(n, {-1: [np.sqrt(S - m * (m - 1)) for m in ms]}))


def number(N, offset=0):
    """src/fock.jl:8-12 — FockBasis(N, offset) has dimension N-offset+1"""
    d = [float(v) for v in range(offset, N + 1)]
    return Op((len(d),), (len(d),), _spdiagm(len(d), {0: d}))


def destroy(N, offset=0):
    """src/fock.jl:22-28"""
    n = N - offset + 1
    return Op((n,), (n,), _spdiagm(n, {1: [np.sqrt(float(v)) for v in range(offset + 1, N + 1)]}))


def create(N, offset=0):
    """src/fock.jl:38-44"""
    n = N - offset + 1
    return Op((n,), (n,), _spdiagm(n, {-1: [np.sqrt(float(v)) for v in range(offset + 1, N + 1)]}))


def transition(n, to, frm):
    """src/nlevel.jl:8-18"""
    if not (1 <= to <= n and 1 <= frm <= n):
        raise IndexError("BoundsError")
    m = sp.lil_matrix((n, n), dtype=C128)
    m[to - 1, frm - 1] = 1.0
    return Op((n,), (n,), sp.csc_matrix(m))


def identityoperator(dl, dr=None):
    """identityoperator(b1, b2) for sparse operators (src/operators_sparse.jl:180-186)"""
    dr = dl if dr is None else dr
    return Op((dl,), (dr,), sp.csc_matrix(sp.eye(dl, dr, dtype=C128)))


def fill_state(n, seed, scale=1.0, offset=0):
    """Counter-based synthetic state shared with the GPU generator (qob_fill_state)."""
    x = np.empty(n, dtype=C128)
    lib().orc_fill_state(_p(x), ctypes.c_int64(offset), ctypes.c_int64(n), ctypes.c_uint64(seed), ctypes.c_double(scale))
    return x


def state_at(seed, index, scale=1.0):
    out = np.zeros(2)
    lib().orc_state_at(ctypes.c_uint64(seed), ctypes.c_double(scale), ctypes.c_int64(index), _p(out))
    return complex(out[0], out[1])


# ------------------------------------------------------------------------------------ next rows of SURVEY §8(f)
# Restated ahead of the device implementation (round 2): the oracle comes first.  Nothing in the product uses these.
def ptrace_op(dims_l, dims_r, data, indices):
    """ptrace(a::DataOperator, indices) — src/operators_dense.jl:191-196, 311-342: sum over equal values of the traced
    subsystems on both sides, result[J_l, J_r] = sum_t a[(J_l, t), (J_r, t)] (subsystem 1 fastest, 1-based indices).
    Returns (dims_l kept, dims_r kept, matrix)."""
    idx = sorted(int(i) - 1 for i in (indices if hasattr(indices, "__iter__") else [indices]))
    n = len(dims_l)
    if len(idx) == 0 or len(idx) >= n or len(set(idx)) != len(idx) or idx[0] < 0 or idx[-1] >= n:
        raise ArgumentError("ptrace: indices must select some but not all subsystems")   # check_ptrace_arguments
    for i in idx:
        if dims_l[i] != dims_r[i]:
            raise ArgumentError("ptrace: traced subsystems need equal left and right dimensions")
    a = np.asarray(_dense_data(data) if not isinstance(data, np.ndarray) else data)
    # column-major composite index: axis k of the reshaped tensor (order="F") is subsystem k+1
    t = a.reshape(tuple(dims_l) + tuple(dims_r), order="F")
    keep = [k for k in range(n) if k not in idx]
    letters = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
    left = [letters[k] for k in range(n)]
    right = [letters[k] if k in idx else letters[n + k] for k in range(n)]
    out = "".join(left[k] for k in keep) + "".join(right[k] for k in keep)
    r = np.einsum("".join(left) + "".join(right) + "->" + out, t)
    kl, kr = tuple(dims_l[k] for k in keep), tuple(dims_r[k] for k in keep)
    return kl, kr, np.asfortranarray(r.reshape(int(np.prod(kl)), int(np.prod(kr)), order="F"))


def ptrace_ket(dims, psi, indices):
    """ptrace(psi::Ket, indices) — src/operators_dense.jl:199-206, 344-362: result[J_l, J_r] = sum_t psi[(J_l,t)] conj(psi[(J_r,t)])"""
    v = np.asarray(psi).reshape(-1)
    return ptrace_op(dims, dims, np.outer(v, np.conj(v)), indices)


def ptrace_bra(dims, psi, indices):
    """ptrace(psi::Bra, indices) — src/operators_dense.jl:208-215, 364-383: result[J_l, J_r] = sum_t conj(psi[(J_l,t)]) psi[(J_r,t)]"""
    v = np.asarray(psi).reshape(-1)
    return ptrace_op(dims, dims, np.outer(np.conj(v), v), indices)


def directsum_mul(result, operators, b, alpha=1.0, beta=0.0, bra=False):
    """mul!(result::Ket, M::LazyDirectSum, b::Ket, alpha, beta) — src/spinors.jl:221-233 (bra=True: :234-247).  `result` and
    `b` are flat arrays over the SumBasis; block i acts on slice i.  Like the reference, BOTH vectors are sliced by the lengths
    of the right bases (`index = cumsum([0; length.(bases_r)...])`), so a non-square block raises DimensionMismatch (the
    reference's `Ket(bases_l[i], result.data[...])` constructor throws)."""
    lens_r = [int(np.prod(o.dims_r)) for o in operators]
    index = np.concatenate([[0], np.cumsum(lens_r)])
    for i, o in enumerate(operators):
        sl = slice(int(index[i]), int(index[i + 1]))
        if int(np.prod(o.dims_l)) != lens_r[i]:
            raise DimensionMismatch("LazyDirectSum block is not square")
        if not bra:
            tmpket = Ket(o.dims_r, np.array(b[sl], dtype=C128))
            tmpres = Ket(o.dims_l, np.array(result[sl], dtype=C128))
            mul(tmpres, o, tmpket, alpha, beta)
        else:
            tmpket = Bra(o.dims_l, np.array(b[sl], dtype=C128))
            tmpres = Bra(o.dims_r, np.array(result[sl], dtype=C128))
            mul(tmpres, tmpket, o, alpha, beta)
        result[sl] = tmpres.data
    return result


def lindblad_rhs(H, J, rho, rates=None):
    """The master-equation right-hand side as the mul! call pattern of test/test_sciml_broadcast_interfaces.jl:36-43 builds
    it, -i[H, rho] + sum_k g_k (J_k rho J_k^+ - (J_k^+ J_k rho + rho J_k^+ J_k)/2), with dense matrices (the fused device
    kernel of round 2 will be checked against this)."""
    H = np.asarray(H if isinstance(H, np.ndarray) else dense(H))
    rho = np.asarray(rho)
    out = -1j * (H @ rho - rho @ H)
    for k, Jk in enumerate(J):
        Jm = np.asarray(Jk if isinstance(Jk, np.ndarray) else dense(Jk))
        g = 1.0 if rates is None else rates[k]
        JdJ = Jm.conj().T @ Jm
        out += g * (Jm @ rho @ Jm.conj().T - 0.5 * (JdJ @ rho + rho @ JdJ))
    return out
