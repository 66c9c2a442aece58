"""Host-side mirror of the reference's operator interface for the `mul!` hot path.

Same names, argument meaning and error behaviour as QuantumOpticsBase.jl (citations relative to
/root/reference): bases (QuantumInterface, src/bases.jl:1-3), `Ket`/`Bra` (src/states.jl:11-32),
`Operator` (src/operators_dense.jl:12-20), `SparseOperator` (src/operators_sparse.jl:5-23),
`LazyTensor` (src/operators_lazytensor.jl:15-58), `LazySum` (src/operators_lazysum.jl:41-72),
`LazyProduct` (src/operators_lazyproduct.jl:32-59) and `mul_` = `mul!(result, a, b, alpha, beta)`.

State data (`Ket.data`, `Bra.data`, a dense `Operator.data` used as a state) are torch complex128
CUDA tensors (column-major); operator DEFINITIONS (site factors, CSC matrices, coefficients) stay on
the host as numpy / scipy objects, exactly like `.data` of the reference's small site operators, and
are compiled once into a libqob200 handle.  All arithmetic happens in libqob200.so; nothing here
computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import (ArgumentError, CudaError, DimensionMismatch, IncompatibleBases, MethodError, c64, lib)

C128 = np.complex128


# ------------------------------------------------------------------------------------ bases
class Basis:
    """QuantumInterface.Basis: `shape` + value equality."""

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in shape)
        self._len = int(np.prod(self.shape)) if self.shape else 1

    def __len__(self):
        return self._len

    def _key(self):
        return (type(self).__name__, self.shape)

    def _ckey(self):
        k = getattr(self, "_cached_key", None)   # bases are immutable: the key is computed once
        if k is None:
            k = self._key()
            self._cached_key = k
        return k

    def __eq__(self, other):
        return self is other or (isinstance(other, Basis) and self._ckey() == other._ckey())

    def __hash__(self):
        return hash(self._ckey())

    def __repr__(self):
        return f"{type(self).__name__}{self._key()[1:]}"


class GenericBasis(Basis):
    def __init__(self, n):
        super().__init__((int(n),))


class SpinBasis(Basis):
    def __init__(self, spinnumber):
        self.spinnumber = float(spinnumber)
        n = 2 * self.spinnumber + 1
        assert abs(n - round(n)) < 1e-12 and n >= 1
        super().__init__((int(round(n)),))

    def _key(self):
        return ("SpinBasis", self.shape, self.spinnumber)


class FockBasis(Basis):
    def __init__(self, N, offset=0):
        self.N, self.offset = int(N), int(offset)
        super().__init__((self.N - self.offset + 1,))

    def _key(self):
        return ("FockBasis", self.shape, self.N, self.offset)


class NLevelBasis(Basis):
    def __init__(self, N):
        self.N = int(N)
        super().__init__((self.N,))


class CompositeBasis(Basis):
    def __init__(self, bases):
        self.bases = list(bases)
        super().__init__(tuple(len(b) for b in self.bases))

    def _key(self):
        return ("CompositeBasis", tuple(b._ckey() for b in self.bases))


class SumBasis(Basis):
    """SumBasis(b1, b2, …) (src/spinors.jl:1-17): direct sum of bases, shape = the lengths of the parts."""

    def __init__(self, *bases):
        if len(bases) == 1 and isinstance(bases[0], (list, tuple)):
            bases = tuple(bases[0])
        self.bases = list(bases)
        Basis.__init__(self, tuple(len(b) for b in self.bases))
        self._len = sum(len(b) for b in self.bases)

    def _key(self):
        return ("SumBasis", tuple(b._ckey() for b in self.bases))


def directsum(*bases):
    out = []
    for b in bases:
        out.extend(b.bases if isinstance(b, SumBasis) else [b])
    return SumBasis(out)


def tensor(*xs):
    """b1 ⊗ b2 ⊗ …  for bases (composite bases are flattened like the reference's `tensor`)."""
    if all(isinstance(x, Basis) for x in xs):
        out = []
        for b in xs:
            out.extend(b.bases if isinstance(b, CompositeBasis) else [b])
        return CompositeBasis(out)
    raise MethodError("tensor of operators/states is host-side construction outside the hot path; "
                      "build LazyTensor terms instead")


def _comp_size(b):
    """`_comp_size` (src/operators_lazytensor.jl:517-518)."""
    return tuple(len(x) for x in b.bases) if isinstance(b, CompositeBasis) else (len(b),)


# ------------------------------------------------------------------------------------ data wrappers
class Eye:
    """FillArrays.Eye(m, n)."""

    def __init__(self, m, n=None):
        self.shape = (int(m), int(m if n is None else n))


class Adjoint:
    """LinearAlgebra.Adjoint wrapper (lazy `dagger`, src/operators_dense.jl:128)."""

    def __init__(self, parent):
        self.parent = parent
        self.shape = (parent.shape[1], parent.shape[0])


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _device_colmajor(arr2d):
    """numpy (rows, cols) -> torch CUDA complex128 tensor of the same logical shape, column-major storage."""
    import torch

    if not torch.cuda.is_available():
        raise CudaError("no CUDA device available: quantumopticsbase.jl_b200 has no CPU fallback")
    a = np.asarray(arr2d, dtype=C128)
    t = torch.from_numpy(np.ascontiguousarray(a.T)).cuda()
    return t.t()


def _device_vector(arr):
    import torch

    if not torch.cuda.is_available():
        raise CudaError("no CUDA device available: quantumopticsbase.jl_b200 has no CPU fallback")
    return torch.from_numpy(np.ascontiguousarray(np.asarray(arr, dtype=C128).reshape(-1))).cuda()


# ------------------------------------------------------------------------------------ states
class StateVector:
    def __init__(self, basis, data=None):
        import torch

        self.basis = basis
        n = len(basis)
        if data is None:
            if not torch.cuda.is_available():
                raise CudaError("no CUDA device available: quantumopticsbase.jl_b200 has no CPU fallback")
            data = torch.zeros(n, dtype=torch.complex128, device="cuda")
        elif not _is_torch(data):
            data = _device_vector(data)
        if data.numel() != n:  # src/states.jl:16-17
            raise DimensionMismatch(f"Tried to assign data of length {data.numel()} to basis of length {n}.")
        assert data.dtype == torch.complex128 and data.is_cuda and data.is_contiguous()
        self.data = data

    def to_host(self):
        return self.data.cpu().numpy()

    def copy(self):
        return type(self)(self.basis, self.data.clone())


class Ket(StateVector):
    """Ket{B,T} with CuPtr-backed ComplexF64 data."""


class Bra(StateVector):
    """Bra{B,T}; `.data` are the plain components (no conjugation), as in the reference."""


# ------------------------------------------------------------------------------------ operators
class AbstractOperator:
    basis_l: Basis
    basis_r: Basis
    _handle = None
    _handle_ctx = None

    @property
    def shape(self):
        return (len(self.basis_l), len(self.basis_r))

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                lib.qob_op_destroy(h)
            except Exception:
                pass


class Operator(AbstractOperator):
    """Operator{BL,BR,T}.  `data`: numpy matrix / scipy sparse / Eye / Adjoint (an operator definition,
    host side) or a torch CUDA tensor (a dense state such as a density matrix, device side)."""

    def __init__(self, basis_l, basis_r=None, data=None):
        if data is None and basis_r is not None and not isinstance(basis_r, Basis):
            basis_r, data = basis_l, basis_r
        if basis_r is None:
            basis_r = basis_l
        self.basis_l, self.basis_r = basis_l, basis_r
        if data is None:
            data = _device_colmajor(np.zeros((len(basis_l), len(basis_r)), dtype=C128))
        if sp.issparse(data):
            data = sp.csc_matrix(data, dtype=C128)
        elif isinstance(data, np.ndarray) or isinstance(data, (list, tuple)):
            data = np.asfortranarray(np.asarray(data, dtype=C128))
        if tuple(data.shape) != (len(basis_l), len(basis_r)):  # src/operators_dense.jl:16-17
            raise DimensionMismatch(f"Tried to assign data of size {tuple(data.shape)} to bases of length "
                                    f"{len(basis_l)} and {len(basis_r)}!")
        if _is_torch(data):
            import torch

            assert data.dtype == torch.complex128 and data.is_cuda
            assert data.stride() == (1, data.shape[0]) or data.numel() <= max(data.shape), \
                "dense device operators must be column-major (use DenseOperator(...))"
        self.data = data

    # classification helpers (the reference's type aliases)
    @property
    def is_device_dense(self):  # DenseOpType holding device memory
        return _is_torch(self.data)

    @property
    def is_sparse(self):  # SparseOpType (pure or adjoint)
        d = self.data
        return sp.issparse(d) or (isinstance(d, Adjoint) and sp.issparse(d.parent))

    @property
    def is_eye(self):
        d = self.data
        return isinstance(d, Eye) or (isinstance(d, Adjoint) and isinstance(d.parent, Eye))

    def to_host(self):
        d = self.data
        if _is_torch(d):
            return np.asfortranarray(d.cpu().numpy())
        return _materialize(d)

    def copy(self):
        d = self.data
        return Operator(self.basis_l, self.basis_r, d.clone() if _is_torch(d) else (d.copy() if hasattr(d, "copy") else d))


def DenseOperator(basis_l, basis_r=None, data=None):
    """Dense operator whose data live on the device (density matrices, Ket batches)."""
    if data is None and basis_r is not None and not isinstance(basis_r, Basis):
        basis_r, data = basis_l, basis_r
    if basis_r is None:
        basis_r = basis_l
    if data is None:
        data = np.zeros((len(basis_l), len(basis_r)), dtype=C128)
    if not _is_torch(data):
        data = _device_colmajor(data)
    return Operator(basis_l, basis_r, data)


def SparseOperator(basis_l, basis_r=None, data=None):
    if data is None and basis_r is not None and not isinstance(basis_r, Basis):
        basis_r, data = basis_l, basis_r
    if basis_r is None:
        basis_r = basis_l
    if data is None:
        data = sp.csc_matrix((len(basis_l), len(basis_r)), dtype=C128)
    if isinstance(data, Operator):
        data = data.data
    return Operator(basis_l, basis_r, sp.csc_matrix(_materialize(data) if not sp.issparse(data) else data, dtype=C128))


def _materialize(d):
    if isinstance(d, Adjoint):
        return np.asfortranarray(_materialize(d.parent).conj().T)
    if isinstance(d, Eye):
        return np.asfortranarray(np.eye(d.shape[0], d.shape[1], dtype=C128))
    if sp.issparse(d):
        return np.asfortranarray(d.toarray().astype(C128))
    if _is_torch(d):
        return np.asfortranarray(d.cpu().numpy())
    return np.asfortranarray(np.asarray(d, dtype=C128))


def dagger(op: Operator):
    """Lazy adjoint (src/operators_dense.jl:128)."""
    d = op.data
    if _is_torch(d):
        raise MethodError("dagger of a device state is outside the mul! hot path")
    return Operator(op.basis_r, op.basis_l, d.parent if isinstance(d, Adjoint) else Adjoint(d))


def dense(op: Operator):
    return Operator(op.basis_l, op.basis_r, _materialize(op.data))


def sparse(op: Operator):
    return Operator(op.basis_l, op.basis_r, sp.csc_matrix(_materialize(op.data)))


def identityoperator(b1, b2=None, kind="sparse"):
    """identityoperator(b1, b2) (src/operators_sparse.jl:180-186); kind='eye' gives the FillArrays form."""
    b2 = b1 if b2 is None else b2
    if kind == "eye":
        return Operator(b1, b2, Eye(len(b1), len(b2)))
    return Operator(b1, b2, sp.csc_matrix(sp.eye(len(b1), len(b2), dtype=C128)))


class LazyTensor(AbstractOperator):
    """LazyTensor(b1[, b2], indices, operators[, factor=1])  (src/operators_lazytensor.jl:4-58)."""

    def __init__(self, basis_l, *args):
        args = list(args)
        if args and isinstance(args[0], Basis):
            basis_r = args.pop(0)
        else:
            basis_r = basis_l
        indices, operators = args[0], args[1]
        factor = args[2] if len(args) > 2 else 1.0
        if not isinstance(basis_l, CompositeBasis):
            basis_l = CompositeBasis([basis_l])
        if not isinstance(basis_r, CompositeBasis):
            basis_r = CompositeBasis([basis_r])
        if isinstance(indices, (int, np.integer)):
            indices, operators = [int(indices)], (operators,)
        if isinstance(operators, list):
            operators = tuple(operators)  # deprecated Vector form (:38-42)
        self.basis_l, self.basis_r = basis_l, basis_r
        self.indices = [int(i) for i in indices]
        self.operators = tuple(operators)
        self.factor = complex(factor)
        N = len(basis_l.bases)
        if N != len(basis_r.bases):
            raise AssertionError("N == length(br.bases)")
        for i in self.indices:  # check_indices
            if not 1 <= i <= N:
                raise ArgumentError("indices out of range")
        if len(set(self.indices)) != len(self.indices):
            raise ArgumentError("indices must be unique")
        if len(self.indices) != len(self.operators):
            raise AssertionError("length(indices) == length(ops)")
        if self.indices != sorted(self.indices):
            raise AssertionError("issorted(indices)")
        for i, o in zip(self.indices, self.operators):
            if not isinstance(o, AbstractOperator):
                raise AssertionError("isa(ops[n], AbstractOperator)")
            if o.basis_l != basis_l.bases[i - 1] or o.basis_r != basis_r.bases[i - 1]:
                raise AssertionError("ops[n].basis_l == bl.bases[indices[n]] && ops[n].basis_r == br.bases[indices[n]]")

    def __mul__(self, x):
        return LazyTensor(self.basis_l, self.basis_r, self.indices, self.operators, self.factor * complex(x))

    __rmul__ = __mul__

    def __neg__(self):
        return self * -1.0


class LazySum(AbstractOperator):
    """LazySum([Tf,] [factors,] operators) / LazySum(basis_l, basis_r, [factors, operators])
    (src/operators_lazysum.jl:13-72).  `factors` may be mutated between calls (TimeDependentSum
    set_time!, src/time_dependent_operator.jl:279-290): they are re-sent on every mul_."""

    def __init__(self, *args):
        args = list(args)
        if args and isinstance(args[0], Basis):
            self.basis_l, self.basis_r = args[0], args[1]
            rest = args[2:]
            factors, operators = (rest[0], rest[1]) if rest else ([], [])
        elif len(args) == 2 and not isinstance(args[0], AbstractOperator):
            factors, operators = args
            if len(operators) == 0:
                raise ArgumentError("LazySum needs a basis, or at least one operator!")
            self.basis_l, self.basis_r = operators[0].basis_l, operators[0].basis_r
        else:
            operators = args
            if len(operators) == 0:
                raise ArgumentError("LazySum needs a basis, or at least one operator!")
            factors = [1.0] * len(operators)
            self.basis_l, self.basis_r = operators[0].basis_l, operators[0].basis_r
        if len(factors) != len(operators):
            raise ArgumentError("LazySum `operators` and `factors` have different lengths.")
        self.factors = [complex(f) for f in factors]
        self.operators = list(operators)
        for o in self.operators:  # _check_bases, :6-11
            if o.basis_l != self.basis_l or o.basis_r != self.basis_r:
                raise IncompatibleBases()


class LazyProduct(AbstractOperator):
    """LazyProduct(operators[, factor=1]) / LazyProduct(op1, op2, …)  (src/operators_lazyproduct.jl:23-59)."""

    def __init__(self, *args):
        if len(args) >= 1 and isinstance(args[0], (list, tuple)):
            operators = list(args[0])
            factor = args[1] if len(args) > 1 else 1.0
        else:
            operators, factor = list(args), 1.0
        if not operators:
            raise ArgumentError("LazyProduct needs at least one operator!")
        for a, b in zip(operators[:-1], operators[1:]):  # check_multiplicable
            if a.basis_r != b.basis_l:
                raise IncompatibleBases()
        self.operators = operators
        self.factor = complex(factor)
        self.basis_l, self.basis_r = operators[0].basis_l, operators[-1].basis_r


class TimeDependentSum(AbstractOperator):
    """TimeDependentSum(coefficients, operators; init_time=0) (src/time_dependent_operator.jl:150-170): a LazySum whose
    factors are numbers or functions of time.  `set_time_` rewrites the static LazySum's factors (:279-290); `mul_` forwards
    to the static operator (:274-277), which re-sends the coefficients to the device plan — a few hundred bytes, no
    replanning (`qob_lazysum_set_coefs`)."""

    def __init__(self, coefficients, operators, init_time=0.0):
        if isinstance(operators, LazySum):
            self.static_op = operators
        else:
            self.static_op = LazySum([0.0] * len(operators), list(operators))
        self.coefficients = list(coefficients)
        if len(self.coefficients) != len(self.static_op.operators):
            raise ArgumentError("TimeDependentSum `coefficients` and `operators` have different lengths.")
        self.basis_l, self.basis_r = self.static_op.basis_l, self.static_op.basis_r
        self.current_time = None
        self.set_time_(init_time)

    def set_time_(self, t):
        if self.current_time != t:
            self.current_time = t
            self.static_op.factors = [complex(c(t)) if callable(c) else complex(c) for c in self.coefficients]
        for o in self.static_op.operators:
            if isinstance(o, TimeDependentSum):
                o.set_time_(t)
        return self


# ------------------------------------------------------------------------------------ handles
def _factor_struct(d, keep):
    """qob_factor for operator data `d`; numpy arrays referenced by the struct are appended to `keep`."""
    f = _lib.Factor()
    trans = _lib.OP_N
    if isinstance(d, Adjoint):
        trans = _lib.OP_C
        d = d.parent
        if isinstance(d, Adjoint):
            raise MethodError("nested Adjoint")
    f.trans = trans
    f.nrows, f.ncols = int(d.shape[0]), int(d.shape[1])
    if isinstance(d, Eye):
        f.kind = _lib.FACTOR_EYE
    elif sp.issparse(d):
        m = sp.csc_matrix(d, dtype=C128)
        colptr = (m.indptr.astype(np.int64) + 1)
        rowval = (m.indices.astype(np.int64) + 1)
        nzval = np.ascontiguousarray(m.data.astype(C128))
        keep.extend([colptr, rowval, nzval])
        f.kind = _lib.FACTOR_CSC
        f.colptr, f.rowval, f.nzval = colptr.ctypes.data, rowval.ctypes.data, nzval.ctypes.data
    elif isinstance(d, np.ndarray):
        a = np.asfortranarray(d.astype(C128))
        keep.append(a)
        f.kind = _lib.FACTOR_DENSE
        f.dense = a.ctypes.data
    else:
        # e.g. a device tensor or an arbitrary AbstractOperator used as a site factor:
        # the reference throws MethodError/ArgumentError (operators_lazytensor.jl:639-641)
        raise MethodError(f"no kernel for a site factor with data of type {type(d).__name__}")
    return f


class LazyDirectSum(AbstractOperator):
    """LazyDirectSum(op1, op2, …) (src/spinors.jl:158-169): block-diagonal operator on SumBases; nested sums are flattened.
    `mul_` exists for Ket and Bra states only, like the reference (src/spinors.jl:221-247)."""

    def __init__(self, *ops):
        flat = []
        for o in ops:
            flat.extend(o.operators if isinstance(o, LazyDirectSum) else [o])
        if not flat:
            raise ArgumentError("LazyDirectSum needs at least one operator")
        self.operators = flat
        self.basis_l = directsum(*[o.basis_l for o in flat])
        self.basis_r = directsum(*[o.basis_r for o in flat])


def handle(op, ctx=None):
    """libqob200 handle of an operator definition (built once, cached on the object)."""
    ctx = _lib.context() if ctx is None else ctx
    if getattr(op, "_handle", None) and op._handle_ctx is ctx:
        if isinstance(op, LazySum):
            _refresh_coefs(op)
        return op._handle
    h = C.c_void_p()
    keep = []
    if isinstance(op, LazyTensor):
        dl = (C.c_int64 * len(op.basis_l.shape))(*op.basis_l.shape)
        dr = (C.c_int64 * len(op.basis_r.shape))(*op.basis_r.shape)
        n = len(op.indices)
        sites = (C.c_int32 * max(n, 1))(*op.indices)
        facs = (_lib.Factor * max(n, 1))()
        for k, o in enumerate(op.operators):
            if not isinstance(o, Operator):
                raise MethodError(f"LazyTensor factor of type {type(o).__name__} has no mul! kernel")
            facs[k] = _factor_struct(o.data, keep)
        _lib.check(lib.qob_lazytensor_create(ctx, len(op.basis_l.shape), dl, dr, n, sites, facs, c64.of(op.factor),
                                             C.byref(h)))
    elif isinstance(op, Operator):
        if op.is_device_dense:
            raise MethodError("dense device x dense device products are BLAS territory (operators_dense.jl:394), "
                              "not part of the lazy/sparse mul! path")
        f = _factor_struct(op.data, keep)
        if op.is_sparse:
            _lib.check(lib.qob_sparse_create(ctx, C.byref(f), C.byref(h)))
        elif op.is_eye:
            one = (C.c_int64 * 1)(len(op.basis_l))
            two = (C.c_int64 * 1)(len(op.basis_r))
            site = (C.c_int32 * 1)(1)
            _lib.check(lib.qob_lazytensor_create(ctx, 1, one, two, 1, site, C.byref(f), c64.of(1.0), C.byref(h)))
        else:
            _lib.check(lib.qob_dense_create(ctx, C.byref(f), C.byref(h)))
    elif isinstance(op, LazySum):
        hs = [handle(o, ctx) for o in op.operators]
        n = len(hs)
        arr = (C.c_void_p * max(n, 1))(*[x.value if isinstance(x, C.c_void_p) else x for x in hs])
        cf = (c64 * max(n, 1))(*[c64.of(f) for f in op.factors])
        _lib.check(lib.qob_lazysum_create(ctx, len(op.basis_l), len(op.basis_r), n, cf, arr, C.byref(h)))
        op._sent_factors = list(op.factors)
    elif isinstance(op, LazyProduct):
        hs = [handle(o, ctx) for o in op.operators]
        arr = (C.c_void_p * len(hs))(*[x.value if isinstance(x, C.c_void_p) else x for x in hs])
        _lib.check(lib.qob_lazyproduct_create(ctx, len(hs), arr, c64.of(op.factor), C.byref(h)))
    elif isinstance(op, LazyDirectSum):
        hs = [handle(o, ctx) for o in op.operators]
        arr = (C.c_void_p * len(hs))(*[x.value if isinstance(x, C.c_void_p) else x for x in hs])
        _lib.check(lib.qob_lazydirectsum_create(ctx, len(hs), arr, C.byref(h)))
    else:
        raise MethodError(f"no mul! method for operator type {type(op).__name__}")
    op._handle, op._handle_ctx = h, ctx
    return h


def _refresh_coefs(op: LazySum):
    if len(op.factors) != len(op.operators):
        raise ArgumentError("LazySum `operators` and `factors` have different lengths.")
    if getattr(op, "_sent_factors", None) != op.factors:
        n = len(op.factors)
        cf = (c64 * max(n, 1))(*[c64.of(f) for f in op.factors])
        _lib.check(lib.qob_lazysum_set_coefs(op._handle, n, cf))
        op._sent_factors = list(op.factors)
    for o in op.operators:
        if isinstance(o, LazySum) and getattr(o, "_handle", None):
            _refresh_coefs(o)


class LindbladRHS:
    """Fused master-equation right-hand side on a dense device rho:

        drho = alpha * ( -i (H rho - rho H) + sum_k r_k ( J_k rho J_k^+ - (J_k^+ J_k rho + rho J_k^+ J_k) / 2 ) ) + beta * drho

    i.e. what the six-mul!-per-jump call pattern of test/test_sciml_broadcast_interfaces.jl:36-43 computes, in ONE kernel
    (`qob_lindblad_create` / `qob_lindblad_apply`).  H and the J_k are `Operator`s with sparse or host-dense data on the same
    basis; `rates` are non-negative reals (default 1)."""

    def __init__(self, H, J=(), rates=None, ctx=None):
        J = list(J)
        for o in [H] + J:
            if not isinstance(o, Operator) or o.is_device_dense:
                raise MethodError("LindbladRHS needs Operators with sparse or host-dense data")
            if not (o.basis_l == H.basis_l and o.basis_r == H.basis_r):
                raise IncompatibleBases("LindbladRHS: H and the jump operators must share their bases")
        if not (H.basis_l == H.basis_r):
            raise IncompatibleBases("LindbladRHS: the Hamiltonian must map a basis to itself")
        if rates is not None and len(rates) != len(J):
            raise ArgumentError("LindbladRHS: one rate per jump operator")
        self.H, self.J, self.rates = H, J, None if rates is None else [float(r) for r in rates]
        self.basis_l = self.basis_r = H.basis_l
        ctx = _lib.context() if ctx is None else ctx
        keep = []
        hf = _factor_struct(H.data, keep)
        n = len(J)
        jf = (_lib.Factor * max(n, 1))()
        for k, o in enumerate(J):
            jf[k] = _factor_struct(o.data, keep)
        rt = None if self.rates is None else (C.c_double * max(n, 1))(*self.rates)
        h = C.c_void_p()
        _lib.check(lib.qob_lindblad_create(ctx, C.byref(hf), n, jf, rt, C.byref(h)))
        self._handle, self._handle_ctx = h, ctx

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                lib.qob_op_destroy(h)
            except Exception:
                pass
            self._handle = None

    def apply_(self, drho, rho, alpha=1.0, beta=0.0):
        """drho <- alpha * L(rho) + beta * drho; both are dense device Operators on the Hamiltonian's basis."""
        for o in (drho, rho):
            if not isinstance(o, Operator) or not o.is_device_dense:
                raise MethodError("LindbladRHS.apply_ needs dense device Operators")
            if not (o.basis_l == self.basis_l and o.basis_r == self.basis_r):
                raise IncompatibleBases("LindbladRHS.apply_: rho and drho must live on the Hamiltonian's basis")
        _lib.check(lib.qob_lindblad_apply(self._handle, c64.of(alpha), C.c_void_p(rho.data.data_ptr()), c64.of(beta),
                                          C.c_void_p(drho.data.data_ptr()), _stream()))
        return drho

    def assembled(self, which):
        """host matrices the kernel works from: 0: H - i/2 sum r_k J_k^+ J_k, 1: H + i/2 sum r_k J_k^+ J_k, 2+k: sqrt(r_k) J_k"""
        D = len(self.basis_l)
        out = np.zeros((D, D), dtype=C128, order="F")
        _lib.check(lib.qob_lindblad_dense(self._handle, int(which), C.c_void_p(out.ctypes.data)))
        return out

    def describe(self):
        buf = C.create_string_buffer(1 << 12)
        _lib.check(lib.qob_op_describe(self._handle, _lib.SIDE_LEFT, 1, buf, len(buf)))
        return buf.value.decode()


def describe(op, side="left", batch=1, ctx=None):
    """Text description of the device plan chosen for `op` (kernels, passes, tiles)."""
    buf = C.create_string_buffer(1 << 16)
    s = _lib.SIDE_LEFT if side in ("left", 0) else _lib.SIDE_RIGHT
    _lib.check(lib.qob_op_describe(handle(op, ctx), s, int(batch), buf, len(buf)))
    return buf.value.decode()


# ------------------------------------------------------------------------------------ mul!
def _stream():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _is_state_op(x):
    return isinstance(x, Operator) and x.is_device_dense


def mul_(result, a, b, alpha=1.0, beta=0.0):
    """mul!(result, a, b, alpha, beta) -> result:  result = alpha*a*b + beta*result.

    Dispatch (who is the operator, who is the state) follows the reference's method table:
      Ket  <- op * Ket                     (operators_lazytensor.jl:539, lazysum:189, lazyproduct:103, sparse:201)
      Bra  <- Bra * op                     (:559, :202, :117, sparse:202)
      Op   <- op * DenseOp                 (:576, :215, :131, sparse:199)
      Op   <- DenseOp * op                 (:593, :227, :148, sparse:200)
    alpha/beta: any Python number (bool/int/float/complex), promoted to ComplexF64."""
    alpha, beta = complex(alpha), complex(beta)
    if isinstance(a, TimeDependentSum):
        a = a.static_op
    if isinstance(b, TimeDependentSum):
        b = b.static_op
    if _is_state_op(a) and (_is_state_op(b) or isinstance(b, Ket)) or (isinstance(a, Bra) and _is_state_op(b)):
        return _dense_device_mul(result, a, b, alpha, beta)
    if isinstance(b, Ket) and isinstance(a, AbstractOperator):
        if not isinstance(result, Ket):
            raise MethodError("result must be a Ket")
        side, op, state, batch = _lib.SIDE_LEFT, a, b, 1
        _check_bases(result.basis, op.basis_l, op.basis_r, state.basis, (len(result.basis),), (len(state.basis),), op)
    elif isinstance(a, Bra) and isinstance(b, AbstractOperator):
        if not isinstance(result, Bra):
            raise MethodError("result must be a Bra")
        side, op, state, batch = _lib.SIDE_RIGHT, b, a, 1
        _check_bases(state.basis, op.basis_l, op.basis_r, result.basis, (len(state.basis),), (len(result.basis),), op,
                     right=True)
    elif _is_state_op(b) and isinstance(a, AbstractOperator) and not _is_state_op(a):
        if not _is_state_op(result):
            raise MethodError("result must be a dense device Operator")
        side, op, state, batch = _lib.SIDE_LEFT, a, b, b.data.shape[1]
        if result.basis_r != b.basis_r:
            raise IncompatibleBases()
        _check_bases(result.basis_l, op.basis_l, op.basis_r, state.basis_l, result.data.shape, state.data.shape, op)
    elif _is_state_op(a) and isinstance(b, AbstractOperator) and not _is_state_op(b):
        if not _is_state_op(result):
            raise MethodError("result must be a dense device Operator")
        side, op, state, batch = _lib.SIDE_RIGHT, b, a, a.data.shape[0]
        if result.basis_l != a.basis_l:
            raise IncompatibleBases()
        _check_bases(state.basis_r, op.basis_l, op.basis_r, result.basis_r, state.data.shape, result.data.shape, op,
                     right=True)
    else:
        raise MethodError(f"no mul! method for ({type(result).__name__}, {type(a).__name__}, {type(b).__name__})")
    if isinstance(op, Operator) and op.is_sparse and batch == 1 and isinstance(state, StateVector) \
            and isinstance(op.data, Adjoint):
        # gemv! exists for SparseOpPureType only (operators_sparse.jl:201-202)
        raise MethodError("mul!(Ket/Bra, adjoint sparse operator) has no sparse gemv! method")
    h = handle(op)
    _lib.check(lib.qob_op_apply(h, side, c64.of(alpha), C.c_void_p(state.data.data_ptr()), c64.of(beta),
                                C.c_void_p(result.data.data_ptr()), int(batch), _stream()))
    return result


def _dense_device_mul(result, a, b, alpha, beta):
    """Dense x dense on device data: the reference forwards `.data` to BLAS (src/operators_dense.jl:394-396), and so does
    this: a plain library GEMM/GEMV (cuBLAS zgemm/zgemv through torch), not one of this repository's kernels.  With
    CuArray-backed data the Julia side needs no glue at all for these three methods (CUDA.jl's mul! is cuBLAS)."""
    import torch

    if _is_state_op(a) and _is_state_op(b):
        if a.basis_r != b.basis_l or result.basis_l != a.basis_l or result.basis_r != b.basis_r:
            raise IncompatibleBases() if a.data.shape[1] == b.data.shape[0] else DimensionMismatch("A and B dimensions do not match")
        if result.data.data_ptr() in (a.data.data_ptr(), b.data.data_ptr()):
            raise ArgumentError("output matrix must not be aliased with input matrix")
        torch.addmm(result.data, a.data, b.data, beta=beta, alpha=alpha, out=result.data)
    elif _is_state_op(a):   # Ket <- dense Operator * Ket
        if a.basis_r != b.basis or result.basis != a.basis_l:
            raise IncompatibleBases() if a.data.shape[1] == b.data.numel() else DimensionMismatch("A and B dimensions do not match")
        torch.addmv(result.data, a.data, b.data, beta=beta, alpha=alpha, out=result.data)
    else:                   # Bra <- Bra * dense Operator: mul!(result.data, transpose(b.data), a.data, alpha, beta)
        if a.basis != b.basis_l or result.basis != b.basis_r:
            raise IncompatibleBases() if a.data.numel() == b.data.shape[0] else DimensionMismatch("A and B dimensions do not match")
        torch.addmv(result.data, b.data.t(), a.data, beta=beta, alpha=alpha, out=result.data)
    return result


def _check_bases(out_basis, op_bl, op_br, in_basis, out_shape, in_shape, op, right=False):
    """Bases are type parameters in the reference: a mismatch never reaches the kernels.  Size
    mismatches raise DimensionMismatch (sparsematrix.jl:100-102, operators_lazytensor.jl:695-703),
    equal sizes with different bases raise IncompatibleBases."""
    if not right:
        ok = (out_basis == op_bl) and (in_basis == op_br)
        size_ok = len(out_basis) == len(op_bl) and len(in_basis) == len(op_br)
    else:
        ok = (out_basis == op_bl) and (in_basis == op_br)
        size_ok = len(out_basis) == len(op_bl) and len(in_basis) == len(op_br)
    if not size_ok:
        raise DimensionMismatch(f"operator is {len(op_bl)}x{len(op_br)}, state/result have sizes {in_shape}/{out_shape}")
    if not ok:
        raise IncompatibleBases()


def apply_host(op, x, side="left", alpha=1.0, beta=0.0, y=None, batch=1):
    """End-to-end call with HOST buffers (numpy, ideally pinned): H2D, apply, D2H inside libqob200
    (`qob_op_apply_host`).  Returns y (numpy complex128, column-major)."""
    s = _lib.SIDE_LEFT if side in ("left", 0) else _lib.SIDE_RIGHT
    dl, dr = len(op.basis_l), len(op.basis_r)
    n_out = (dl if s == _lib.SIDE_LEFT else dr) * batch
    xh = np.ascontiguousarray(np.asarray(x, dtype=C128).reshape(-1, order="F"))
    if y is None:
        y = np.zeros(n_out, dtype=C128)
    assert y.dtype == C128 and y.flags.c_contiguous and y.size == n_out
    _lib.check(lib.qob_op_apply_host(handle(op), s, c64.of(alpha), C.c_void_p(xh.ctypes.data), c64.of(beta),
                                     C.c_void_p(y.ctypes.data), int(batch)))
    return y


# ------------------------------------------------------------------------------------ site operators
def _spdiagm(n, offsets):
    m = sp.lil_matrix((n, n), dtype=C128)
    for off, vals in offsets.items():
        for t, v in enumerate(vals):
            i, j = (t, t + off) if off >= 0 else (t - off, t)
            m[i, j] = v
    return sp.csc_matrix(m)


def sigmax(b: SpinBasis):
    """src/spin.jl:14-20"""
    n, s = len(b), b.spinnumber
    d = [math.sqrt((s + 1) * 2 * a - a * (a + 1)) for a in range(1, n)]
    return Operator(b, b, _spdiagm(n, {1: d, -1: d}))


def sigmay(b: SpinBasis):
    """src/spin.jl:34-40"""
    n, s = len(b), b.spinnumber
    d = [1j * math.sqrt((s + 1) * 2 * a - a * (a + 1)) for a in range(1, n)]
    return Operator(b, b, _spdiagm(n, {-1: d, 1: [-v for v in d]}))


def sigmaz(b: SpinBasis):
    """src/spin.jl:54-60"""
    n, s = len(b), b.spinnumber
    return Operator(b, b, _spdiagm(n, {0: [2 * (s - t) for t in range(n)]}))


def sigmap(b: SpinBasis):
    """src/spin.jl:68-75"""
    n, s = len(b), b.spinnumber
    S = (s + 1) * s
    return Operator(b, b, _spdiagm(n, {1: [math.sqrt(S - m * (m + 1)) for m in [s - 1 - t for t in range(n - 1)]]}))


def sigmam(b: SpinBasis):
    """src/spin.jl:83-90"""
    n, s = len(b), b.spinnumber
    S = (s + 1) * s
    return Operator(b, b, _spdiagm(n, {-1: [math.sqrt(S - m * (m - 1)) for m in [s - t for t in range(n - 1)]]}))


def number(b: FockBasis):
    """src/fock.jl:8-12"""
    return Operator(b, b, _spdiagm(len(b), {0: [float(v) for v in range(b.offset, b.N + 1)]}))


def destroy(b: FockBasis):
    """src/fock.jl:22-28"""
    return Operator(b, b, _spdiagm(len(b), {1: [math.sqrt(float(v)) for v in range(b.offset + 1, b.N + 1)]}))


def create(b: FockBasis):
    """src/fock.jl:38-44"""
    return Operator(b, b, _spdiagm(len(b), {-1: [math.sqrt(float(v)) for v in range(b.offset + 1, b.N + 1)]}))


def transition(b: NLevelBasis, to, frm):
    """src/nlevel.jl:8-18"""
    if not (1 <= to <= b.N and 1 <= frm <= b.N):
        raise IndexError("BoundsError: transition indices must be between 1 and b.N")
    m = sp.lil_matrix((b.N, b.N), dtype=C128)
    m[to - 1, frm - 1] = 1.0
    return Operator(b, b, sp.csc_matrix(m))


def randstate(b, seed=0, normalize=True):
    """randstate(b) (src/state_definitions.jl:6-10) with the counter-based generator shared with the
    oracle (qob_fill_state): re/im uniform in [-1, 1), then normalised on the device."""
    import torch

    k = Ket(b)
    fill_state(k.data, seed)
    if normalize:
        k.data /= math.sqrt(norm2(k.data))
    return k


def fill_state(t, seed, scale=1.0, offset=0):
    _lib.check(lib.qob_fill_state(C.c_void_p(t.data_ptr()), int(offset), t.numel(), C.c_uint64(seed), float(scale),
                                  _stream()))
    return t


def norm2(t):
    out = C.c_double()
    _lib.check(lib.qob_norm2(C.c_void_p(t.data_ptr()), t.numel(), C.byref(out), _stream()))
    return out.value


def dot(x, y):
    out = c64()
    _lib.check(lib.qob_dot(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), x.numel(), C.byref(out), _stream()))
    return complex(out.re, out.im)


def launch_count(family=None):
    """kernel launches issued by the library in this process; `family`: 1 round-1 tile kernel, 2 round-2 tile kernel,
    3 round-2 peer-addressed (exchange), 4 round-1 peer-addressed"""
    return int(lib.qob_launch_count() if family is None else lib.qob_launch_count_of(family))


def expect(op, state):
    """expect(op, state) (src/operators.jl:119-142): <psi|op|psi> for a Ket (`qob_expect`: mul! into the handle's scratch and one
    deterministic device reduction, only the scalar crosses to the host), tr(op*rho) for a dense device Operator."""
    import torch

    if isinstance(state, Ket):
        if op.basis_r != state.basis or op.basis_l != state.basis:
            raise IncompatibleBases()
        out = c64()
        _lib.check(lib.qob_expect(handle(op), C.c_void_p(state.data.data_ptr()), C.byref(out), _stream()))
        return complex(out.re, out.im)
    if _is_state_op(state):
        tmp = DenseOperator(op.basis_l, state.basis_r)
        mul_(tmp, op, state, 1.0, 0.0)
        return complex(torch.diagonal(tmp.data).sum().item())
    raise MethodError(f"expect: unsupported state type {type(state).__name__}")


def variance(op, state):
    """variance(op, state) = psi'(op(op psi)) - (psi'(op psi))^2 (src/operators.jl:139-142): `qob_variance`, two applications
    like the reference."""
    if not isinstance(state, Ket):
        raise MethodError("variance: only Ket states are supported on the device path")
    if op.basis_r != state.basis or op.basis_l != state.basis:
        raise IncompatibleBases()
    out = c64()
    _lib.check(lib.qob_variance(handle(op), C.c_void_p(state.data.data_ptr()), C.byref(out), _stream()))
    return complex(out.re, out.im)


def ptrace(a, indices):
    """ptrace(a, indices) (src/operators_dense.jl:191-215): partial trace of a dense device Operator, a Ket or a Bra over the
    1-based subsystem `indices`; the result is a dense device Operator on the remaining subsystems."""
    import torch

    idx = [int(indices)] if isinstance(indices, (int, np.integer)) else [int(i) for i in indices]
    tr = (C.c_int32 * max(len(idx), 1))(*idx)
    ctx = _lib.context()

    def reduced_basis(b):
        if not isinstance(b, CompositeBasis):
            raise ArgumentError("Partial trace can only be applied onto operators with composite bases.")
        kept = [bb for k, bb in enumerate(b.bases) if (k + 1) not in idx]
        return kept[0] if len(kept) == 1 else CompositeBasis(kept)

    if isinstance(a, (Ket, Bra)):
        b = a.basis
        if not isinstance(b, CompositeBasis):
            raise ArgumentError("Partial trace can only be applied onto states with composite bases.")
        dims = (C.c_int64 * len(b.shape))(*b.shape)
        # argument errors (all subsystems traced, bad index) are raised by the library before anything is allocated
        m = 1
        for k, d in enumerate(b.shape):
            if (k + 1) not in idx:
                m *= d
        res = torch.empty((m, m), dtype=torch.complex128, device=a.data.device).t()   # column-major m x m
        _lib.check(lib.qob_ptrace_state(ctx, len(b.shape), dims, len(idx), tr, 1 if isinstance(a, Bra) else 0,
                                        C.c_void_p(a.data.data_ptr()), C.c_void_p(res.data_ptr()), _stream()))
        rb = reduced_basis(b) if len(idx) < len(b.shape) else b
        return DenseOperator(rb, rb, res)
    if isinstance(a, Operator) and a.is_device_dense:
        bl, br = a.basis_l, a.basis_r
        if not isinstance(bl, CompositeBasis) or not isinstance(br, CompositeBasis):
            raise ArgumentError("Partial trace can only be applied onto operators with composite bases.")
        if len(bl.shape) != len(br.shape):
            raise ArgumentError("Partial trace can only be applied onto operators wich have the same number of subsystems in the "
                                "left basis and right basis.")
        dl = (C.c_int64 * len(bl.shape))(*bl.shape)
        dr = (C.c_int64 * len(br.shape))(*br.shape)
        ml = mr = 1
        for k in range(len(bl.shape)):
            if (k + 1) not in idx:
                ml *= bl.shape[k]
                mr *= br.shape[k]
        res = torch.empty((mr, ml), dtype=torch.complex128, device=a.data.device).t()   # column-major ml x mr
        _lib.check(lib.qob_ptrace_op(ctx, len(bl.shape), dl, dr, len(idx), tr, C.c_void_p(a.data.data_ptr()), C.c_void_p(res.data_ptr()),
                                     _stream()))
        return DenseOperator(reduced_basis(bl), reduced_basis(br), res)
    raise MethodError(f"ptrace: unsupported argument type {type(a).__name__}")


def reduced(a, indices):
    """reduced(a, indices) (QuantumInterface): the state of the subsystems `indices` = ptrace over their complement."""
    idx = [int(indices)] if isinstance(indices, (int, np.integer)) else [int(i) for i in indices]
    b = a.basis if isinstance(a, (Ket, Bra)) else a.basis_l
    n = len(b.shape)
    return ptrace(a, [k for k in range(1, n + 1) if k not in idx])


def profile_enable(on=True):
    _lib.check(lib.qob_profile_enable(1 if on else 0))


def profile_read(max_entries=4096):
    """[(ms, pass_index, algorithmic_bytes)] of every tile-pass launch since profile_enable, in launch order."""
    ms = (C.c_float * max_entries)()
    pi = (C.c_int32 * max_entries)()
    by = (C.c_double * max_entries)()
    n = C.c_int32()
    _lib.check(lib.qob_profile_read(max_entries, ms, pi, by, C.byref(n)))
    return [(float(ms[i]), int(pi[i]), float(by[i])) for i in range(n.value)]
