"""quantumopticsbase.jl_b200 — B200-native `mul!` hot path of QuantumOpticsBase.jl.

Import as `qob200` (the directory name is not a valid Python identifier; `qob200.py` at the repo root
loads this package under that name).  Contents: `csrc/` (CUDA kernels + C ABI -> libqob200.so),
`_lib.py` (ctypes binding), `operators.py` (host mirror of the reference's operator interface),
`dist.py` (sharded multi-GPU apply).
"""
from ._lib import (ArgumentError, CudaError, DimensionMismatch, IncompatibleBases, MethodError, LIB_PATH, EXPORTED,
                   context, lib)
from .operators import (Adjoint, Basis, Bra, CompositeBasis, DenseOperator, Eye, FockBasis, GenericBasis, Ket,
                        LazyDirectSum, LazyProduct, LazySum, LazyTensor, LindbladRHS, SumBasis, directsum, ptrace, reduced, NLevelBasis, Operator, SparseOperator, SpinBasis, TimeDependentSum,
                        apply_host, create, dagger, dense, describe, destroy, dot, expect, fill_state, handle,
                        identityoperator, launch_count, mul_, norm2, number, profile_enable, profile_read, randstate, sigmam, sigmap, sigmax,
                        sigmay, sigmaz, sparse, tensor, transition, variance)

mul = mul_  # `mul!`
