"""ctypes binding of libqob200.so (include/qob200.h).

The library is built in-tree by `csrc/build.sh` (see __graft_entry__.build).  There is no CPU
fallback anywhere in this package: if the shared library is missing, importing raises; if no CUDA
device is present, every compute call raises `CudaError`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqob200.so")


class DimensionMismatch(Exception):
    """Julia DimensionMismatch (QOB_STATUS_DIM_MISMATCH)."""


class IncompatibleBases(Exception):
    """QuantumInterface.IncompatibleBases — bases are type parameters in the reference."""


class ArgumentError(Exception):
    """Julia ArgumentError (QOB_STATUS_INVALID_ARG / QOB_STATUS_ALIASING)."""


class MethodError(Exception):
    """Julia MethodError: no kernel for these operand types (QOB_STATUS_UNSUPPORTED)."""


class CudaError(RuntimeError):
    """CUDA / NCCL failure, or no device: the product has no CPU fallback."""


class c64(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]

    @staticmethod
    def of(z):
        z = complex(z)
        return c64(z.real, z.imag)


class Factor(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("trans", C.c_int32),
        ("nrows", C.c_int64),
        ("ncols", C.c_int64),
        ("dense", C.c_void_p),
        ("colptr", C.c_void_p),
        ("rowval", C.c_void_p),
        ("nzval", C.c_void_p),
    ]


FACTOR_DENSE, FACTOR_CSC, FACTOR_EYE = 0, 1, 2
OP_N, OP_T, OP_C = 0, 1, 2
SIDE_LEFT, SIDE_RIGHT = 0, 1

STATUS_EXC = {
    1: DimensionMismatch,
    2: ArgumentError,
    3: ArgumentError,
    4: MethodError,
    5: CudaError,
    6: CudaError,
    7: MemoryError,
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with quantumopticsbase.jl_b200/csrc/build.sh "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no fallback path."
    )
lib = C.CDLL(LIB_PATH)

_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
_sigs = {
    "qob_version": (C.c_int, []),
    "qob_last_error": (C.c_char_p, []),
    "qob_status_string": (C.c_char_p, [C.c_int]),
    "qob_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "qob_ctx_destroy": (C.c_int, [_vp]),
    "qob_ctx_scratch_bytes": (C.c_int, [_vp, C.POINTER(_i64)]),
    "qob_ctx_clear_scratch": (C.c_int, [_vp]),
    "qob_lazytensor_create": (C.c_int, [_vp, _i32, C.POINTER(_i64), C.POINTER(_i64), _i32, C.POINTER(_i32),
                                        C.POINTER(Factor), c64, C.POINTER(_vp)]),
    "qob_sparse_create": (C.c_int, [_vp, C.POINTER(Factor), C.POINTER(_vp)]),
    "qob_dense_create": (C.c_int, [_vp, C.POINTER(Factor), C.POINTER(_vp)]),
    "qob_lazysum_create": (C.c_int, [_vp, _i64, _i64, _i32, C.POINTER(c64), C.POINTER(_vp), C.POINTER(_vp)]),
    "qob_lazysum_set_coefs": (C.c_int, [_vp, _i32, C.POINTER(c64)]),
    "qob_lazyproduct_create": (C.c_int, [_vp, _i32, C.POINTER(_vp), c64, C.POINTER(_vp)]),
    "qob_lindblad_create": (C.c_int, [_vp, C.POINTER(Factor), _i32, C.POINTER(Factor), C.POINTER(C.c_double), C.POINTER(_vp)]),
    "qob_lindblad_apply": (C.c_int, [_vp, c64, _vp, c64, _vp, _vp]),
    "qob_lindblad_dense": (C.c_int, [_vp, _i32, _vp]),
    "qob_op_destroy": (C.c_int, [_vp]),
    "qob_op_dims": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "qob_op_apply": (C.c_int, [_vp, _i32, c64, _vp, c64, _vp, _i64, _vp]),
    "qob_op_apply_host": (C.c_int, [_vp, _i32, c64, _vp, c64, _vp, _i64]),
    "qob_launch_count": (_i64, []),
    "qob_launch_count_of": (_i64, [_i32]),
    "qob_op_describe": (C.c_int, [_vp, _i32, _i64, C.c_char_p, _i64]),
    "qob_fill_state": (C.c_int, [_vp, _i64, _i64, C.c_uint64, C.c_double, _vp]),
    "qob_norm2": (C.c_int, [_vp, _i64, C.POINTER(C.c_double), _vp]),
    "qob_dot": (C.c_int, [_vp, _vp, _i64, C.POINTER(c64), _vp]),
    "qob_profile_enable": (C.c_int, [_i32]),
    "qob_profile_read": (C.c_int, [_i32, C.POINTER(C.c_float), C.POINTER(_i32), C.POINTER(C.c_double), C.POINTER(_i32)]),
    "qob_lazysum_term_masks": (C.c_int, [_vp, _i32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "qob_layout_plan_create": (C.c_int, [_vp, _i32, C.POINTER(_i32), C.c_uint64, C.POINTER(C.c_uint8), C.POINTER(_i32)]),
    "qob_layout_plan_apply": (C.c_int, [_vp, _i32, c64, _vp, c64, _vp, _vp]),
    "qob_layout_plan_describe": (C.c_int, [_vp, _i32, C.c_char_p, _i64]),
    "qob_layout_plan_apply_ex": (C.c_int, [_vp, _i32, c64, _vp, c64, _vp, _vp, _i32, C.POINTER(_vp), C.POINTER(_vp), _i32, _i32,
                                           _i32, _i32, _vp]),
    "qob_layout_plan_info": (C.c_int, [_vp, _i32, C.POINTER(_i32), C.POINTER(C.c_uint64)]),
    "qob_layout_plan_set_chunk_bits": (C.c_int, [_vp, _i32, C.c_uint64]),
    "qob_set_sm_budget": (C.c_int, [_i32]),
    "qob_dist_alloc": (C.c_int, [_vp, _i64, C.POINTER(_vp)]),
    "qob_dist_free": (C.c_int, [_vp, _vp]),
    "qob_ipc_export": (C.c_int, [_vp, C.POINTER(C.c_uint8)]),
    "qob_ipc_open": (C.c_int, [_vp, C.POINTER(C.c_uint8), C.POINTER(_vp)]),
    "qob_ipc_close": (C.c_int, [_vp, _vp]),
    "qob_dist_create": (C.c_int, [_vp, _i32, _i32, C.POINTER(_vp)]),
    "qob_dist_info": (C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64), C.POINTER(_i64)]),
    "qob_dist_bind": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "qob_dist_apply": (C.c_int, [_vp, c64, c64, _vp, _vp]),
    "qob_dist_exchange_timing": (C.c_int, [_vp, _i32]),
    "qob_dist_exchange_ms": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(_i32), C.POINTER(_i64)]),
    "qob_dist_direct_capable": (C.c_int, [_vp, C.POINTER(_i32)]),
    "qob_dist_bind_result": (C.c_int, [_vp, C.POINTER(_vp)]),
    "qob_dist_describe": (C.c_int, [_vp, C.c_char_p, _i64]),
    "qob_dist_destroy": (C.c_int, [_vp]),
    "qob_lazydirectsum_create": (C.c_int, [_vp, _i32, C.POINTER(_vp), C.POINTER(_vp)]),
    "qob_expect": (C.c_int, [_vp, _vp, C.POINTER(c64), _vp]),
    "qob_variance": (C.c_int, [_vp, _vp, C.POINTER(c64), _vp]),
    "qob_ptrace_op": (C.c_int, [_vp, _i32, C.POINTER(_i64), C.POINTER(_i64), _i32, C.POINTER(_i32), _vp, _vp, _vp]),
    "qob_ptrace_state": (C.c_int, [_vp, _i32, C.POINTER(_i64), _i32, C.POINTER(_i32), _i32, _vp, _vp, _vp]),
}
for _name, (_res, _args) in _sigs.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args
EXPORTED = sorted(_sigs)


def check(status: int):
    if status != 0:
        msg = lib.qob_last_error().decode("utf-8", "replace")
        raise STATUS_EXC.get(status, RuntimeError)(f"{lib.qob_status_string(status).decode()}: {msg}")


_ctx_cache = {}


def context(device=None):
    """qob_ctx for `device` (an int, or None = current torch CUDA device; -1 = planning-only, no GPU)."""
    if device is None:
        import torch

        if not torch.cuda.is_available():
            raise CudaError("no CUDA device available: quantumopticsbase.jl_b200 has no CPU fallback")
        device = torch.cuda.current_device()
    if device not in _ctx_cache:
        h = _vp()
        check(lib.qob_ctx_create(int(device), C.byref(h)))
        _ctx_cache[device] = h
    return _ctx_cache[device]
