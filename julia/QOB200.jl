# QOB200.jl — thin Julia glue that makes libqob200.so a drop-in for QuantumOpticsBase's `mul!` hot path.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: neither this image nor the GPU boxes have a Julia toolchain
# (SURVEY.md "Three facts", §8b).  It is written to the reference's method table so that a maintainer can
# load it next to QuantumOpticsBase + CUDA.jl; the same C ABI is exercised from Python (ctypes) by tests/.
#
# What it does
#   * device-resident states: `Ket{B,<:CuVector{ComplexF64}}`, `Operator{BL,BR,<:CuMatrix{ComplexF64}}`
#     (made by `Adapt.adapt(CuArray, x)`, reference src/states.jl:315-316, src/operators_dense.jl:47);
#   * operator definitions (LazyTensor / LazySum / LazyProduct / SparseOperator / dense site operators) stay
#     host objects; the first `mul!` compiles them into a libqob200 handle cached per (operator object, device) in a
#     WeakKeyDict, so the handle (and its device tables) is released when the operator is collected;
#     the definition is SNAPSHOTTED at that point: cheap scalars (`LazySum.factors`) are re-sent on every call, a change of
#     `LazyTensor.factor` / `LazyProduct.factor` or of the operator lists is detected by a fingerprint and rebuilds the handle,
#     in-place edits of a site operator's `.data` need `QOB200.invalidate!(op)`;
#   * `mul!(result, op, state, alpha, beta)` methods with the SAME signatures as
#     src/operators_lazytensor.jl:539-609, src/operators_lazysum.jl:189-238,
#     src/operators_lazyproduct.jl:103-163, src/operators_sparse.jl:199-202 forward to `qob_op_apply`.
module QOB200

using LinearAlgebra, SparseArrays
import LinearAlgebra: mul!
using CUDA
using FillArrays: Eye
using QuantumOpticsBase
using QuantumOpticsBase: Ket, Bra, Operator, LazyTensor, LazySum, LazyProduct, DataOperator, AbstractOperator,
                         Basis, CompositeBasis, IncompatibleBases, LazyDirectSum
import QuantumOpticsBase: expect, variance, ptrace

const libqob200 = get(ENV, "LIBQOB200", joinpath(@__DIR__, "..", "quantumopticsbase.jl_b200", "libqob200.so"))

struct C64
    re::Float64
    im::Float64
end
C64(z::Number) = (c = ComplexF64(z); C64(real(c), imag(c)))   # Bool/Int/Float/Complex are all promoted

# qob_factor (include/qob200.h)
struct QobFactor
    kind::Int32
    trans::Int32
    nrows::Int64
    ncols::Int64
    dense::Ptr{ComplexF64}
    colptr::Ptr{Int64}
    rowval::Ptr{Int64}
    nzval::Ptr{ComplexF64}
end
const FACTOR_DENSE, FACTOR_CSC, FACTOR_EYE = Int32(0), Int32(1), Int32(2)
const OP_N, OP_T, OP_C = Int32(0), Int32(1), Int32(2)
const SIDE_LEFT, SIDE_RIGHT = Int32(0), Int32(1)

function check(status::Integer)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:qob_last_error, libqob200), Cstring, ()))
    status == 1 && throw(DimensionMismatch(msg))
    (status == 2 || status == 3) && throw(ArgumentError(msg))
    status == 4 && throw(MethodError(mul!, ()))           # unsupported factor / operand types
    error("libqob200 (status $status): $msg")
end

# ---------------------------------------------------------------- context (one per device)
const CONTEXTS = Dict{Int,Ptr{Cvoid}}()
function context()
    dev = Int(CUDA.deviceid(CUDA.device()))
    get!(CONTEXTS, dev) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qob_ctx_create, libqob200), Cint, (Cint, Ref{Ptr{Cvoid}}), dev, h))
        h[]
    end
end

# the reference's cache controls (src/operators_lazytensor.jl:241-279) map onto the scratch pool
lazytensor_cachesize() = (b = Ref{Int64}(0); check(ccall((:qob_ctx_scratch_bytes, libqob200), Cint, (Ptr{Cvoid}, Ref{Int64}), context(), b)); b[])
lazytensor_clear_cache() = check(ccall((:qob_ctx_clear_scratch, libqob200), Cint, (Ptr{Cvoid},), context()))

# ---------------------------------------------------------------- factors
# returns (QobFactor, objects that must stay alive during the ccall)
function factor(d::Matrix{ComplexF64}, trans=OP_N)
    QobFactor(FACTOR_DENSE, trans, size(d, 1), size(d, 2), pointer(d), C_NULL, C_NULL, C_NULL), (d,)
end
factor(d::AbstractMatrix{<:Number}, trans=OP_N) = factor(Matrix{ComplexF64}(d), trans)
function factor(d::SparseMatrixCSC, trans=OP_N)
    m = SparseMatrixCSC{ComplexF64,Int64}(d)
    QobFactor(FACTOR_CSC, trans, size(m, 1), size(m, 2), C_NULL, pointer(m.colptr), pointer(m.rowval), pointer(m.nzval)), (m,)
end
factor(d::Eye, trans=OP_N) = (QobFactor(FACTOR_EYE, trans, size(d, 1), size(d, 2), C_NULL, C_NULL, C_NULL, C_NULL), ())
factor(d::Adjoint, trans=OP_N) = factor(parent(d), trans == OP_N ? OP_C : error("nested adjoint"))
factor(d::Transpose, trans=OP_N) = factor(parent(d), trans == OP_N ? OP_T : error("nested transpose"))

# ---------------------------------------------------------------- handles, cached per (operator object, device)
mutable struct Handle
    ptr::Ptr{Cvoid}
    children::Vector{Any}      # child handles: kept alive as long as the parent handle lives
    fingerprint::UInt          # of the cheap, mutable parts of the definition (see `fingerprint`)
    function Handle(p, children=Any[], fp=UInt(0))
        h = new(p, children, fp)
        finalizer(x -> ccall((:qob_op_destroy, libqob200), Cint, (Ptr{Cvoid},), x.ptr), h)
        h
    end
end
# WeakKeyDict: an entry does not keep its operator alive, so operators built per time step are collected together with
# their device tables (the reference's LazyTensor / LazySum / LazyProduct are mutable structs: valid weak keys).  One
# dictionary per device: a handle belongs to the context (device) it was created on.
const HANDLES = Dict{Int,WeakKeyDict{Any,Handle}}()
const HANDLES_LOCK = ReentrantLock()
handles() = lock(HANDLES_LOCK) do
    get!(() -> WeakKeyDict{Any,Handle}(), HANDLES, Int(CUDA.deviceid(CUDA.device())))
end
"Forget the compiled handle of `op` (call after editing a site operator's `.data` in place)."
invalidate!(op) = (lock(HANDLES_LOCK) do; for d in values(HANDLES); delete!(d, op); end; end; op)

# what may change between two `mul!` calls without the object changing identity
fingerprint(op::LazyTensor) = hash((op.factor, op.indices, map(objectid, op.operators)))
fingerprint(op::LazyProduct) = hash((op.factor, map(objectid, op.operators)))
fingerprint(op::LazySum) = hash(map(objectid, op.operators))          # the factors are re-sent on every call
fingerprint(op::LazyDirectSum) = hash(map(objectid, op.operators))
fingerprint(op) = UInt(0)

function cached(build, op)
    d = handles()
    fp = fingerprint(op)
    lock(HANDLES_LOCK) do
        h = get(d, op, nothing)
        if h === nothing || h.fingerprint != fp
            h = build()
            h.fingerprint = fp
            d[op] = h
        end
        h
    end
end

shape(b::CompositeBasis) = Int64[length(x) for x in b.bases]
shape(b::Basis) = Int64[length(b)]

function handle(op::LazyTensor)
    cached(op) do
        facs = QobFactor[]
        keep = Any[]
        for o in op.operators
            o isa DataOperator || throw(MethodError(mul!, (op,)))
            f, k = factor(o.data)
            push!(facs, f); push!(keep, k)
        end
        dl, dr = shape(op.basis_l), shape(op.basis_r)
        sites = Int32[i for i in op.indices]
        out = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve keep facs dl dr sites begin
            check(ccall((:qob_lazytensor_create, libqob200), Cint,
                        (Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{Int32}, Ptr{QobFactor}, C64, Ref{Ptr{Cvoid}}),
                        context(), length(dl), dl, dr, length(sites), sites, facs, C64(op.factor), out))
        end
        Handle(out[])
    end
end

function handle(op::Operator)   # SparseOperator / dense operator definition (mutable struct in the reference: a valid weak key)
    cached(op) do
        f, keep = factor(op.data)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        fn = f.kind == FACTOR_CSC ? :qob_sparse_create : :qob_dense_create
        GC.@preserve keep begin
            if f.kind == FACTOR_CSC
                check(ccall((:qob_sparse_create, libqob200), Cint, (Ptr{Cvoid}, Ref{QobFactor}, Ref{Ptr{Cvoid}}), context(), Ref(f), out))
            else
                check(ccall((:qob_dense_create, libqob200), Cint, (Ptr{Cvoid}, Ref{QobFactor}, Ref{Ptr{Cvoid}}), context(), Ref(f), out))
            end
        end
        Handle(out[])
    end
end

function handle(op::LazySum)
    h = cached(op) do
        hs = [handle(o) for o in op.operators]
        ptrs = Ptr{Cvoid}[x.ptr for x in hs]
        coefs = C64[C64(f) for f in op.factors]
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qob_lazysum_create, libqob200), Cint,
                    (Ptr{Cvoid}, Int64, Int64, Int32, Ptr{C64}, Ptr{Ptr{Cvoid}}, Ref{Ptr{Cvoid}}),
                    context(), length(op.basis_l), length(op.basis_r), length(ptrs), coefs, ptrs, out))
        Handle(out[], hs)
    end
    # TimeDependentSum rewrites `factors` at every set_time! (src/time_dependent_operator.jl:279-290): resend, it is cheap
    coefs = C64[C64(f) for f in op.factors]
    check(ccall((:qob_lazysum_set_coefs, libqob200), Cint, (Ptr{Cvoid}, Int32, Ptr{C64}), h.ptr, length(coefs), coefs))
    h
end

function handle(op::LazyProduct)
    cached(op) do
        hs = [handle(o) for o in op.operators]
        ptrs = Ptr{Cvoid}[x.ptr for x in hs]
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qob_lazyproduct_create, libqob200), Cint, (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}, C64, Ref{Ptr{Cvoid}}),
                    context(), length(ptrs), ptrs, C64(op.factor), out))
        Handle(out[], hs)
    end
end

function handle(op::LazyDirectSum)
    cached(op) do
        hs = [handle(o) for o in op.operators]
        ptrs = Ptr{Cvoid}[x.ptr for x in hs]
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:qob_lazydirectsum_create, libqob200), Cint, (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}, Ref{Ptr{Cvoid}}),
                    context(), length(ptrs), ptrs, out))
        Handle(out[], hs)
    end
end

# ---------------------------------------------------------------- the one call
function apply!(y::CuArray{ComplexF64}, op, side::Int32, x::CuArray{ComplexF64}, alpha, beta, batch::Integer)
    h = handle(op)
    GC.@preserve h check(ccall((:qob_op_apply, libqob200), Cint,
                               (Ptr{Cvoid}, Int32, C64, CuPtr{Cvoid}, C64, CuPtr{Cvoid}, Int64, Ptr{Cvoid}),
                               h.ptr, side, C64(alpha), pointer(x), C64(beta), pointer(y), batch, CUDA.stream().handle))
    y
end

const DevVec = CuVector{ComplexF64}
const DevMat = CuMatrix{ComplexF64}
const DevOp{BL,BR} = Operator{BL,BR,<:DevMat}
const LazyOps{BL,BR} = Union{LazySum{BL,BR},LazyProduct{BL,BR}}
const HostSparseOrDense{BL,BR} = Operator{BL,BR,<:Union{SparseMatrixCSC,Adjoint{<:Number,<:SparseMatrixCSC},Matrix,Adjoint{<:Number,<:Matrix}}}

# LazyTensor — same type constraints as src/operators_lazytensor.jl:539,559,576,593 to avoid ambiguities
mul!(result::Ket{B1,<:DevVec}, a::LazyTensor{B1,B2,F,I,T}, b::Ket{B2,<:DevVec}, alpha, beta) where {B1<:Basis,B2<:Basis,F,I,T<:Tuple{Vararg{DataOperator}}} =
    (apply!(result.data, a, SIDE_LEFT, b.data, alpha, beta, 1); result)
mul!(result::Bra{B2,<:DevVec}, a::Bra{B1,<:DevVec}, b::LazyTensor{B1,B2,F,I,T}, alpha, beta) where {B1<:Basis,B2<:Basis,F,I,T<:Tuple{Vararg{DataOperator}}} =
    (apply!(result.data, b, SIDE_RIGHT, a.data, alpha, beta, 1); result)
mul!(result::DevOp{B1,B3}, a::LazyTensor{B1,B2,F,I,T}, b::DevOp{B2,B3}, alpha, beta) where {B1<:Basis,B2<:Basis,B3<:Basis,F,I,T<:Tuple{Vararg{DataOperator}}} =
    (apply!(result.data, a, SIDE_LEFT, b.data, alpha, beta, size(b.data, 2)); result)
mul!(result::DevOp{B1,B3}, a::DevOp{B1,B2}, b::LazyTensor{B2,B3,F,I,T}, alpha, beta) where {B1<:Basis,B2<:Basis,B3<:Basis,F,I,T<:Tuple{Vararg{DataOperator}}} =
    (apply!(result.data, b, SIDE_RIGHT, a.data, alpha, beta, size(a.data, 1)); result)

# LazySum / LazyProduct (src/operators_lazysum.jl:189-238, src/operators_lazyproduct.jl:103-163): ONE call, the
# per-term loop and the LazyProduct buffers live behind the ABI (device temporaries, not the host ket_l/bra_r)
mul!(result::Ket{B1,<:DevVec}, a::LazyOps{B1,B2}, b::Ket{B2,<:DevVec}, alpha, beta) where {B1,B2} =
    (apply!(result.data, a, SIDE_LEFT, b.data, alpha, beta, 1); result)
mul!(result::Bra{B2,<:DevVec}, a::Bra{B1,<:DevVec}, b::LazyOps{B1,B2}, alpha, beta) where {B1,B2} =
    (apply!(result.data, b, SIDE_RIGHT, a.data, alpha, beta, 1); result)
mul!(result::DevOp{B1,B3}, a::LazyOps{B1,B2}, b::DevOp{B2,B3}, alpha, beta) where {B1,B2,B3} =
    (apply!(result.data, a, SIDE_LEFT, b.data, alpha, beta, size(b.data, 2)); result)
mul!(result::DevOp{B1,B3}, a::DevOp{B1,B2}, b::LazyOps{B2,B3}, alpha, beta) where {B1,B2,B3} =
    (apply!(result.data, b, SIDE_RIGHT, a.data, alpha, beta, size(a.data, 1)); result)

# SparseOperator / dense operator definitions on the host applied to device states (src/operators_sparse.jl:199-202,
# src/operators_dense.jl:394-396).  Without these, a device `Operator` falls to the column-by-column fallback
# (src/operators_dense.jl:400-418).
mul!(result::DevOp{B1,B3}, M::HostSparseOrDense{B1,B2}, b::DevOp{B2,B3}, alpha, beta) where {B1,B2,B3} =
    (apply!(result.data, M, SIDE_LEFT, b.data, alpha, beta, size(b.data, 2)); result)
mul!(result::DevOp{B1,B3}, a::DevOp{B1,B2}, M::HostSparseOrDense{B2,B3}, alpha, beta) where {B1,B2,B3} =
    (apply!(result.data, M, SIDE_RIGHT, a.data, alpha, beta, size(a.data, 1)); result)
mul!(result::Ket{B1,<:DevVec}, M::HostSparseOrDense{B1,B2}, b::Ket{B2,<:DevVec}, alpha, beta) where {B1,B2} =
    (apply!(result.data, M, SIDE_LEFT, b.data, alpha, beta, 1); result)
mul!(result::Bra{B2,<:DevVec}, b::Bra{B1,<:DevVec}, M::HostSparseOrDense{B1,B2}, alpha, beta) where {B1,B2} =
    (apply!(result.data, M, SIDE_RIGHT, b.data, alpha, beta, 1); result)

# LazyDirectSum (src/spinors.jl:221-247): Ket and Bra only, like the reference
mul!(result::Ket{B1,<:DevVec}, M::LazyDirectSum{B1,B2}, b::Ket{B2,<:DevVec}, alpha, beta) where {B1,B2} =
    (apply!(result.data, M, SIDE_LEFT, b.data, alpha, beta, 1); result)
mul!(result::Bra{B2,<:DevVec}, b::Bra{B1,<:DevVec}, M::LazyDirectSum{B1,B2}, alpha, beta) where {B1,B2} =
    (apply!(result.data, M, SIDE_RIGHT, b.data, alpha, beta, 1); result)

# expect / variance of a lazy or sparse operator in a device Ket (src/operators.jl:119,139-142): one ccall each, only the scalar
# crosses to the host
const LazyOrSparse{B} = Union{LazyTensor{B,B},LazySum{B,B},LazyProduct{B,B},HostSparseOrDense{B,B}}
function expect(op::LazyOrSparse{B}, state::Ket{B,<:DevVec}) where B
    h = handle(op); out = Ref(C64(0))
    GC.@preserve h check(ccall((:qob_expect, libqob200), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, Ref{C64}, Ptr{Cvoid}),
                               h.ptr, pointer(state.data), out, CUDA.stream().handle))
    ComplexF64(out[].re, out[].im)
end
function variance(op::LazyOrSparse{B}, state::Ket{B,<:DevVec}) where B
    h = handle(op); out = Ref(C64(0))
    GC.@preserve h check(ccall((:qob_variance, libqob200), Cint, (Ptr{Cvoid}, CuPtr{Cvoid}, Ref{C64}, Ptr{Cvoid}),
                               h.ptr, pointer(state.data), out, CUDA.stream().handle))
    ComplexF64(out[].re, out[].im)
end

# ptrace of device-resident states (src/operators_dense.jl:191-215): the result is a device-resident dense Operator
function ptrace(a::DevOp, indices)
    QuantumOpticsBase.check_ptrace_arguments(a, indices)
    dl, dr = shape(a.basis_l), shape(a.basis_r)
    tr = Int32[i for i in indices]
    bl, br = ptrace(a.basis_l, indices), ptrace(a.basis_r, indices)
    res = CUDA.zeros(ComplexF64, length(bl), length(br))
    check(ccall((:qob_ptrace_op, libqob200), Cint,
                (Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{Int32}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}),
                context(), length(dl), dl, dr, length(tr), tr, pointer(a.data), pointer(res), CUDA.stream().handle))
    Operator(bl, br, res)
end
function _ptrace_state(psi, indices, bra::Bool)
    QuantumOpticsBase.check_ptrace_arguments(psi, indices)
    d = shape(psi.basis)
    tr = Int32[i for i in indices]
    b = ptrace(psi.basis, indices)
    res = CUDA.zeros(ComplexF64, length(b), length(b))
    check(ccall((:qob_ptrace_state, libqob200), Cint,
                (Ptr{Cvoid}, Int32, Ptr{Int64}, Int32, Ptr{Int32}, Int32, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}),
                context(), length(d), d, length(tr), tr, Int32(bra), pointer(psi.data), pointer(res), CUDA.stream().handle))
    Operator(b, b, res)
end
ptrace(psi::Ket{B,<:DevVec}, indices) where B = _ptrace_state(psi, indices, false)
ptrace(psi::Bra{B,<:DevVec}, indices) where B = _ptrace_state(psi, indices, true)

# ---- fused master-equation right-hand side (include/qob200.h: qob_lindblad_*; no single reference function — it replaces the
# mul! sequence of test/test_sciml_broadcast_interfaces.jl:36-43 / QuantumOptics.jl's dmaster_h!)
struct LindbladRHS
    handle::Handle
    keep::Vector{Any}
end
function LindbladRHS(H::HostSparseOrDense{B,B}, J::Vector; rates=nothing) where B
    keep = Any[]
    fH, kH = factor(H.data); push!(keep, kH)
    fJ = QobFactor[]
    for j in J
        f, k = factor(j.data)
        push!(fJ, f); push!(keep, k)
    end
    r = rates === nothing ? C_NULL : convert(Vector{Float64}, rates)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep fJ r begin
        check(ccall((:qob_lindblad_create, libqob200), Cint,
                    (Ptr{Cvoid}, Ref{QobFactor}, Int32, Ptr{QobFactor}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                    context(), Ref(fH), Int32(length(fJ)), fJ, r, h))
    end
    LindbladRHS(Handle(h[]), keep)
end
"drho = alpha * L(rho) + beta * drho on device-resident dense operators"
function apply!(drho::DevOp{B,B}, L::LindbladRHS, rho::DevOp{B,B}, alpha=true, beta=false) where B
    GC.@preserve L check(ccall((:qob_lindblad_apply, libqob200), Cint, (Ptr{Cvoid}, C64, CuPtr{Cvoid}, C64, CuPtr{Cvoid}, Ptr{Cvoid}),
                               L.handle.ptr, C64(alpha), pointer(rho.data), C64(beta), pointer(drho.data), CUDA.stream().handle))
    drho
end

# ---- sharded apply across processes / GPUs (include/qob200.h: qob_dist_*).  The state of a LazySum over 2-level sites is cut
# into `world` slabs on its most significant index bits, one process per GPU.  The library plans the exchange, maps the peers'
# slabs through CUDA IPC and runs the fused exchange; the host only has to move 192 bytes per rank once.  `allgather` is any
# function Vector{UInt8} -> Vector{Vector{UInt8}} ordered by rank (MPI.Allgather, a socket, a shared file) and `barrier` any
# zero-argument function that returns once all ranks called it.
mutable struct ShardedLazySum
    dist::Ptr{Cvoid}
    op::LazySum
    rank::Int
    world::Int
    nbits_local::Int
    x::CuVector{ComplexF64}       # this rank's slab of the state (library memory, visible to the peers)
    y::Union{Nothing,CuVector{ComplexF64}}   # direct mode: this rank's slab of the result (also visible to the peers)
    own::Vector{Ptr{Cvoid}}
    peers::Vector{Ptr{Cvoid}}
end
function _dist_alloc(bytes)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qob_dist_alloc, libqob200), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), context(), bytes, p))
    p[]
end
# direct = true (default when the plan allows it): the exchange adds into the owners' result slabs `sh.y`; a rank holds 2 slabs
# instead of 3 (x, contributions, y) and `apply!(sh)` writes `sh.y`.  direct = false: any CuVector can receive the result.
function ShardedLazySum(op::LazySum, rank::Integer, world::Integer; allgather, barrier, direct::Bool=true)
    h = handle(op)
    d = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:qob_dist_create, libqob200), Cint, (Ptr{Cvoid}, Int32, Int32, Ref{Ptr{Cvoid}}), h.ptr, rank, world, d))
    nloc, nrem, nch = Ref{Int32}(0), Ref{Int32}(0), Ref{Int32}(0)
    slab, flagb = Ref{Int64}(0), Ref{Int64}(0)
    check(ccall((:qob_dist_info, libqob200), Cint, (Ptr{Cvoid}, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ref{Int64}, Ref{Int64}),
                d[], nloc, nrem, nch, slab, flagb))
    own, peers = Ptr{Cvoid}[], Ptr{Cvoid}[]
    px = _dist_alloc(slab[]); push!(own, px)
    x = unsafe_wrap(CuArray, CuPtr{ComplexF64}(UInt(px)), slab[] >> 4)
    y = nothing
    cap = Ref{Int32}(0)
    check(ccall((:qob_dist_direct_capable, libqob200), Cint, (Ptr{Cvoid}, Ref{Int32}), d[], cap))
    direct = direct && cap[] != 0 && nrem[] > 0 && world > 1
    tx, tz, tf = fill(px, world), nothing, nothing
    if nrem[] > 0 && world > 1
        # second slab: the result (direct mode) or the contribution buffer
        pz, pf = _dist_alloc(slab[]), _dist_alloc(flagb[]); push!(own, pz, pf)
        direct && (y = unsafe_wrap(CuArray, CuPtr{ComplexF64}(UInt(pz)), slab[] >> 4))
        fill!(unsafe_wrap(CuArray, CuPtr{UInt8}(UInt(pf)), flagb[]), 0x00); CUDA.synchronize()
        mine = Vector{UInt8}(undef, 192)
        for (k, p) in enumerate((px, pz, pf))
            GC.@preserve mine check(ccall((:qob_ipc_export, libqob200), Cint, (Ptr{Cvoid}, Ptr{UInt8}), p, pointer(mine, 64k - 63)))
        end
        all = allgather(mine)
        tabs = (Ptr{Cvoid}[], Ptr{Cvoid}[], Ptr{Cvoid}[])
        for q in 0:world-1, k in 1:3
            if q == rank
                push!(tabs[k], (px, pz, pf)[k])
            else
                pp = Ref{Ptr{Cvoid}}(C_NULL)
                hb = all[q+1][64k-63:64k]
                check(ccall((:qob_ipc_open, libqob200), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ref{Ptr{Cvoid}}), context(), hb, pp))
                push!(peers, pp[]); push!(tabs[k], pp[])
            end
        end
        tx, tz, tf = tabs
        barrier()   # every pad is zeroed and every mapping exists before the first apply
    end
    check(ccall((:qob_dist_bind, libqob200), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}),
                d[], tx, (direct || tz === nothing) ? C_NULL : tz, tf === nothing ? C_NULL : tf))
    direct && check(ccall((:qob_dist_bind_result, libqob200), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), d[], tz))
    sh = ShardedLazySum(d[], op, rank, world, nloc[], x, y, own, peers)
    finalizer(close!, sh)
end
"y_local = alpha * (op * x)_local + beta * y_local with x = `sh.x` on every rank; collective over the ranks"
apply!(sh::ShardedLazySum, alpha=true, beta=false) = apply!(sh.y, sh, alpha, beta)   # direct mode: the result slab `sh.y`
function apply!(y::DevVec, sh::ShardedLazySum, alpha=true, beta=false)
    handle(sh.op)   # re-sends mutated coefficients (time-dependent sums)
    GC.@preserve sh check(ccall((:qob_dist_apply, libqob200), Cint, (Ptr{Cvoid}, C64, C64, CuPtr{Cvoid}, Ptr{Cvoid}),
                                sh.dist, C64(alpha), C64(beta), pointer(y), CUDA.stream().handle))
    y
end
function close!(sh::ShardedLazySum)
    sh.dist == C_NULL && return
    CUDA.synchronize()
    ccall((:qob_dist_destroy, libqob200), Cint, (Ptr{Cvoid},), sh.dist); sh.dist = C_NULL
    foreach(p -> ccall((:qob_ipc_close, libqob200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), context(), p), sh.peers)
    foreach(p -> ccall((:qob_dist_free, libqob200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), context(), p), sh.own)
    empty!(sh.peers); empty!(sh.own)
    nothing
end

end # module
