# HierarchicalMatricesB200.jl -- the reference-side binding of libhmb200.so.
#
# `using HierarchicalMatrices, HierarchicalMatricesB200` adds more specific methods of
# the reference's `mul!` / `*` for Float64 `KernelMatrix` and `HierarchicalMatrix`
# operators (reference: src/KernelMatrix.jl:5-45, src/HierarchicalMatrix.jl:5-52) and
# forwards them through `ccall` to the C ABI of include/hmb200.h.  The Julia side only
# walks `assigned` and pushes leaves (the planner front end); packing, partitioning and
# all arithmetic happen in the library.
#
# NOTE: Julia is not installed in the build image, so this file has not been executed
# there; the identical ABI calls are exercised from Python (api.py, tests/).
module HierarchicalMatricesB200

using LinearAlgebra
using Libdl
using HierarchicalMatrices
import HierarchicalMatrices: KernelMatrix, HierarchicalMatrix, LowRankMatrix, BarycentricMatrix2D,
                             EvenBarycentricMatrix, blocksize, Block

const libhm = get(ENV, "HMB200_LIB", joinpath(@__DIR__, "..", "lib", "libhmb200.so"))

struct HmError <: Exception
    status::Int32
    msg::String
end
Base.showerror(io::IO, e::HmError) = print(io, "libhmb200 status ", e.status, ": ", e.msg)

function check(st::Int32)
    st == 0 && return nothing
    msg = unsafe_string(ccall((:hm_last_error, libhm), Cstring, ()))
    st == 9 && throw(BoundsError())          # HM_ERR_REFERENCE: what the reference throws
    st == 3 && throw(DimensionMismatch(msg)) # HM_ERR_SHAPE
    throw(HmError(st, msg))
end

# ---------------------------------------------------------------------------
# plan handle, freed by the GC
# ---------------------------------------------------------------------------
mutable struct Plan
    ptr::Ptr{Cvoid}
    function Plan(ptr::Ptr{Cvoid})
        p = new(ptr)
        finalizer(p) do q
            q.ptr == C_NULL || ccall((:hm_plan_destroy, libhm), Int32, (Ptr{Cvoid},), q.ptr)
            q.ptr = C_NULL
        end
        p
    end
end

device() = parse(Int32, get(ENV, "HMB200_DEVICE", "0"))

# ---------------------------------------------------------------------------
# planner front end: the reference's own walk (KernelMatrix.jl:24-41), run once
# ---------------------------------------------------------------------------
function push_leaves!(b::Ptr{Cvoid}, H, i0::Int, j0::Int)
    M, N = blocksize(H)
    p = 0
    for m = 1:M
        q = 0
        for n = 1:N
            Hmn = H.assigned[m, n]
            if Hmn == 1
                push_leaves!(b, getfield(H, 1)[m, n], i0 + p, j0 + q)
            elseif Hmn == 2
                push_leaf!(b, getfield(H, 2)[m, n], i0 + p, j0 + q)
            elseif Hmn == 3
                push_leaf!(b, getfield(H, 3)[m, n], i0 + p, j0 + q)
            end
            q += blocksize(H, 1, n, 2)
        end
        p += blocksize(H, m, N, 1)
    end
end

function push_leaf!(b, A::Matrix{Float64}, i0, j0)
    m, n = size(A)
    GC.@preserve A check(ccall((:hm_builder_add_dense, libhm), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64, Int64, Int64),
        b, A, m, n, max(stride(A, 2), 1), i0, j0))
end

function push_leaf!(b, L::LowRankMatrix{Float64}, i0, j0)
    U, S, V = L.U, L.Σ.diag, L.V
    m, n, r = size(U, 1), size(V, 1), length(S)
    GC.@preserve U S V check(ccall((:hm_builder_add_lowrank, libhm), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Int64, Int64, Int64),
        b, U, max(stride(U, 2), 1), S, V, max(stride(V, 2), 1), m, n, r, i0, j0))
end

function push_leaf!(b, B::BarycentricMatrix2D{Float64}, i0, j0)
    U, F, V = B.U, B.B.F, B.V
    m, n, r = size(U, 1), size(V, 1), size(F, 1)
    GC.@preserve U F V check(ccall((:hm_builder_add_bary2d, libhm), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Int64, Int64, Int64, Int64),
        b, U, max(stride(U, 2), 1), F, max(stride(F, 2), 1), V, max(stride(V, 2), 1), m, n, r, i0, j0))
end

"""
    plan(H) -> Plan

Flatten `H` and pack it on the GPU.  The plan is a device-resident snapshot of `H`, cached per
operator and rebuilt when `H` may have changed:

* the cache is a `WeakKeyDict` keyed on `H.assigned` (the hierarchical types are immutable structs,
  their `assigned` matrix is the mutable object that lives exactly as long as the node): when `H`
  is collected the entry disappears, the `Plan` finalizer runs and the device memory is freed;
* `H[Block(m), Block(n)] = A`, `scale!` and `add_col!` on any Float64 hierarchical operator bump a
  global mutation epoch (more specific methods below that forward to the reference's own); a cached
  plan built in an older epoch is rebuilt on the next `mul!`.  The epoch is global because a nested
  node does not know its parents.  `rmul!` / `lmul!` with a `Diagonal` patch the plan in place
  instead.  Writing into a leaf array directly (`H.Matrixblocks[1,1][i,j] = v`) cannot be seen:
  call `invalidate!(H)` afterwards.
"""
mutable struct CachedPlan
    plan::Plan
    epoch::Int
end
const PLANS = WeakKeyDict{Matrix{Int},CachedPlan}()
const EPOCH = Ref(0)
const HOp = Union{KernelMatrix{Float64},HierarchicalMatrix{Float64}}

function build_plan(H::HOp)
    nr, nc = size(H)
    b = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:hm_builder_create, libhm), Int32, (Ref{Ptr{Cvoid}}, Int64, Int64, Int32, Int32),
                b, nr, nc, 0, device()))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    try
        push_leaves!(b[], H, 0, 0)
        dev = Int32[device()]
        check(ccall((:hm_plan_finalize, libhm), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}),
                    b[], dev, 1, out))
    finally
        ccall((:hm_builder_destroy, libhm), Int32, (Ptr{Cvoid},), b[])
    end
    Plan(out[])
end

function plan(H::HOp)
    c = get(PLANS, H.assigned, nothing)
    if c === nothing || c.epoch != EPOCH[]
        c = CachedPlan(build_plan(H), EPOCH[])
        PLANS[H.assigned] = c
    end
    c.plan
end

invalidate!(H) = (delete!(PLANS, H.assigned); H)

# Mutators of the reference (src/hierarchical.jl:149-172, src/HierarchicalMatrix.jl:54-139): the
# Float64 specialisations forward to the generic methods and mark every cached snapshot stale.
function Base.setindex!(H::HierarchicalMatrix{Float64}, A::AbstractMatrix{Float64}, B1::Block, B2::Block)
    EPOCH[] += 1
    invoke(setindex!, Tuple{HierarchicalMatrix{S},AbstractMatrix{S},Block,Block} where S, H, A, B1, B2)
end
function Base.setindex!(H::KernelMatrix{Float64}, A::AbstractMatrix{Float64}, B1::Block, B2::Block)
    EPOCH[] += 1
    invoke(setindex!, Tuple{KernelMatrix{S},AbstractMatrix{S},Block,Block} where S, H, A, B1, B2)
end
function HierarchicalMatrices.scale!(H::HierarchicalMatrix{Float64}, b::AbstractVector, jstart::Int)
    EPOCH[] += 1
    invoke(HierarchicalMatrices.scale!, Tuple{HierarchicalMatrix,AbstractVector,Int}, H, b, jstart)
end
function HierarchicalMatrices.scale!(b::AbstractVector, H::HierarchicalMatrix{Float64}, istart::Int)
    EPOCH[] += 1
    invoke(HierarchicalMatrices.scale!, Tuple{AbstractVector,HierarchicalMatrix,Int}, b, H, istart)
end
function HierarchicalMatrices.add_col!(H::HierarchicalMatrix{Float64}, u::Vector{Float64}, istart::Int, j::Int)
    EPOCH[] += 1
    invoke(HierarchicalMatrices.add_col!, Tuple{HierarchicalMatrix{S},Vector{S},Int,Int} where S, H, u, istart, j)
end

"""
    assemble(f, x, y, a, b, c, d) -> Plan

`KernelMatrix(f, x, y, a, b, c, d)` (KernelMatrix.jl:47) assembled on the GPU; `f` is one
of the kernels of examples/Kernel.jl (`:cauchy`, `:coulomb`, `:coulombprime`, `:log`).
"""
function assemble(f::Symbol, x::Vector{Float64}, y::Vector{Float64}, a, b, c, d; matrix_free::Bool = false,
                  part::Integer = 0, nparts::Integer = 1)
    id = Dict(:cauchy => 0, :coulomb => 1, :coulombprime => 2, :log => 3)[f]
    out = Ref{Ptr{Cvoid}}(C_NULL)
    # matrix_free: nothing but the r x r cores is stored, entries are evaluated inside every mul!
    entry = Libdl.dlsym(Libdl.dlopen(libhm), matrix_free ? :hm_assemble_kernel_free : :hm_assemble_kernel)
    GC.@preserve x y check(ccall(entry, Int32,
        (Ptr{Float64}, Int64, Ptr{Float64}, Int64, Float64, Float64, Float64, Float64, Int32, Int32, Int32, Int32,
         Ref{Ptr{Cvoid}}),
        x, length(x), y, length(y), a, b, c, d, id, device(), part, nparts, out))
    Plan(out[])
end

# Any kernel function (KernelMatrix.jl:47 takes any f::Function): U and V are filled on the device,
# the r x r cores and the dense leaves -- the only parts that depend on f -- are evaluated by `f`
# on the host through a batch callback (hm_assemble_kernel_fn).
function kernel_batch(px::Ptr{Float64}, py::Ptr{Float64}, n::Int64, pout::Ptr{Float64}, user::Ptr{Cvoid})
    f = (unsafe_pointer_to_objref(user)::Base.RefValue{Any})[]
    for i = 1:n
        unsafe_store!(pout, Float64(f(unsafe_load(px, i), unsafe_load(py, i))), i)
    end
    nothing
end

function assemble(f::Function, x::Vector{Float64}, y::Vector{Float64}, a, b, c, d;
                  part::Integer = 0, nparts::Integer = 1)
    fref = Ref{Any}(f)
    cb = @cfunction(kernel_batch, Cvoid, (Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Cvoid}))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve fref x y check(ccall((:hm_assemble_kernel_fn, libhm), Int32,
        (Ptr{Float64}, Int64, Ptr{Float64}, Int64, Float64, Float64, Float64, Float64, Ptr{Cvoid}, Ptr{Cvoid},
         Int32, Int32, Int32, Ref{Ptr{Cvoid}}),
        x, length(x), y, length(y), a, b, c, d, cb, pointer_from_objref(fref), device(), part, nparts, out))
    Plan(out[])
end

# ---------------------------------------------------------------------------
# mul!: y[istart + (i-1)INCY] += sum_j H[i,j] x[jstart + (j-1)INCX]
# (1-based offsets become pointer offsets, as in src/blas.jl:12)
# ---------------------------------------------------------------------------
function matvec!(y::StridedArray{Float64}, P::Plan, x::StridedArray{Float64}, istart::Int, jstart::Int,
                 INCX::Int, INCY::Int, accumulate::Bool)
    GC.@preserve x y check(ccall((:hm_matvec, libhm), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int32),
        P.ptr, pointer(x, jstart), INCX, pointer(y, istart), INCY, accumulate ? 1 : 0))
    y
end

function checkbounds_mul(y, H, x, istart, jstart, INCX, INCY)
    nr, nc = size(H)
    (nr == 0 || istart + (nr - 1) * INCY <= length(y)) || throw(BoundsError(y, istart + (nr - 1) * INCY))
    (nc == 0 || jstart + (nc - 1) * INCX <= length(x)) || throw(BoundsError(x, jstart + (nc - 1) * INCX))
end

# KernelMatrix: src/KernelMatrix.jl:14,17 (vectors only, unit stride)
function HierarchicalMatrices.mul!(u::Vector{Float64}, H::KernelMatrix{Float64}, v::StridedVector{Float64},
                                   istart::Int, jstart::Int)
    checkbounds_mul(u, H, v, istart, jstart, 1, 1)
    matvec!(u, plan(H), v, istart, jstart, 1, 1, true)
end
LinearAlgebra.mul!(u::Vector{Float64}, H::KernelMatrix{Float64}, v::StridedVector{Float64}) =
    HierarchicalMatrices.mul!(u, H, v, 1, 1)

# HierarchicalMatrix: src/HierarchicalMatrix.jl:14,19,24 (linear indexing, strides)
function HierarchicalMatrices.mul!(y::StridedVecOrMat{Float64}, H::HierarchicalMatrix{Float64},
                                   x::StridedVecOrMat{Float64}, istart::Int, jstart::Int, INCX::Int, INCY::Int)
    checkbounds_mul(y, H, x, istart, jstart, INCX, INCY)
    matvec!(y, plan(H), x, istart, jstart, INCX, INCY, true)
end
HierarchicalMatrices.mul!(y::StridedVecOrMat{Float64}, H::HierarchicalMatrix{Float64},
                          x::StridedVecOrMat{Float64}, istart::Int, jstart::Int) =
    HierarchicalMatrices.mul!(y, H, x, istart, jstart, 1, 1)
LinearAlgebra.mul!(y::StridedVector{Float64}, H::HierarchicalMatrix{Float64}, x::StridedVector{Float64}) =
    HierarchicalMatrices.mul!(y, H, x, 1, 1, 1, 1)

# EvenBarycentricMatrix: src/algebra.jl:166-239.  The active parity class depends on the absolute
# offsets, so one single-leaf plan is kept per parity of (istart-1)+(jstart-1).
# (EvenBarycentricMatrix is an immutable struct of arrays; the cache is keyed on its W factor.)
const EVEN_PLANS = WeakKeyDict{Matrix{Float64},Vector{Union{Nothing,Plan}}}()

function plan(B::EvenBarycentricMatrix{Float64}, parity::Int)
    slots = get!(() -> Union{Nothing,Plan}[nothing, nothing], EVEN_PLANS, B.W)
    slots[parity + 1] === nothing || return slots[parity + 1]
    m, n = size(B)
    W, F = B.W, B.F
    b = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:hm_builder_create, libhm), Int32, (Ref{Ptr{Cvoid}}, Int64, Int64, Int32, Int32),
                b, m, n, 0, device()))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    try
        GC.@preserve W F check(ccall((:hm_builder_add_evenbary, libhm), Int32,
            (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Int64, Int64, Int64, Int64, Int32),
            b[], W, max(stride(W, 2), 1), F, max(stride(F, 2), 1), m, n, size(W, 1), 0, 0, parity))
        dev = Int32[device()]
        check(ccall((:hm_plan_finalize, libhm), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}),
                    b[], dev, 1, out))
    finally
        ccall((:hm_builder_destroy, libhm), Int32, (Ptr{Cvoid},), b[])
    end
    slots[parity + 1] = Plan(out[])
end

function HierarchicalMatrices.mul!(u::Vector{Float64}, B::EvenBarycentricMatrix{Float64},
                                   v::StridedVector{Float64}, istart::Int, jstart::Int)
    checkbounds_mul(u, B, v, istart, jstart, 1, 1)
    matvec!(u, plan(B, (istart + jstart) & 1), v, istart, jstart, 1, 1, true)
end

# rmul!(H, Diagonal(b)) / lmul!(Diagonal(b), H): src/HierarchicalMatrix.jl:15-16, 54-108.
# The reference methods update the Julia blocks; afterwards the cached device plan (if any)
# is updated in place by the library's streaming kernels instead of being rebuilt.
function scale_plan!(H, b::Vector{Float64}, side::Integer, epoch_before::Int)
    c = get(PLANS, H.assigned, nothing)
    (c === nothing || c.epoch != epoch_before) && return H   # no snapshot, or already stale
    GC.@preserve b check(ccall((:hm_plan_scale, libhm), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32),
                               c.plan.ptr, b, 1, side))
    c.epoch = EPOCH[]                                        # this snapshot followed the update
    H
end
function LinearAlgebra.rmul!(H::HierarchicalMatrix{Float64}, b::Diagonal{Float64,Vector{Float64}})
    e = EPOCH[]
    HierarchicalMatrices.scale!(H, b.diag, 1)
    scale_plan!(H, b.diag, 0, e)
end
function LinearAlgebra.lmul!(b::Diagonal{Float64,Vector{Float64}}, H::HierarchicalMatrix{Float64})
    e = EPOCH[]
    HierarchicalMatrices.scale!(b.diag, H, 1)
    scale_plan!(H, b.diag, 1, e)
    H
end
scale!(P::Plan, b::Vector{Float64}, side::Integer) =
    (GC.@preserve b check(ccall((:hm_plan_scale, libhm), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32),
                                P.ptr, b, 1, side)); P)

# Adjoint apply (no hierarchical adjoint exists in the reference; leaf rules of
# src/algebra.jl:52-82, 138-159): y += H' x through the same packed operator.
function adjoint_mul!(y::StridedVector{Float64}, H::Union{KernelMatrix{Float64},HierarchicalMatrix{Float64}},
                      x::StridedVector{Float64}; accumulate::Bool = true)
    nr, nc = size(H)
    (length(x) >= nr && length(y) >= nc) || throw(DimensionMismatch("adjoint_mul!"))
    GC.@preserve x y check(ccall((:hm_matvec_adjoint, libhm), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int32),
        plan(H).ptr, x, stride(x, 1), y, stride(y, 1), accumulate ? 1 : 0))
    y
end

# Many right-hand sides: Y (+)= H X through hm_matmat (FP64 tensor-core panel kernels).
# Reference entry points: `*(H, x::AbstractMatrix)` src/HierarchicalMatrix.jl:9-12,
# src/KernelMatrix.jl:9-12.  NOTE: the reference's HierarchicalMatrix method forwards to the
# linear-index `mul!(y, H, x, 1, 1, 1, 1)`, which fills only the first column of the result (the
# multi-column form it tests goes through the stride pair, test/runtests.jl:23-25); `H * X` here
# returns the full product, the 7-argument `mul!` above keeps the reference's linear indexing.
function matmat!(Y::StridedMatrix{Float64}, P::Plan, X::StridedMatrix{Float64}; accumulate::Bool = true)
    st = stats(P)
    (size(X, 1) == st.ncols && size(Y, 1) == st.nrows && size(X, 2) == size(Y, 2)) ||
        throw(DimensionMismatch("matmat!: Y $(size(Y)), H ($(st.nrows), $(st.ncols)), X $(size(X))"))
    (stride(X, 1) == 1 && stride(Y, 1) == 1) || throw(ArgumentError("matmat!: columns must be contiguous"))
    GC.@preserve X Y check(ccall((:hm_matmat, libhm), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, Int32),
        P.ptr, X, max(stride(X, 2), 1), Y, max(stride(Y, 2), 1), size(X, 2), accumulate ? 1 : 0))
    Y
end
matmat!(Y::StridedMatrix{Float64}, H::HOp, X::StridedMatrix{Float64}; accumulate::Bool = true) =
    matmat!(Y, plan(H), X; accumulate = accumulate)
Base.:*(H::KernelMatrix{Float64}, X::StridedMatrix{Float64}) =
    matmat!(zeros(size(H, 1), size(X, 2)), plan(H), X; accumulate = false)
Base.:*(H::HierarchicalMatrix{Float64}, X::StridedMatrix{Float64}) =
    matmat!(zeros(size(H, 1), size(X, 2)), plan(H), X; accumulate = false)
Base.:*(P::Plan, X::StridedMatrix{Float64}) = matmat!(zeros(stats(P).nrows, size(X, 2)), P, X; accumulate = false)

# Plans built by `assemble` act as operators themselves
Base.:*(P::Plan, v::Vector{Float64}) = begin
    st = stats(P)
    matvec!(zeros(st.nrows), P, v, 1, 1, 1, 1, false)
end
function adjoint_mul!(y::StridedVector{Float64}, P::Plan, x::StridedVector{Float64}; accumulate::Bool = true)
    st = stats(P)
    (length(x) >= st.nrows && length(y) >= st.ncols) || throw(DimensionMismatch("adjoint_mul!"))
    GC.@preserve x y check(ccall((:hm_matvec_adjoint, libhm), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int32),
        P.ptr, x, stride(x, 1), y, stride(y, 1), accumulate ? 1 : 0))
    y
end

# ---------------------------------------------------------------------------
# Multi-GPU (include/hmb200.h, hm_dist_*): one Julia process per GPU (e.g. under MPI.jl), rank r
# assembles / finalises block-row part r of nranks and joins the exchange.
#   id = rank == 0 ? dist_unique_id() : nothing;  id = MPI.bcast(id, 0, comm)
#   P  = assemble(:cauchy, x, y, 1.0, -1.0, 1.0, -1.0; part = rank, nparts = nranks)
#   dist_init!(P, id, nranks, rank)
#   dist_mul!(u, P, v)        # v read on rank 0, the whole u on every rank
# ---------------------------------------------------------------------------
function dist_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:hm_dist_get_id, libhm), Int32, (Ptr{UInt8},), id))
    id
end
function dist_init!(P::Plan, id::Vector{UInt8}, nranks::Integer, rank::Integer)
    length(id) == 128 || throw(ArgumentError("id must be the 128 bytes of dist_unique_id()"))
    check(ccall((:hm_dist_init, libhm), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), P.ptr, id, nranks, rank))
    P
end
function dist_mul!(u::Union{Nothing,StridedVector{Float64}}, P::Plan, v::Union{Nothing,StridedVector{Float64}};
                   root::Integer = 0, accumulate::Bool = false)
    pu = u === nothing ? Ptr{Float64}(C_NULL) : pointer(u)
    pv = v === nothing ? Ptr{Float64}(C_NULL) : pointer(v)
    GC.@preserve u v check(ccall((:hm_dist_matvec, libhm), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int32, Int32),
        P.ptr, pv, v === nothing ? 1 : stride(v, 1), pu, u === nothing ? 1 : stride(u, 1), root, accumulate ? 1 : 0))
    u
end

struct Stats
    nrows::Int64; ncols::Int64; n_dense::Int64; n_lowrank::Int64; n_bary2d::Int64
    dense_words::Int64; lowrank_words::Int64; core_words::Int64; algorithmic_bytes::Int64
    row_begin::Int64; row_end::Int64; part_words::Int64; stored_bytes::Int64
    v_stream_bytes::Int64; u_stream_bytes::Int64; partial_bytes::Int64
    n_stage1_items::Int64; n_stage2_blocks::Int64; n_stage3_items::Int64; n_stage3_rounds::Int64
    part_algorithmic_bytes::Int64; part_v_words::Int64; part_core_words::Int64; part_u_words::Int64
    part_dense_words::Int64
end

function stats(P::Plan)
    s = Ref{Stats}()
    check(ccall((:hm_plan_stats, libhm), Int32, (Ptr{Cvoid}, Ref{Stats}), P.ptr, s))
    s[]
end

end # module
