# LeastSquaresOptimB200.jl — Julia-side glue that puts liblsob200.so behind LeastSquaresOptim.jl's own plugin
# surface.  NOT executable in this repository's image (no Julia here or on the GPU box): it is a SKETCH of the binding a
# maintainer adds, kept next to the C header it binds (include/lsob200.h) and never run — treat it as documentation of
# which ABI entry point each reference call site binds, not as tested code.  The Python package
# `leastsquaresoptim.jl_b200/` mirrors exactly these types and methods and is what the tests drive.
#
# Nothing in LeastSquaresOptim.jl changes: the package dispatches on the solver type
# (`AbstractAllocatedSolver(nls, optimizer)` at src/types.jl:156, `ldiv!` at levenberg_marquardt.jl:87 and
# dogleg.jl:115), so new `AbstractSolver` subtypes + methods are enough.
module LeastSquaresOptimB200

using LinearAlgebra, SparseArrays
import LeastSquaresOptim
import LeastSquaresOptim: AbstractSolver, AbstractAllocatedSolver, LeastSquaresProblem, Dogleg, LevenbergMarquardt,
                          colsumabs2!, wdot

const LIB = get(ENV, "LSOB200_LIB", "liblsob200.so")

# ---- status -> the exception the reference would have thrown --------------------------------------------------
function check(st::Cint, ctx = C_NULL)
    st == 0 && return
    msg = unsafe_string(ccall((:lso_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
    st > 0 && occursin("RankDeficient", msg) && throw(LinearAlgebra.RankDeficientException(Int(st)))
    st > 0 && throw(LinearAlgebra.PosDefException(Int(st)))          # dense_cholesky.jl:57
    st == -1 && throw(DimensionMismatch(msg))                         # dense_qr.jl:10,61; dense_cholesky.jl:10,50
    st == -6 && throw(LeastSquaresOptim.IsFiniteException(Int[]))     # utils.jl:63-78
    error("lsob200 error $st: $msg")
end

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:lso_ctx_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r))
        c = new(r[])
        finalizer(c -> ccall((:lso_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), c)
    end
end
const CTX = Ref{Context}()
ctx() = (isassigned(CTX) || (CTX[] = Context()); CTX[])

# ---- solver markers (src/types.jl:79-86) ------------------------------------------------------------------------
struct B200QR <: AbstractSolver end
struct B200Cholesky <: AbstractSolver end
struct B200LSMR <: AbstractSolver end

# ---- dense workspaces: DenseQRAllocatedSolver / DenseCholeskyAllocatedSolver -------------------------------------
mutable struct B200DenseWorkspace <: AbstractAllocatedSolver
    h::Ptr{Cvoid}
    kind::Cint
end
function B200DenseWorkspace(m, n, kind, damped)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lso_dense_ws_create, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Cint, Cint, Ref{Ptr{Cvoid}}),
                ctx().h, m, n, kind, damped, r), ctx().h)
    w = B200DenseWorkspace(r[], kind)
    finalizer(w -> ccall((:lso_dense_ws_destroy, LIB), Cint, (Ptr{Cvoid},), w.h), w)
end
# cf. src/solver/dense_qr.jl:25-28 (Dogleg) and :50-54 (LevenbergMarquardt)
AbstractAllocatedSolver(nls::LeastSquaresProblem, ::Dogleg{B200QR}) =
    B200DenseWorkspace(length(nls.y), length(nls.x), 1, 0)
AbstractAllocatedSolver(nls::LeastSquaresProblem, ::LevenbergMarquardt{B200QR}) =
    B200DenseWorkspace(length(nls.y), length(nls.x), 1, 1)
# cf. src/solver/dense_cholesky.jl:19-21
AbstractAllocatedSolver(nls::LeastSquaresProblem, ::Dogleg{B200Cholesky}) =
    B200DenseWorkspace(length(nls.y), length(nls.x), 2, 0)
AbstractAllocatedSolver(nls::LeastSquaresProblem, ::LevenbergMarquardt{B200Cholesky}) =
    B200DenseWorkspace(length(nls.y), length(nls.x), 2, 1)

# Julia Arrays are pageable: J, x, y live for the whole optimize! run (types.jl:141-157), so they are page-locked ONCE,
# when the allocated problem is built; every later H2D copy of J then runs at pinned speed (28 ms instead of 95 ms per
# step at 100 000 x 1 000).  Call `unpin!` when the problem is dropped.
pin!(a::Array{Float64}) = (check(ccall((:lso_host_register, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx().h, a, sizeof(a)), ctx().h); a)
unpin!(a::Array{Float64}) = (ccall((:lso_host_unregister, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx().h, a); a)

# host-Array form: `ldiv!` as called at dogleg.jl:115 and levenberg_marquardt.jl:87.  Returns (x, n_mul) like
# dense_qr.jl:41,87 / dense_cholesky.jl:34,58.
function _solve_host(A::B200DenseWorkspace, x, J::StridedMatrix{Float64}, y, damp)
    rank = Ref{Cint}(0)
    GC.@preserve x J y damp begin
        pd = damp === nothing ? Ptr{Float64}(C_NULL) : pointer(damp)
        st = A.kind == 1 ?
            ccall((:lso_qr_solve_host, LIB), Cint,
                  (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Cint}),
                  A.h, J, stride(J, 2), y, pd, x, rank) :
            ccall((:lso_chol_solve_host, LIB), Cint,
                  (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                  A.h, J, stride(J, 2), y, pd, x)
        check(st, ctx().h)
    end
    return x, 1
end
LinearAlgebra.ldiv!(x::Vector{Float64}, J::StridedMatrix{Float64}, y::Vector{Float64}, A::B200DenseWorkspace) =
    _solve_host(A, x, J, y, nothing)
LinearAlgebra.ldiv!(x::Vector{Float64}, J::StridedMatrix{Float64}, y::Vector{Float64}, damp::Vector{Float64},
                    A::B200DenseWorkspace) = _solve_host(A, x, J, y, damp)

# ---- resident mode: device vector / matrix duck types (README.md:37-43, src/utils/lsmr.jl:24-44) ------------------
mutable struct B200Vector <: AbstractVector{Float64}
    p::Ptr{Float64}
    n::Int
    function B200Vector(n::Integer)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:lso_dev_alloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx().h, 8n, r), ctx().h)
        v = new(Ptr{Float64}(r[]), n)
        finalizer(v -> ccall((:lso_dev_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx().h, v.p), v)
    end
end
Base.size(v::B200Vector) = (v.n,)
Base.similar(v::B200Vector) = B200Vector(v.n)
# Scalar indexing: the reference's loops touch single elements only in the box projection (LM:89-98, dogleg:148-157) and
# in the bounds check `all(x .>= lower)` (LM:51, dogleg:52) — both are overridden below by ONE kernel each, so these two
# methods exist for completeness (show, tests, user code) and cost a PCIe round trip per element.
function Base.getindex(v::B200Vector, i::Int)
    @boundscheck checkbounds(v, i)
    r = Ref{Float64}(0)
    check(ccall((:lso_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx().h, r, v.p + 8 * (i - 1), 8), ctx().h)
    r[]
end
function Base.setindex!(v::B200Vector, a, i::Int)
    @boundscheck checkbounds(v, i)
    r = Ref{Float64}(a)
    check(ccall((:lso_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx().h, v.p + 8 * (i - 1), r, 8), ctx().h)
    a
end
Base.IndexStyle(::Type{B200Vector}) = IndexLinear()
B200Vector(h::Vector{Float64}) = copyto!(B200Vector(length(h)), h)
Base.copyto!(d::B200Vector, h::Vector{Float64}) =
    (GC.@preserve h check(ccall((:lso_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx().h, d.p, h, 8d.n)); d)
Base.copyto!(h::Vector{Float64}, d::B200Vector) =
    (GC.@preserve h check(ccall((:lso_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx().h, h, d.p, 8d.n)); h)
Base.copyto!(d::B200Vector, s::B200Vector) =
    (check(ccall((:lso_vec_copy, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}), ctx().h, d.n, d.p, s.p)); d)
Base.fill!(v::B200Vector, a) = (check(ccall((:lso_vec_fill, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Float64), ctx().h, v.n, v.p, a)); v)
LinearAlgebra.rmul!(v::B200Vector, a::Number) = (check(ccall((:lso_vec_scal, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Float64), ctx().h, v.n, v.p, a)); v)
LinearAlgebra.axpy!(a::Number, x::B200Vector, y::B200Vector) =
    (check(ccall((:lso_vec_axpy, LIB), Cint, (Ptr{Cvoid}, Int64, Float64, Ptr{Float64}, Ptr{Float64}), ctx().h, y.n, a, x.p, y.p)); y)
function _scalar(f::Symbol, v::B200Vector, extra...)
    r = Ref{Float64}(0)
    check(ccall((f, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ref{Float64}), ctx().h, v.n, v.p, r)); r[]
end
LinearAlgebra.norm(v::B200Vector) = _scalar(:lso_vec_nrm2, v)
Base.sum(::typeof(abs2), v::B200Vector) = _scalar(:lso_vec_sumabs2, v)
Base.sum(v::B200Vector) = _scalar(:lso_vec_sum, v)
Base.maximum(::typeof(abs), v::B200Vector) = _scalar(:lso_vec_maxabs, v)
Base.clamp!(v::B200Vector, lo, hi) = (check(ccall((:lso_vec_clamp, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Float64, Float64), ctx().h, v.n, v.p, lo, hi)); v)
Base.map!(::typeof(/), o::B200Vector, x::B200Vector, y::B200Vector) =
    (check(ccall((:lso_vec_div, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), ctx().h, o.n, o.p, x.p, y.p)); o)
Base.map!(::typeof(*), o::B200Vector, x::B200Vector, y::B200Vector) =
    (check(ccall((:lso_vec_mul, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), ctx().h, o.n, o.p, x.p, y.p)); o)
Base.map!(::typeof(sqrt), o::B200Vector, x::B200Vector) = (o === x || copyto!(o, x);
    check(ccall((:lso_vec_sqrt, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), ctx().h, o.n, o.p)); o)
function wdot(x::B200Vector, y::B200Vector, w::B200Vector)          # src/utils/utils.jl:165-175
    r = Ref{Float64}(0)
    check(ccall((:lso_vec_wdot, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}),
                ctx().h, x.n, x.p, y.p, w.p, r)); r[]
end
# maxabs_projected_gradient(g, x, lower, upper) (src/utils/utils.jl:38-55) and the box projection of the step
# (levenberg_marquardt.jl:89-98, dogleg.jl:148-157: δx[i] = min(δx[i], x[i] - lower[i]) / max(…, x[i] - upper[i])), one
# kernel each instead of an element loop.  `lower` / `upper` are B200Vectors (or empty Vectors = no bound).
_bptr(b) = (b isa B200Vector && length(b) > 0) ? b.p : Ptr{Float64}(C_NULL)
function LeastSquaresOptim.maxabs_projected_gradient(g::B200Vector, x::B200Vector, lower, upper)
    r = Ref{Float64}(0)
    check(ccall((:lso_vec_maxabs_projected, LIB), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}),
                ctx().h, g.n, g.p, x.p, _bptr(lower), _bptr(upper), r), ctx().h); r[]
end
box_project!(δx::B200Vector, x::B200Vector, lower, upper) =
    (check(ccall((:lso_vec_box_project, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                 ctx().h, δx.n, δx.p, x.p, _bptr(lower), _bptr(upper)), ctx().h); δx)
# the four reductions of an iteration with ONE synchronisation (LM:104,110,114-117; utils.jl:21), see lso_lm_step_tail
function step_tail(J, δx::B200Vector, fcur::B200Vector, ftrial::B200Vector, fpredict::B200Vector)
    out = zeros(4)
    dJ, ld, csc = J isa B200Matrix ? (J.p, J.m, C_NULL) : (Ptr{Float64}(C_NULL), 0, J.h)
    check(ccall((:lso_lm_step_tail, LIB), Cint,
                (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}),
                ctx().h, size(J, 1), size(J, 2), dJ, ld, csc, δx.p, fcur.p, ftrial.p, fpredict.p, 0, out), ctx().h)
    (trial_ssr = out[1], predicted_ssr = out[2], maxabs_dx = out[3], maxabs_gr = out[4])
end
LeastSquaresOptim.check_isfinite(v::B200Vector) = begin                # src/utils/utils.jl:70-78
    bad = Ref{Int64}(-1)
    st = ccall((:lso_vec_check_finite, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ref{Int64}), ctx().h, v.n, v.p, bad)
    st == -6 && throw(LeastSquaresOptim.IsFiniteException([Int(bad[]) + 1]))
    check(st)
end

mutable struct B200Matrix <: AbstractMatrix{Float64}      # device image of a Matrix{Float64}, ld = m
    p::Ptr{Float64}
    m::Int
    n::Int
end
Base.size(A::B200Matrix) = (A.m, A.n)
function colsumabs2!(v::B200Vector, A::B200Matrix)                      # src/utils/utils.jl:139-144
    check(ccall((:lso_dense_colsumabs2, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}),
                ctx().h, A.m, A.n, A.p, A.m, v.p)); v
end
LinearAlgebra.mul!(y::B200Vector, A::B200Matrix, x::B200Vector, α::Number, β::Number) =   # LM:114, dogleg:109,171
    (check(ccall((:lso_dense_gemv_n, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Float64, Ptr{Float64}, Int64, Ptr{Float64}, Float64, Ptr{Float64}),
                 ctx().h, A.m, A.n, α, A.p, A.m, x.p, β, y.p)); y)
LinearAlgebra.mul!(x::B200Vector, At::Adjoint{Float64,B200Matrix}, y::B200Vector, α::Number, β::Number) =   # LM:102, dogleg:99
    (A = parent(At); check(ccall((:lso_dense_gemv_t, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Float64, Ptr{Float64}, Int64, Ptr{Float64}, Float64, Ptr{Float64}),
                 ctx().h, A.m, A.n, α, A.p, A.m, y.p, β, x.p)); x)
function LinearAlgebra.ldiv!(x::B200Vector, J::B200Matrix, y::B200Vector, damp::B200Vector, A::B200DenseWorkspace)
    rank = Ref{Cint}(0)
    st = A.kind == 1 ?
        ccall((:lso_qr_solve, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Cint}),
              A.h, J.p, J.m, y.p, damp.p, x.p, rank) :
        ccall((:lso_chol_solve, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
              A.h, J.p, J.m, y.p, damp.p, x.p)
    check(st, ctx().h); return x, 1
end
function LinearAlgebra.ldiv!(x::B200Vector, J::B200Matrix, y::B200Vector, A::B200DenseWorkspace)
    rank = Ref{Cint}(0)
    st = A.kind == 1 ?
        ccall((:lso_qr_solve, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Cint}),
              A.h, J.p, J.m, y.p, C_NULL, x.p, rank) :
        ccall((:lso_chol_solve, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
              A.h, J.p, J.m, y.p, C_NULL, x.p)
    check(st, ctx().h); return x, 1
end

# ---- sparse operator + LSMR (src/solver/iterative_lsmr.jl:161-259) ------------------------------------------------
mutable struct B200SparseMatrixCSC <: AbstractMatrix{Float64}   # device image of a SparseMatrixCSC{Float64,Int64}
    h::Ptr{Cvoid}
    host::SparseMatrixCSC{Float64,Int64}
end
Base.size(A::B200SparseMatrixCSC) = size(A.host)
Base.getindex(A::B200SparseMatrixCSC, i::Int, j::Int) = A.host[i, j]     # host mirror (display only; values as of the last refresh!)
function B200SparseMatrixCSC(J::SparseMatrixCSC{Float64,Int64})
    r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve J check(ccall((:lso_csc_create, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ref{Ptr{Cvoid}}),
                               ctx().h, size(J, 1), size(J, 2), nnz(J), J.colptr, J.rowval, r), ctx().h)
    A = B200SparseMatrixCSC(r[], J)
    finalizer(A -> ccall((:lso_csc_destroy, LIB), Cint, (Ptr{Cvoid},), A.h), A)
end
# after g!(J, x) rewrote nonzeros(J) (test/nonlinearleastsquares.jl:47-86): values-only refresh
refresh!(A::B200SparseMatrixCSC) = check(ccall((:lso_csc_set_values_host, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), A.h, nonzeros(A.host)))
# g! stored a DIFFERENT pattern (setindex! into a sparse J, test/nonlinearsolvers.jl:526-530): re-import colptr / rowval
function refresh_pattern!(A::B200SparseMatrixCSC)
    J = A.host
    GC.@preserve J check(ccall((:lso_csc_update_pattern, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}),
                               A.h, nnz(J), J.colptr, J.rowval), ctx().h)
    refresh!(A)
end
colsumabs2!(v::B200Vector, A::B200SparseMatrixCSC) = (check(ccall((:lso_csc_colsumabs2, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), A.h, v.p)); v)
LinearAlgebra.mul!(y::B200Vector, A::B200SparseMatrixCSC, x::B200Vector, α::Number, β::Number) =      # LM:114, dogleg:109,171
    (check(ccall((:lso_csc_mul_n, LIB), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}, Float64, Ptr{Float64}), A.h, α, x.p, β, y.p)); y)
LinearAlgebra.mul!(x::B200Vector, At::Adjoint{Float64,B200SparseMatrixCSC}, y::B200Vector, α::Number, β::Number) =   # LM:102, dogleg:99
    (A = parent(At); check(ccall((:lso_csc_mul_t, LIB), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}, Float64, Ptr{Float64}), A.h, α, y.p, β, x.p)); x)

mutable struct B200LSMRWorkspace <: AbstractAllocatedSolver
    h::Ptr{Cvoid}
    damped::Bool
end
function B200LSMRWorkspace(m, n, damped)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:lso_lsmr_ws_create, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Cint, Ref{Ptr{Cvoid}}), ctx().h, m, n, damped, r), ctx().h)
    w = B200LSMRWorkspace(r[], damped)
    finalizer(w -> ccall((:lso_lsmr_ws_destroy, LIB), Cint, (Ptr{Cvoid},), w.h), w)
end
AbstractAllocatedSolver(nls::LeastSquaresProblem, ::Dogleg{B200LSMR}) = B200LSMRWorkspace(length(nls.y), length(nls.x), false)   # :173-177
AbstractAllocatedSolver(nls::LeastSquaresProblem, ::LevenbergMarquardt{B200LSMR}) = B200LSMRWorkspace(length(nls.y), length(nls.x), true)  # :233-236
function _lsmr(A::B200LSMRWorkspace, x::B200Vector, J, y::B200Vector, damp, btol)
    iters, istop = Ref{Int64}(0), Ref{Cint}(0)
    csc = J isa B200SparseMatrixCSC ? J.h : C_NULL
    dJ, ld = J isa B200Matrix ? (J.p, J.m) : (Ptr{Float64}(C_NULL), 0)
    check(ccall((:lso_lsmr_solve, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                 Float64, Float64, Float64, Int64, Ref{Int64}, Ref{Cint}),
                A.h, csc, dJ, ld, y.p, damp === nothing ? C_NULL : damp.p, x.p, 1e-6, btol, 1e8, 0, iters, istop), ctx().h)
    return x, 2 * Int(iters[])                                      # ch.mvps (src/utils/lsmr.jl:236)
end
LinearAlgebra.ldiv!(x::B200Vector, J, y::B200Vector, A::B200LSMRWorkspace) = _lsmr(A, x, J, y, nothing, 1e-6)            # :179-198
LinearAlgebra.ldiv!(x::B200Vector, J, y::B200Vector, damp::B200Vector, A::B200LSMRWorkspace) = _lsmr(A, x, J, y, damp, 0.5)  # :238-259

export B200QR, B200Cholesky, B200LSMR, B200Vector, B200Matrix, B200SparseMatrixCSC, refresh!, refresh_pattern!,
       box_project!, step_tail, pin!, unpin!
end # module
