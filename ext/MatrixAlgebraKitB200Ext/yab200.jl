# YAB200 — "yet another" thin binding layer, the B200 twin of YACUSOLVER
# (ext/MatrixAlgebraKitCUDAExt/yacusolver.jl of the reference).  Zero numerical logic: every
# function is argument checking + one `ccall` into libmakb200 (include/makb200.h).
#
# NOTE: this file cannot be executed in the build environment (no Julia in the image); the
# executable twin of this binding is matrixalgebrakit.jl_b200/_lib.py (ctypes, same symbols).
module YAB200

using CUDA
using CUDA: CuPtr, CuMatrix, CuVector, StridedCuMatrix, StridedCuVector
using LinearAlgebra: BlasFloat, checksquare, chkstride1

const libmakb200 = get(ENV, "MAKB200_LIB", "libmakb200")
const B200Float = Union{Float64, ComplexF64}

const MAKB200_F64, MAKB200_C128 = Cint(0), Cint(1)
const QR_COMPACT, QR_FULL = Cint(0), Cint(1)
dtypecode(::Type{Float64}) = MAKB200_F64
dtypecode(::Type{ComplexF64}) = MAKB200_C128

# ---- handle: one per task/device, like cuSOLVER.dense_handle() (yacusolver.jl:76) ------------
const HANDLES = Dict{Tuple{Int, UInt}, Ptr{Cvoid}}()
function handle()
    dev = CUDA.deviceid(CUDA.device())
    key = (dev, objectid(current_task()))
    h = get!(HANDLES, key) do
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:makb200_create, libmakb200), Cint, (Ref{Ptr{Cvoid}}, Cint), ref, dev)
        rc == 0 || error("makb200_create failed with code $rc (needs an sm_100 device)")
        ref[]
    end
    ccall((:makb200_set_stream, libmakb200), Cint, (Ptr{Cvoid}, CUDA.CUstream), h, CUDA.stream())
    return h
end

function chkargsok(rc::Cint, what)
    rc == 0 && return nothing
    rc < 0 && throw(ArgumentError("$what: invalid value in argument $(-rc)"))  # LAPACK info < 0
    msg = unsafe_string(ccall((:makb200_last_error, libmakb200), Cstring, (Ptr{Cvoid},), handle()))
    error("$what failed with code $rc: $msg")
end

# workspace: caller-provided, borrowed from CUDA.jl's pool for the duration of the call
with_workspace(f, nbytes) = (buf = CuVector{UInt8}(undef, max(nbytes, 1)); try f(buf) finally CUDA.unsafe_free!(buf) end)

# ---- QR: qr_householder!(::B200, A, Q, R) in one fused call ------------------------------------
function qr!(A::StridedCuMatrix{T}, Q::StridedCuMatrix{T}, R::StridedCuMatrix{T}; full::Bool, positive::Bool) where {T <: B200Float}
    chkstride1(A, Q)
    m, n = size(A)
    h = handle()
    mode = full ? QR_FULL : QR_COMPACT
    lw = ccall((:makb200_qr_worksize, libmakb200), Csize_t, (Ptr{Cvoid}, Cint, Cint, Cint, Cint), h, dtypecode(T), mode, m, n)
    computeR = length(R) > 0
    with_workspace(lw) do work
        rc = ccall((:makb200_qr, libmakb200), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, CuPtr{T}, Cint, CuPtr{T}, Cint, CuPtr{T}, Cint, CuPtr{UInt8}, Csize_t),
            h, dtypecode(T), mode, positive, m, n, A, max(1, stride(A, 2)), Q, max(1, stride(Q, 2)),
            computeR ? pointer(R) : CU_NULL, computeR ? max(1, stride(R, 2)) : 0, work, lw)
        chkargsok(rc, "makb200_qr")
    end
    return Q, R
end

# ---- L1 shims: geqrf! / ungqr! (same call shapes as YACUSOLVER, yacusolver.jl:12-14) ----------
function geqrf!(A::StridedCuMatrix{T}, tau::StridedCuVector{T} = similar(A, min(size(A)...))) where {T <: B200Float}
    m, n = size(A)
    h = handle()
    lw = ccall((:makb200_geqrf_worksize, libmakb200), Csize_t, (Ptr{Cvoid}, Cint, Cint, Cint), h, dtypecode(T), m, n)
    with_workspace(lw) do work
        rc = ccall((:makb200_geqrf, libmakb200), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{T}, Cint, CuPtr{T}, CuPtr{UInt8}, Csize_t),
            h, dtypecode(T), m, n, A, max(1, stride(A, 2)), tau, work, lw)
        chkargsok(rc, "makb200_geqrf")
    end
    return A, tau
end

function ungqr!(A::StridedCuMatrix{T}, tau::StridedCuVector{T}, Q::StridedCuMatrix{T}) where {T <: B200Float}
    m, k = size(A, 1), length(tau)
    ncols = size(Q, 2)
    h = handle()
    lw = ccall((:makb200_orgqr_worksize, libmakb200), Csize_t, (Ptr{Cvoid}, Cint, Cint, Cint, Cint), h, dtypecode(T), m, ncols, k)
    with_workspace(lw) do work
        rc = ccall((:makb200_orgqr, libmakb200), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Cint, CuPtr{T}, Cint, CuPtr{T}, CuPtr{T}, Cint, CuPtr{UInt8}, Csize_t),
            h, dtypecode(T), m, ncols, k, A, max(1, stride(A, 2)), tau, Q, max(1, stride(Q, 2)), work, lw)
        chkargsok(rc, "makb200_orgqr")
    end
    return Q
end

# ---- eigh: heevd!(A, W, V) (yacusolver.jl:766-810 call shape) -----------------------------------
function heevd!(A::StridedCuMatrix{T}, W::StridedCuVector{Float64}, V::StridedCuMatrix{T}; fixgauge::Bool = false) where {T <: B200Float}
    n = checksquare(A)
    vectors = length(V) > 0   # job 'N' when V is empty (yalapack.jl:1192-1195, 1293-1298): Sturm K-section, no vectors
    h = handle()
    lw = ccall((:makb200_eigh_worksize, libmakb200), Csize_t, (Ptr{Cvoid}, Cint, Cint), h, dtypecode(T), n)
    with_workspace(lw) do work
        rc = ccall((:makb200_eigh, libmakb200), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{T}, Cint, CuPtr{Float64}, CuPtr{T}, Cint, CuPtr{UInt8}, Csize_t, CuPtr{Cint}),
            h, dtypecode(T), fixgauge, n, A, max(1, stride(A, 2)), W, vectors ? pointer(V) : CU_NULL,
            vectors ? max(1, stride(V, 2)) : 0, work, lw, CU_NULL)
        chkargsok(rc, "makb200_eigh")
    end
    return W, V
end

# (||(A-A')/2||_F^2, max|A_ij|) in one pass: device half of check_hermitian
function hermitian_defect(A::StridedCuMatrix{T}) where {T <: B200Float}
    n = checksquare(A)
    out = CUDA.zeros(Float64, 2)
    rc = ccall((:makb200_hermitian_defect, libmakb200), Cint, (Ptr{Cvoid}, Cint, Cint, CuPtr{T}, Cint, CuPtr{Float64}),
        handle(), dtypecode(T), n, A, max(1, stride(A, 2)), out)
    chkargsok(rc, "makb200_hermitian_defect")
    d2, mx = Array(out)
    return sqrt(d2), mx
end

# ---- gemm!: C = α op(A) op(B) + β C on the DMMA kernel (the `mul!` of polar.jl:63,88) ----------
opcode(c::AbstractChar) = c == 'N' ? Cint(0) : c == 'T' ? Cint(1) : c == 'C' ? Cint(2) : throw(ArgumentError("op $c"))
function gemm!(opa::AbstractChar, opb::AbstractChar, α::T, A::StridedCuMatrix{T}, B::StridedCuMatrix{T}, β::T,
        C::StridedCuMatrix{T}) where {T <: B200Float}
    m, n = size(C)
    k = opa == 'N' ? size(A, 2) : size(A, 1)
    rc = ccall((:makb200_gemm, libmakb200), Cint,
        (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, Cint, Ref{T}, CuPtr{T}, Cint, CuPtr{T}, Cint, Ref{T}, CuPtr{T}, Cint),
        handle(), dtypecode(T), opcode(opa), opcode(opb), m, n, k, Ref(α), A, max(1, stride(A, 2)), B, max(1, stride(B, 2)),
        Ref(β), C, max(1, stride(C, 2)))
    chkargsok(rc, "makb200_gemm")
    return C
end

# B = (A ± Aᴴ)/2 in one launch; B may be A (in place)
function project_hermitian!(A::StridedCuMatrix{T}, B::StridedCuMatrix{T}, anti::Bool) where {T <: B200Float}
    n = checksquare(A)
    size(B) == (n, n) || throw(DimensionMismatch("B must be $n x $n"))
    rc = ccall((:makb200_project_hermitian, libmakb200), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{T}, Cint, CuPtr{T}, Cint),
        handle(), dtypecode(T), anti, n, A, max(1, stride(A, 2)), B, max(1, stride(B, 2)))
    chkargsok(rc, "makb200_project_hermitian")
    return B
end

# ‖A‖_F² in one launch (‖W‖_F² = n for an isometric polar factor, = rank for QDWH's partial isometry)
function fro2(A::StridedCuMatrix{T}) where {T <: B200Float}
    m, n = size(A)
    out = CUDA.zeros(Float64, 1)
    rc = ccall((:makb200_fro2, libmakb200), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{T}, Cint, CuPtr{Float64}),
        handle(), dtypecode(T), m, n, A, max(1, stride(A, 2)), out)
    chkargsok(rc, "makb200_fro2")
    return Array(out)[1]
end

# one! (mode 0), uppertriangular! (1), lowertriangular! (2) in one launch
function tri_init!(A::StridedCuMatrix{T}, mode::Integer) where {T <: B200Float}
    m, n = size(A)
    rc = ccall((:makb200_tri_init, libmakb200), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, CuPtr{T}, Cint),
        handle(), dtypecode(T), mode, m, n, A, max(1, stride(A, 2)))
    chkargsok(rc, "makb200_tri_init")
    return A
end

# (‖part that must vanish‖_F, max|A_ij|, ‖A‖_F, #exact mismatches): every ingredient of ishermitian / isantihermitian
function hermitian_props(A::StridedCuMatrix{T}, anti::Bool) where {T <: B200Float}
    n = checksquare(A)
    out = CUDA.zeros(Float64, 4)
    rc = ccall((:makb200_hermitian_props, libmakb200), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{T}, Cint, CuPtr{Float64}),
        handle(), dtypecode(T), anti, n, A, max(1, stride(A, 2)), out)
    chkargsok(rc, "makb200_hermitian_props")
    d2, mx, f2, bad = Array(out)
    return sqrt(d2), mx, sqrt(f2), Int(bad)
end

# (‖P‖_F, ‖P − I‖_F) of a Gram matrix
function gram_defect(P::StridedCuMatrix{T}) where {T <: B200Float}
    n = checksquare(P)
    out = CUDA.zeros(Float64, 2)
    rc = ccall((:makb200_gram_defect, libmakb200), Cint, (Ptr{Cvoid}, Cint, Cint, CuPtr{T}, Cint, CuPtr{Float64}),
        handle(), dtypecode(T), n, P, max(1, stride(P, 2)), out)
    chkargsok(rc, "makb200_gram_defect")
    p2, d2 = Array(out)
    return sqrt(p2), sqrt(d2)
end

# ---- svd: gesvdp!(A, S, U, Vᴴ) (QDWH + eigh; yacusolver.jl:101-181 call shape) ------------------
function gesvdp!(A::StridedCuMatrix{T}, S::StridedCuVector{Float64}, U::StridedCuMatrix{T}, Vᴴ::StridedCuMatrix{T};
        fixgauge::Bool = false, l0::Float64 = 0.0) where {T <: B200Float}
    m, n = size(A)
    vectors = length(U) > 0 && length(Vᴴ) > 0   # job 'N' when both are empty (yalapack.jl:2105-2106)
    h = handle()
    lw = ccall((:makb200_svd_worksize, libmakb200), Csize_t, (Ptr{Cvoid}, Cint, Cint, Cint), h, dtypecode(T), m, n)
    with_workspace(lw) do work
        rc = ccall((:makb200_svd, libmakb200), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Cint, CuPtr{T}, Cint, CuPtr{Float64}, CuPtr{T}, Cint, CuPtr{T}, Cint, Cdouble, CuPtr{UInt8}, Csize_t, CuPtr{Cint}),
            h, dtypecode(T), fixgauge, m, n, A, max(1, stride(A, 2)), S,
            vectors ? pointer(U) : CU_NULL, vectors ? max(1, stride(U, 2)) : 0,
            vectors ? pointer(Vᴴ) : CU_NULL, vectors ? max(1, stride(Vᴴ, 2)) : 0, l0, work, lw, CU_NULL)
        chkargsok(rc, "makb200_svd")
    end
    return S, U, Vᴴ
end

# leading-r SVD (svd_trunc! with truncrank(r)): S gets all k values, U is m×r, Vᴴ is r×n
function svd_leading!(A::StridedCuMatrix{T}, r::Int, S::StridedCuVector{Float64}, U::StridedCuMatrix{T}, Vᴴ::StridedCuMatrix{T};
        fixgauge::Bool = true, l0::Float64 = 0.0) where {T <: B200Float}
    m, n = size(A)
    size(U) == (m, r) && size(Vᴴ) == (r, n) && length(S) == min(m, n) || throw(DimensionMismatch("svd_leading!: U m×r, S min(m,n), Vᴴ r×n"))
    h = handle()
    lw = ccall((:makb200_svd_worksize, libmakb200), Csize_t, (Ptr{Cvoid}, Cint, Cint, Cint), h, dtypecode(T), m, n)
    with_workspace(lw) do work
        rc = ccall((:makb200_svd_leading, libmakb200), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, CuPtr{T}, Cint, CuPtr{Float64}, CuPtr{T}, Cint, CuPtr{T}, Cint, Cdouble, CuPtr{UInt8}, Csize_t, CuPtr{Cint}),
            h, dtypecode(T), fixgauge, m, n, r, A, max(1, stride(A, 2)), S, U, max(1, stride(U, 2)), Vᴴ, max(1, stride(Vᴴ, 2)),
            l0, work, lw, CU_NULL)
        chkargsok(rc, "makb200_svd_leading")
    end
    return U, S, Vᴴ
end

# rank and truncation error of every block of a batch from one launch (findtruncated_svd + truncation_error!,
# implementations/truncation.jl:54-102,168-174); spec mirrors `makb200_trunc_spec`
struct TruncSpec
    maxrank::Cint; minrank::Cint; by_value::Cint; by_error::Cint
    vatol::Cdouble; vrtol::Cdouble; vp::Cdouble
    eatol::Cdouble; ertol::Cdouble; ep::Cdouble
end
function trunc_select_batched(Ss::Vector{<:StridedCuVector{Float64}}, spec::TruncSpec; maxranks::Union{Nothing, Vector{<:Integer}} = nothing)
    b = length(Ss)
    ks = Cint[length(S) for S in Ss]
    caps = maxranks === nothing ? Ptr{Cint}(C_NULL) : Cint.(maxranks)   # per-block rank caps (config 3: n_i ÷ 2)
    ptrs = [reinterpret(Ptr{Cvoid}, pointer(S)) for S in Ss]
    rank, eps = CUDA.zeros(Cint, b), CUDA.zeros(Float64, b)
    h = handle()
    lw = ccall((:makb200_trunc_select_batched_worksize, libmakb200), Csize_t, (Ptr{Cvoid}, Cint), h, b)
    with_workspace(lw) do work
        rc = ccall((:makb200_trunc_select_batched, libmakb200), Cint,
            (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Ptr{Cvoid}}, Ref{TruncSpec}, Ptr{Cint}, CuPtr{Cint}, CuPtr{Float64}, CuPtr{UInt8}, Csize_t),
            h, b, ks, ptrs, Ref(spec), caps, rank, eps, work, lw)
        chkargsok(rc, "makb200_trunc_select_batched")
    end
    return Array(rank), Array(eps)
end

# ---- polar: QDWH ----------------------------------------------------------------------------------
function polar_qdwh!(A::StridedCuMatrix{T}, W::StridedCuMatrix{T}, P::StridedCuMatrix{T}; l0::Float64 = 0.0, maxiter::Int = 12) where {T <: B200Float}
    m, n = size(A)
    wantP = length(P) > 0
    h = handle()
    lw = ccall((:makb200_polar_worksize, libmakb200), Csize_t, (Ptr{Cvoid}, Cint, Cint, Cint), h, dtypecode(T), m, n)
    iters = Ref{Cint}(0)
    with_workspace(lw) do work
        rc = ccall((:makb200_polar_qdwh, libmakb200), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{T}, Cint, CuPtr{T}, Cint, CuPtr{T}, Cint, Cdouble, Cint, CuPtr{UInt8}, Csize_t, Ref{Cint}, CuPtr{Cint}),
            h, dtypecode(T), m, n, A, max(1, stride(A, 2)), W, max(1, stride(W, 2)),
            wantP ? pointer(P) : CU_NULL, wantP ? max(1, stride(P, 2)) : 0, l0, maxiter, work, lw, iters, CU_NULL)
        chkargsok(rc, "makb200_polar_qdwh")
    end
    return W, P
end

# ---- batched qr_compact! over a vector of blocks ---------------------------------------------------
function qr_batched!(As::Vector{<:StridedCuMatrix{T}}, Qs::Vector{<:StridedCuMatrix{T}}, Rs::Vector{<:StridedCuMatrix{T}}) where {T <: B200Float}
    b = length(As)
    m = Cint[size(A, 1) for A in As]; n = Cint[size(A, 2) for A in As]
    lda = Cint[max(1, stride(A, 2)) for A in As]; ldq = Cint[max(1, stride(Q, 2)) for Q in Qs]
    ldr = Cint[length(R) > 0 ? max(1, stride(R, 2)) : 0 for R in Rs]
    Ap = CuPtr{T}[pointer(A) for A in As]; Qp = CuPtr{T}[pointer(Q) for Q in Qs]
    Rp = CuPtr{T}[length(R) > 0 ? pointer(R) : CU_NULL for R in Rs]
    h = handle()
    lw = ccall((:makb200_qr_batched_worksize, libmakb200), Csize_t, (Ptr{Cvoid}, Cint, Cint, Ptr{Cint}, Ptr{Cint}), h, dtypecode(T), b, m, n)
    with_workspace(lw) do work
        rc = ccall((:makb200_qr_batched, libmakb200), Cint,
            (Ptr{Cvoid}, Cint, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{CuPtr{T}}, Ptr{Cint}, Ptr{CuPtr{T}}, Ptr{Cint}, Ptr{CuPtr{T}}, Ptr{Cint}, CuPtr{Cint}, CuPtr{UInt8}, Csize_t),
            h, dtypecode(T), b, m, n, Ap, lda, Qp, ldq, Rp, ldr, CU_NULL, work, lw)
        chkargsok(rc, "makb200_qr_batched")
    end
    return Qs, Rs
end

# ---- batched QR plan: block structure classified and uploaded once, run! only launches ------------
mutable struct QRBatchedPlan{T}
    ptr::Ptr{Cvoid}
    work::CuVector{UInt8}          # owned by the plan: descriptors live there between runs
    keep::Any                      # the block arrays (their pointers are baked into the plan)
end
function qr_batched_plan(As::Vector{<:StridedCuMatrix{T}}, Qs::Vector{<:StridedCuMatrix{T}}, Rs::Vector{<:StridedCuMatrix{T}}) where {T <: B200Float}
    b = length(As)
    m = Cint[size(A, 1) for A in As]; n = Cint[size(A, 2) for A in As]
    lda = Cint[max(1, stride(A, 2)) for A in As]; ldq = Cint[max(1, stride(Q, 2)) for Q in Qs]
    ldr = Cint[length(R) > 0 ? max(1, stride(R, 2)) : 0 for R in Rs]
    Ap = CuPtr{T}[pointer(A) for A in As]; Qp = CuPtr{T}[pointer(Q) for Q in Qs]
    Rp = CuPtr{T}[length(R) > 0 ? pointer(R) : CU_NULL for R in Rs]
    h = handle()
    lw = ccall((:makb200_qr_batched_worksize, libmakb200), Csize_t, (Ptr{Cvoid}, Cint, Cint, Ptr{Cint}, Ptr{Cint}), h, dtypecode(T), b, m, n)
    work = CuVector{UInt8}(undef, max(lw, 256))
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:makb200_qr_batched_plan_create, libmakb200), Cint,
        (Ptr{Cvoid}, Cint, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{CuPtr{T}}, Ptr{Cint}, Ptr{CuPtr{T}}, Ptr{Cint}, Ptr{CuPtr{T}}, Ptr{Cint}, CuPtr{UInt8}, Csize_t, Ref{Ptr{Cvoid}}),
        h, dtypecode(T), b, m, n, Ap, lda, Qp, ldq, Rp, ldr, work, length(work), ref)
    chkargsok(rc, "makb200_qr_batched_plan_create")
    plan = QRBatchedPlan{T}(ref[], work, (As, Qs, Rs))
    finalizer(p -> ccall((:makb200_qr_batched_plan_destroy, libmakb200), Cint, (Ptr{Cvoid},), p.ptr), plan)
    return plan
end
function run!(plan::QRBatchedPlan)
    rc = ccall((:makb200_qr_batched_plan_run, libmakb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, CuPtr{Cint}), handle(), plan.ptr, CU_NULL)
    chkargsok(rc, "makb200_qr_batched_plan_run")
    return plan.keep[2], plan.keep[3]
end

# ---- batched svd_compact! / eigh_full! over a vector of blocks (per-block semantics incl. gauge) ---
function svd_batched!(As::Vector{<:StridedCuMatrix{T}}, Us::Vector{<:StridedCuMatrix{T}}, Ss::Vector{<:StridedCuVector{Float64}},
        Vhs::Vector{<:StridedCuMatrix{T}}; fixgauge::Bool = true) where {T <: B200Float}
    b = length(As)
    m = Cint[size(A, 1) for A in As]; n = Cint[size(A, 2) for A in As]
    lda = Cint[max(1, stride(A, 2)) for A in As]; ldu = Cint[max(1, stride(U, 2)) for U in Us]
    ldvh = Cint[max(1, stride(V, 2)) for V in Vhs]
    Ap = CuPtr{T}[pointer(A) for A in As]; Up = CuPtr{T}[pointer(U) for U in Us]
    Sp = CuPtr{Float64}[pointer(S) for S in Ss]; Vp = CuPtr{T}[pointer(V) for V in Vhs]
    h = handle()
    lw = ccall((:makb200_svd_batched_worksize, libmakb200), Csize_t, (Ptr{Cvoid}, Cint, Cint, Ptr{Cint}, Ptr{Cint}), h, dtypecode(T), b, m, n)
    with_workspace(lw) do work
        rc = ccall((:makb200_svd_batched, libmakb200), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cint}, Ptr{Cint}, Ptr{CuPtr{T}}, Ptr{Cint}, Ptr{CuPtr{Float64}}, Ptr{CuPtr{T}}, Ptr{Cint},
             Ptr{CuPtr{T}}, Ptr{Cint}, CuPtr{Cint}, CuPtr{UInt8}, Csize_t),
            h, dtypecode(T), fixgauge, b, m, n, Ap, lda, Sp, Up, ldu, Vp, ldvh, CU_NULL, work, lw)
        chkargsok(rc, "makb200_svd_batched")
    end
    return Us, Ss, Vhs
end
function eigh_batched!(As::Vector{<:StridedCuMatrix{T}}, Ws::Vector{<:StridedCuVector{Float64}}, Vs::Vector{<:StridedCuMatrix{T}};
        fixgauge::Bool = true) where {T <: B200Float}
    b = length(As)
    n = Cint[checksquare(A) for A in As]
    lda = Cint[max(1, stride(A, 2)) for A in As]; ldv = Cint[max(1, stride(V, 2)) for V in Vs]
    Ap = CuPtr{T}[pointer(A) for A in As]; Wp = CuPtr{Float64}[pointer(W) for W in Ws]; Vp = CuPtr{T}[pointer(V) for V in Vs]
    h = handle()
    lw = ccall((:makb200_eigh_batched_worksize, libmakb200), Csize_t, (Ptr{Cvoid}, Cint, Cint, Ptr{Cint}), h, dtypecode(T), b, n)
    with_workspace(lw) do work
        rc = ccall((:makb200_eigh_batched, libmakb200), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cint}, Ptr{CuPtr{T}}, Ptr{Cint}, Ptr{CuPtr{Float64}}, Ptr{CuPtr{T}}, Ptr{Cint}, CuPtr{Cint},
             CuPtr{UInt8}, Csize_t),
            h, dtypecode(T), fixgauge, b, n, Ap, lda, Wp, Vp, ldv, CU_NULL, work, lw)
        chkargsok(rc, "makb200_eigh_batched")
    end
    return Ws, Vs
end

# ---- TSQR local step (row shard of a tall-skinny matrix) and adjoint (lq_via_qr!, svd_via_adjoint!) -
function tsqr_local!(A::StridedCuMatrix{T}, Q::StridedCuMatrix{T}, R::StridedCuMatrix{T}) where {T <: B200Float}
    m, n = size(A); h = handle()
    lw = ccall((:makb200_tsqr_local_worksize, libmakb200), Csize_t, (Ptr{Cvoid}, Cint, Cint, Cint), h, dtypecode(T), m, n)
    info = CUDA.zeros(Cint, 1)
    with_workspace(lw) do work
        rc = ccall((:makb200_tsqr_local, libmakb200), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{T}, Cint, CuPtr{T}, Cint, CuPtr{T}, Cint, CuPtr{UInt8}, Csize_t, CuPtr{Cint}),
            h, dtypecode(T), m, n, A, max(1, stride(A, 2)), Q, max(1, stride(Q, 2)), R, max(1, stride(R, 2)), work, lw, info)
        chkargsok(rc, "makb200_tsqr_local")
    end
    return Q, R, info        # info[1] != 0: Cholesky breakdown (kappa too large for CholeskyQR2); read lazily
end
function adjoint!(B::StridedCuMatrix{T}, A::StridedCuMatrix{T}) where {T <: B200Float}
    m, n = size(A)
    size(B) == (n, m) || throw(DimensionMismatch("adjoint!: B must be $n x $m"))
    rc = ccall((:makb200_adjoint, libmakb200), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{T}, Cint, CuPtr{T}, Cint),
        handle(), dtypecode(T), m, n, A, max(1, stride(A, 2)), B, max(1, stride(B, 2)))
    chkargsok(rc, "makb200_adjoint")
    return B
end

# ---- L1 shim: unmqr!(side, trans, A, tau, C) (yalapack.jl:688-735; yacusolver.jl:14) ---------------------------
function unmqr!(side::AbstractChar, trans::AbstractChar, A::StridedCuMatrix{T}, tau::StridedCuVector{T}, C::StridedCuMatrix{T}) where {T <: B200Float}
    side == 'L' || throw(ArgumentError("the B200 driver provides unmqr! for side = 'L' only"))
    chkstride1(A, C)
    m, n = size(C)
    k = length(tau)
    size(A, 1) == m || throw(DimensionMismatch("A has $(size(A, 1)) rows, C has $m"))
    h = handle()
    lw = ccall((:makb200_ormqr_worksize, libmakb200), Csize_t, (Ptr{Cvoid}, Cint, Cint, Cint, Cint), h, dtypecode(T), m, n, k)
    with_workspace(lw) do work
        rc = ccall((:makb200_ormqr, libmakb200), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, Cint, CuPtr{T}, Cint, CuPtr{T}, CuPtr{T}, Cint, CuPtr{UInt8}, Csize_t),
            h, dtypecode(T), 0, opcode(trans), m, n, k, A, max(1, stride(A, 2)), tau, C, max(1, stride(C, 2)), work, lw)
        chkargsok(rc, "makb200_ormqr")
    end
    return C
end

# ---- gaugefix!(eigh_full!, V) as one launch (common/gauge.jl:38-45) ------------------------------------------------
function gauge_columns!(V::StridedCuMatrix{T}) where {T <: B200Float}
    rc = ccall((:makb200_gauge_columns, libmakb200), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{T}, Cint),
        handle(), dtypecode(T), size(V, 1), size(V, 2), V, max(1, stride(V, 2)))
    chkargsok(rc, "makb200_gauge_columns")
    return V
end

# ---- multi-GPU TSQR: one call, NCCL communicator in rank order = row order of the shards -------------------------
# `comm` is an `NCCL.Communicator` (NCCL.jl); its raw `ncclComm_t` is `comm.handle`.  Hosts without NCCL.jl can build
# one with `comm_create` (the 128-byte id from `nccl_unique_id()` on rank 0, shipped by any side channel).
function nccl_unique_id()
    id = Vector{UInt8}(undef, 128)
    rc = ccall((:makb200_nccl_unique_id, libmakb200), Cint, (Ptr{UInt8},), id)
    rc == 0 || error("makb200_nccl_unique_id failed with code $rc (libnccl.so.2 not loadable)")
    return id
end
function comm_create(nranks::Integer, rank::Integer, id::Vector{UInt8})
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:makb200_comm_create, libmakb200), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{UInt8}), ref, nranks, rank, id)
    rc == 0 || error("makb200_comm_create failed with code $rc")
    return ref[]
end
comm_destroy(comm::Ptr{Cvoid}) = ccall((:makb200_comm_destroy, libmakb200), Cint, (Ptr{Cvoid},), comm)

function tsqr!(comm::Ptr{Cvoid}, nranks::Integer, A::StridedCuMatrix{T}, Q::StridedCuMatrix{T}, R::StridedCuMatrix{T}) where {T <: B200Float}
    chkstride1(A, Q, R)
    m, n = size(A)
    h = handle()
    lw = ccall((:makb200_tsqr_worksize, libmakb200), Csize_t, (Ptr{Cvoid}, Cint, Cint, Cint, Cint), h, dtypecode(T), m, n, nranks)
    info = CUDA.zeros(Cint, 1)
    with_workspace(lw) do work
        rc = ccall((:makb200_tsqr, libmakb200), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, CuPtr{T}, Cint, CuPtr{T}, Cint, CuPtr{T}, Cint, CuPtr{UInt8}, Csize_t, CuPtr{Cint}),
            h, comm, dtypecode(T), m, n, A, max(1, stride(A, 2)), Q, max(1, stride(Q, 2)), R, max(1, stride(R, 2)), work, lw, info)
        chkargsok(rc, "makb200_tsqr")
    end
    Array(info)[1] == 0 || error("tsqr!: Cholesky breakdown in the local factorization (kappa(A) too large for CholeskyQR2)")
    return Q, R
end

end # module
