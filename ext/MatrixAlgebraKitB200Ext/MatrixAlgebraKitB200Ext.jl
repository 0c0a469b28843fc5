# MatrixAlgebraKitB200Ext — B200-native backend for MatrixAlgebraKit.jl (fork: the `CUDA` weak-dep
# trigger of Project.toml:21 is re-pointed from MatrixAlgebraKitCUDAExt to this module; both cannot
# be loaded on the same trigger without method overwrites).
#
# Mirrors ext/MatrixAlgebraKitCUDAExt/MatrixAlgebraKitCUDAExt.jl:19-62 (defaults + L1 shims) and,
# for whole-op fusion, ext/MatrixAlgebraKitGenericLinearAlgebraExt.jl:56-82 (L2 override of
# `qr_householder!`).  Core-package additions this file relies on are listed in ext/CORE_PATCH.md.
#
# NOTE: not executable in the build environment (no Julia); the tested twin is the Python host
# layer matrixalgebrakit.jl_b200/{qr,svd,eigh,polar}.py over the same C ABI.
module MatrixAlgebraKitB200Ext

using MatrixAlgebraKit
using MatrixAlgebraKit: @algdef, Algorithm, check_input, one!, zero!, diagview
using MatrixAlgebraKit: B200, Householder, DivideAndConquer, SafeDivideAndConquer, RobustRepresentations, SVDViaPolar, B200_QDWH, TruncationByValue
using MatrixAlgebraKit: default_qr_algorithm, default_svd_algorithm, default_eigh_algorithm
import MatrixAlgebraKit: geqrf!, ungqr!, unmqr!, gesvdp!, gesdd!, gesdvd!, gesvd!, gesvdj!, heevd!, heevr!, heev!, heevj!
import MatrixAlgebraKit: qr_householder!, qr_null_householder!, left_polar!
using CUDA
using LinearAlgebra
using LinearAlgebra: BlasFloat

include("yab200.jl")

const B200Mat = StridedCuMatrix{<:YAB200.B200Float}

# ---- defaults (pattern: MatrixAlgebraKitCUDAExt.jl:19-29) -----------------------------------------
MatrixAlgebraKit.default_driver(::Type{TA}) where {TA <: StridedCuVecOrMat{<:BlasFloat}} = B200()
MatrixAlgebraKit.default_svd_algorithm(::Type{T}; kwargs...) where {T <: StridedCuVecOrMat{<:BlasFloat}} = SVDViaPolar(; kwargs...)
MatrixAlgebraKit.default_eigh_algorithm(::Type{T}; kwargs...) where {T <: StridedCuVecOrMat{<:BlasFloat}} = DivideAndConquer(; kwargs...)
MatrixAlgebraKit.default_algorithm(::typeof(left_polar!), ::Type{T}; kwargs...) where {T <: StridedCuVecOrMat{<:BlasFloat}} = B200_QDWH(; kwargs...)

# Float32/ComplexF32 are not provided (SURVEY A8): fail with an ArgumentError instead of a MethodError
_chk_eltype(A) = eltype(A) <: YAB200.B200Float || throw(ArgumentError("the B200 driver provides Float64 and ComplexF64 only, got $(eltype(A))"))

# ---- L1 shims (pattern: MatrixAlgebraKitCUDAExt.jl:32-34,52-62) -----------------------------------
geqrf!(::B200, A::StridedCuMatrix, args...) = (_chk_eltype(A); YAB200.geqrf!(A, args...))
ungqr!(::B200, A::StridedCuMatrix, tau, Q = similar(A)) = (_chk_eltype(A); YAB200.ungqr!(A, tau, Q))
heevd!(::B200, A::StridedCuMatrix, Dd::StridedCuVector, V::StridedCuMatrix; kwargs...) = (_chk_eltype(A); YAB200.heevd!(A, Dd, V))
gesvdp!(::B200, A::StridedCuMatrix, S::StridedCuVector, U::StridedCuMatrix, Vᴴ::StridedCuMatrix; kwargs...) = (_chk_eltype(A); YAB200.gesvdp!(A, S, U, Vᴴ))
unmqr!(::B200, side, trans, A::StridedCuMatrix, tau, C::StridedCuMatrix) = (_chk_eltype(A); YAB200.unmqr!(side, trans, A, tau, C))

# The LAPACK-named tags with `driver = B200()` (what the deprecated aliases `LAPACK_DivideAndConquer(; driver = B200())`
# and `LAPACK_MultipleRelativelyRobustRepresentations(; driver = B200())` expand to, svd.jl:352-368, eigh.jl:210-217) are
# the same kernels: the tag names the contract (sorted values, gauge-fixed vectors), the driver the implementation.
# Reference bodies that dispatch on these shims: implementations/svd.jl:196-201, eigh.jl:150-156.
gesdd!(::B200, A::StridedCuMatrix, S::StridedCuVector, U::StridedCuMatrix, Vᴴ::StridedCuMatrix; kwargs...) = (_chk_eltype(A); YAB200.gesvdp!(A, S, U, Vᴴ))
gesdvd!(::B200, A::StridedCuMatrix, S::StridedCuVector, U::StridedCuMatrix, Vᴴ::StridedCuMatrix; kwargs...) = (_chk_eltype(A); YAB200.gesvdp!(A, S, U, Vᴴ))
heevr!(::B200, A::StridedCuMatrix, Dd::StridedCuVector, V::StridedCuMatrix; kwargs...) = (_chk_eltype(A); YAB200.heevd!(A, Dd, V))
# ... and the ones this driver does not provide throw (no MethodError, SURVEY A8)
for f in (:gesvd!, :gesvdj!, :heev!, :heevj!)
    @eval $f(d::B200, args...; kwargs...) = throw(ArgumentError(LazyString("driver ", d, " does not provide `", $(QuoteNode(f)), "`")))
end

# svd_full! (job 'A', implementations/svd.jl:202-212): the compact decomposition in the leading columns / rows, the
# complement by qr_null_householder! (k reflectors applied to [0; I]), each extra vector with its own gauge
MatrixAlgebraKit.supports_svd_full(::B200, f::Symbol) = f in (:svd_polar, :divide_and_conquer, :safe_divide_and_conquer)
function _svd_full_b200!(A, U, S, Vᴴ; fixgauge::Bool = true)
    isempty(A) && return one!(U), zero!(S), one!(Vᴴ)
    _chk_eltype(A)
    m, n = size(A)
    k = min(m, n)
    zero!(S)
    Sd = CUDA.zeros(Float64, k)
    Uc = view(U, :, 1:k)
    Vc = m >= n ? Vᴴ : similar(A, (k, n))
    YAB200.gesvdp!(A, Sd, Uc, Vc; fixgauge)
    m < n && copyto!(view(Vᴴ, 1:k, :), Vc)
    diagview(S) .= Sd
    if m > k
        N = qr_null_householder!(B200(), copy(Uc), view(U, :, (k + 1):m))
        fixgauge && YAB200.gauge_columns!(N)
    end
    if n > k
        Nt = qr_null_householder!(B200(), YAB200.adjoint!(similar(A, (n, k)), Vc), similar(A, (n, n - k)))
        fixgauge && YAB200.gauge_columns!(Nt)      # column gauge before the adjoint = the reference's row rule (gauge.jl:61-64)
        YAB200.adjoint!(view(Vᴴ, (k + 1):n, :), Nt)
    end
    return U, S, Vᴴ
end
MatrixAlgebraKit.svd_full_svd_polar!(::B200, A, U, S, Vᴴ; kwargs...) = _svd_full_b200!(A, U, S, Vᴴ; kwargs...)
MatrixAlgebraKit.svd_full_divide_and_conquer!(::B200, A, U, S, Vᴴ; kwargs...) = _svd_full_b200!(A, U, S, Vᴴ; kwargs...)
MatrixAlgebraKit.svd_full_safe_divide_and_conquer!(::B200, A, U, S, Vᴴ; kwargs...) = _svd_full_b200!(A, U, S, Vᴴ; kwargs...)

# qr_null! the reference's way (implementations/qr.jl:236-262: N = [0; I], geqrf!, unmqr!('L','N')); the LQ family and
# left_orth!/right_orth!/left_null!/right_null! reach it through lq_via_qr! / lq_null_via_qr! (lq.jl:130-131,303-327)
function qr_null_householder!(driver::B200, A::AbstractMatrix, N::AbstractMatrix;
        positive::Bool = true, pivoted::Bool = false, blocksize::Int = 0)
    _chk_eltype(A)
    blocksize <= 1 || throw(ArgumentError(lazy"$driver does not provide a blocked QR decomposition"))
    pivoted && throw(ArgumentError(lazy"$driver does not provide a pivoted QR decomposition"))
    m, n = size(A)
    minmn = min(m, n)
    zero!(N)
    one!(view(N, (minmn + 1):m, 1:(m - minmn)))
    A, τ = geqrf!(driver, A)
    return unmqr!(driver, 'L', 'N', A, τ, N)
end

# ---- batched public entry points (SURVEY 8f rank 4): one C call for a vector of blocks, per-block semantics -------
function MatrixAlgebraKit.qr_compact!(As::Vector{<:B200Mat}, QRs::Vector{<:Tuple}, alg::Householder)
    Qs, Rs = [qr[1] for qr in QRs], [qr[2] for qr in QRs]
    foreach((A, qr) -> check_input(qr_compact!, A, qr, alg), As, QRs)
    YAB200.qr_batched!(As, Qs, Rs)
    return QRs
end
function MatrixAlgebraKit.svd_compact!(As::Vector{<:B200Mat}, USVs::Vector{<:Tuple}, alg::Union{SVDViaPolar, DivideAndConquer})
    foreach((A, usv) -> check_input(svd_compact!, A, usv, alg), As, USVs)
    YAB200.svd_batched!(As, [u[1] for u in USVs], [diagview(u[2]) for u in USVs], [u[3] for u in USVs];
        fixgauge = get(alg.kwargs, :fixgauge, true))
    return USVs
end
function MatrixAlgebraKit.eigh_full!(As::Vector{<:B200Mat}, DVs::Vector{<:Tuple}, alg::DivideAndConquer)
    foreach((A, dv) -> check_input(eigh_full!, A, dv, alg), As, DVs)
    YAB200.eigh_batched!(As, [diagview(dv[1]) for dv in DVs], [dv[2] for dv in DVs];
        fixgauge = get(alg.kwargs, :fixgauge, MatrixAlgebraKit.default_fixgauge()))
    return DVs
end

# ---- L2: whole-op QR (factorization + R extraction + Q formation + gauge in one C call) ---------
function qr_householder!(driver::B200, A::AbstractMatrix, Q::AbstractMatrix, R::AbstractMatrix;
        positive::Bool = true, pivoted::Bool = false, blocksize::Int = 0)
    _chk_eltype(A)
    # capability negatives are thrown, not ignored (implementations/qr.jl:140-145)
    blocksize <= 1 || throw(ArgumentError(lazy"$driver does not provide a blocked QR decomposition"))
    pivoted && throw(ArgumentError(lazy"$driver does not provide a pivoted QR decomposition"))
    Q === A && throw(ArgumentError("inplace Q is not supported by the B200 driver"))
    m = size(A, 1)
    YAB200.qr!(A, Q, R; full = size(Q, 2) == m && size(Q, 2) != min(size(A)...), positive)
    return Q, R
end

# ---- L2: fused gauge for eigh / svd (more specific than the generated bodies, eigh.jl:150, svd.jl:196)
function MatrixAlgebraKit.eigh_full_divide_and_conquer!(::B200, A, DV; fixgauge::Bool = MatrixAlgebraKit.default_fixgauge(), kwargs...)
    D, V = DV
    _chk_eltype(A)
    YAB200.heevd!(A, diagview(D), V; fixgauge)
    return DV
end
function MatrixAlgebraKit.eigh_full_robust_representations!(::B200, A, DV; fixgauge::Bool = MatrixAlgebraKit.default_fixgauge(), kwargs...)
    D, V = DV
    _chk_eltype(A)
    YAB200.heevd!(A, diagview(D), V; fixgauge)
    return DV
end
function MatrixAlgebraKit.svd_compact_divide_and_conquer!(d::B200, A, U, S, Vᴴ; kwargs...)
    return MatrixAlgebraKit.svd_compact_svd_polar!(d, A, U, S, Vᴴ; kwargs...)
end
function MatrixAlgebraKit.svd_compact_safe_divide_and_conquer!(d::B200, A, U, S, Vᴴ; kwargs...)
    return MatrixAlgebraKit.svd_compact_svd_polar!(d, A, U, S, Vᴴ; kwargs...)
end
function MatrixAlgebraKit.svd_compact_svd_polar!(::B200, A, U, S, Vᴴ; fixgauge::Bool = true, kwargs...)
    isempty(A) && return one!(U), zero!(S), one!(Vᴴ)
    _chk_eltype(A)
    YAB200.gesvdp!(A, diagview(S), U, Vᴴ; fixgauge)
    return U, S, Vᴴ
end

# Hermitian tests in one device pass instead of `A == A'` / the allocating fallback
# (MatrixAlgebraKitCUDAExt.jl:147-152; common/matrixproperties.jl:86-112)
function MatrixAlgebraKit.ishermitian_approx(A::B200Mat; atol, rtol, kwargs...)
    defect, _, fro, _ = YAB200.hermitian_props(A, false)
    return defect <= max(atol, rtol * fro)
end
function MatrixAlgebraKit.isantihermitian_approx(A::B200Mat; atol, rtol, kwargs...)
    defect, _, fro, _ = YAB200.hermitian_props(A, true)
    return defect <= max(atol, rtol * fro)
end
MatrixAlgebraKit.ishermitian_exact(A::B200Mat; kwargs...) = YAB200.hermitian_props(A, false)[4] == 0
MatrixAlgebraKit.isantihermitian_exact(A::B200Mat; kwargs...) = YAB200.hermitian_props(A, true)[4] == 0

# project_hermitian! / project_antihermitian!: the whole NativeBlocked walk (projections.jl:86-107, per-tile
# @cuda launches in MatrixAlgebraKitCUDAExt.jl:130-143) as ONE launch; B === A is the default output
MatrixAlgebraKit.project_hermitian_native!(A::B200Mat, B::B200Mat, ::Val{anti}; kwargs...) where {anti} =
    YAB200.project_hermitian!(A, B, anti)

# one! / uppertriangular! / lowertriangular! (common/initialization.jl:11-36; n `zero!` launches on a CuArray): one launch
MatrixAlgebraKit.one!(A::B200Mat) = (isempty(A) ? A : YAB200.tri_init!(A, 0))
MatrixAlgebraKit.uppertriangular!(A::B200Mat) = YAB200.tri_init!(A, 1)
MatrixAlgebraKit.lowertriangular!(A::B200Mat) = YAB200.tri_init!(A, 2)

# is_left_isometric (matrixproperties.jl:53-58): Gram matrix on the DMMA GEMM, both norms from one reduction
function MatrixAlgebraKit.is_left_isometric(A::B200Mat; atol::Real = 0, rtol::Real = MatrixAlgebraKit.defaulttol(A), kwargs...)
    P = similar(A, (size(A, 2), size(A, 2)))
    YAB200.gemm!('C', 'N', one(eltype(A)), A, A, zero(eltype(A)), P)
    nP, dP = YAB200.gram_defect(P)
    return dP <= max(atol, rtol * nP)
end

# ---- left_polar!(A, (W,P), ::B200_QDWH) ------------------------------------------------------------
function left_polar!(A::AbstractMatrix, WP, alg::B200_QDWH)
    check_input(left_polar!, A, WP, alg)
    _chk_eltype(A)
    W, P = WP
    isempty(A) && return W, P
    Ac = copy(A)   # QDWH destroys A; kept for the singular case below
    YAB200.polar_qdwh!(A, W, P; l0 = get(alg.kwargs, :l0, 0.0), maxiter = get(alg.kwargs, :maxiter, 12))
    # singular A: QDWH converges to a partial isometry (‖W‖_F² = rank); take the PolarViaSVD recipe (polar.jl:59-70)
    # on the rank-robust B200 SVD so that W is isometric for any input, as with LAPACK
    abs(YAB200.fro2(W) - size(A, 2)) <= 1.0e-9 * size(A, 2) || return left_polar!(Ac, WP, MatrixAlgebraKit.PolarViaSVD(SVDViaPolar()))
    return W, P
end

# svd_trunc!(A, alg) with truncrank(r) and no caller-provided outputs: only the r leading triplets' vectors are
# formed (implementations/svd.jl:226-237 computes the full compact SVD and slices).  Same values, same ϵ.
function _leading(A::B200Mat, alg::MatrixAlgebraKit.TruncatedAlgorithm{<:SVDViaPolar, <:MatrixAlgebraKit.TruncationByOrder})
    t = alg.trunc
    r, k = t.howmany, min(size(A)...)
    (t.by === abs && t.rev && 0 < r < k) || return nothing
    S = CUDA.zeros(Float64, k)
    U, Vᴴ = similar(A, (size(A, 1), r)), similar(A, (r, size(A, 2)))
    YAB200.svd_leading!(A, r, S, U, Vᴴ; fixgauge = get(alg.alg.kwargs, :fixgauge, true))
    return U, S, Vᴴ, r
end
function MatrixAlgebraKit.svd_trunc!(A::B200Mat, alg::MatrixAlgebraKit.TruncatedAlgorithm{<:SVDViaPolar, <:MatrixAlgebraKit.TruncationByOrder})
    res = _leading(A, alg)
    res === nothing && return MatrixAlgebraKit.svd_trunc!(A, MatrixAlgebraKit.initialize_output(svd_trunc!, A, alg), alg)
    U, S, Vᴴ, r = res
    ϵ = norm(view(S, (r + 1):length(S)))
    return U, Diagonal(S[1:r]), Vᴴ, ϵ
end

# sorted-values truncation search on the device vector (same as MatrixAlgebraKitCUDAExt.jl:64-66)
MatrixAlgebraKit.findtruncated_svd(values::StridedCuVector, strategy::TruncationByValue) =
    MatrixAlgebraKit.findtruncated(values, strategy)

end # module
