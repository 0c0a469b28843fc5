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
using MatrixAlgebraKit: B200, Householder, DivideAndConquer, SVDViaPolar, B200_QDWH, TruncationByValue
using MatrixAlgebraKit: default_qr_algorithm, default_svd_algorithm, default_eigh_algorithm
import MatrixAlgebraKit: geqrf!, ungqr!, gesvdp!, heevd!, qr_householder!, left_polar!
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
MatrixAlgebraKit.supports_svd_full(::B200, f::Symbol) = false

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
