// Single-launch kernels for the steps just before and after the factorizations:
//   project_hermitian! / project_antihermitian!   (implementations/projections.jl:60-139)
//   ishermitian / isantihermitian, exact and approx (common/matrixproperties.jl:77-195)
//   isisometric                                     (common/matrixproperties.jl:53-58)
// The reference walks 32x32 blocks on the host, which on a CuArray is O((n/32)^2) launches; here each
// is ONE launch, 32x32 tiles staged through shared memory so both the (r,c) and the (c,r) side of a
// pair are read and written coalesced.  Device code only, written against the subset of CUDA that
// tests/cpu_harness/cuda_emu.h emulates (tests/test_emu_kernels_cpu.py runs it with g++).
#pragma once
#include "devutil.cuh"

namespace mak {

// (a + conj(b))/2 or (a - conj(b))/2: `_project_hermitian` (projections.jl:109-111).  With b the
// mirror entry this one expression is the reference's value on BOTH sides of the diagonal (its
// lower entry is +-adjoint(val), which equals the expression with the roles swapped bit for bit) and
// on the diagonal ((a + conj(a))/2 = real(a), (a - conj(a))/2 = i imag(a) exactly).
template <bool ANTI, typename T>
__device__ __forceinline__ T herm_part(T a, T b) {
    return scale_(ANTI ? sub_(a, conj_(b)) : add_(a, conj_(b)), 0.5);
}

// B = (A +- A^H)/2.  One CTA (32 x 8 threads) per PAIR of mirror tiles (bi <= bj): both tiles are
// read into shared memory before either is written, so B may be A itself (the reference's default:
// initialize_output returns A, projections.jl:38-43).
template <typename T, bool ANTI>
__global__ void project_herm_kernel(int n, const T* A, int lda, T* B, int ldb) {
    __shared__ T tu[32][33];   // tile (bi, bj): tu[k][t] = A[bi*32 + t, bj*32 + k]
    __shared__ T tl[32][33];   // tile (bj, bi): tl[k][t] = A[bj*32 + t, bi*32 + k]
    const int bi = blockIdx.x, bj = blockIdx.y;
    if (bi > bj) return;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int k = ty; k < 32; k += 8) {
        const int ru = bi * 32 + tx, cu = bj * 32 + k;
        tu[k][tx] = (ru < n && cu < n) ? A[(size_t)cu * lda + ru] : zero<T>();
        if (bi != bj) {
            const int rl = bj * 32 + tx, cl = bi * 32 + k;
            tl[k][tx] = (rl < n && cl < n) ? A[(size_t)cl * lda + rl] : zero<T>();
        }
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int ru = bi * 32 + tx, cu = bj * 32 + k;
        // mirror of (ru, cu) is (cu, ru): row k, column tx of tile (bj, bi)
        const T mu = (bi != bj) ? tl[tx][k] : tu[tx][k];
        if (ru < n && cu < n) B[(size_t)cu * ldb + ru] = herm_part<ANTI>(tu[k][tx], mu);
        if (bi != bj) {
            const int rl = bj * 32 + tx, cl = bi * 32 + k;
            if (rl < n && cl < n) B[(size_t)cl * ldb + rl] = herm_part<ANTI>(tl[k][tx], tu[tx][k]);
        }
    }
}

// One pass over A for every Hermitian / anti-Hermitian test:
//   out[0] += || (A -+ A^H)/2 ||_F^2   the part that must vanish (ANTI = false: the anti-Hermitian part)
//   out[1]  = max |A_ij|               (norm(A, Inf) of default_hermitian_tol, common/defaults.jl:44)
//   out[2] += || A ||_F^2              (for rtol)
//   out[3] += number of entries (i <= j) with A_ij != +-conj(A_ji)   (the exact test, matrixproperties.jl:115-150)
// out must be zeroed by the caller.  Atomics: the results only feed threshold tests.
template <typename T, bool ANTI>
__global__ void herm_props_kernel(int n, const T* __restrict__ A, int lda, double* out) {
    __shared__ T tile[32][33];
    __shared__ double red[4][8];
    const int bi = blockIdx.x, bj = blockIdx.y;
    if (bj > bi) return;   // lower block pairs only
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int k = ty; k < 32; k += 8) {
        const int r = bj * 32 + tx, c = bi * 32 + k;
        tile[k][tx] = (r < n && c < n) ? A[(size_t)c * lda + r] : zero<T>();
    }
    __syncthreads();
    double part = 0.0, mx = 0.0, fro = 0.0, bad = 0.0;
    for (int k = ty; k < 32; k += 8) {
        const int r = bi * 32 + tx, c = bj * 32 + k;   // entry (r, c) of tile (bi, bj); its mirror (c, r) = tile[tx][k]
        if (r < n && c < n && (bi != bj || r >= c)) {
            const T a = A[(size_t)c * lda + r];
            const T b = tile[tx][k];
            const T dlt = herm_part<!ANTI>(a, b);
            const double w = (r == c) ? 1.0 : 2.0;    // every mirror pair is visited once
            part += w * abs2_(dlt);
            fro += (r == c) ? abs2_(a) : abs2_(a) + abs2_(b);
            mx = fmax(mx, fmax(sqrt(abs2_(a)), sqrt(abs2_(b))));
            const T want = ANTI ? neg_(conj_(b)) : conj_(b);
            if (real_(a) != real_(want) || imag_(a) != imag_(want)) bad += 1.0;
        }
    }
    const int t = ty * 32 + tx;
    part = warp_sum(part);
    fro = warp_sum(fro);
    bad = warp_sum(bad);
    mx = warp_max(mx);
    if ((t & 31) == 0) {
        red[0][t >> 5] = part;
        red[1][t >> 5] = mx;
        red[2][t >> 5] = fro;
        red[3][t >> 5] = bad;
    }
    __syncthreads();
    if (t == 0) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int i = 0; i < 8; ++i) {
            s0 += red[0][i];
            s1 = fmax(s1, red[1][i]);
            s2 += red[2][i];
            s3 += red[3][i];
        }
        atomicAdd(out, s0);
        // non-negative doubles order like their bit patterns
        atomicMax((unsigned long long*)(out + 1), (unsigned long long)__double_as_longlong(s1));
        atomicAdd(out + 2, s2);
        atomicAdd(out + 3, s3);
    }
}

// Isometry test on the Gram matrix P = A^H A (is_left_isometric, matrixproperties.jl:53-58):
//   out[0] += ||P||_F^2,  out[1] += ||P - I||_F^2.   Grid-stride over the n x n entries, 256 threads.
template <typename T>
__global__ void gram_defect_kernel(int n, const T* __restrict__ P, int ldp, double* out) {
    __shared__ double red[2][8];
    double a = 0.0, b = 0.0;
    const size_t total = (size_t)n * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx % n), c = (int)(idx / n);
        const T p = P[(size_t)c * ldp + r];
        a += abs2_(p);
        b += abs2_(r == c ? sub_(p, one<T>()) : p);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    const int t = threadIdx.x;
    if ((t & 31) == 0) {
        red[0][t >> 5] = a;
        red[1][t >> 5] = b;
    }
    __syncthreads();
    if (t == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
            s0 += red[0][i];
            s1 += red[1][i];
        }
        atomicAdd(out, s0);
        atomicAdd(out + 1, s1);
    }
}

// one! / uppertriangular! / lowertriangular! (src/common/initialization.jl:11-36; on a CuArray the reference's
// uppertriangular! is one `zero!` launch PER COLUMN): one grid-stride launch over the m x n entries.
//   mode 0: A = I (rectangular identity)   1: zero below the diagonal   2: zero above the diagonal
template <typename T>
__global__ void tri_init_kernel(int mode, int m, int n, T* __restrict__ A, int lda) {
    const size_t total = (size_t)m * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx % m), c = (int)(idx / m);
        T* p = A + (size_t)c * lda + r;
        if (mode == 0) *p = (r == c) ? one<T>() : zero<T>();
        else if (mode == 1) { if (r > c) *p = zero<T>(); }
        else { if (r < c) *p = zero<T>(); }
    }
}

// out[0] += ||A||_F^2 of an m x n matrix (grid-stride, one atomic per CTA; feeds threshold tests only, e.g.
// ||W||_F^2 = n for an isometric polar factor, = rank for the partial isometry QDWH returns on singular input)
template <typename T>
__global__ void fro2_atomic_kernel(int m, int n, const T* __restrict__ A, int lda, double* out) {
    __shared__ double red[8];
    double a = 0.0;
    const size_t total = (size_t)m * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx % m), c = (int)(idx / m);
        a += abs2_(A[(size_t)c * lda + r]);
    }
    a = warp_sum(a);
    const int t = threadIdx.x;
    if ((t & 31) == 0) red[t >> 5] = a;
    __syncthreads();
    if (t == 0) {
        double s0 = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s0 += red[i];
        atomicAdd(out, s0);
    }
}

// upper <- conj(lower) of an n x n matrix, in place, tiled so both sides are coalesced; the diagonal is made
// real.  One CTA per tile pair (bi >= bj) reads the lower tile into shared memory and writes its adjoint into
// the mirror tile.  Used by the lower-triangle variant of the dense -> band reduction (sy2sb_t, qr.cu), whose
// rank-2b update only computes the tiles on and below the diagonal.
template <typename T>
__global__ void mirror_lower_kernel(int n, T* __restrict__ A, int lda) {
    __shared__ T tile[32][33];
    const int bi = blockIdx.x, bj = blockIdx.y;   // tile rows bi*32.., columns bj*32..
    if (bi < bj) return;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int k = ty; k < 32; k += 8) {
        const int r = bi * 32 + tx, c = bj * 32 + k;
        tile[k][tx] = (r < n && c < n) ? A[(size_t)c * lda + r] : zero<T>();
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        // target (r2, c2) in tile (bj, bi): r2 = bj*32 + tx, c2 = bi*32 + k;  source (c2, r2) = tile[tx][k]
        const int r2 = bj * 32 + tx, c2 = bi * 32 + k;
        if (r2 < n && c2 < n) {
            if (r2 < c2) A[(size_t)c2 * lda + r2] = conj_(tile[tx][k]);
            else if (r2 == c2) A[(size_t)c2 * lda + r2] = mk<T>(real_(tile[tx][k]));
        }
    }
}

// out[0] = max_j | ||U(:, j)||^2 - 1 |  (one warp per column; non-negative doubles order like their bit patterns).
// For a full-rank A the left factor U = W V is isometric to rounding; for a rank-deficient A the QDWH polar
// factor W is only a partial isometry and the columns of U that belong to zero singular values collapse.
template <typename T>
__global__ void col_norm_defect_kernel(int m, int ncols, const T* __restrict__ U, int ldu, double* out) {
    const int lane = threadIdx.x & 31, j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= ncols) return;
    const T* u = U + (size_t)j * ldu;
    double s = 0.0;
    for (int r = lane; r < m; r += 32) s += abs2_(u[r]);
    s = warp_sum(s);
    if (lane == 0) {
        double dft = fabs(s - 1.0);
        if (!(dft == dft)) dft = 1e300;   // NaN counts as a defect
        atomicMax((unsigned long long*)out, (unsigned long long)__double_as_longlong(dft));
    }
}

}  // namespace mak
