// left_polar! by QDWH (Nakatsukasa, Bai, Gygi 2010; Nakatsukasa & Higham 2013) and
// svd_compact! = polar + eigh (the reference's `SVDViaPolar` tag, cuSOLVER gesvdp class).
//
// The reference has no QDWH (its polar is PolarViaSVD / PolarNewton, implementations/polar.jl:
// 59-166); the contract is the reference's RESULT: W isometric, P Hermitian PSD, A = W P.
// Every step is GEMM-shaped and runs on the DMMA GEMM:
//   QR-type iteration   (c > 100):  [sqrt(c) X; I] = [Q1;Q2] R,  X <- (b/c) X + (a-b/c)/sqrt(c) Q1 Q2^H
//   Cholesky iteration  (c <= 100): Z = I + c X^H X = L L^H,     X <- (b/c) X + (a-b/c) (X L^-H) L^-1
// The iteration schedule (a,b,c per step) depends only on the scalar lower bound l0 and is
// computed on the host, so the device loop runs without a host round trip.
#include "polar.cuh"
#include "potf2.cuh"
#include "eigh.cuh"
#include "gemm.cuh"
#include "qr.cuh"
#include "projections.cuh"
#include "nccl_dl.h"
#include <vector>
#include <cmath>

namespace mak {

constexpr size_t POTRF_AUX_BYTES = (size_t)64 * 128 * 128 * sizeof(double);

#define MAK_GEMM2(h, ...)                                              \
    do {                                                               \
        cudaError_t _e = gemm<T>(__VA_ARGS__);                         \
        if (_e != cudaSuccess) return cuda_fail(h, _e, "gemm");        \
    } while (0)

static inline int grid_for2(size_t total, int num_sms) {
    size_t b = (total + 255) / 256, cap = (size_t)num_sms * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------------------------------
// elementwise helpers
// ---------------------------------------------------------------------------------------
// out[0] = sum |A_ij|^2 (deterministic two-stage reduction: partials then a single block)
template <typename T>
__global__ void fro2_partial_kernel(int m, int n, const T* __restrict__ A, int lda, double* __restrict__ partial) {
    __shared__ double red[32];
    size_t total = (size_t)m * n;
    double s = 0.0;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % m), c = (int)(idx / m);
        s += abs2_(A[(size_t)c * lda + r]);
    }
    double t = block_sum<double>(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
__global__ void fro2_final_kernel(int np, const double* __restrict__ partial, double* out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) s += partial[i];
    double t = block_sum<double>(s, red);
    if (threadIdx.x == 0) out[0] = t;
}

// max |re|, |im| over the matrix (two-stage), for the overflow/underflow-safe scaling of X0: the plain sum of
// squares underflows to 0 for ||A||_F < 1e-154 and overflows for entries > 1e154 (LAPACK's lassq scales; so do we)
template <typename T>
__global__ void absmax_partial_kernel(int m, int n, const T* __restrict__ A, int lda, double* __restrict__ partial) {
    __shared__ double red[32];
    size_t total = (size_t)m * n;
    double mx = 0.0;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % m), c = (int)(idx / m);
        const T a = A[(size_t)c * lda + r];
        const double v = fmax(fabs(real_(a)), fabs(imag_(a)));
        mx = (v > mx || v != v) ? v : mx;     // NaN propagates
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t = fmax(t, red[i]);
        partial[blockIdx.x] = t;
    }
}
// out[0] = 1 / max (1 for a zero or non-finite matrix)
__global__ void absmax_final_kernel(int np, const double* __restrict__ partial, double* out) {
    __shared__ double red[32];
    double mx = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) mx = fmax(mx, partial[i]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t = fmax(t, red[i]);
        out[0] = (t > 0.0 && isfinite(t)) ? 1.0 / t : 1.0;
    }
}
// sum |A_ij * pre[0]|^2 partials
template <typename T>
__global__ void fro2_pre_partial_kernel(int m, int n, const T* __restrict__ A, int lda, const double* __restrict__ pre,
                                        double* __restrict__ partial) {
    __shared__ double red[32];
    const double f = pre[0];
    size_t total = (size_t)m * n;
    double s = 0.0;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % m), c = (int)(idx / m);
        s += abs2_(scale_(A[(size_t)c * lda + r], f));
    }
    double t = block_sum<double>(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
// X = (A * pre[0]) / sqrt(norm2[0])
template <typename T>
__global__ void scale_copy_pre_kernel(int m, int n, const T* __restrict__ A, int lda, T* __restrict__ X, int ldx,
                                      const double* __restrict__ pre, const double* __restrict__ norm2) {
    const double nn = norm2[0], f = pre[0];
    const double inv = nn > 0.0 ? 1.0 / sqrt(nn) : 1.0;
    size_t total = (size_t)m * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % m), c = (int)(idx / m);
        X[(size_t)c * ldx + r] = scale_(scale_(A[(size_t)c * lda + r], f), inv);
    }
}

// X = A / sqrt(norm2[0])   (X0 = A / ||A||_F); a zero matrix is copied unchanged
template <typename T>
__global__ void scale_copy_kernel(int m, int n, const T* __restrict__ A, int lda, T* __restrict__ X, int ldx,
                                  const double* __restrict__ norm2) {
    const double nn = norm2[0];
    const double inv = nn > 0.0 ? 1.0 / sqrt(nn) : 1.0;
    size_t total = (size_t)m * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % m), c = (int)(idx / m);
        X[(size_t)c * ldx + r] = scale_(A[(size_t)c * lda + r], inv);
    }
}

// B = [sqrt(c) X ; I]  ((m+n) x n)
template <typename T>
__global__ void stack_kernel(int m, int n, const T* __restrict__ X, int ldx, T* __restrict__ B, int ldb, double sc) {
    size_t total = (size_t)(m + n) * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % (m + n)), c = (int)(idx / (m + n));
        T v;
        if (r < m) v = scale_(X[(size_t)c * ldx + r], sc);
        else v = (r - m == c) ? one<T>() : zero<T>();
        B[(size_t)c * ldb + r] = v;
    }
}

// Z = I  (n x n)
template <typename T>
__global__ void eye_kernel(int n, T* __restrict__ Z, int ldz) {
    size_t total = (size_t)n * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % n), c = (int)(idx / n);
        Z[(size_t)c * ldz + r] = (r == c) ? one<T>() : zero<T>();
    }
}

// X = alpha*X + beta*Y
template <typename T>
__global__ void axpby_kernel(int m, int n, double alpha, T* __restrict__ X, int ldx, double beta,
                             const T* __restrict__ Y, int ldy) {
    size_t total = (size_t)m * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % m), c = (int)(idx / m);
        T* px = X + (size_t)c * ldx + r;
        *px = add_(scale_(*px, alpha), scale_(Y[(size_t)c * ldy + r], beta));
    }
}

template <typename T>
__global__ void copy2d_kernel(int m, int n, const T* __restrict__ S, int lds, T* __restrict__ D, int ldd) {
    size_t total = (size_t)m * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % m), c = (int)(idx / m);
        D[(size_t)c * ldd + r] = S[(size_t)c * lds + r];
    }
}

// X *= factor / sqrt(val[0])   (val on the device: no host round trip)
template <typename T>
__global__ void scale_by_dev_kernel(int m, int n, T* __restrict__ X, int ldx, const double* __restrict__ val,
                                    double factor) {
    const double v = val[0];
    const double f = (v > 0.0 && isfinite(v)) ? factor / sqrt(v) : 1.0;
    size_t total = (size_t)m * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % m), c = (int)(idx / m);
        T* p = X + (size_t)c * ldx + r;
        *p = scale_(*p, f);
    }
}

__global__ void sqrt_dev_kernel(double* v) { v[0] = sqrt(v[0]); }

// deterministic +-1 entries (Hutchinson probes / power-iteration start): hash of the index
template <typename T>
__global__ void rademacher_kernel(int m, int n, T* __restrict__ G, int ldg, unsigned seed) {
    size_t total = (size_t)m * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        unsigned hsh = (unsigned)idx * 2654435761u + seed;
        hsh ^= hsh >> 16; hsh *= 0x85ebca6bu; hsh ^= hsh >> 13; hsh *= 0xc2b2ae35u; hsh ^= hsh >> 16;
        int r = (int)(idx % m), c = (int)(idx / m);
        G[(size_t)c * ldg + r] = mk<T>((hsh & 1u) ? 1.0 : -1.0);
    }
}

// y[c] = sum_r conj(Z[r,c]) v[r]  (= (Z v)[c] for Hermitian Z): one warp per column, coalesced
template <typename T>
__global__ void herm_matvec_kernel(int n, const T* __restrict__ Z, int ldz, const T* __restrict__ v, T* __restrict__ y) {
    const int lane = threadIdx.x & 31, c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n) return;
    const T* col = Z + (size_t)c * ldz;
    T s = zero<T>();
    for (int r = lane; r < n; r += 32) fmac_(s, col[r], v[r]);
    s = warp_sum(s);
    if (lane == 0) y[c] = s;
}

// Z *= factor / val[0]  (all n x n entries)
template <typename T>
__global__ void scale_by_dev_lin_kernel(int n, T* __restrict__ Z, int ldz, const double* __restrict__ val, double factor) {
    const double v = val[0];
    const double f = (v > 0.0 && isfinite(v)) ? factor / v : 1.0;
    size_t total = (size_t)n * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % n), c = (int)(idx / n);
        T* p = Z + (size_t)c * ldz + r;
        *p = scale_(*p, f);
    }
}

// Z (lower part) = I + c * Z0   (reuses the Gram matrix X^H X computed for the estimate)
template <typename T>
__global__ void eye_plus_scaled_kernel(int n, T* __restrict__ Z, int ldz, const T* __restrict__ Z0, int ld0, double c) {
    size_t total = (size_t)n * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % n), cc = (int)(idx / n);
        if (r < cc) continue;
        T v = scale_(Z0[(size_t)cc * ld0 + r], c);
        if (r == cc) v = add_(v, one<T>());
        Z[(size_t)cc * ldz + r] = v;
    }
}

// P = (H + H^H)/2, tiled so both sides are coalesced (project_hermitian!, one launch)
template <typename T>
__global__ void symmetrize_kernel(int n, const T* __restrict__ H, int ldh, T* __restrict__ P, int ldp) {
    __shared__ T tile[32][33];
    const int bi = blockIdx.x, bj = blockIdx.y, tx = threadIdx.x, ty = threadIdx.y;
    for (int k = ty; k < 32; k += 8) {
        int r = bj * 32 + tx, c = bi * 32 + k;
        tile[k][tx] = (r < n && c < n) ? H[(size_t)c * ldh + r] : zero<T>();
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        int r = bi * 32 + tx, c = bj * 32 + k;
        if (r < n && c < n) {
            T a = H[(size_t)c * ldh + r];
            T b = conj_(tile[tx][k]);
            T v = scale_(add_(a, b), 0.5);
            if (r == c) v = mk<T>(real_(v));
            P[(size_t)c * ldp + r] = v;
        }
    }
}

// D = S^H  (S: m x n  ->  D: n x m)
template <typename T>
__global__ void adjoint_kernel(int m, int n, const T* __restrict__ S, int lds, T* __restrict__ D, int ldd) {
    __shared__ T tile[32][33];
    const int bi = blockIdx.x, bj = blockIdx.y, tx = threadIdx.x, ty = threadIdx.y;
    for (int k = ty; k < 32; k += 8) {
        int r = bi * 32 + tx, c = bj * 32 + k;
        tile[k][tx] = (r < m && c < n) ? S[(size_t)c * lds + r] : zero<T>();
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        int r = bj * 32 + tx, c = bi * 32 + k;  // D[r, c] = conj(S[c, r]) = conj(tile[tx][k])
        if (r < n && c < m) D[(size_t)c * ldd + r] = conj_(tile[tx][k]);
    }
}

template <typename T>
static int potf2_init(makb200_handle* h) {
    constexpr int nb = CholNB<T>::value;
    MAK_CUDA(h, cudaFuncSetAttribute(potf2_inv_kernel<T, nb>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(sizeof(T) * nb * (nb + 1))));
    return 0;
}

int polar_init(makb200_handle* h) {
    int rc = potf2_init<double>(h);
    if (rc) return rc;
    return potf2_init<cplx>(h);
}

// left-looking blocked Cholesky: Z (n x n Hermitian, lower part read, destroyed) -> L (lower; only
// blocks strictly below the block diagonal are referenced later) and Linv (nb x nb per block).
// Per block column: (a) update of the nb x nb diagonal block, (b) its one-CTA factorization + inverse (~150 us of
// pure latency), (c) update of the panel below it, (d) panel times the inverse.  (b) is the serial chain: n/nb of them
// were ~40 % of the 23 ms this took at n = 8192.  With look-ahead (a) and (b) run on the handle's high-priority
// auxiliary stream while (c) - the bulk of the flops - runs on the caller's stream; (d) joins the two.
static bool potrf_lookahead() {
    static const bool v = []() { const char* e = getenv("MAKB200_POTRF_LOOKAHEAD"); return !(e && e[0] == '0'); }();
    return v;
}
template <typename T>
static int potrf_blocked(makb200_handle* h, int n, T* Z, int ldz, T* L, int ldl, T* Linv, int* info,
                         void* ws = nullptr, size_t ws_bytes = 0) {
    constexpr int nb = CholNB<T>::value;
    cudaStream_t s = h->stream;
    const T one_ = one<T>(), zero_ = zero<T>(), mone = neg_(one<T>());
    // the diagonal-block product on the auxiliary stream is ONE tile with K up to n: it needs split-K, hence its own
    // scratch - the tail of `ws` (the carve sites add POTRF_AUX_BYTES for it)
    constexpr size_t aux_bytes = (size_t)64 * nb * nb * sizeof(T);
    const bool la = potrf_lookahead() && !h->no_lookahead && n >= 8 * nb && ws && ws_bytes > 2 * aux_bytes;
    void* ws_aux = nullptr;
    if (la) { ws_bytes -= aux_bytes; ws_aux = (char*)ws + ws_bytes; }
    cudaStream_t sA = la ? h->aux_stream : s;
    for (int j0 = 0, b = 0; j0 < n; j0 += nb, ++b) {
        const int jb = (n - j0 < nb) ? (n - j0) : nb;
        const int mr = n - j0 - jb;
        if (la) {
            // fork: everything queued so far (the previous column's L) is visible to the auxiliary stream
            MAK_CUDA(h, cudaEventRecord(h->ev[6], s));
            MAK_CUDA(h, cudaStreamWaitEvent(sA, h->ev[6], 0));
        }
        if (j0 > 0) {
            if (la) {
                // (a) diagonal block on the auxiliary stream (no split-K scratch: `ws` belongs to the main stream)
                MAK_GEMM2(h, sA, h->num_sms, MAKB200_OP_N, MAKB200_OP_C, jb, jb, j0, mone, L + j0, ldl, L + j0, ldl, one_,
                          Z + (size_t)j0 * ldz + j0, ldz, ws_aux, aux_bytes);
                // (c) panel below it on the caller's stream
                if (mr > 0)
                    MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_C, mr, jb, j0, mone, L + j0 + jb, ldl, L + j0, ldl, one_,
                              Z + (size_t)j0 * ldz + j0 + jb, ldz, ws, ws_bytes);
            } else {
                MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_C, n - j0, jb, j0, mone, L + j0, ldl, L + j0, ldl,
                          one_, Z + (size_t)j0 * ldz + j0, ldz, ws, ws_bytes);
            }
        }
        potf2_inv_kernel<T, nb><<<1, 256, sizeof(T) * nb * (nb + 1), sA>>>(jb, Z + (size_t)j0 * ldz + j0, ldz,
                                                                       L + (size_t)j0 * ldl + j0, ldl,
                                                                       Linv + (size_t)b * nb * nb, nb, info);
        count_launch();
        MAK_LAUNCH_CHECK(h, "potf2_inv_kernel");
        if (la) {
            // join before (d)
            MAK_CUDA(h, cudaEventRecord(h->ev[7], sA));
            MAK_CUDA(h, cudaStreamWaitEvent(s, h->ev[7], 0));
        }
        if (mr > 0)
            MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_C, mr, jb, jb, one_, Z + (size_t)j0 * ldz + j0 + jb,
                      ldz, Linv + (size_t)b * nb * nb, nb, zero_, L + (size_t)j0 * ldl + j0 + jb, ldl, nullptr, 0);
    }
    return 0;
}

// Y = X L^-H (conjtrans) or Y = X L^-1, X m x n untouched, Tmp m x nb scratch
template <typename T>
static int trsm_right(makb200_handle* h, bool conjtrans, int m, int n, const T* X, int ldx, const T* L, int ldl,
                      const T* Linv, T* Y, int ldy, T* Tmp, void* ws = nullptr, size_t ws_bytes = 0) {
    constexpr int nb = CholNB<T>::value;
    cudaStream_t s = h->stream;
    const T one_ = one<T>(), zero_ = zero<T>(), mone = neg_(one<T>());
    const int nblk = (n + nb - 1) / nb;
    for (int bb = 0; bb < nblk; ++bb) {
        const int b = conjtrans ? bb : (nblk - 1 - bb);
        const int j0 = b * nb, jb = (n - j0 < nb) ? (n - j0) : nb;
        copy2d_kernel<T><<<grid_for2((size_t)m * jb, h->num_sms), 256, 0, s>>>(m, jb, X + (size_t)j0 * ldx, ldx, Tmp, m);
        count_launch();
        if (conjtrans) {
            // (Y L^H)_j = sum_{i<=j} Y_i L_ji^H
            if (j0 > 0)
                MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_C, m, jb, j0, mone, Y, ldy, L + j0, ldl, one_, Tmp,
                          m, ws, ws_bytes);
            MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_C, m, jb, jb, one_, Tmp, m,
                      Linv + (size_t)b * nb * nb, nb, zero_, Y + (size_t)j0 * ldy, ldy, nullptr, 0);
        } else {
            // (Y L)_j = sum_{i>=j} Y_i L_ij
            const int j1 = j0 + jb, rest = n - j1;
            if (rest > 0)
                MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_N, m, jb, rest, mone, Y + (size_t)j1 * ldy, ldy,
                          L + (size_t)j0 * ldl + j1, ldl, one_, Tmp, m, ws, ws_bytes);
            MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_N, m, jb, jb, one_, Tmp, m,
                      Linv + (size_t)b * nb * nb, nb, zero_, Y + (size_t)j0 * ldy, ldy, nullptr, 0);
        }
    }
    MAK_LAUNCH_CHECK(h, "trsm_right");
    return 0;
}

// ---------------------------------------------------------------------------------------
// QDWH
// ---------------------------------------------------------------------------------------
template <typename T>
struct PolarWork {
    T *X, *B, *Q, *Z, *L, *Linv, *Y, *Y2, *Tmp, *R0, *Q0;
    double* scal;      // [8]
    double* partial;   // [1024]
    int* info;
    void* sub;         // QR workspace
    size_t sub_bytes;
    void* ws;          // split-K scratch for the skinny GEMMs of potrf / trsm
    size_t ws_bytes;
};

template <typename T, typename AR>
static void polar_carve(makb200_handle* h, AR& ar, int m, int n, bool tall, PolarWork<T>* w) {
    constexpr int nb = CholNB<T>::value;
    const int ms = tall ? n : m;  // rows of the iterated matrix
    size_t nn = (size_t)(n > 0 ? n : 1), mm = (size_t)(ms > 0 ? ms : 1);
    w->X = ar.template get<T>(mm * nn);
    w->B = ar.template get<T>((mm + nn) * nn);
    w->Q = ar.template get<T>((mm + nn) * nn);
    w->Z = ar.template get<T>(nn * nn);
    w->L = ar.template get<T>(nn * nn);
    w->Linv = ar.template get<T>((size_t)nb * nb * ((nn + nb - 1) / nb));
    w->Y = ar.template get<T>(mm * nn);
    w->Y2 = ar.template get<T>(mm * nn);
    w->Tmp = ar.template get<T>((mm + nn) * nb);
    w->R0 = tall ? ar.template get<T>(nn * nn) : nullptr;
    w->Q0 = tall ? ar.template get<T>((size_t)m * nn) : nullptr;
    w->scal = ar.template get<double>(8);
    w->partial = ar.template get<double>(1024);
    w->info = ar.template get<int>(4);
    size_t a = qr_worksize_t<T>(h, (int)(mm + nn), n, n);
    size_t b = tall ? qr_worksize_t<T>(h, m, n, n) : 0;
    w->sub_bytes = a > b ? a : b;
    w->sub = ar.template get<char>(w->sub_bytes);
    w->ws_bytes = (size_t)h->num_sms * 128 * 128 * sizeof(double);
    if (nn <= 1024 && w->ws_bytes < 64 * nn * nn * sizeof(T)) w->ws_bytes = 64 * nn * nn * sizeof(T);   // split-K up to 64
    w->ws_bytes += POTRF_AUX_BYTES;   // split-K scratch of the Cholesky look-ahead stream (potrf_blocked)
    w->ws = ar.template get<char>(w->ws_bytes);
}

static inline bool polar_tall(int m, int n) { return m > n + n / 8; }

template <typename T>
size_t polar_worksize_t(makb200_handle* h, int m, int n) {
    ArenaSize ar;
    PolarWork<T> w;
    polar_carve<T>(h, ar, m, n, polar_tall(m, n), &w);
    return ar.off + 256;
}

// polar factor of a (ms x n, ms >= n) matrix held in w.X (already scaled so that ||X||_2 <= 1)
template <typename T>
static int qdwh_iterate(makb200_handle* h, int ms, int n, PolarWork<T>& w, double l0, int maxiter, int* iters_out,
                        const T* Z0 = nullptr) {
    cudaStream_t s = h->stream;
    static const bool raise_c = []() { const char* e = getenv("MAKB200_QDWH_CMAX_N8"); return !(e && e[0] == '0'); }();
    const double cmax = (raise_c && n / 8.0 > 100.0) ? n / 8.0 : 100.0;
    std::vector<QdwhStep> sched = qdwh_schedule(l0, maxiter, cmax);
    if (iters_out) *iters_out = (int)sched.size();
    PhaseTimer pt(s);
    pt.mark("start");
    bool first = true;
    auto gram = [&](double c) -> int {
        // Z (lower) = I + c X^H X; the first step reuses the Gram matrix of the estimate
        if (first && Z0) {
            eye_plus_scaled_kernel<T><<<grid_for2((size_t)n * n, h->num_sms), 256, 0, s>>>(n, w.Z, n, Z0, n, c);
            count_launch();
            return 0;
        }
        eye_kernel<T><<<grid_for2((size_t)n * n, h->num_sms), 256, 0, s>>>(n, w.Z, n);
        count_launch();
        MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_C, MAKB200_OP_N, n, n, ms, mk<T>(c), w.X, ms, w.X, ms, one<T>(), w.Z, n,
                  nullptr, 0, true);
        return 0;
    };
    for (const QdwhStep& st : sched) {
        if (st.qr) {
            const int mb = ms + n;
            stack_kernel<T><<<grid_for2((size_t)mb * n, h->num_sms), 256, 0, s>>>(ms, n, w.X, ms, w.B, mb, sqrt(st.c));
            count_launch();
            MAK_LAUNCH_CHECK(h, "stack_kernel");
            const T* Qf;
            if (st.c <= QDWH_CHOLQR_MAX_C) {
                // cond([sqrt(c)X; I]) <= sqrt(1+c) <= 1e6: the orthonormal basis can be formed by
                // CholeskyQR2 (all DMMA GEMMs) instead of Householder QR.
                // pass 1: Z = B^H B = c X^H X + I (lower), L1, Q' = B L1^-H
                int rc = gram(st.c);
                if (rc) return rc;
                rc = potrf_blocked<T>(h, n, w.Z, n, w.L, n, w.Linv, w.info, w.ws, w.ws_bytes);
                if (rc) return rc;
                rc = trsm_right<T>(h, true, mb, n, w.B, mb, w.L, n, w.Linv, w.Q, mb, w.Tmp, w.ws, w.ws_bytes);
                if (rc) return rc;
                // pass 2: Z = Q'^H Q' (lower), L2, Q = Q' L2^-H  (into B's storage)
                MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_C, MAKB200_OP_N, n, n, mb, one<T>(), w.Q, mb, w.Q, mb, zero<T>(),
                          w.Z, n, nullptr, 0, true);
                rc = potrf_blocked<T>(h, n, w.Z, n, w.L, n, w.Linv, w.info, w.ws, w.ws_bytes);
                if (rc) return rc;
                rc = trsm_right<T>(h, true, mb, n, w.Q, mb, w.L, n, w.Linv, w.B, mb, w.Tmp, w.ws, w.ws_bytes);
                if (rc) return rc;
                Qf = w.B;
            } else {
                int rc = qr_fused_t<T>(h, MAKB200_QR_COMPACT, mb, n, w.B, mb, w.Q, mb, (T*)nullptr, 0, w.sub,
                                       w.sub_bytes);
                if (rc) return rc;
                Qf = w.Q;
            }
            const double al = (st.a - st.b / st.c) / sqrt(st.c), be = st.b / st.c;
            MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_C, ms, n, n, mk<T>(al), Qf, mb, Qf + ms, mb,
                      mk<T>(be), w.X, ms, nullptr, 0);
            pt.mark(st.c <= QDWH_CHOLQR_MAX_C ? "cholqrstep" : "qrstep");
        } else {
            // Z is Hermitian and only its lower triangle is read by potrf: tiles above it are skipped
            int rc = gram(st.c);
            if (rc) return rc;
            pt.mark("gram");
            rc = potrf_blocked<T>(h, n, w.Z, n, w.L, n, w.Linv, w.info, w.ws, w.ws_bytes);
            if (rc) return rc;
            pt.mark("potrf");
            rc = trsm_right<T>(h, true, ms, n, w.X, ms, w.L, n, w.Linv, w.Y, ms, w.Tmp, w.ws, w.ws_bytes);
            if (rc) return rc;
            rc = trsm_right<T>(h, false, ms, n, w.Y, ms, w.L, n, w.Linv, w.Y2, ms, w.Tmp, w.ws, w.ws_bytes);
            if (rc) return rc;
            pt.mark("trsm2");
            axpby_kernel<T><<<grid_for2((size_t)ms * n, h->num_sms), 256, 0, s>>>(ms, n, st.b / st.c, w.X, ms,
                                                                                  st.a - st.b / st.c, w.Y2, ms);
            count_launch();
            MAK_LAUNCH_CHECK(h, "axpby_kernel");
        }
        first = false;
    }
    pt.report("qdwh steps");
    return 0;
}

template <typename T>
int polar_qdwh_t(makb200_handle* h, int m, int n, T* A, int lda, T* W, int ldw, T* P, int ldp, double l0,
                 int maxiter, void* work, size_t lwork, int* iters_host, int* info_dev) {
    if (n <= 0 || m <= 0) return 0;
    cudaStream_t s = h->stream;
    const bool tall = polar_tall(m, n);
    Arena ar(work, lwork);
    PolarWork<T> w;
    polar_carve<T>(h, ar, m, n, tall, &w);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    MAK_CUDA(h, cudaMemsetAsync(w.info, 0, sizeof(int) * 4, s));
    PhaseTimer pt(s);
    pt.mark("start");
    const T* src = A;
    int lds = lda, ms = m;
    if (tall) {
        // A = Q0 R0 first (as PolarNewton does, implementations/polar.jl:131-135); iterate on R0
        int rc = qr_fused_t<T>(h, MAKB200_QR_COMPACT, m, n, A, lda, w.Q0, m, w.R0, n, w.sub, w.sub_bytes);
        if (rc) return rc;
        src = w.R0; lds = n; ms = n;
    }
    const int np = grid_for2((size_t)ms * n, h->num_sms) < 1024 ? grid_for2((size_t)ms * n, h->num_sms) : 1024;
    // X0 = A / ||A||_F, with the entries pre-scaled by 1 / max|a_ij| (scal[7]) so that neither a tiny nor a huge A
    // loses the sum of squares; W is scale-invariant and P is formed from the unscaled A, so nothing is undone later
    absmax_partial_kernel<T><<<np, 256, 0, s>>>(ms, n, src, lds, w.partial);
    absmax_final_kernel<<<1, 256, 0, s>>>(np, w.partial, w.scal + 7);
    fro2_pre_partial_kernel<T><<<np, 256, 0, s>>>(ms, n, src, lds, w.scal + 7, w.partial);
    fro2_final_kernel<<<1, 256, 0, s>>>(np, w.partial, w.scal);
    scale_copy_pre_kernel<T><<<grid_for2((size_t)ms * n, h->num_sms), 256, 0, s>>>(ms, n, src, lds, w.X, ms, w.scal + 7, w.scal);
    count_launch(5);
    MAK_LAUNCH_CHECK(h, "scale_copy_kernel");
    pt.mark("prep");
    // ---- scaling and conditioning estimates (large matrices, default l0 only) -------------------
    //  sigma_max: 8 power iterations on X^H X through skinny GEMMs -> X /= 1.1 sigma_max (on device)
    //  sigma_min: Z0 = X^H X = L L^H, tr(Z0^-1) = ||L^-1||_F^2 by Hutchinson (8 probes)
    //             => sigma_min >= 1/sqrt(tr); l0 = 0.3/sqrt(tr_est).  One device->host read.
    //  Falls back to l0 = eps (valid for any kappa <= 1e16) if the Cholesky of Z0 breaks down.
    const T* Z0 = nullptr;
    static const bool est_on = []() { const char* e = getenv("MAKB200_QDWH_ESTIMATE"); return !(e && e[0] == '0'); }();
    if (est_on && l0 <= 2.3e-16 && n >= 1024) {
        const T one_ = one<T>(), zero_ = zero<T>();
        // Z0 = X^H X (full Hermitian: the power iteration needs both triangles) in Y2's storage
        T* Z0w = w.Y2;
        MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_C, MAKB200_OP_N, n, n, ms, one_, w.X, ms, w.X, ms, zero_, Z0w, n, nullptr, 0);
        pt.mark("est_gram");
        T* v = w.Tmp;                 // n
        T* y = w.Tmp + n;             // n
        rademacher_kernel<T><<<grid_for2((size_t)n, h->num_sms), 256, 0, s>>>(n, 1, v, n, 12345u);
        count_launch();
        const int mvgrid = (n + 7) / 8;
        for (int it = 0; it < 10; ++it) {
            // y = Z0 v ; lambda ~ ||y|| for unit v ; v = y / ||y||
            herm_matvec_kernel<T><<<mvgrid, 256, 0, s>>>(n, Z0w, n, v, y);
            fro2_partial_kernel<T><<<32, 256, 0, s>>>(n, 1, y, n, w.partial);
            fro2_final_kernel<<<1, 256, 0, s>>>(32, w.partial, w.scal + 1);
            scale_copy_kernel<T><<<grid_for2((size_t)n, h->num_sms), 256, 0, s>>>(n, 1, y, n, v, n, w.scal + 1);
            count_launch(4);
        }
        // after the loop scal[1] = ||Z0 v||^2 for the last unit v, i.e. (lambda_max estimate)^2:
        // sigma_max^2 ~ sqrt(scal[1]).  X /= 1.1 sigma_max  and  Z0 /= 1.21 sigma_max^2
        fro2_final_kernel<<<1, 256, 0, s>>>(32, w.partial, w.scal + 1);
        sqrt_dev_kernel<<<1, 1, 0, s>>>(w.scal + 1);   // scal[1] <- lambda_max estimate = sigma_max^2
        scale_by_dev_kernel<T><<<grid_for2((size_t)ms * n, h->num_sms), 256, 0, s>>>(ms, n, w.X, ms, w.scal + 1, 1.0 / 1.1);
        scale_by_dev_lin_kernel<T><<<grid_for2((size_t)n * n, h->num_sms), 256, 0, s>>>(n, Z0w, n, w.scal + 1, 1.0 / 1.21);
        count_launch(4);
        pt.mark("est_power");
        copy2d_kernel<T><<<grid_for2((size_t)n * n, h->num_sms), 256, 0, s>>>(n, n, Z0w, n, w.Z, n);
        count_launch();
        int rc0 = potrf_blocked<T>(h, n, w.Z, n, w.L, n, w.Linv, w.info + 1, w.ws, w.ws_bytes);
        if (rc0) return rc0;
        pt.mark("est_potrf");
        constexpr int NPROBE = 8;
        T* Gp = w.Y;                               // NPROBE x n probes, then NPROBE x n solution
        T* Yp = w.Y + (size_t)NPROBE * n;
        rademacher_kernel<T><<<grid_for2((size_t)NPROBE * n, h->num_sms), 256, 0, s>>>(NPROBE, n, Gp, NPROBE, 777u);
        count_launch();
        rc0 = trsm_right<T>(h, true, NPROBE, n, Gp, NPROBE, w.L, n, w.Linv, Yp, NPROBE, w.Tmp, w.ws, w.ws_bytes);   // split-K: 8-row products with K up to n
        if (rc0) return rc0;
        fro2_partial_kernel<T><<<64, 256, 0, s>>>(NPROBE, n, Yp, NPROBE, w.partial);
        fro2_final_kernel<<<1, 256, 0, s>>>(64, w.partial, w.scal + 2);
        count_launch(2);
        double hs[4];
        int hinfo[2];
        MAK_CUDA(h, cudaMemcpyAsync(hs, w.scal, sizeof(double) * 3, cudaMemcpyDeviceToHost, s));
        MAK_CUDA(h, cudaMemcpyAsync(hinfo, w.info, sizeof(int) * 2, cudaMemcpyDeviceToHost, s));
        MAK_CUDA(h, cudaStreamSynchronize(s));
        const double tr = hs[2] / NPROBE;
        if (hinfo[1] == 0 && tr > 0.0 && std::isfinite(tr)) {
            double est = 0.3 / sqrt(tr);
            if (est > 1e-7) {   // kappa small enough for the Gram-based estimate to be meaningful
                l0 = est < 0.9 ? est : 0.9;
                Z0 = Z0w;
            }
        }
        pt.mark("est_probe");
        if (pt.on) fprintf(stderr, "[makb200 profile] qdwh estimate: sigma_max^2(X0)=%.3e tr=%.3e l0=%.3e info=%d\n", hs[1], tr, l0, hinfo[1]);
    }
    int rc = qdwh_iterate<T>(h, ms, n, w, l0, maxiter, iters_host, Z0);
    if (rc) return rc;
    pt.mark("qdwh");
    // W, P
    if (tall) {
        MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_N, m, n, n, one<T>(), w.Q0, m, w.X, n, zero<T>(), W, ldw,
                  nullptr, 0);
    } else {
        copy2d_kernel<T><<<grid_for2((size_t)m * n, h->num_sms), 256, 0, s>>>(m, n, w.X, ms, W, ldw);
        count_launch();
    }
    if (P && ldp > 0) {
        // P = sym(W^H A) = sym(X^H src)   (project_hermitian!, polar.jl:108)
        MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_C, MAKB200_OP_N, n, n, ms, one<T>(), w.X, ms, src, lds, zero<T>(), w.Z, n,
                  nullptr, 0);
        int nb32 = (n + 31) / 32;
        symmetrize_kernel<T><<<dim3(nb32, nb32), dim3(32, 8), 0, s>>>(n, w.Z, n, P, ldp);
        count_launch();
        MAK_LAUNCH_CHECK(h, "symmetrize_kernel");
    }
    pt.mark("WP");
    pt.report("polar");
    if (info_dev) MAK_CUDA(h, cudaMemcpyAsync(info_dev, w.info, sizeof(int), cudaMemcpyDeviceToDevice, s));
    return 0;
}

// ---------------------------------------------------------------------------------------
// SVD via polar + eigh
// ---------------------------------------------------------------------------------------
// S[j] = max(w[n-1-j], 0) for j < n;  Vh[j, :] = conj(V[:, n-1-j]) for j < k   (descending order)
template <typename T>
__global__ void svd_reorder_kernel(int n, int k, const double* __restrict__ w, const T* __restrict__ V, int ldv,
                                   double* __restrict__ S, T* __restrict__ Vh, int ldvh) {
    __shared__ T tile[32][33];
    const int bi = blockIdx.x, bj = blockIdx.y, tx = threadIdx.x, ty = threadIdx.y;
    // tile of V: rows bi*32.. (index i), source columns n-1-(bj*32+kk)
    for (int kk = ty; kk < 32; kk += 8) {
        int i = bi * 32 + tx, j = bj * 32 + kk;
        tile[kk][tx] = (i < n && j < k) ? V[(size_t)(n - 1 - j) * ldv + i] : zero<T>();
    }
    __syncthreads();
    for (int kk = ty; kk < 32; kk += 8) {
        int j = bj * 32 + tx, i = bi * 32 + kk;  // Vh[j, i] = conj(V[i, n-1-j]) = conj(tile[tx][kk])
        if (j < k && i < n && Vh) Vh[(size_t)i * ldvh + j] = conj_(tile[tx][kk]);
    }
    if (bi == 0 && ty == 0) {
        int j = bj * 32 + tx;
        if (j < n) { double v = w[n - 1 - j]; S[j] = v > 0.0 ? v : 0.0; }
    }
}

template <typename T>
struct SvdWork {
    T *Wp, *P, *V, *At, *Ut, *Vht;
    double* flag;
    double* wv;
    void* sub;
    size_t sub_bytes;
};

template <typename T, typename AR>
static void svd_carve(makb200_handle* h, AR& ar, int m, int n, SvdWork<T>* w) {
    const bool wide = m < n;
    const int mm = wide ? n : m, nn = wide ? m : n;  // work on the tall orientation
    size_t N = (size_t)(nn > 0 ? nn : 1), M = (size_t)(mm > 0 ? mm : 1);
    w->Wp = ar.template get<T>(M * N);
    w->P = ar.template get<T>(N * N);
    w->V = ar.template get<T>(N * N);
    w->wv = ar.template get<double>(N);
    w->flag = ar.template get<double>(2);
    w->At = wide ? ar.template get<T>(M * N) : nullptr;
    w->Ut = wide ? ar.template get<T>(M * N) : nullptr;
    w->Vht = wide ? ar.template get<T>(N * N) : nullptr;
    size_t a = polar_worksize_t<T>(h, mm, nn);
    size_t b = eigh_worksize_t<T>(h, nn);
    size_t c = qr_worksize_t<T>(h, mm, nn, nn);   // re-orthonormalisation of U for rank-deficient input
    w->sub_bytes = a > b ? a : b;
    if (c > w->sub_bytes) w->sub_bytes = c;
    w->sub = ar.template get<char>(w->sub_bytes);
}

template <typename T>
size_t svd_worksize_t(makb200_handle* h, int m, int n) {
    ArenaSize ar;
    SvdWork<T> w;
    svd_carve<T>(h, ar, m, n, &w);
    return ar.off + 256;
}

template <typename T>
static int svd_tail(makb200_handle* h, int m, int n, int r, double* S, T* U, int ldu, T* Vh, int ldvh, int fixgauge,
                    SvdWork<T>& w, const TrdPre<T>* pre, PhaseTimer& pt);

// tall/square core: A (m x n, m >= n) -> U (m x r), S (n, always all values), Vh (r x n); U/Vh may be null
// (values).  r = n is the compact SVD; r < n (svd_trunc! with a rank known up front) back-transforms
// only the r leading eigenvectors of P and forms U with an m x r x n product.
template <typename T>
static int svd_tall(makb200_handle* h, int m, int n, int r, T* A, int lda, double* S, T* U, int ldu, T* Vh, int ldvh,
                    int fixgauge, double l0, SvdWork<T>& w, int* info_dev) {
    cudaStream_t s = h->stream;
    int iters = 0;
    PhaseTimer pt(s);
    pt.mark("start");
    int rc = polar_qdwh_t<T>(h, m, n, A, lda, w.Wp, m, w.P, n, l0, 12, w.sub, w.sub_bytes, &iters, info_dev);
    if (rc) return rc;
    pt.mark("polar");
    return svd_tail<T>(h, m, n, r, S, U, ldu, Vh, ldvh, fixgauge, w, nullptr, pt);
}

// second half of the SVD: W (w.Wp) and P (w.P) are in place; `pre`: P was already tridiagonalised (batched path)
template <typename T>
static int svd_tail(makb200_handle* h, int m, int n, int r, double* S, T* U, int ldu, T* Vh, int ldvh, int fixgauge,
                    SvdWork<T>& w, const TrdPre<T>* pre, PhaseTimer& pt) {
    cudaStream_t s = h->stream;
    int rc = 0;
    const bool vectors = (U != nullptr && Vh != nullptr);
    // values only (job 'N'): eigenvalues of P by Sturm K-section, no eigenvectors
    rc = eigh_t<T>(h, n, w.P, n, w.wv, vectors ? w.V : (T*)nullptr, n, 0, w.sub, w.sub_bytes, nullptr, r, pre);
    if (rc) return rc;
    pt.mark("eigh");
    int nb32 = (n + 31) / 32;
    T* Vh_eff = vectors ? Vh : nullptr;
    svd_reorder_kernel<T><<<dim3(nb32, nb32), dim3(32, 8), 0, s>>>(n, r, w.wv, w.V, n, S, Vh_eff, ldvh);
    count_launch();
    MAK_LAUNCH_CHECK(h, "svd_reorder_kernel");
    if (vectors) {
        // U = W * V_desc = W * Vh^H
        MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_C, m, r, n, one<T>(), w.Wp, m, Vh, ldvh, zero<T>(), U, ldu,
                  nullptr, 0);
        // Rank-deficient A: W is a partial isometry, so the columns of U that belong to (numerically) zero
        // singular values are not unit vectors.  LAPACK returns an isometric U for any input, so detect the
        // case (one 8-byte read, like the reference's info read per call) and replace U by the Q of its
        // positive-diagonal Householder QR: the leading isometric columns are reproduced, the collapsed
        // ones become an orthonormal completion.
        MAK_CUDA(h, cudaMemsetAsync(w.flag, 0, sizeof(double), s));
        col_norm_defect_kernel<T><<<(r + 7) / 8, 256, 0, s>>>(m, r, U, ldu, w.flag);
        count_launch();
        MAK_LAUNCH_CHECK(h, "col_norm_defect_kernel");
        double defect = 0.0;
        if (h->defect_dev) {
            // graph-replayed batched path (capi.cu): no host read inside the captured sequence; the caller collects the
            // indicator of every block with one read and sends deficient blocks through this function again, uncaptured
            MAK_CUDA(h, cudaMemcpyAsync(h->defect_dev, w.flag, sizeof(double), cudaMemcpyDeviceToDevice, s));
        } else {
            MAK_CUDA(h, cudaMemcpyAsync(&defect, w.flag, sizeof(double), cudaMemcpyDeviceToHost, s));
            MAK_CUDA(h, cudaStreamSynchronize(s));
        }
        if (defect > 1e-6) {
            rc = qr_fused_t<T>(h, MAKB200_QR_COMPACT, m, r, U, ldu, w.Wp, m, (T*)nullptr, 0, w.sub, w.sub_bytes);
            if (rc) return rc;
            copy2d_kernel<T><<<grid_for2((size_t)m * r, h->num_sms), 256, 0, s>>>(m, r, w.Wp, m, U, ldu);
            count_launch();
            MAK_LAUNCH_CHECK(h, "copy2d_kernel");
            pt.mark("U_repair");
        }
        if (fixgauge) {
            rc = gauge_columns<T>(h, m, r, U, ldu, Vh, ldvh, n);
            if (rc) return rc;
        }
    }
    pt.mark("UV");
    pt.report("svd");
    return 0;
}

// ---- phased SVD of one block for the batched path (m >= n): phase 1 = QDWH into caller-held W (m x n, ld m) and P
// (n x n, ld n); the caller then tridiagonalises the P of ALL blocks in one launch (bhetrd_batched_t); phase 2 = the
// eigensolve from (d, e, tau) and U, S, Vh.  `scratch` is per stream slot, the W/P/V/wv/flag buffers are per block. ----
template <typename T>
size_t svd_phase_scratch_t(makb200_handle* h, int m, int n) {
    size_t a = polar_worksize_t<T>(h, m, n), b = eigh_worksize_t<T>(h, n), c = qr_worksize_t<T>(h, m, n, n);
    size_t v = a > b ? a : b;
    return (v > c ? v : c) + 256;
}
template <typename T>
int svd_phase1_t(makb200_handle* h, int m, int n, T* A, int lda, T* Wp, T* P, double l0, void* scratch, size_t lscratch) {
    int iters = 0;
    return polar_qdwh_t<T>(h, m, n, A, lda, Wp, m, P, n, l0, 12, scratch, lscratch, &iters, nullptr);
}
template <typename T>
int svd_phase2_t(makb200_handle* h, int m, int n, T* Wp, T* P, T* V, double* wv, double* flag, const TrdPre<T>* pre,
                 double* S, T* U, int ldu, T* Vh, int ldvh, int fixgauge, void* scratch, size_t lscratch) {
    SvdWork<T> w{};
    w.Wp = Wp; w.P = P; w.V = V; w.wv = wv; w.flag = flag;
    w.sub = scratch; w.sub_bytes = lscratch;
    PhaseTimer pt(h->stream);
    pt.mark("start");
    return svd_tail<T>(h, m, n, n, S, U, ldu, Vh, ldvh, fixgauge, w, pre, pt);
}
template size_t svd_phase_scratch_t<double>(makb200_handle*, int, int);
template size_t svd_phase_scratch_t<cplx>(makb200_handle*, int, int);
template int svd_phase1_t<double>(makb200_handle*, int, int, double*, int, double*, double*, double, void*, size_t);
template int svd_phase1_t<cplx>(makb200_handle*, int, int, cplx*, int, cplx*, cplx*, double, void*, size_t);
template int svd_phase2_t<double>(makb200_handle*, int, int, double*, double*, double*, double*, double*, const TrdPre<double>*,
                                  double*, double*, int, double*, int, int, void*, size_t);
template int svd_phase2_t<cplx>(makb200_handle*, int, int, cplx*, cplx*, cplx*, double*, double*, const TrdPre<cplx>*, double*,
                                cplx*, int, cplx*, int, int, void*, size_t);

template <typename T>
int svd_t(makb200_handle* h, int m, int n, T* A, int lda, double* S, T* U, int ldu, T* Vh, int ldvh, int fixgauge,
          double l0, void* work, size_t lwork, int* info_dev, int r) {
    if (m <= 0 || n <= 0) return 0;
    const int kk = m < n ? m : n;
    if (r <= 0 || r > kk) r = kk;   // leading triplets wanted (all singular values are always returned)
    cudaStream_t s = h->stream;
    Arena ar(work, lwork);
    SvdWork<T> w;
    svd_carve<T>(h, ar, m, n, &w);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    if (m >= n) return svd_tall<T>(h, m, n, r, A, lda, S, U, ldu, Vh, ldvh, fixgauge, l0, w, info_dev);
    // wide: SVD of A^H (svd_via_adjoint!, implementations/svd.jl:134-142): A^H = U' S V'^H
    //   => A = V' S U'^H : U = (Vh')^H (m x m), Vh = (U')^H (m x n)
    const bool vectors = (U != nullptr && Vh != nullptr);
    dim3 g((m + 31) / 32, (n + 31) / 32);
    adjoint_kernel<T><<<g, dim3(32, 8), 0, s>>>(m, n, A, lda, w.At, n);
    count_launch();
    int rc = svd_tall<T>(h, n, m, r, w.At, n, S, vectors ? w.Ut : nullptr, n, vectors ? w.Vht : nullptr, m, 0, l0, w,
                         info_dev);
    if (rc) return rc;
    if (vectors) {
        // Vht (r x m, ld m) -> U (m x r);  Ut (n x r, ld n) -> Vh (r x n)
        dim3 g1((r + 31) / 32, (m + 31) / 32);
        adjoint_kernel<T><<<g1, dim3(32, 8), 0, s>>>(r, m, w.Vht, m, U, ldu);
        dim3 g2((n + 31) / 32, (r + 31) / 32);
        adjoint_kernel<T><<<g2, dim3(32, 8), 0, s>>>(n, r, w.Ut, n, Vh, ldvh);
        count_launch(2);
        MAK_LAUNCH_CHECK(h, "adjoint_kernel");
        if (fixgauge) return gauge_columns<T>(h, m, r, U, ldu, Vh, ldvh, n);
    }
    return 0;
}


// ---------------------------------------------------------------------------------------
// tall-skinny local QR for TSQR: CholeskyQR2 (Fukaya et al. 2014) on the DMMA GEMM
//   G = A^H A = L L^H,  Q1 = A L^-H;  G2 = Q1^H Q1 = L2 L2^H,  Q = Q1 L2^-H,  R = L2^H L^H
// Two passes make ||Q^H Q - I|| = O(eps) provided kappa(A) <~ 1e7 (first Cholesky must succeed;
// info_dev reports a breakdown).  diag(R) > 0 by construction, i.e. the QR gauge holds.
// A (m x n) is overwritten (scratch for the second pass); rows are processed in slabs so the
// scratch stays small for m in the tens of millions.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void rmul_upper_kernel(int n, const T* __restrict__ L2, const T* __restrict__ L1, T* __restrict__ R, int ldr) {
    // R = L2^H * L1^H  (both upper triangular after the adjoint); one thread per entry
    int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r >= n || c >= n) return;
    T s = zero<T>();
    if (r <= c)
        for (int p = r; p <= c; ++p) {
            // (L2^H)[r,p] = conj(L2[p,r]); (L1^H)[p,c] = conj(L1[c,p])
            T a = conj_(L2[(size_t)r * n + p]), b = conj_(L1[(size_t)p * n + c]);
            fma_(s, a, b);
        }
    R[(size_t)c * ldr + r] = s;
}

template <typename T>
__global__ void lower_clean_kernel(int n, T* __restrict__ L) {
    int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r < n && c < n && r < c) L[(size_t)c * n + r] = zero<T>();
}

// shifted CholeskyQR (Fukaya, Kannan, Nakatsukasa, Yamamoto, Yanagisawa 2020): G += s I with
// s = coef * trace(G), coef = 11 (mn + n(n+1)) u  (trace(G) = ||X||_F^2 >= ||X||_2^2).  One CTA.
template <typename T>
__global__ void add_shift_kernel(int n, T* __restrict__ G, double coef) {
    __shared__ double red[256];
    double t = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) t += real_(G[(size_t)i * n + i]);
    red[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    const double sft = coef * red[0];
    for (int i = threadIdx.x; i < n; i += 256) {
        T* d = G + (size_t)i * n + i;
        *d = add_(*d, mk<T>(sft));
    }
}
// R (upper, ld n) = L^H
template <typename T>
__global__ void upper_from_lower_kernel(int n, const T* __restrict__ L, T* __restrict__ R) {
    int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r >= n || c >= n) return;
    R[(size_t)c * n + r] = (r <= c) ? conj_(L[(size_t)r * n + c]) : zero<T>();
}
// Rout (upper, ld ldo) = L^H * Rin   (L lower, Rin upper with ld n)
template <typename T>
__global__ void rmul_upper2_kernel(int n, const T* __restrict__ L, const T* __restrict__ Rin, T* __restrict__ Rout,
                                   int ldo) {
    int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (r >= n || c >= n) return;
    T s = zero<T>();
    if (r <= c)
        for (int p = r; p <= c; ++p) fma_(s, conj_(L[(size_t)r * n + p]), Rin[(size_t)c * n + p]);
    Rout[(size_t)c * ldo + r] = s;
}

constexpr int CQR_SLAB = 1 << 20;

template <typename T>
struct CqrWork {
    T *G, *L1, *L2, *R2, *Linv, *Tmp;
    int* info;
    void* ws;
    size_t ws_bytes;
};

template <typename T, typename AR>
static void cqr_carve(makb200_handle* h, AR& ar, int m, int n, CqrWork<T>* w) {
    constexpr int nb = CholNB<T>::value;
    size_t nn = (size_t)(n > 0 ? n : 1);
    size_t slab = (size_t)(m < CQR_SLAB ? (m > 0 ? m : 1) : CQR_SLAB);
    w->G = ar.template get<T>(nn * nn);
    w->L1 = ar.template get<T>(nn * nn);
    w->L2 = ar.template get<T>(nn * nn);
    w->R2 = ar.template get<T>(nn * nn);
    w->Linv = ar.template get<T>((size_t)nb * nb * ((nn + nb - 1) / nb));
    w->Tmp = ar.template get<T>(slab * nb);
    w->info = ar.template get<int>(4);
    w->ws_bytes = (size_t)h->num_sms * 128 * 128 * sizeof(double);
    if (nn <= 1024 && w->ws_bytes < 64 * nn * nn * sizeof(T)) w->ws_bytes = 64 * nn * nn * sizeof(T);   // split-K up to 64
    w->ws_bytes += POTRF_AUX_BYTES;   // split-K scratch of the Cholesky look-ahead stream (potrf_blocked)
    w->ws = ar.template get<char>(w->ws_bytes);
}

template <typename T>
size_t cholqr2_worksize_t(makb200_handle* h, int m, int n) {
    ArenaSize ar;
    CqrWork<T> w;
    cqr_carve<T>(h, ar, m, n, &w);
    return ar.off + 256;
}

template <typename T>
static int cholqr_pass(makb200_handle* h, int m, int n, const T* X, int ldx, T* Y, int ldy, T* L, CqrWork<T>& w,
                       double shift_coef = 0.0) {
    cudaStream_t s = h->stream;
    // G = X^H X (split-K over the long dimension)
    // (only the lower triangle of G is read by the Cholesky factorization: upper tiles are skipped)
    MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_C, MAKB200_OP_N, n, n, m, one<T>(), X, ldx, X, ldx, zero<T>(), w.G, n, w.ws,
              w.ws_bytes, true);
    if (shift_coef > 0.0) {
        add_shift_kernel<T><<<1, 256, 0, s>>>(n, w.G, shift_coef);
        count_launch();
    }
    int rc = potrf_blocked<T>(h, n, w.G, n, L, n, w.Linv, w.info);
    if (rc) return rc;
    for (int r0 = 0; r0 < m; r0 += CQR_SLAB) {
        int mr = (m - r0 < CQR_SLAB) ? (m - r0) : CQR_SLAB;
        rc = trsm_right<T>(h, true, mr, n, X + r0, ldx, L, n, w.Linv, Y + r0, ldy, w.Tmp);
        if (rc) return rc;
    }
    return 0;
}

template <typename T>
int cholqr2_t(makb200_handle* h, int m, int n, T* A, int lda, T* Q, int ldq, T* R, int ldr, void* work, size_t lwork,
              int* info_dev, int nshift) {
    if (m <= 0 || n <= 0) return 0;
    cudaStream_t s = h->stream;
    Arena ar(work, lwork);
    CqrWork<T> w;
    cqr_carve<T>(h, ar, m, n, &w);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    MAK_CUDA(h, cudaMemsetAsync(w.info, 0, sizeof(int) * 4, s));
    PhaseTimer pt(s);
    pt.mark("start");
    if (nshift > 0) {
        // shifted CholeskyQR: `nshift` preconditioning passes with G + sI (each divides kappa by
        // ~1/sqrt(coef)), then the two plain passes.  R = L_last^H ... L_0^H accumulated pass by pass.
        const double u = 1.1102230246251565e-16;
        const double coef = 11.0 * ((double)m * n + (double)n * (n + 1)) * u;
        const int npass = nshift + 2;
        T* src = A; int lds = lda;
        T* dst = Q; int ldd = ldq;
        T* Racc = nullptr;
        dim3 g((n + 127) / 128, n);
        for (int p = 0; p < npass; ++p) {
            int rc = cholqr_pass<T>(h, m, n, src, lds, dst, ldd, w.L1, w, p < nshift ? coef : 0.0);
            if (rc) return rc;
            lower_clean_kernel<T><<<g, 128, 0, s>>>(n, w.L1);
            if (p == 0) {
                Racc = w.L2;
                upper_from_lower_kernel<T><<<g, 128, 0, s>>>(n, w.L1, Racc);
            } else {
                T* Rn = (Racc == w.L2) ? w.R2 : w.L2;
                rmul_upper2_kernel<T><<<g, 128, 0, s>>>(n, w.L1, Racc, Rn, n);
                Racc = Rn;
            }
            count_launch(2);
            T* t = src; src = dst; dst = t;
            int tl = lds; lds = ldd; ldd = tl;
        }
        // after the swap `src` holds the last result
        if (src != Q) {
            copy2d_kernel<T><<<grid_for2((size_t)m * n, h->num_sms), 256, 0, s>>>(m, n, src, lds, Q, ldq);
            count_launch();
        }
        if (R && ldr > 0) {
            copy2d_kernel<T><<<grid_for2((size_t)n * n, h->num_sms), 256, 0, s>>>(n, n, Racc, n, R, ldr);
            count_launch();
        }
        MAK_LAUNCH_CHECK(h, "shifted cholqr tail");
        pt.mark("shifted");
        pt.report("cholqr3");
        if (info_dev) MAK_CUDA(h, cudaMemcpyAsync(info_dev, w.info, sizeof(int), cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    int rc = cholqr_pass<T>(h, m, n, A, lda, Q, ldq, w.L1, w);   // Q1 -> Q
    if (rc) return rc;
    pt.mark("pass1");
    rc = cholqr_pass<T>(h, m, n, Q, ldq, A, lda, w.L2, w);       // Q2 -> A
    if (rc) return rc;
    pt.mark("pass2");
    copy2d_kernel<T><<<grid_for2((size_t)m * n, h->num_sms), 256, 0, s>>>(m, n, A, lda, Q, ldq);
    count_launch();
    if (R && ldr > 0) {
        // the diagonal blocks of L were written with zeros above the diagonal by potf2; blocks
        // above the block diagonal were never written: clean before the triangular product
        dim3 g((n + 127) / 128, n);
        lower_clean_kernel<T><<<g, 128, 0, s>>>(n, w.L1);
        lower_clean_kernel<T><<<g, 128, 0, s>>>(n, w.L2);
        rmul_upper_kernel<T><<<g, 128, 0, s>>>(n, w.L2, w.L1, R, ldr);
        count_launch(3);
    }
    MAK_LAUNCH_CHECK(h, "cholqr2 tail");
    pt.mark("finish");
    pt.report("cholqr2");
    if (info_dev) MAK_CUDA(h, cudaMemcpyAsync(info_dev, w.info, sizeof(int), cudaMemcpyDeviceToDevice, s));
    return 0;
}

// ---------------------------------------------------------------------------------------
// multi-GPU TSQR (BASELINE config 4): qr_compact! of a row-sharded tall-skinny matrix
// ---------------------------------------------------------------------------------------
// Rank p holds A_p (m_loc x n).  Local: CholeskyQR2 up to - but not including - the last triangular solve:
//   A_p = Q1_p L1^H,  Q1_p^H Q1_p = L2 L2^H,  R_p = L2^H L1^H   (Q_p^loc = Q1_p L2^-H is never formed)
// Across ranks: binary tree over R_p.  Each round one n x n block travels (ncclSend/ncclRecv on the handle's
// stream, 512 KiB for n = 256 f64), the receiver takes the Householder QR of the stacked pair [R; R_b] and
// keeps the 2n x n orthogonal factor.  Down-sweep: the path product T_p (n x n) of those factors reaches
// every rank, and the ONE remaining tall GEMM applies the local solve and the tree together:
//   Q_p = Q1_p (L2^-H T_p)
// so a multi-GPU run does no tall-matrix work beyond the single-GPU algorithm (round 1 paid an extra
// m_loc x n x n product).  R (positive diagonal by construction of the tree QRs) is broadcast from rank 0.
constexpr int TSQR_MAX_LEVELS = 16;

template <typename T>
struct TsqrWork {
    CqrWork<T> c;
    T *Rcur, *Rb, *S, *Tcur, *Tb, *Tn, *Minv, *M, *Eye;
    T* Qs[TSQR_MAX_LEVELS];
    void* sub;
    size_t sub_bytes;
};

template <typename T, typename AR>
static void tsqr_carve(makb200_handle* h, AR& ar, int m, int n, int nranks, TsqrWork<T>* w) {
    cqr_carve<T>(h, ar, m, n, &w->c);
    size_t nn = (size_t)(n > 0 ? n : 1) * (size_t)(n > 0 ? n : 1);
    w->Rcur = ar.template get<T>(nn);
    w->Rb = ar.template get<T>(nn);
    w->S = ar.template get<T>(2 * nn);
    w->Tcur = ar.template get<T>(nn);
    w->Tb = ar.template get<T>(nn);
    w->Tn = ar.template get<T>(nn);
    w->Minv = ar.template get<T>(nn);
    w->M = ar.template get<T>(nn);
    w->Eye = ar.template get<T>(nn);
    int levels = 0;
    for (int s = 1; s < nranks; s *= 2) ++levels;
    for (int l = 0; l < TSQR_MAX_LEVELS; ++l) w->Qs[l] = l < levels ? ar.template get<T>(2 * nn) : nullptr;
    w->sub_bytes = nranks > 1 ? cholqr2_worksize_t<T>(h, 2 * n, n) : 0;
    w->sub = ar.template get<char>(w->sub_bytes);
}

__global__ void info_or_kernel(int* __restrict__ dst, const int* __restrict__ src) {
    if (src[0] != 0) dst[0] = src[0];
}

template <typename T>
size_t tsqr_worksize_t(makb200_handle* h, int m, int n, int nranks) {
    ArenaSize ar;
    TsqrWork<T> w;
    tsqr_carve<T>(h, ar, m, n, nranks, &w);
    return ar.off + 256;
}

// S (2n x n) = [Ra; Rb]
template <typename T>
__global__ void stack2_kernel(int n, const T* __restrict__ Ra, const T* __restrict__ Rb, T* __restrict__ S) {
    size_t total = (size_t)2 * n * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % (2 * n)), c = (int)(idx / (2 * n));
        S[idx] = r < n ? Ra[(size_t)c * n + r] : Rb[(size_t)c * n + (r - n)];
    }
}

#define MAK_NCCL(h, api, call)                                                               \
    do {                                                                                     \
        ncclResult_t _r = (call);                                                            \
        if (_r != ncclSuccess) {                                                             \
            snprintf(h->err, sizeof(h->err), "%s: %s", #call, api->GetErrorString(_r));      \
            return MAKB200_ERR_NCCL;                                                         \
        }                                                                                    \
    } while (0)

template <typename T>
int tsqr_t(makb200_handle* h, const NcclApi* api, ncclComm_t comm, int m, int n, T* A, int lda, T* Q, int ldq, T* R,
           int ldr, void* work, size_t lwork, int* info_dev) {
    int nranks = 1, rank = 0;
    if (comm) {
        MAK_NCCL(h, api, api->CommCount(comm, &nranks));
        MAK_NCCL(h, api, api->CommUserRank(comm, &rank));
    }
    if (n <= 0) return 0;
    cudaStream_t s = h->stream;
    Arena ar(work, lwork);
    TsqrWork<T> w;
    tsqr_carve<T>(h, ar, m, n, nranks, &w);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    MAK_CUDA(h, cudaMemsetAsync(w.c.info, 0, sizeof(int) * 4, s));
    PhaseTimer pt(s);
    pt.mark("start");
    const size_t nn = (size_t)n * n;
    const size_t cnt = nn * (is_cplx<T>::value ? 2 : 1);   // doubles per n x n block on the wire
    const dim3 g((n + 127) / 128, n);
    const int gnn = grid_for2(nn, h->num_sms);
    int rc;
    // ---- local: two Gram/Cholesky passes; the second solve is deferred ----
    if (m > 0) {
        rc = cholqr_pass<T>(h, m, n, A, lda, A, lda, w.c.L1, w.c);     // Q1 over A (block order makes it safe in place)
        if (rc) return rc;
        pt.mark("pass1");
        MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_C, MAKB200_OP_N, n, n, m, one<T>(), A, lda, A, lda, zero<T>(), w.c.G, n,
                  w.c.ws, w.c.ws_bytes, true);
        rc = potrf_blocked<T>(h, n, w.c.G, n, w.c.L2, n, w.c.Linv, w.c.info);
        if (rc) return rc;
        lower_clean_kernel<T><<<g, 128, 0, s>>>(n, w.c.L1);
        lower_clean_kernel<T><<<g, 128, 0, s>>>(n, w.c.L2);
        rmul_upper_kernel<T><<<g, 128, 0, s>>>(n, w.c.L2, w.c.L1, w.Rcur, n);
        count_launch(3);
        pt.mark("gram2");
    } else {
        MAK_CUDA(h, cudaMemsetAsync(w.Rcur, 0, nn * sizeof(T), s));   // a rank without rows contributes R = 0
    }
    if (nranks == 1) {
        for (int r0 = 0; r0 < m; r0 += CQR_SLAB) {
            int mr = (m - r0 < CQR_SLAB) ? (m - r0) : CQR_SLAB;
            rc = trsm_right<T>(h, true, mr, n, A + r0, lda, w.c.L2, n, w.c.Linv, Q + r0, ldq, w.c.Tmp);
            if (rc) return rc;
        }
        if (R && ldr > 0) {
            copy2d_kernel<T><<<gnn, 256, 0, s>>>(n, n, w.Rcur, n, R, ldr);
            count_launch();
        }
        MAK_LAUNCH_CHECK(h, "tsqr (1 rank)");
        pt.mark("solve");
        pt.report("tsqr");
        if (info_dev) MAK_CUDA(h, cudaMemcpyAsync(info_dev, w.c.info, sizeof(int), cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    // ---- up-sweep ----
    int nfac = 0, fac_partner[TSQR_MAX_LEVELS];
    int sent_to = -1;
    bool active = true;
    for (int stride = 1; stride < nranks; stride *= 2) {
        if (!active) continue;
        if (rank % (2 * stride) == 0) {
            int partner = rank + stride;
            if (partner < nranks) {
                MAK_NCCL(h, api, api->Recv(w.Rb, cnt, ncclDouble, partner, comm, s));
                stack2_kernel<T><<<grid_for2(2 * nn, h->num_sms), 256, 0, s>>>(n, w.Rcur, w.Rb, w.S);
                count_launch();
                // QR of the stacked pair: CholeskyQR2 again (kappa of the pair = kappa of the rows it stands for, the
                // assumption the local step already makes): a dozen small launches, ~0.15 ms, where the blocked Householder
                // QR of a 512 x 256 matrix took 1.7 ms per tree level - at 8 ranks that was 5 of the 6 ms over the ideal
                rc = cholqr2_t<T>(h, 2 * n, n, w.S, 2 * n, w.Qs[nfac], 2 * n, w.Rcur, n, w.sub, w.sub_bytes, w.c.info + 2, 0);
                if (rc) return rc;
                info_or_kernel<<<1, 1, 0, s>>>(w.c.info, w.c.info + 2);
                count_launch();
                fac_partner[nfac++] = partner;
            }
        } else {
            sent_to = rank - stride;
            MAK_NCCL(h, api, api->Send(w.Rcur, cnt, ncclDouble, sent_to, comm, s));
            active = false;
        }
    }
    pt.mark("up");
    // ---- down-sweep: T_p = product of the tree factors on the path root -> p ----
    T* Tc = w.Tcur;
    T* Tn = w.Tn;
    if (rank == 0) {
        eye_kernel<T><<<gnn, 256, 0, s>>>(n, Tc, n);
        count_launch();
    } else {
        MAK_NCCL(h, api, api->Recv(Tc, cnt, ncclDouble, sent_to, comm, s));
    }
    for (int f = nfac - 1; f >= 0; --f) {
        const T* Qs = w.Qs[f];
        MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_N, n, n, n, one<T>(), Qs + n, 2 * n, Tc, n, zero<T>(), w.Tb, n,
                  nullptr, 0);
        MAK_NCCL(h, api, api->Send(w.Tb, cnt, ncclDouble, fac_partner[f], comm, s));
        MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_N, n, n, n, one<T>(), Qs, 2 * n, Tc, n, zero<T>(), Tn, n,
                  nullptr, 0);
        T* t = Tc; Tc = Tn; Tn = t;
    }
    pt.mark("down");
    // ---- Q_p = Q1_p (L2^-H T_p): explicit L2^-H from the blocked solver applied to I, then two products ----
    if (m > 0) {
        eye_kernel<T><<<gnn, 256, 0, s>>>(n, w.Eye, n);
        count_launch();
        rc = trsm_right<T>(h, true, n, n, w.Eye, n, w.c.L2, n, w.c.Linv, w.Minv, n, w.c.Tmp);
        if (rc) return rc;
        MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_N, n, n, n, one<T>(), w.Minv, n, Tc, n, zero<T>(), w.M, n,
                  nullptr, 0);
        for (int r0 = 0; r0 < m; r0 += CQR_SLAB) {
            int mr = (m - r0 < CQR_SLAB) ? (m - r0) : CQR_SLAB;
            MAK_GEMM2(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_N, mr, n, n, one<T>(), A + r0, lda, w.M, n, zero<T>(),
                      Q + r0, ldq, nullptr, 0);
        }
    }
    pt.mark("apply");
    // ---- R to everybody ----
    MAK_NCCL(h, api, api->Broadcast(w.Rcur, w.Rcur, cnt, ncclDouble, 0, comm, s));
    if (R && ldr > 0) {
        copy2d_kernel<T><<<gnn, 256, 0, s>>>(n, n, w.Rcur, n, R, ldr);
        count_launch();
    }
    MAK_LAUNCH_CHECK(h, "tsqr tail");
    pt.mark("bcast");
    pt.report("tsqr");
    if (info_dev) MAK_CUDA(h, cudaMemcpyAsync(info_dev, w.c.info, sizeof(int), cudaMemcpyDeviceToDevice, s));
    return 0;
}
template size_t tsqr_worksize_t<double>(makb200_handle*, int, int, int);
template size_t tsqr_worksize_t<cplx>(makb200_handle*, int, int, int);
template int tsqr_t<double>(makb200_handle*, const NcclApi*, ncclComm_t, int, int, double*, int, double*, int, double*, int,
                            void*, size_t, int*);
template int tsqr_t<cplx>(makb200_handle*, const NcclApi*, ncclComm_t, int, int, cplx*, int, cplx*, int, cplx*, int, void*,
                          size_t, int*);

template <typename T>
int adjoint_t(makb200_handle* h, int m, int n, const T* A, int lda, T* B, int ldb) {
    if (m <= 0 || n <= 0) return 0;
    dim3 g((m + 31) / 32, (n + 31) / 32);
    adjoint_kernel<T><<<g, dim3(32, 8), 0, h->stream>>>(m, n, A, lda, B, ldb);
    count_launch();
    MAK_LAUNCH_CHECK(h, "adjoint_kernel");
    return 0;
}
template int adjoint_t<double>(makb200_handle*, int, int, const double*, int, double*, int);
template int adjoint_t<cplx>(makb200_handle*, int, int, const cplx*, int, cplx*, int);

#define INSTP(T)                                                                                             \
    template size_t polar_worksize_t<T>(makb200_handle*, int, int);                                          \
    template int polar_qdwh_t<T>(makb200_handle*, int, int, T*, int, T*, int, T*, int, double, int, void*,   \
                                 size_t, int*, int*);                                                        \
    template size_t svd_worksize_t<T>(makb200_handle*, int, int);                                            \
    template int svd_t<T>(makb200_handle*, int, int, T*, int, double*, T*, int, T*, int, int, double, void*, \
                          size_t, int*, int);                                                                     \
    template size_t cholqr2_worksize_t<T>(makb200_handle*, int, int);                                        \
    template int cholqr2_t<T>(makb200_handle*, int, int, T*, int, T*, int, T*, int, void*, size_t, int*, int);
INSTP(double)
INSTP(cplx)

}  // namespace mak
