// DMMA GEMM kernels (see gemm.cuh).
#include "gemm.cuh"

namespace mak {

unsigned long long g_launches = 0;
double g_gemm_flops = 0.0;  // real flops issued through gemm() since the last makb200_kernel_timing()
KernelClock g_clock_dots;
KernelClock g_clock_gemm;
KernelClock g_clock_w;

// ---------------------------------------------------------------------------------------
// tile configuration
// ---------------------------------------------------------------------------------------
// Config<T, V>: V = 0 default; V = 1 alternative (512 threads, 32x32 warp tiles) for f64
template <typename T, int V> struct Cfg;
template <> struct Cfg<double, 0> {
    static constexpr int BM = 128, BN = 128, BK = 16, WM = 64, WN = 32, STAGES = 4, THREADS = 256;
    // strides (in elements) chosen so that the 16 lanes of a half-warp hit 16 distinct
    // 8-byte banks for both fragment patterns: S == 4 (mod 8)
    static constexpr int SA_MN = BM + 4, SB_MN = BN + 4, S_K = BK + 4;
};
template <> struct Cfg<double, 1> {
    static constexpr int BM = 128, BN = 128, BK = 16, WM = 32, WN = 32, STAGES = 4, THREADS = 512;
    static constexpr int SA_MN = BM + 4, SB_MN = BN + 4, S_K = BK + 4;
};
template <> struct Cfg<cplx, 0> {
    static constexpr int BM = 64, BN = 128, BK = 8, WM = 32, WN = 32, STAGES = 4, THREADS = 256;
    // 16-byte elements: quarter-warp (8 lanes) must hit 8 distinct 16-byte banks:
    // [k][mn] layout needs S == 2 (mod 4), [mn][k] layout needs S == 4 (mod 8)
    static constexpr int SA_MN = BM + 2, SB_MN = BN + 2, S_K = BK + 4;
};
template <> struct Cfg<cplx, 1> : Cfg<cplx, 0> {};
// V = 2: small tiles, 128 threads, several resident CTAs per SM -- for the grouped launches of the
// batched paths (thousands of small problems with short K: latency is hidden by co-resident CTAs,
// not by a deep pipeline)
template <> struct Cfg<double, 2> {
    static constexpr int BM = 64, BN = 64, BK = 16, WM = 32, WN = 32, STAGES = 3, THREADS = 128, MINB = 4;
    static constexpr int SA_MN = BM + 4, SB_MN = BN + 4, S_K = BK + 4;
};
template <> struct Cfg<cplx, 2> {
    static constexpr int BM = 64, BN = 64, BK = 8, WM = 32, WN = 32, STAGES = 3, THREADS = 128, MINB = 3;
    static constexpr int SA_MN = BM + 2, SB_MN = BN + 2, S_K = BK + 4;
};
// V = 3: 128 x 64 tile, 256 threads, TWO resident CTAs per SM.  For the rank-k updates of the blocked factorizations
// (K = 64..256: her2k of hetrd, compact-WY trailing updates, Cholesky panel updates) the C tile is read and written
// once per 8..16 k-tiles, and with one resident CTA that epilogue is not overlapped with anybody's main loop: the
// 8192 x 8192 x 128 update ran at 15 TF/s against cuBLAS's 31 (profiles/r2_gemm_shapes.txt).
template <> struct Cfg<double, 3> {
    static constexpr int BM = 128, BN = 64, BK = 16, WM = 32, WN = 32, STAGES = 4, THREADS = 256, MINB = 2;
    static constexpr int SA_MN = BM + 4, SB_MN = BN + 4, S_K = BK + 4;
};
template <> struct Cfg<cplx, 3> : Cfg<cplx, 2> {};
template <typename C, typename = void> struct MinBlocks { static constexpr int value = 1; };
template <typename C> struct MinBlocks<C, decltype((void)C::MINB)> { static constexpr int value = C::MINB; };

template <typename T, typename C, bool TA> struct ATile {
    static constexpr int STRIDE = TA ? C::S_K : C::SA_MN;
    static constexpr int ROWS = TA ? C::BM : C::BK;
    static constexpr int CONTIG = TA ? C::BK : C::BM;
    static constexpr int ELEMS = STRIDE * ROWS;
};
template <typename T, typename C, bool TB> struct BTile {
    static constexpr int STRIDE = TB ? C::SB_MN : C::S_K;
    static constexpr int ROWS = TB ? C::BK : C::BN;
    static constexpr int CONTIG = TB ? C::BN : C::BK;
    static constexpr int ELEMS = STRIDE * ROWS;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* g, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Load a CONTIG x ROWS tile: element (c, r) <- g[(c0+c) + (r0+r)*ld] if in range else 0,
// stored at smem[r*STRIDE + c].
template <typename T, int GEMM_THREADS, int CONTIG, int ROWS, int STRIDE>
__device__ __forceinline__ void load_tile(T* smem, const T* __restrict__ g, int ld, int c0, int r0,
                                          int cmax, int rmax, bool vec16) {
    if constexpr (is_cplx<T>::value) {
        constexpr int TOTAL = CONTIG * ROWS;
#pragma unroll
        for (int i = 0; i < (TOTAL + GEMM_THREADS - 1) / GEMM_THREADS; ++i) {
            int e = threadIdx.x + i * GEMM_THREADS;
            if (TOTAL % GEMM_THREADS == 0 || e < TOTAL) {
                int r = e / CONTIG, c = e % CONTIG;
                bool ok = (c0 + c < cmax) && (r0 + r < rmax);
                const T* src = ok ? g + (size_t)(c0 + c) + (size_t)(r0 + r) * ld : g;
                cp_async16(smem + r * STRIDE + c, src, ok ? 16 : 0);
            }
        }
    } else {
        if (vec16) {
            constexpr int CH = CONTIG / 2, TOTAL = CH * ROWS;
#pragma unroll
            for (int i = 0; i < (TOTAL + GEMM_THREADS - 1) / GEMM_THREADS; ++i) {
                int e = threadIdx.x + i * GEMM_THREADS;
                if (TOTAL % GEMM_THREADS == 0 || e < TOTAL) {
                    int r = e / CH, c = (e % CH) * 2;
                    int valid = 0;
                    if (r0 + r < rmax) valid = min(2, max(0, cmax - (c0 + c)));
                    const T* src = valid ? g + (size_t)(c0 + c) + (size_t)(r0 + r) * ld : g;
                    cp_async16(smem + r * STRIDE + c, src, valid * 8);
                }
            }
        } else {
            constexpr int TOTAL = CONTIG * ROWS;
#pragma unroll
            for (int i = 0; i < (TOTAL + GEMM_THREADS - 1) / GEMM_THREADS; ++i) {
                int e = threadIdx.x + i * GEMM_THREADS;
                if (TOTAL % GEMM_THREADS == 0 || e < TOTAL) {
                    int r = e / CONTIG, c = e % CONTIG;
                    bool ok = (c0 + c < cmax) && (r0 + r < rmax);
                    const T* src = ok ? g + (size_t)(c0 + c) + (size_t)(r0 + r) * ld : g;
                    cp_async8(smem + r * STRIDE + c, src, ok ? 8 : 0);
                }
            }
        }
    }
}

// accumulator fragment for an 8x8 tile
template <typename T> struct Acc;
template <> struct Acc<double> { double c0, c1; };
template <> struct Acc<cplx> { double r0, r1, i0, i1; };

template <typename T, typename C, bool TA, bool TB>
__global__ void __launch_bounds__(C::THREADS, MinBlocks<C>::value)
gemm_kernel(const GemmProblem<T> p0, const GemmProblem<T>* __restrict__ plist, int splitk,
            T* __restrict__ ws) {
    using AT = ATile<T, C, TA>;
    using BT = BTile<T, C, TB>;
    constexpr int BM = C::BM, BN = C::BN, BK = C::BK, WM = C::WM, WN = C::WN, ST = C::STAGES;
    constexpr int MT = WM / 8, NT = WN / 8;
    constexpr int WARPS_M = BM / WM;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sA = reinterpret_cast<T*>(smem_raw);
    T* sB = sA + ST * AT::ELEMS;

    const GemmProblem<T> p = plist ? plist[blockIdx.z] : p0;
    const int M = p.m, N = p.n, K = p.k;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (M <= 0 || N <= 0 || m0 >= M || n0 >= N) return;
    if (p.lower && n0 >= m0 + BM) return;  // tile entirely above the diagonal

    int ktiles = K > 0 ? (K + BK - 1) / BK : 0;
    int kt_beg = 0, kt_end = ktiles;
    if (!plist && splitk > 1) {
        int per = (ktiles + splitk - 1) / splitk;
        kt_beg = blockIdx.z * per;
        kt_end = min(ktiles, kt_beg + per);
        if (kt_end < kt_beg) kt_end = kt_beg;
    }
    const int nkt = kt_end - kt_beg;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = (warp % WARPS_M) * WM, wn = (warp / WARPS_M) * WN;
    const int lr = lane >> 2, lc = lane & 3;

    bool vecA = false, vecB = false;
    if constexpr (!is_cplx<T>::value) {
        vecA = ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0) && ((p.lda & 1) == 0);
        vecB = ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0) && ((p.ldb & 1) == 0);
    }

    auto load_stage = [&](int stage, int kt) {
        const int k0 = kt * BK;
        T* a = sA + stage * AT::ELEMS;
        T* b = sB + stage * BT::ELEMS;
        if constexpr (TA) load_tile<T, C::THREADS, AT::CONTIG, AT::ROWS, AT::STRIDE>(a, p.A, p.lda, k0, m0, K, M, vecA);
        else load_tile<T, C::THREADS, AT::CONTIG, AT::ROWS, AT::STRIDE>(a, p.A, p.lda, m0, k0, M, K, vecA);
        if constexpr (TB) load_tile<T, C::THREADS, BT::CONTIG, BT::ROWS, BT::STRIDE>(b, p.B, p.ldb, n0, k0, N, K, vecB);
        else load_tile<T, C::THREADS, BT::CONTIG, BT::ROWS, BT::STRIDE>(b, p.B, p.ldb, k0, n0, K, N, vecB);
    };

    Acc<T> acc[MT][NT];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j] = Acc<T>{};

#pragma unroll
    for (int s = 0; s < ST - 1; ++s) {
        if (s < nkt) load_stage(s, kt_beg + s);
        cp_async_commit();
    }

    const double sgnA = p.conja ? -1.0 : 1.0, sgnB = p.conjb ? -1.0 : 1.0;

    for (int it = 0; it < nkt; ++it) {
        cp_async_wait<ST - 2>();
        __syncthreads();
        {
            int nx = it + ST - 1;
            if (nx < nkt) load_stage(nx % ST, kt_beg + nx);
            cp_async_commit();
        }
        const T* a = sA + (it % ST) * AT::ELEMS;
        const T* b = sB + (it % ST) * BT::ELEMS;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            T af[MT], bf[NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                int mm = wm + i * 8 + lr, k = kk + lc;
                af[i] = TA ? a[mm * AT::STRIDE + k] : a[k * AT::STRIDE + mm];
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                int nn = wn + j * 8 + lr, k = kk + lc;
                bf[j] = TB ? b[k * BT::STRIDE + nn] : b[nn * BT::STRIDE + k];
            }
            if constexpr (is_cplx<T>::value) {
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    double ar = af[i].re, ai = af[i].im * sgnA, nai = -ai;
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        double br = bf[j].re, bi = bf[j].im * sgnB;
                        dmma(acc[i][j].r0, acc[i][j].r1, ar, br);
                        dmma(acc[i][j].r0, acc[i][j].r1, nai, bi);
                        dmma(acc[i][j].i0, acc[i][j].i1, ar, bi);
                        dmma(acc[i][j].i0, acc[i][j].i1, ai, br);
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma(acc[i][j].c0, acc[i][j].c1, af[i], bf[j]);
            }
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds C[row = lr][cols = 2*lc, 2*lc+1] of each 8x8 tile.  beta != 0: the C values of one row
    // block are all loaded before the first store - C may alias nothing the compiler can prove, so a fused
    // load/modify/store loop serialises into NT*2 L2 round trips per row block (the rank-k updates with K = 128 ran at
    // 15 TF/s because of it, profiles/r2_gemm_shapes.txt)
    const bool partial = (!plist && splitk > 1);
    const bool beta0 = is_zero(p.beta);
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        int r = m0 + wm + i * 8 + lr;
        if (r >= M) continue;
        if (partial) {
#pragma unroll
            for (int j = 0; j < NT; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    int c = n0 + wn + j * 8 + lc * 2 + e;
                    if (c >= N) continue;
                    T v;
                    if constexpr (is_cplx<T>::value) v = cplx{e ? acc[i][j].r1 : acc[i][j].r0, e ? acc[i][j].i1 : acc[i][j].i0};
                    else v = e ? acc[i][j].c1 : acc[i][j].c0;
                    ws[(size_t)blockIdx.z * M * N + (size_t)c * M + r] = v;
                }
            }
            continue;
        }
        T old[NT][2];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int c = n0 + wn + j * 8 + lc * 2 + e;
                old[j][e] = (!beta0 && c < N) ? p.C[(size_t)c * p.ldc + r] : zero<T>();
            }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                int c = n0 + wn + j * 8 + lc * 2 + e;
                if (c >= N) continue;
                T v;
                if constexpr (is_cplx<T>::value) v = cplx{e ? acc[i][j].r1 : acc[i][j].r0, e ? acc[i][j].i1 : acc[i][j].i0};
                else v = e ? acc[i][j].c1 : acc[i][j].c0;
                T out = mul_(p.alpha, v);
                if (!beta0) out = add_(out, mul_(p.beta, old[j][e]));
                p.C[(size_t)c * p.ldc + r] = out;
            }
        }
    }
}

template <typename T>
__global__ void splitk_reduce_kernel(int M, int N, int splitk, const T* __restrict__ ws, T alpha, T beta,
                                     T* __restrict__ Cm, int ldc) {
    size_t total = (size_t)M * N;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % M), c = (int)(idx / M);
        T s = zero<T>();
        for (int z = 0; z < splitk; ++z) s = add_(s, ws[(size_t)z * total + idx]);
        T* dst = Cm + (size_t)c * ldc + r;
        T out = mul_(alpha, s);
        if (!is_zero(beta)) out = add_(out, mul_(beta, *dst));
        *dst = out;
    }
}

// C = beta*C for degenerate k == 0
template <typename T>
__global__ void scale_c_kernel(int M, int N, T beta, T* __restrict__ Cm, int ldc) {
    size_t total = (size_t)M * N;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % M), c = (int)(idx / M);
        T* dst = Cm + (size_t)c * ldc + r;
        *dst = is_zero(beta) ? zero<T>() : mul_(beta, *dst);
    }
}

}  // namespace mak
#include "gemm_tma.cuh"
namespace mak {

static int gemm_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MAKB200_GEMM_VARIANT"); v = e ? atoi(e) : 1; if (v != 0 && v != 3) v = 1; }
    return v;
}
// K at or below which the two-CTA-per-SM configuration is used (0 disables)
static int gemm_shortk() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MAKB200_GEMM_SHORTK"); v = e ? atoi(e) : 0; if (v < 0) v = 0; }
    return v;
}

template <typename T, typename C, bool TA, bool TB>
static cudaError_t launch(cudaStream_t stream, dim3 grid, const GemmProblem<T>& p,
                          const GemmProblem<T>* plist, int splitk, T* ws) {
    constexpr size_t smem = (size_t)C::STAGES * (ATile<T, C, TA>::ELEMS + BTile<T, C, TB>::ELEMS) * sizeof(T);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_kernel<T, C, TA, TB>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    g_clock_gemm.begin(stream);
    gemm_kernel<T, C, TA, TB><<<grid, C::THREADS, smem, stream>>>(p, plist, splitk, ws);
    g_clock_gemm.end(stream);
    count_launch();
    return cudaGetLastError();
}

template <typename T, typename C>
static cudaError_t dispatch2(cudaStream_t stream, bool ta, bool tb, dim3 grid, const GemmProblem<T>& p,
                             const GemmProblem<T>* plist, int splitk, T* ws) {
    if (!ta && !tb) return launch<T, C, false, false>(stream, grid, p, plist, splitk, ws);
    if (ta && !tb) return launch<T, C, true, false>(stream, grid, p, plist, splitk, ws);
    if (!ta && tb) return launch<T, C, false, true>(stream, grid, p, plist, splitk, ws);
    return launch<T, C, true, true>(stream, grid, p, plist, splitk, ws);
}

template <typename T>
static cudaError_t dispatch(cudaStream_t stream, bool ta, bool tb, dim3 grid, const GemmProblem<T>& p,
                            const GemmProblem<T>* plist, int splitk, T* ws) {
    if constexpr (!is_cplx<T>::value) {
        if (gemm_variant() == 1) return dispatch2<T, Cfg<T, 1>>(stream, ta, tb, grid, p, plist, splitk, ws);
    }
    return dispatch2<T, Cfg<T, 0>>(stream, ta, tb, grid, p, plist, splitk, ws);
}

template <typename T>
cudaError_t gemm(cudaStream_t stream, int num_sms, int opa, int opb, int m, int n, int k, T alpha,
                 const T* A, int lda, const T* B, int ldb, T beta, T* Cm, int ldc, void* ws,
                 size_t ws_bytes, bool lower) {
    using C = Cfg<T, 0>;  // both variants share the CTA tile
    if (m <= 0 || n <= 0) return cudaSuccess;
    if (k > 0 && !is_zero(alpha) && (gemm_variant() == 3 || (k <= gemm_shortk() && (size_t)m * n >= (size_t)1 << 20))) {
        // rank-k update on a large C: two resident CTAs per SM so that one CTA's C read/write overlaps the other's
        // main loop; no split-K (the output grid fills the machine)
        using C3 = Cfg<T, 3>;
        GemmProblem<T> p3;
        p3.m = m; p3.n = n; p3.k = k;
        p3.A = A; p3.lda = lda; p3.B = B; p3.ldb = ldb; p3.C = Cm; p3.ldc = ldc;
        p3.alpha = alpha; p3.beta = beta;
        p3.conja = (opa == MAKB200_OP_C); p3.conjb = (opb == MAKB200_OP_C);
        p3.lower = lower ? 1 : 0;
        dim3 g3((m + C3::BM - 1) / C3::BM, (n + C3::BN - 1) / C3::BN, 1);
        if (g_clock_gemm.on) {
            double mn = (double)m * (double)n;
            if (lower) {
                mn = 0.0;
                for (int bx = 0; bx < (int)g3.x; ++bx) {
                    int rows = min(C3::BM, m - bx * C3::BM);
                    int ncols = min(n, ((bx * C3::BM + C3::BM + C3::BN - 1) / C3::BN) * C3::BN);
                    mn += (double)rows * (double)ncols;
                }
            }
            const double fl = 2.0 * mn * (double)k * (is_cplx<T>::value ? 4.0 : 1.0);
            g_gemm_flops += fl;
            g_clock_gemm.tag(m, n, k, (opa != MAKB200_OP_N ? 1 : 0) | (opb != MAKB200_OP_N ? 2 : 0) | (lower ? 4 : 0) | 8, fl);
        }
        return dispatch2<T, C3>(stream, opa != MAKB200_OP_N, opb != MAKB200_OP_N, g3, p3, nullptr, 1, nullptr);
    }
    if (k <= 0 || is_zero(alpha)) {
        if (is_one(beta)) return cudaSuccess;
        scale_c_kernel<T><<<min(1024, (int)(((size_t)m * n + 255) / 256)), 256, 0, stream>>>(m, n, beta, Cm, ldc);
        return cudaGetLastError();
    }
    GemmProblem<T> p;
    p.m = m; p.n = n; p.k = k;
    p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = Cm; p.ldc = ldc;
    p.alpha = alpha; p.beta = beta;
    p.conja = (opa == MAKB200_OP_C); p.conjb = (opb == MAKB200_OP_C);
    p.lower = lower ? 1 : 0;
    dim3 grid((m + C::BM - 1) / C::BM, (n + C::BN - 1) / C::BN, 1);
    // split-K when the output grid cannot fill the machine and K is long
    int splitk = 1;
    int tiles = grid.x * grid.y, ktiles = (k + C::BK - 1) / C::BK;
    if (lower) {
        // only the tiles that touch the lower triangle run: split K by THAT count (the Gram matrices of the tall-skinny
        // QR are 2 x 2 tiles of which 3 run: 3 x 49 CTAs fill 147 of 148 SMs, 4 x 32 left 52 idle)
        tiles = 0;
        for (int bx = 0; bx < (int)grid.x; ++bx) tiles += min((int)grid.y, (bx * C::BM + C::BM + C::BN - 1) / C::BN);
    }
    if (ws && tiles * 2 <= num_sms && ktiles >= 16) {
        splitk = min(min(num_sms / tiles, ktiles / 8), 64);
        size_t need = (size_t)splitk * m * n * sizeof(T);
        while (splitk > 1 && need > ws_bytes) { --splitk; need = (size_t)splitk * m * n * sizeof(T); }
        if (splitk < 2) splitk = 1;
    }
    grid.z = splitk;
    if (g_clock_gemm.on) {
        // flops of the tiles that actually run: a `lower` launch skips every tile entirely above the diagonal
        double mn = (double)m * (double)n;
        if (lower) {
            mn = 0.0;
            for (int bx = 0; bx < (int)grid.x; ++bx) {
                int rows = min(C::BM, m - bx * C::BM);
                int ncols = min(n, ((bx * C::BM + C::BM + C::BN - 1) / C::BN) * C::BN);
                mn += (double)rows * (double)ncols;
            }
        }
        const double fl = 2.0 * mn * (double)k * (is_cplx<T>::value ? 4.0 : 1.0);
        g_gemm_flops += fl;
        // flags: bit0 opa != N, bit1 opb != N, bit2 lower, bits 4.. split-K factor
        g_clock_gemm.tag(m, n, k, (opa != MAKB200_OP_N ? 1 : 0) | (opb != MAKB200_OP_N ? 2 : 0) | (lower ? 4 : 0) | (splitk << 4), fl);
    }
    cudaError_t e = cudaSuccess;
    bool done = false;
    // TMA-fed warp-specialised kernel when both operands can be described by a tensor map (16-byte aligned base
    // and leading dimension); otherwise (Float64 with an odd leading dimension or a view starting at an odd row) cp.async
    done = gemm_tma_try(stream, opa != MAKB200_OP_N, opb != MAKB200_OP_N, grid, p, splitk, (T*)ws, &e);
    if (!done) e = dispatch<T>(stream, opa != MAKB200_OP_N, opb != MAKB200_OP_N, grid, p, nullptr, splitk, (T*)ws);
    if (e != cudaSuccess) return e;
    if (splitk > 1) {
        size_t total = (size_t)m * n;
        int blocks = (int)min((size_t)num_sms * 8, (total + 255) / 256);
        splitk_reduce_kernel<T><<<blocks, 256, 0, stream>>>(m, n, splitk, (const T*)ws, alpha, beta, Cm, ldc);
        count_launch();
        e = cudaGetLastError();
    }
    return e;
}

template <typename T>
cudaError_t gemm_grouped(cudaStream_t stream, int opa, int opb, int count, int max_m, int max_n,
                         const GemmProblem<T>* problems_dev) {
    if (count <= 0 || max_m <= 0 || max_n <= 0) return cudaSuccess;
    GemmProblem<T> dummy{};
    static const bool small = []() { const char* e = getenv("MAKB200_GROUPED_SMALL"); return !(e && e[0] == '0'); }();
    if (small) {
        using C = Cfg<T, 2>;
        dim3 grid((max_m + C::BM - 1) / C::BM, (max_n + C::BN - 1) / C::BN, count);
        return dispatch2<T, C>(stream, opa != MAKB200_OP_N, opb != MAKB200_OP_N, grid, dummy, problems_dev, 1, nullptr);
    }
    using C = Cfg<T, 0>;
    dim3 grid((max_m + C::BM - 1) / C::BM, (max_n + C::BN - 1) / C::BN, count);
    return dispatch<T>(stream, opa != MAKB200_OP_N, opb != MAKB200_OP_N, grid, dummy, problems_dev, 1, nullptr);
}

template cudaError_t gemm<double>(cudaStream_t, int, int, int, int, int, int, double, const double*, int,
                                  const double*, int, double, double*, int, void*, size_t, bool);
template cudaError_t gemm<cplx>(cudaStream_t, int, int, int, int, int, int, cplx, const cplx*, int, const cplx*,
                                int, cplx, cplx*, int, void*, size_t, bool);
template cudaError_t gemm_grouped<double>(cudaStream_t, int, int, int, int, int, const GemmProblem<double>*);
template cudaError_t gemm_grouped<cplx>(cudaStream_t, int, int, int, int, int, const GemmProblem<cplx>*);

}  // namespace mak
