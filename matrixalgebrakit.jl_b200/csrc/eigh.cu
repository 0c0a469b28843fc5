// Hermitian eigensolver: eigh_full!  (replaces heevd!/heevr!, yalapack.jl:1164-1362,
// yacusolver.jl:766-810).
//   1. mirror the upper triangle (LAPACK is called with uplo='U': only it is read)
//   2. hetrd: blocked Householder tridiagonalisation; per column two kernels
//        (a) all column dot products of the step in ONE pass over the trailing matrix
//            (y = A22 v, W^H v, V^H v, y^H v) — HBM-bound, the dominant kernel;
//        (b) w, write-back of v, left-looking update of the next column and its norm;
//      per panel one DMMA GEMM  A22 -= [V W][W V]^H  (her2k as a single K = 2*nb product)
//   3. stedc (stedc.cu): tridiagonal divide and conquer, GEMM-rich merges
//   4. back-transform V = Q Z with compact-WY block reflectors (DMMA GEMMs, qr.cu)
//   5. eigenvector gauge (common/gauge.jl:38-45) fused into one launch
#include <cuda.h>
#include <cudaTypedefs.h>
#include "eigh.cuh"
#include "gauge.cuh"
#include "sbr.cuh"
#include "gemm.cuh"
#include "qr.cuh"
#include "stedc.cuh"
#include "sturm_core.h"
#include "projections.cuh"
#include "bhetrd.cuh"

namespace mak {

constexpr int TRD_NB = 64;

// ---------------------------------------------------------------------------------------
// Hermitian defect and mirroring
// ---------------------------------------------------------------------------------------
// out[0] += sum |(A - A^H)/2|^2 ; out[1] = max |A_ij|  (atomics: only used for a threshold test)
template <typename T>
__global__ void herm_defect_kernel(int n, const T* __restrict__ A, int lda, double* out) {
    __shared__ T tile[32][33];
    __shared__ double red[32];
    const int bi = blockIdx.x, bj = blockIdx.y;
    if (bj > bi) return;  // lower block pairs only
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    // tile (bj, bi): rows bj*32.., cols bi*32..  -> stored transposed-access friendly
    for (int k = ty; k < 32; k += 8) {
        int r = bj * 32 + tx, c = bi * 32 + k;
        tile[k][tx] = (r < n && c < n) ? A[(size_t)c * lda + r] : zero<T>();
    }
    __syncthreads();
    double part = 0.0, mx = 0.0;
    for (int k = ty; k < 32; k += 8) {
        int r = bi * 32 + tx, c = bj * 32 + k;  // element (r,c) of tile (bi,bj); partner (c,r) = tile[tx][k]
        if (r < n && c < n) {
            T a = A[(size_t)c * lda + r];
            T b = conj_(tile[tx][k]);
            T dlt = sub_(a, b);
            double w = (bi == bj) ? 1.0 : 2.0;  // off-diagonal block pairs count twice
            part += w * 0.25 * abs2_(dlt);
            mx = fmax(mx, fmax(sqrt(abs2_(a)), sqrt(abs2_(b))));
        }
    }
    const int t = ty * 32 + tx;
    part = warp_sum(part);
    mx = warp_max(mx);
    if ((t & 31) == 0) red[t >> 5] = part;
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[i];
        atomicAdd(out, s);
    }
    __syncthreads();
    if ((t & 31) == 0) red[t >> 5] = mx;
    __syncthreads();
    if (t == 0) {
        double m2 = 0.0;
        for (int i = 0; i < 8; ++i) m2 = fmax(m2, red[i]);
        // non-negative doubles order like their bit patterns
        atomicMax((unsigned long long*)(out + 1), (unsigned long long)__double_as_longlong(m2));
    }
}

// lower <- conj(upper); diagonal made real
template <typename T>
__global__ void mirror_upper_kernel(int n, T* __restrict__ A, int lda) {
    __shared__ T tile[32][33];
    const int bi = blockIdx.x, bj = blockIdx.y;  // upper block (bi <= bj): rows bi*32, cols bj*32
    if (bi > bj) return;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int k = ty; k < 32; k += 8) {
        int r = bi * 32 + tx, c = bj * 32 + k;
        tile[k][tx] = (r < n && c < n) ? A[(size_t)c * lda + r] : zero<T>();
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        int r = bj * 32 + tx, c = bi * 32 + k;  // lower element (r,c) = conj(upper (c,r)) = conj(tile[tx][k])
        if (r < n && c < n) {
            if (r > c) A[(size_t)c * lda + r] = conj_(tile[tx][k]);
            else if (r == c) A[(size_t)c * lda + r] = mk<T>(real_(tile[tx][k]));
        }
    }
}

// ---------------------------------------------------------------------------------------
// hetrd
// ---------------------------------------------------------------------------------------
template <typename T>
struct TrdCtx {
    int n;
    T* A; int lda;
    T* P; int ldp;   // panel [V | W | V], each pw columns wide
    int pw;
    T* y;            // [nstrips][maxseg][CW] : column parts of y = A22 v per (strip, row segment)
    int maxseg;
    int seg;         // rows per segment
    int prefetch;    // 1: symv CTAs prefetch their strip into L2 before waiting on the previous kernel (PDL)
    T* ypart; int ldy;  // [nstrips][n] : row parts of y, one slice per column strip
    T* t;            // 2*TRD_NB : t1 = W^H v, t2 = V^H v
    double* pn;      // partial tail norms of the current column
    T* pyv;          // partial y^H v
    T* tau; double* d; double* e;
    // persistent TMA column kernel (trd2.cuh): fixed-slot partial buffers, panel-dot partials, step scalars, ticket
    T* yrowp; T* ycolp; T* tpart; int tpld;
    int v2;          // 1: trd_symv2_kernel / trd_w2_kernel
};

constexpr int TRD_K1_THREADS = 256;   // 8 warps x 4 columns
constexpr int TRD_K1_COLS = 32;
constexpr int TRD_K2_ROWS = 32;       // rows per CTA of trd_w_kernel (64 left the GPU under-filled: mt/64 < 148 CTAs)
constexpr int TRD_K2_NG = 256 / TRD_K2_ROWS;   // threads (p-groups) per row

// Programmatic dependent launch (default on, MAKB200_PDL=0 disables): the two kernels of a column are launched with
// programmatic stream serialization, so the CTAs of the next launch are scheduled while the previous
// grid drains; every kernel first waits for its prerequisite grid to complete and flush (so no read
// or write is reordered across the dependency), then lets its own dependent launch.  Without the
// launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait_then_trigger() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
static bool trd_pdl() {
    // measured on B200 (round 1, eigh 8192 f64): hetrd 406 -> 379 ms; MAKB200_PDL=0 disables
    static const bool v = []() { const char* e = getenv("MAKB200_PDL"); return !(e && e[0] == '0'); }();
    return v;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// partial tail norms of column c (rows >= c+2), used at panel starts
template <typename T>
__global__ void trd_colnorm_kernel(TrdCtx<T> x, int c) {
    __shared__ double red[32];
    const int r = c + 1 + blockIdx.x * TRD_K2_ROWS + (threadIdx.x % TRD_K2_ROWS);
    double part = 0.0;
    if (threadIdx.x < TRD_K2_ROWS && r < x.n && r >= c + 2) part = abs2_(x.A[(size_t)c * x.lda + r]);
    double tot = block_sum<double>(part, red);
    if (threadIdx.x == 0) x.pn[blockIdx.x] = tot;
}

template <typename T>
__global__ void __launch_bounds__(TRD_K1_THREADS)
trd_dots_kernel(TrdCtx<T> x, int c, int i, int npn) {
    __shared__ T redT[32];
    const int n = x.n, row0 = c + 1, mt = n - row0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // reflector scalars (every thread, identical arithmetic)
    double sigma = 0.0;
    for (int q = 0; q < npn; ++q) sigma += x.pn[q];
    const T* acol = x.A + (size_t)c * x.lda + row0;
    const T alpha = acol[0];
    double beta; T tau, scale;
    larfgp_scalars<T>(alpha, sigma, beta, tau, scale);

    const int ncols = mt + 2 * i;
    const int q0 = (blockIdx.x * 8 + warp) * 4;
    const T* col[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int q = q0 + k;
        if (q < mt) col[k] = x.A + (size_t)(row0 + q) * x.lda + row0;
        else if (q < mt + i) col[k] = x.P + (size_t)(x.pw + (q - mt)) * x.ldp + row0;
        else if (q < ncols) col[k] = x.P + (size_t)(q - mt - i) * x.ldp + row0;
        else col[k] = nullptr;
    }
    T acc[4] = {zero<T>(), zero<T>(), zero<T>(), zero<T>()};
    if (q0 < ncols) {
        for (int r = lane; r < mt; r += 32) {
            T vr = (r == 0) ? one<T>() : mul_(acol[r], scale);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (col[k]) fmac_(acc[k], col[k][r], vr);
        }
    }
    T pyv = zero<T>();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        acc[k] = warp_sum(acc[k]);
        int q = q0 + k;
        if (lane == 0 && q < ncols) {
            if (q < mt) {
                x.y[row0 + q] = acc[k];
                T vq = (q == 0) ? one<T>() : mul_(acol[q], scale);
                fmac_(pyv, acc[k], vq);  // conj(y_q) * v_q
                x.P[(size_t)i * x.ldp + row0 + q] = vq;                  // V(:, i)
                x.P[(size_t)(2 * x.pw + i) * x.ldp + row0 + q] = vq;     // second copy
            } else if (q < mt + i) {
                x.t[q - mt] = acc[k];                 // t1 = W^H v
            } else {
                x.t[TRD_NB + (q - mt - i)] = acc[k];  // t2 = V^H v
            }
        }
    }
    // CTA partial of y^H v (lane 0 of each warp holds its share)
    __syncthreads();
    if (lane == 0) redT[warp] = pyv;
    __syncthreads();
    if (threadIdx.x == 0) {
        T s = zero<T>();
        for (int w = 0; w < TRD_K1_THREADS / 32; ++w) s = add_(s, redT[w]);
        x.pyv[blockIdx.x] = s;
        if (blockIdx.x == 0) {
            x.tau[c] = tau;
            x.e[c] = beta;
            x.d[c] = real_(x.A[(size_t)c * x.lda + c]);
        }
    }
}

// Symmetric single-pass kernel: reads only the LOWER triangle of the trailing matrix (half the
// HBM traffic of a full gemv).  CTA b owns the column strip [cb, cb+CW) and streams it once:
// lane = row, registers hold one partial dot product per strip column (column part, reduced at
// the end) while the row part of the strip goes to ypart[b][r] (summed over strips by trd_w_kernel).
// v^H A22 v (needed for the rank-2 correction) is accumulated on the fly.  CTAs beyond the strips
// do the panel dot products W^H v, V^H v.
constexpr int TRD_SEG_MIN = 512;  // smallest row segment (sizes the partial buffers)
// rows per CTA of a column strip (balances the triangular work); env override for tuning
static int trd_seg() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MAKB200_TRD_SEG");
        v = e ? atoi(e) : 1024;
        if (v < TRD_SEG_MIN) v = TRD_SEG_MIN;
        v = (v + 31) / 32 * 32;
    }
    return v;
}
template <typename T> struct SymvCW { static constexpr int value = 16; };
template <> struct SymvCW<cplx> { static constexpr int value = 8; };

template <typename T>
__global__ void __launch_bounds__(256, 2)
trd_symv_kernel(TrdCtx<T> x, int c, int i, int npn, int nstrips) {
    constexpr int CW = SymvCW<T>::value;
    const int SEG = x.seg;
    const int sgi = blockIdx.y;
    // Under programmatic dependent launch this CTA may be resident while the previous kernel (the w
    // kernel of column c-1) is still running.  The trailing matrix it is about to stream is not touched
    // by that kernel, so pull this CTA's part of the strip into L2 before waiting on the dependency:
    // the first wave of the launch (296 CTAs x 128 KB = 39 MB, fits L2) then starts from L2 hits.
    // prefetch has no architectural effect, so this is safe even when the prerequisite is the panel GEMM.
    if (x.prefetch && (int)blockIdx.x < nstrips) {
        const int row0p = c + 1, mtp = x.n - row0p, cbp = blockIdx.x * CW;
        const int cwp = (mtp - cbp < CW) ? (mtp - cbp) : CW;
        const int rsp = cbp + sgi * SEG;
        const int rep = (rsp + SEG < mtp) ? (rsp + SEG) : mtp;
        constexpr int PER_LINE = 128 / (int)sizeof(T);
        const int ln = threadIdx.x & 31, wp = threadIdx.x >> 5;
        if ((ln % PER_LINE) == 0) {
            const T* Abp = x.A + (size_t)(row0p + cbp) * x.lda + row0p;
            for (int rt = rsp + 32 * wp; rt < rep; rt += 256) {
                const int r = rt + ln;
                if (r < rep) {
#pragma unroll
                    for (int k = 0; k < CW; ++k)
                        if (k < cwp) asm volatile("prefetch.global.L2 [%0];" ::"l"(Abp + (size_t)k * x.lda + r));
                }
            }
        }
    }
    pdl_wait_then_trigger();
    const int slot = sgi * gridDim.x + blockIdx.x;
    __shared__ T s_vcw[8][CW];   // per-warp copy of v over the strip columns (no block barrier needed)
    __shared__ T s_col[8][CW];
    __shared__ double s_q[8];
    const int n = x.n, row0 = c + 1, mt = n - row0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tail norm of the column: every warp sums the partials itself in the same fixed order (identical
    // result in every warp of every CTA, and no block barrier before the first load is issued)
    double sigma = 0.0;
    for (int q = lane; q < npn; q += 32) sigma += x.pn[q];
    sigma = warp_sum(sigma);
    const T* acol = x.A + (size_t)c * x.lda + row0;
    const T alpha = acol[0];
    double beta; T tau, scale;
    larfgp_scalars<T>(alpha, sigma, beta, tau, scale);

    if ((int)blockIdx.x >= nstrips) {
        if (sgi > 0) { if (tid == 0) x.pyv[slot] = zero<T>(); return; }
        // ---- panel columns: t1 = W^H v, t2 = V^H v ----
        const int q0 = (((int)blockIdx.x - nstrips) * 8 + warp) * 4;
        const T* col[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int q = q0 + k;
            if (q < i) col[k] = x.P + (size_t)(x.pw + q) * x.ldp + row0;
            else if (q < 2 * i) col[k] = x.P + (size_t)(q - i) * x.ldp + row0;
            else col[k] = nullptr;
        }
        T acc[4] = {zero<T>(), zero<T>(), zero<T>(), zero<T>()};
        if (q0 < 2 * i) {
            for (int r = lane; r < mt; r += 32) {
                T vr = (r == 0) ? one<T>() : mul_(acol[r], scale);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (col[k]) fmac_(acc[k], col[k][r], vr);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc[k] = warp_sum(acc[k]);
            int q = q0 + k;
            if (lane == 0 && q < 2 * i) {
                if (q < i) x.t[q] = acc[k];
                else x.t[TRD_NB + (q - i)] = acc[k];
            }
        }
        if (tid == 0) x.pyv[slot] = zero<T>();
        return;
    }

    const int cb = blockIdx.x * CW;
    const int cw = (mt - cb < CW) ? (mt - cb) : CW;
    const int rs = cb + sgi * SEG;                         // this CTA's row segment of the strip
    if (rs >= mt) { if (tid == 0) x.pyv[slot] = zero<T>(); return; }
    const int re = (rs + SEG < mt) ? (rs + SEG) : mt;
    T* s_vc = s_vcw[warp];
    if (lane < CW) {
        T vq = zero<T>();
        if (lane < cw) {
            int r = cb + lane;
            vq = (r == 0) ? one<T>() : mul_(acol[r], scale);
            if (sgi == 0 && warp == 0) {
                x.P[(size_t)i * x.ldp + row0 + r] = vq;                  // V(:, i)
                x.P[(size_t)(2 * x.pw + i) * x.ldp + row0 + r] = vq;     // second copy
            }
        }
        s_vc[lane] = vq;
    }
    __syncwarp();
    T colacc[CW];
#pragma unroll
    for (int k = 0; k < CW; ++k) colacc[k] = zero<T>();
    double qacc = 0.0;
    const T* Ab = x.A + (size_t)(row0 + cb) * x.lda + row0;  // A22(:, cb), local row index
    T* Yp = x.ypart + (size_t)blockIdx.x * x.ldy + row0;
    for (int rt = rs + 32 * warp; rt < re; rt += 256) {
        const int r = rt + lane;
        if (r >= re) continue;
        const T vr = (r == 0) ? one<T>() : mul_(acol[r], scale);
        T rowoff = zero<T>();
        double ad = 0.0;
        if (rt >= cb + CW && cw == CW) {
            // strictly below the diagonal block of a full strip: batches of CW/2 independent loads
            // per lane (register budget for two resident CTAs per SM), then two FMAs per element
            constexpr int HB = CW;
#pragma unroll
            for (int hb = 0; hb < 1; ++hb) {
                T av[HB];
#pragma unroll
                for (int k = 0; k < HB; ++k) av[k] = Ab[(size_t)(hb * HB + k) * x.lda + r];
#pragma unroll
                for (int k = 0; k < HB; ++k) {
                    fma_(rowoff, av[k], s_vc[hb * HB + k]);
                    fmac_(colacc[hb * HB + k], av[k], vr);
                }
            }
        } else if (rt >= cb + CW) {
#pragma unroll
            for (int k = 0; k < CW; ++k) {
                if (k < cw) {
                    T a = Ab[(size_t)k * x.lda + r];
                    fma_(rowoff, a, s_vc[k]);
                    fmac_(colacc[k], a, vr);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < CW; ++k) {
                const int cc = cb + k;
                if (k < cw) {
                    if (r > cc) {
                        T a = Ab[(size_t)k * x.lda + r];
                        fma_(rowoff, a, s_vc[k]);
                        fmac_(colacc[k], a, vr);
                    } else if (r == cc) {
                        ad = real_(Ab[(size_t)k * x.lda + r]);  // Hermitian: real diagonal
                    }
                }
            }
        }
        Yp[r] = add_(rowoff, scale_(vr, ad));
        T cv = zero<T>();
        fmac_(cv, vr, rowoff);               // conj(v_r) * (off-diagonal row part)
        qacc += 2.0 * real_(cv) + ad * abs2_(vr);
    }
    // column parts: reduce over lanes, then over warps
#pragma unroll
    for (int k = 0; k < CW; ++k) {
        T v = warp_sum(colacc[k]);
        if (lane == 0) s_col[warp][k] = v;
    }
    qacc = warp_sum(qacc);
    if (lane == 0) s_q[warp] = qacc;
    __syncthreads();
    if (tid < cw) {
        T v = zero<T>();
#pragma unroll
        for (int w = 0; w < 8; ++w) v = add_(v, s_col[w][tid]);
        x.y[((size_t)blockIdx.x * x.maxseg + sgi) * CW + tid] = v;
    }
    if (tid == 0) {
        double q = 0.0;
        for (int w = 0; w < 8; ++w) q += s_q[w];
        x.pyv[slot] = mk<T>(q);   // v^H A22 v is real
        if (blockIdx.x == 0 && sgi == 0) {
            x.tau[c] = tau;
            x.e[c] = beta;
            x.d[c] = real_(x.A[(size_t)c * x.lda + c]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// TMA-fed variant of the symmetric column kernel.  A dedicated producer warp streams the strip
// through a 3-stage shared-memory ring with cp.async.bulk.tensor (2-D tiled tensor map over A,
// 256 rows x CW columns = 32 KB per stage, zero register cost, ~190 KB of loads in flight per SM
// with two resident CTAs); eight consumer warps do the two FMAs per element from shared memory.
// mbarrier full/empty handshakes replace block barriers in the main loop.
// ---------------------------------------------------------------------------------------
constexpr int TMA_ROWS = 256;   // matrix rows per stage
constexpr int TMA_NST = 3;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
            "r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

template <typename T>
__global__ void __launch_bounds__(288, 2)
trd_symv_tma_kernel(const __grid_constant__ CUtensorMap tmap, TrdCtx<T> x, int c, int i, int npn, int nstrips) {
    constexpr int CW = SymvCW<T>::value;
    constexpr int NBOX = sizeof(T) / 8;            // boxes (of 256 doubles x CW) per stage
    constexpr int RB = TMA_ROWS / NBOX;            // matrix rows per box
    constexpr unsigned STAGE_BYTES = TMA_ROWS * CW * sizeof(T);
    const int SEG = x.seg;
    const int sgi = blockIdx.y;
    const int slot = sgi * gridDim.x + blockIdx.x;
    __shared__ T s_vc[CW];
    __shared__ T s_col[8][CW];
    __shared__ double s_q[8];
    __shared__ double s_sigma;
    __shared__ __align__(8) uint64_t full_bar[TMA_NST], empty_bar[TMA_NST];
    extern __shared__ __align__(128) unsigned char tma_smem_raw[];
    unsigned char* tma_smem = tma_smem_raw + ((128u - (smem_u32(tma_smem_raw) & 127u)) & 127u);  // TMA dst alignment
    const int n = x.n, row0 = c + 1, mt = n - row0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {
        double sg = 0.0;
        for (int q = lane; q < npn; q += 32) sg += x.pn[q];
        sg = warp_sum(sg);
        if (lane == 0) s_sigma = sg;
    }
    if (tid == 32) {
        for (int st = 0; st < TMA_NST; ++st) { mbar_init(&full_bar[st], 1); mbar_init(&empty_bar[st], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const double sigma = s_sigma;
    const T* acol = x.A + (size_t)c * x.lda + row0;
    const T alpha = acol[0];
    double beta; T tau, scale;
    larfgp_scalars<T>(alpha, sigma, beta, tau, scale);

    if ((int)blockIdx.x >= nstrips) {
        if (sgi > 0 || warp == 8) { if (tid == 0 && sgi > 0) x.pyv[slot] = zero<T>(); return; }
        // ---- panel columns: t1 = W^H v, t2 = V^H v ----
        const int q0 = (((int)blockIdx.x - nstrips) * 8 + warp) * 4;
        const T* col[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int q = q0 + k;
            if (q < i) col[k] = x.P + (size_t)(x.pw + q) * x.ldp + row0;
            else if (q < 2 * i) col[k] = x.P + (size_t)(q - i) * x.ldp + row0;
            else col[k] = nullptr;
        }
        T acc[4] = {zero<T>(), zero<T>(), zero<T>(), zero<T>()};
        if (q0 < 2 * i) {
            for (int r = lane; r < mt; r += 32) {
                T vr = (r == 0) ? one<T>() : mul_(acol[r], scale);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (col[k]) fmac_(acc[k], col[k][r], vr);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc[k] = warp_sum(acc[k]);
            int q = q0 + k;
            if (lane == 0 && q < 2 * i) {
                if (q < i) x.t[q] = acc[k];
                else x.t[TRD_NB + (q - i)] = acc[k];
            }
        }
        if (tid == 0) x.pyv[slot] = zero<T>();
        return;
    }

    const int cb = blockIdx.x * CW;
    const int cw = (mt - cb < CW) ? (mt - cb) : CW;
    const int rs = cb + sgi * SEG;
    if (rs >= mt) { if (tid == 0) x.pyv[slot] = zero<T>(); return; }
    const int re = (rs + SEG < mt) ? (rs + SEG) : mt;
    if (tid < CW) {
        T vq = zero<T>();
        if (tid < cw) {
            int r = cb + tid;
            vq = (r == 0) ? one<T>() : mul_(acol[r], scale);
            if (sgi == 0) {
                x.P[(size_t)i * x.ldp + row0 + r] = vq;
                x.P[(size_t)(2 * x.pw + i) * x.ldp + row0 + r] = vq;
            }
        }
        s_vc[tid] = vq;
    }
    __syncthreads();
    // TMA needs a 16-byte aligned box start: for 8-byte elements an odd global row is reached by
    // starting one row early and masking that row
    const int shift = (NBOX == 1) ? ((row0 + rs) & 1) : 0;
    const int ntile = (re - rs + shift + TMA_ROWS - 1) / TMA_ROWS;
    T colacc[CW];
#pragma unroll
    for (int k = 0; k < CW; ++k) colacc[k] = zero<T>();
    double qacc = 0.0;
    if (warp == 8) {
        // ---- producer: one lane streams the strip through the ring ----
        if (lane == 0) {
            for (int t = 0; t < ntile; ++t) {
                const int st = t % TMA_NST, use = t / TMA_NST;
                if (use > 0) mbar_wait(&empty_bar[st], (unsigned)((use - 1) & 1));
                mbar_expect_tx(&full_bar[st], STAGE_BYTES);
                unsigned char* dst = tma_smem + (size_t)st * STAGE_BYTES;
                const int grow = (row0 + rs - shift + t * TMA_ROWS) * NBOX;   // row coordinate in doubles
#pragma unroll
                for (int bx = 0; bx < NBOX; ++bx)
                    tma_load_2d(dst + (size_t)bx * (STAGE_BYTES / NBOX), &tmap, &full_bar[st], grow + bx * 256,
                                row0 + cb);
            }
        }
    } else {
        // ---- consumers: warp w owns rows 32w..32w+31 of every tile ----
        T* Yp = x.ypart + (size_t)blockIdx.x * x.ldy + row0;
        for (int t = 0; t < ntile; ++t) {
            const int st = t % TMA_NST, use = t / TMA_NST;
            mbar_wait(&full_bar[st], (unsigned)(use & 1));
            const T* tile = reinterpret_cast<const T*>(tma_smem + (size_t)st * STAGE_BYTES);
            const int rr = 32 * warp + lane;                 // row inside the tile
            const int rt = rs - shift + t * TMA_ROWS + 32 * warp;    // first row of this warp's slice
            const int r = rt + lane;
            // element (rr, k): box rr / RB, row rr % RB inside it
            const T* src = tile + (size_t)(rr / RB) * (CW * RB) + (rr % RB);
            if (r < re && r >= rs) {
                const T vr = (r == 0) ? one<T>() : mul_(acol[r], scale);
                T rowoff = zero<T>();
                double ad = 0.0;
                if (rt >= cb + CW && cw == CW) {
#pragma unroll
                    for (int k = 0; k < CW; ++k) {
                        const T a = src[k * RB];
                        fma_(rowoff, a, s_vc[k]);
                        fmac_(colacc[k], a, vr);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < CW; ++k) {
                        const int cc = cb + k;
                        if (k < cw) {
                            if (r > cc) {
                                const T a = src[k * RB];
                                fma_(rowoff, a, s_vc[k]);
                                fmac_(colacc[k], a, vr);
                            } else if (r == cc) {
                                ad = real_(src[k * RB]);
                            }
                        }
                    }
                }
                Yp[r] = add_(rowoff, scale_(vr, ad));
                T cv = zero<T>();
                fmac_(cv, vr, rowoff);
                qacc += 2.0 * real_(cv) + ad * abs2_(vr);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
        }
#pragma unroll
        for (int k = 0; k < CW; ++k) {
            T v = warp_sum(colacc[k]);
            if (lane == 0) s_col[warp][k] = v;
        }
        qacc = warp_sum(qacc);
        if (lane == 0) s_q[warp] = qacc;
    }
    __syncthreads();
    if (tid < cw) {
        T v = zero<T>();
#pragma unroll
        for (int w = 0; w < 8; ++w) v = add_(v, s_col[w][tid]);
        x.y[((size_t)blockIdx.x * x.maxseg + sgi) * CW + tid] = v;
    }
    if (tid == 0) {
        double q = 0.0;
        for (int w = 0; w < 8; ++w) q += s_q[w];
        x.pyv[slot] = mk<T>(q);
        if (blockIdx.x == 0 && sgi == 0) {
            x.tau[c] = tau;
            x.e[c] = beta;
            x.d[c] = real_(x.A[(size_t)c * x.lda + c]);
        }
    }
}

// host: 2-D tiled tensor map over A viewed as doubles (rows = n * sizeof(T)/8 contiguous, n columns)
template <typename T>
static bool make_symv_tmap(CUtensorMap* tm, const T* A, int n, int lda, bool force = false) {
    static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    }
    // measured on B200 (round 1): the register-landing kernel is ~8 % faster than the TMA ring for this
    // access pattern, so TMA is opt-in (MAKB200_SYMV_TMA=1)
    const char* e = getenv("MAKB200_SYMV_TMA");
    if (!force && !(e && e[0] == '1')) return false;
    if (!enc) return false;
    if ((reinterpret_cast<uintptr_t>(A) & 15) != 0) return false;
    if (((size_t)lda * sizeof(T)) % 16 != 0) return false;
    constexpr int NBOX = sizeof(T) / 8;
    cuuint64_t gdim[2] = {(cuuint64_t)n * NBOX, (cuuint64_t)n};
    cuuint64_t gstr[1] = {(cuuint64_t)lda * sizeof(T)};
    cuuint32_t box[2] = {256, (cuuint32_t)SymvCW<T>::value};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)A, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// column part of y for global row r: sum over the row segments of the strip that owns column r
template <typename T>
__device__ __forceinline__ T trd_ycol(const TrdCtx<T>& x, int row0, int mt, int r) {
    constexpr int CW = SymvCW<T>::value;
    const int rl = r - row0, b = rl / CW, k = rl - b * CW;
    const int nseg = (mt - b * CW + x.seg - 1) / x.seg;
    T s = zero<T>();
    for (int sg = 0; sg < nseg; ++sg) s = add_(s, x.y[((size_t)b * x.maxseg + sg) * CW + k]);
    return s;
}

}  // namespace mak
#include "trd2.cuh"
namespace mak {

// 288 threads: warps 0..7 make ONE pass over the panel rows (V[r,p], W[r,p] feed both the w update
// and the left-looking update of the next column) and sum the row parts of y; warp 8 computes the
// step's scalars (y^H v, alpha2, first row of w) concurrently.
template <typename T>
__global__ void __launch_bounds__(288)
trd_w_kernel(TrdCtx<T> x, int c, int i, int npyv, int do_next) {
    pdl_wait_then_trigger();
    __shared__ T sm[TRD_K2_NG][TRD_K2_ROWS];
    __shared__ T sm2[TRD_K2_NG][TRD_K2_ROWS];
    __shared__ T st[2 * TRD_NB];       // t1, t2
    __shared__ T srow[2 * TRD_NB + 2]; // conj(W[c1,p]), conj(V[c1,p])
    __shared__ T sscal[4];
    __shared__ double red[32];
    const int n = x.n, row0 = c + 1, c1 = c + 1;
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool scalar_warp = (warp == 8);
    const int tx = tid % TRD_K2_ROWS, ty = (tid / TRD_K2_ROWS) % TRD_K2_NG;
    const int r = row0 + blockIdx.x * TRD_K2_ROWS + tx;
    const bool live = !scalar_warp && r < n;
    const T tauc = x.tau[c];
    for (int p = tid; p < i; p += blockDim.x) {
        st[p] = x.t[p];
        st[TRD_NB + p] = x.t[TRD_NB + p];
        srow[p] = conj_(x.P[(size_t)(x.pw + p) * x.ldp + c1]);          // conj(W[c1,p])
        srow[TRD_NB + 1 + p] = conj_(x.P[(size_t)p * x.ldp + c1]);      // conj(V[c1,p])
    }
    __syncthreads();
    T part = zero<T>(), part2 = zero<T>();
    if (scalar_warp) {
        // alpha2 = -(tau/2) * w^H v,  w^H v = conj(tau) * (y^H v - t1^H t2 - t2^H t1)
        const int lane = tid & 31;
        T yhv = zero<T>();
        for (int q = lane; q < npyv; q += 32) yhv = add_(yhv, x.pyv[q]);
        T s12 = zero<T>();
        for (int p = lane; p < i; p += 32) {
            fmac_(s12, st[p], st[TRD_NB + p]);
            fmac_(s12, st[TRD_NB + p], st[p]);
        }
        // first row of w (row c+1): y - V t1 - W t2, y[row0] = column part + row part of strip 0
        T sf = zero<T>();
        if (lane == 0) sf = neg_(x.ypart[row0]);
        for (int p = lane; p < i; p += 32) {
            fma_(sf, x.P[(size_t)p * x.ldp + row0], st[p]);
            fma_(sf, x.P[(size_t)(x.pw + p) * x.ldp + row0], st[TRD_NB + p]);
        }
        yhv = warp_sum(yhv);
        s12 = warp_sum(s12);
        sf = warp_sum(sf);
        if (lane == 0) {
            T whv = mul_(conj_(tauc), sub_(yhv, s12));
            T alpha2 = neg_(scale_(mul_(tauc, whv), 0.5));
            T wfirst = add_(mul_(tauc, sub_(trd_ycol<T>(x, row0, n - row0, row0), sf)), alpha2);  // v[row0] = 1
            sscal[0] = alpha2;
            sscal[1] = wfirst;
        }
    } else if (live) {
        for (int p = ty; p < i; p += TRD_K2_NG) {
            const T vp = x.P[(size_t)p * x.ldp + r], wp = x.P[(size_t)(x.pw + p) * x.ldp + r];
            fma_(part, vp, st[p]);
            fma_(part, wp, st[TRD_NB + p]);
            fma_(part2, vp, srow[p]);
            fma_(part2, wp, srow[TRD_NB + 1 + p]);
        }
        // row parts of y from every strip at or left of this row (subtracted: w = tau (y - part)).
        // Up to n/16 strips per row (the bottom-row CTAs are the critical path of the launch).
        const int nsb = (r - row0) / SymvCW<T>::value + 1;
        const T* yp = x.ypart + r;
        // eight independent partial sums: the launch is bound by loads in flight per SM, not by bytes
        constexpr int YU = 8;
        T ya[YU];
#pragma unroll
        for (int u = 0; u < YU; ++u) ya[u] = zero<T>();
        int b = ty;
        for (; b + (YU - 1) * TRD_K2_NG < nsb; b += YU * TRD_K2_NG) {
#pragma unroll
            for (int u = 0; u < YU; ++u) ya[u] = add_(ya[u], yp[(size_t)(b + u * TRD_K2_NG) * x.ldy]);
        }
        for (; b < nsb; b += TRD_K2_NG) ya[0] = add_(ya[0], yp[(size_t)b * x.ldy]);
#pragma unroll
        for (int u = YU / 2; u > 0; u >>= 1)
#pragma unroll
            for (int v = 0; v < u; ++v) ya[v] = add_(ya[v], ya[v + u]);
        part = sub_(part, ya[0]);
    }
    if (!scalar_warp) { sm[ty][tx] = part; sm2[ty][tx] = part2; }
    __syncthreads();
    const T alpha2 = sscal[0], wfirst = sscal[1];
    double nrm = 0.0;
    if (live && ty == 0) {
        T s = sm[0][tx];
#pragma unroll
        for (int g = 1; g < TRD_K2_NG; ++g) s = add_(s, sm[g][tx]);
        const T vr = x.P[(size_t)i * x.ldp + r];
        const T wr = add_(mul_(tauc, sub_(trd_ycol<T>(x, row0, n - row0, r), s)), mul_(alpha2, vr));
        x.P[(size_t)(x.pw + i) * x.ldp + r] = wr;                              // W(:, i)
        x.A[(size_t)c * x.lda + r] = (r == row0) ? mk<T>(x.e[c]) : vr;          // reflector storage
        if (do_next) {
            // left-looking update of column c1 = c+1: previous panel columns + the new one
            T s2 = sm2[0][tx];
#pragma unroll
            for (int g = 1; g < TRD_K2_NG; ++g) s2 = add_(s2, sm2[g][tx]);
            fma_(s2, vr, conj_(wfirst));   // V[r,i] conj(W[c1,i])
            s2 = add_(s2, wr);             // W[r,i] conj(V[c1,i]), V[c1,i] = 1
            T a = sub_(x.A[(size_t)c1 * x.lda + r], s2);
            if (r == c1) a = mk<T>(real_(a));
            x.A[(size_t)c1 * x.lda + r] = a;
            if (r >= c1 + 2) nrm = abs2_(a);
        }
    }
    if (!do_next) return;
    double tot = block_sum<double>(nrm, red);
    if (tid == 0) x.pn[blockIdx.x] = tot;
}

template <typename T>
__global__ void trd_last_d_kernel(TrdCtx<T> x) {
    x.d[x.n - 1] = real_(x.A[(size_t)(x.n - 1) * x.lda + (x.n - 1)]);
}

// grid of the persistent column kernel: one wave of two CTAs per SM
static int trd2_grid(makb200_handle* h) { return 2 * h->num_sms; }
static bool trd2_early() {
    static const bool v = []() { const char* e = getenv("MAKB200_SYMV_EARLY"); return !(e && e[0] == '0'); }();
    return v;
}
static bool trd2_enabled() {
    static const bool v = []() { const char* e = getenv("MAKB200_SYMV_V2"); return !(e && e[0] == '0'); }();
    return v;
}

template <typename T, typename AR>
static void trd_carve(AR& ar, int n, TrdCtx<T>* x, int G = 0) {
    size_t nn = (size_t)(n > 0 ? n : 1);
    x->n = n;
    x->v2 = 0;
    x->yrowp = x->ycolp = x->tpart = nullptr;
    x->tpld = 0;
    // below ~1.5k columns a step is a chain of latencies, not bytes, and the lighter round-1 kernels win
    // (c128 n = 1024: 13.8 vs 16.4 ms; n = 4096: 149 vs 110 ms): MAKB200_SYMV_V2_MIN moves the switch
    // (read per call: tests/test_gpu_eigh.py runs the small-n sweep with the switch at 2)
    const int v2_min = []() { const char* e = getenv("MAKB200_SYMV_V2_MIN"); int v = e ? atoi(e) : 1536; return v < 2 ? 2 : v; }();
    if (G >= 2 * TRD_NB && trd2_enabled() && n >= v2_min && n <= 64 * G) {
        x->v2 = 1;
        x->yrowp = ar.template get<T>((size_t)G * nn);
        x->ycolp = ar.template get<T>((size_t)((nn + TRD2_BH - 1) / TRD2_BH) * nn);
        x->tpld = (int)(nn / TRD_K2_ROWS + 2);
        x->tpart = ar.template get<T>((size_t)x->tpld * 2 * TRD_NB);
    }
    x->P = ar.template get<T>(nn * 3 * TRD_NB);
    x->ldp = n > 0 ? n : 1;
    x->maxseg = (int)((nn + TRD_SEG_MIN - 1) / TRD_SEG_MIN);
    x->seg = trd_seg();
    {
        static const bool pf = []() { const char* e = getenv("MAKB200_SYMV_PREFETCH"); return e && e[0] == '1'; }();   // measured: 380 vs 375 ms hetrd -> opt-in
        x->prefetch = (pf && trd_pdl()) ? 1 : 0;
    }
    x->y = ar.template get<T>((nn + SymvCW<T>::value) * (size_t)x->maxseg);
    x->ldy = n > 0 ? n : 1;
    x->ypart = ar.template get<T>(nn * ((nn + SymvCW<T>::value - 1) / SymvCW<T>::value));
    x->t = ar.template get<T>(2 * TRD_NB);
    x->pn = ar.template get<double>(nn / TRD_K2_ROWS + 2);
    {
        size_t npyv = (nn / SymvCW<T>::value + 2 * TRD_NB / 32 + 8) * (size_t)(x->maxseg + 1);
        if (npyv < (size_t)G) npyv = (size_t)G;     // the persistent kernel writes one slot per CTA
        x->pyv = ar.template get<T>(npyv);
    }
    x->tau = ar.template get<T>(nn);
    x->d = ar.template get<double>(nn);
    x->e = ar.template get<double>(nn);
}

template <typename T>
static int hetrd(makb200_handle* h, TrdCtx<T>& x) {
    const int n = x.n;
    cudaStream_t s = h->stream;
    CUtensorMap tmap;
    const bool use_tma = make_symv_tmap<T>(&tmap, x.A, n, x.lda);
    constexpr size_t tma_smem_bytes = (size_t)TMA_NST * TMA_ROWS * SymvCW<T>::value * sizeof(T) + 128;
    static bool tma_configured = false;
    if (use_tma && !tma_configured) {
        MAK_CUDA(h, cudaFuncSetAttribute(trd_symv_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)tma_smem_bytes));
        tma_configured = true;
    }
    if (n == 1) {
        trd_last_d_kernel<T><<<1, 1, 0, s>>>(x);
        return 0;
    }
    // persistent TMA column kernels (trd2.cuh): need a 16-byte aligned A / leading dimension for the tensor map
    CUtensorMap tmap2;
    const bool v2 = x.v2 && !use_tma && make_symv_tmap<T>(&tmap2, x.A, n, x.lda, true);
    constexpr size_t v2_smem = (size_t)TRD2_NST * TRD2_BH * SymvCW<T>::value * sizeof(T) + 128;
    const int G2 = trd2_grid(h);
    if (v2) {
        static bool v2_configured = false;
        if (!v2_configured) {
            MAK_CUDA(h, cudaFuncSetAttribute(trd_symv2_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v2_smem));
            v2_configured = true;
        }
    }
    for (int j0 = 0; j0 < n - 1; j0 += TRD_NB) {
        const int ncols = (n - 1 - j0 < TRD_NB) ? (n - 1 - j0) : TRD_NB;
        x.pw = ncols;
        int npn;  // grid size of whichever kernel produced the partial norms of the current column
        {
            int mt = n - j0 - 1;
            npn = (mt + TRD_K2_ROWS - 1) / TRD_K2_ROWS;
            trd_colnorm_kernel<T><<<npn, 256, 0, s>>>(x, j0);
            count_launch();
        }
        for (int i = 0; i < ncols; ++i) {
            const int c = j0 + i, mt = n - c - 1;
            const int g2 = (mt + TRD_K2_ROWS - 1) / TRD_K2_ROWS;
            const int nstrips = (mt + SymvCW<T>::value - 1) / SymvCW<T>::value;
            const int gx = nstrips + (2 * i + 31) / 32, gy = (mt + x.seg - 1) / x.seg;
            const int g1 = gx * gy;  // pyv slots written by this launch
            if (v2) {
                const int do_next2 = (i + 1 < ncols) ? 1 : 0;
                const bool pdl2 = trd_pdl() && !g_clock_dots.on;
                g_clock_dots.begin(s);
                if (pdl2) {
                    cudaError_t e = launch_pdl(trd_symv2_kernel<T>, dim3(G2), dim3(288), v2_smem, s, tmap2, x, c, i, npn, (i > 0 && trd2_early()) ? 1 : 0);
                    if (e != cudaSuccess) return cuda_fail(h, e, "trd_symv2_kernel (PDL)");
                } else trd_symv2_kernel<T><<<G2, 288, v2_smem, s>>>(tmap2, x, c, i, npn, 0);
                g_clock_dots.end(s);
                g_clock_w.begin(s);
                if (pdl2) {
                    cudaError_t e = launch_pdl(trd_w2_kernel<T>, dim3(g2), dim3(288), 0, s, x, c, i, G2, do_next2);
                    if (e != cudaSuccess) return cuda_fail(h, e, "trd_w2_kernel (PDL)");
                } else trd_w2_kernel<T><<<g2, 288, 0, s>>>(x, c, i, G2, do_next2);
                g_clock_w.end(s);
                count_launch(2);
                npn = g2;
                continue;
            }
            g_clock_dots.begin(s);
            const bool pdl = trd_pdl() && !use_tma && !g_clock_dots.on;
            if (use_tma) trd_symv_tma_kernel<T><<<dim3(gx, gy), 288, tma_smem_bytes, s>>>(tmap, x, c, i, npn, nstrips);
            else if (pdl) {
                cudaError_t e = launch_pdl(trd_symv_kernel<T>, dim3(gx, gy), dim3(256), 0, s, x, c, i, npn, nstrips);
                if (e != cudaSuccess) return cuda_fail(h, e, "trd_symv_kernel (PDL)");
            } else trd_symv_kernel<T><<<dim3(gx, gy), 256, 0, s>>>(x, c, i, npn, nstrips);
            g_clock_dots.end(s);
            const int do_next = (i + 1 < ncols) ? 1 : 0;
            g_clock_w.begin(s);
            if (pdl) {
                cudaError_t e = launch_pdl(trd_w_kernel<T>, dim3(g2), dim3(288), 0, s, x, c, i, g1, do_next);
                if (e != cudaSuccess) return cuda_fail(h, e, "trd_w_kernel (PDL)");
            } else trd_w_kernel<T><<<g2, 288, 0, s>>>(x, c, i, g1, do_next);
            g_clock_w.end(s);
            count_launch(2);
            npn = g2;
        }
        MAK_LAUNCH_CHECK(h, "hetrd column kernels");
        // trailing update: A[t0:, t0:] -= [V W][W V]^H  (rows >= t0 of the panel)
        const int t0 = j0 + ncols, mtr = n - t0;
        if (mtr > 0) {
            // only the lower triangle of the trailing matrix is ever read: skip tiles above it
            cudaError_t e = gemm<T>(s, h->num_sms, MAKB200_OP_N, MAKB200_OP_C, mtr, mtr, 2 * ncols, neg_(one<T>()),
                                    x.P + t0, x.ldp, x.P + (size_t)ncols * x.ldp + t0, x.ldp, one<T>(),
                                    x.A + (size_t)t0 * x.lda + t0, x.lda, nullptr, 0, true);
            if (e != cudaSuccess) return cuda_fail(h, e, "hetrd gemm");
        }
    }
    trd_last_d_kernel<T><<<1, 1, 0, s>>>(x);
    MAK_LAUNCH_CHECK(h, "trd_last_d_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------
// gauge: V[:, j] *= conj(sign(first entry of maximal modulus))   (common/gauge.jl:12-14,38-45)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void gauge_cols_kernel(int m, int ncols, T* __restrict__ V, int ldv, T* __restrict__ other, int ldo,
                                  int other_rows_n) {
    // one CTA per column; `other` (optional): rows of a second matrix scaled by sign (svd V^H)
    for (int j = blockIdx.x; j < ncols; j += gridDim.x) gauge_column_body<T>(m, j, V, ldv, other, ldo, other_rows_n);
}

template <typename T>
int gauge_columns(makb200_handle* h, int m, int ncols, T* V, int ldv, T* other, int ldo, int other_n) {
    if (ncols <= 0) return 0;
    gauge_cols_kernel<T><<<ncols < 4096 ? ncols : 4096, 256, 0, h->stream>>>(m, ncols, V, ldv, other, ldo, other_n);
    MAK_LAUNCH_CHECK(h, "gauge_cols_kernel");
    return 0;
}
template int gauge_columns<double>(makb200_handle*, int, int, double*, int, double*, int, int);
template int gauge_columns<cplx>(makb200_handle*, int, int, cplx*, int, cplx*, int, int);

__global__ void real_to_T_kernel(int n, const double* __restrict__ Z, int ldz, cplx* __restrict__ V, int ldv) {
    size_t total = (size_t)n * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % n), c = (int)(idx / n);
        V[(size_t)c * ldv + r] = cplx{Z[(size_t)c * ldz + r], 0.0};
    }
}

// ---------------------------------------------------------------------------------------
// drivers
// ---------------------------------------------------------------------------------------
template <typename T>
int herm_defect_t(makb200_handle* h, int n, const T* A, int lda, double* out2) {
    MAK_CUDA(h, cudaMemsetAsync(out2, 0, 2 * sizeof(double), h->stream));
    if (n <= 0) return 0;
    int nb = (n + 31) / 32;
    herm_defect_kernel<T><<<dim3(nb, nb), dim3(32, 8), 0, h->stream>>>(n, A, lda, out2);
    MAK_LAUNCH_CHECK(h, "herm_defect_kernel");
    return 0;
}
template int herm_defect_t<double>(makb200_handle*, int, const double*, int, double*);
template int herm_defect_t<cplx>(makb200_handle*, int, const cplx*, int, double*);

// project_hermitian! / project_antihermitian! (B may be A), the Hermitian property pass and the isometry
// defect of a Gram matrix: one launch each (projections.cuh)
template <typename T>
int project_herm_t(makb200_handle* h, int anti, int n, const T* A, int lda, T* B, int ldb) {
    if (n <= 0) return 0;
    const int nb = (n + 31) / 32;
    if (anti) project_herm_kernel<T, true><<<dim3(nb, nb), dim3(32, 8), 0, h->stream>>>(n, A, lda, B, ldb);
    else project_herm_kernel<T, false><<<dim3(nb, nb), dim3(32, 8), 0, h->stream>>>(n, A, lda, B, ldb);
    count_launch();
    MAK_LAUNCH_CHECK(h, "project_herm_kernel");
    return 0;
}
template <typename T>
int herm_props_t(makb200_handle* h, int anti, int n, const T* A, int lda, double* out4) {
    MAK_CUDA(h, cudaMemsetAsync(out4, 0, 4 * sizeof(double), h->stream));
    if (n <= 0) return 0;
    const int nb = (n + 31) / 32;
    if (anti) herm_props_kernel<T, true><<<dim3(nb, nb), dim3(32, 8), 0, h->stream>>>(n, A, lda, out4);
    else herm_props_kernel<T, false><<<dim3(nb, nb), dim3(32, 8), 0, h->stream>>>(n, A, lda, out4);
    count_launch();
    MAK_LAUNCH_CHECK(h, "herm_props_kernel");
    return 0;
}
template <typename T>
int gram_defect_t(makb200_handle* h, int n, const T* P, int ldp, double* out2) {
    MAK_CUDA(h, cudaMemsetAsync(out2, 0, 2 * sizeof(double), h->stream));
    if (n <= 0) return 0;
    const size_t total = (size_t)n * n;
    const size_t want = (total + 255) / 256, cap = (size_t)h->num_sms * 8;
    gram_defect_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, h->stream>>>(n, P, ldp, out2);
    count_launch();
    MAK_LAUNCH_CHECK(h, "gram_defect_kernel");
    return 0;
}
template <typename T>
int tri_init_t(makb200_handle* h, int mode, int m, int n, T* A, int lda) {
    if (m <= 0 || n <= 0) return 0;
    const size_t total = (size_t)m * n;
    const size_t want = (total + 255) / 256, cap = (size_t)h->num_sms * 16;
    tri_init_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, h->stream>>>(mode, m, n, A, lda);
    count_launch();
    MAK_LAUNCH_CHECK(h, "tri_init_kernel");
    return 0;
}
template <typename T>
int fro2_t(makb200_handle* h, int m, int n, const T* A, int lda, double* out1) {
    MAK_CUDA(h, cudaMemsetAsync(out1, 0, sizeof(double), h->stream));
    if (m <= 0 || n <= 0) return 0;
    const size_t total = (size_t)m * n;
    const size_t want = (total + 255) / 256, cap = (size_t)h->num_sms * 8;
    fro2_atomic_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, h->stream>>>(m, n, A, lda, out1);
    count_launch();
    MAK_LAUNCH_CHECK(h, "fro2_atomic_kernel");
    return 0;
}
#define INST_PROJ(T)                                                                          \
    template int project_herm_t<T>(makb200_handle*, int, int, const T*, int, T*, int);        \
    template int herm_props_t<T>(makb200_handle*, int, int, const T*, int, double*);          \
    template int gram_defect_t<T>(makb200_handle*, int, const T*, int, double*);             \
    template int tri_init_t<T>(makb200_handle*, int, int, int, T*, int);                    \
    template int fro2_t<T>(makb200_handle*, int, int, const T*, int, double*);
INST_PROJ(double)
INST_PROJ(cplx)

// EXPERIMENTAL two-stage path (MAKB200_EIGH_TWOSTAGE=<bandwidth 8..64>, "1" = 64; default off):
// dense -> band (sy2sb, qr.cu) -> tridiagonal (bulge chasing, sbr.cu) -> D&C -> Q2 (diamond blocks)
// -> Q1 (compact-WY back-transform).  Round-2 item 1 of DESIGN.md section 7; every stage is
// parity-tested on its own, the assembled path is opt-in until it beats the one-stage reduction.
static int eigh_twostage_b() {
    static const int v = []() {
        const char* e = getenv("MAKB200_EIGH_TWOSTAGE");
        if (!e || e[0] == '0') return 0;
        int b = atoi(e);
        if (b == 1) b = 64;
        if (b < 8) b = 8;
        if (b > 64) b = 64;
        return b;
    }();
    return v;
}
// sweeps per diamond block of the Q2 application (<= bandwidth); MAKB200_Q2_G overrides
static int eigh_twostage_g(int b) {
    static const int v = []() { const char* e = getenv("MAKB200_Q2_G"); return e ? atoi(e) : 0; }();
    int g = v > 0 ? v : b;
    if (g > b) g = b;
    if (g < 1) g = 1;
    return g;
}
// ---------------------------------------------------------------------------------------
// values-only tridiagonal solver: thread k brackets the k-th eigenvalue (sturm_core.h)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) sturm_eigvals_kernel(int n, const double* __restrict__ d,
                                                           const double* __restrict__ e, double* __restrict__ W) {
    const int k = blockIdx.x * 32 + threadIdx.x;
    if (k >= n) return;
    const sturm::Bounds b = sturm::bounds(n, d, e);
    W[k] = sturm::kth_eigenvalue(n, d, e, b, k);
}

template <typename T>
struct TwoStageWork {
    T* tau1;   // n
    T* V2;     // n x n
    T* tau2;   // ldt x n
    int ldt;
};

template <typename T, typename AR>
static void eigh_carve(makb200_handle* h, AR& ar, int n, TrdCtx<T>* x, double** Zreal, void** sub, size_t* sub_bytes,
                       TwoStageWork<T>* ts = nullptr) {
    trd_carve<T>(ar, n, x, trd2_grid(h));
    size_t nn = (size_t)(n > 0 ? n : 1);
    *Zreal = is_cplx<T>::value ? ar.template get<double>(nn * nn) : nullptr;
    size_t a = stedc_worksize(n);
    size_t b = ormqr_worksize_t<T>(h, n > 1 ? n - 1 : 1, n > 1 ? n - 1 : 1, n);
    *sub_bytes = a > b ? a : b;
    const int b2 = eigh_twostage_b();
    if (b2 > 0 && n > 2 * b2) {
        TwoStageWork<T> tmp;
        TwoStageWork<T>* t = ts ? ts : &tmp;
        t->ldt = (n + b2 - 1) / b2 + 1;
        t->tau1 = ar.template get<T>(nn);
        t->V2 = ar.template get<T>(nn * nn);
        t->tau2 = ar.template get<T>((size_t)t->ldt * nn);
        size_t c = sy2sb_worksize_t<T>(h, n, b2), d = sbr_chase_worksize_t<T>(n, b2),
               e = sbr_apply_q2_worksize_t<T>(n, b2, eigh_twostage_g(b2), n), f = ormqr_worksize_t<T>(h, n - b2, n - b2, n);
        if (c > *sub_bytes) *sub_bytes = c;
        if (d > *sub_bytes) *sub_bytes = d;
        if (e > *sub_bytes) *sub_bytes = e;
        if (f > *sub_bytes) *sub_bytes = f;
    }
    *sub = ar.template get<char>(*sub_bytes);
}

// opt-in dynamic shared memory of the one-CTA tridiagonalisation kernels (both entry points, once per type)
template <typename T>
static int bhetrd_configure(makb200_handle* h, size_t smem) {
    static bool done = false;   // benign race: the attribute call is idempotent
    if (smem > 48 * 1024 && !done) {
        const int mx = (int)(bhetrd_smem_elems(BHETRD_MAX_N) * sizeof(T));
        MAK_CUDA(h, cudaFuncSetAttribute(bhetrd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        MAK_CUDA(h, cudaFuncSetAttribute(bhetrd_one_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        done = true;
    }
    return 0;
}
static bool bhetrd_single() {
    const char* e = getenv("MAKB200_BHETRD");   // read per call: the bring-up tests toggle it
    return e && e[0] == '2';
}

template <typename T>
size_t eigh_worksize_t(makb200_handle* h, int n) {
    ArenaSize ar;
    TrdCtx<T> x;
    double* zr;
    void* sub;
    size_t sb;
    eigh_carve<T>(h, ar, n, &x, &zr, &sub, &sb);
    return ar.off + 256;
}

template <typename T>
int eigh_t(makb200_handle* h, int n, T* A, int lda, double* W, T* V, int ldv, int fixgauge, void* work, size_t lwork,
           int* info_dev, int top, const TrdPre<T>* pre) {
    if (n <= 0) return 0;
    // top in (0, n): only the eigenvectors of the `top` largest eigenvalues (the last columns of V) are
    // back-transformed and gauged; the leading n - top columns of V are left holding scratch
    const int c0 = (top > 0 && top < n) ? n - top : 0;
    const int nc = n - c0;
    cudaStream_t s = h->stream;
    Arena ar(work, lwork);
    TrdCtx<T> x;
    double* Zreal;
    void* sub;
    size_t sb;
    TwoStageWork<T> ts{};
    eigh_carve<T>(h, ar, n, &x, &Zreal, &sub, &sb, &ts);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    x.A = A;
    x.lda = lda;
    int nb32 = (n + 31) / 32;
    PhaseTimer pt(s);
    pt.mark("start");
    if (!pre) {
        mirror_upper_kernel<T><<<dim3(nb32, nb32), dim3(32, 8), 0, s>>>(n, A, lda);
        MAK_LAUNCH_CHECK(h, "mirror_upper_kernel");
    }
    const int b2 = eigh_twostage_b();
    const bool two_stage = !pre && b2 > 0 && n > 2 * b2;
    int rc = 0;
    if (pre) {
        // tridiagonalised in place by bhetrd_batched_t: d, e, tau and the reflectors below the sub-diagonal of A
        x.d = pre->d;
        x.e = pre->e;
        x.tau = pre->tau;
    } else if (two_stage) {
        rc = sy2sb_t<T>(h, n, b2, A, lda, ts.tau1, sub, sb);
        if (rc) return rc;
        pt.mark("sy2sb");
        rc = sbr_chase_t<T>(h, n, b2, A, lda, x.d, x.e, ts.V2, n, ts.tau2, ts.ldt, sub, sb);
        if (rc) return rc;
        pt.mark("chase");
    } else if (bhetrd_single() && n <= BHETRD_MAX_N && n >= 3) {
        // EXPERIMENTAL (MAKB200_BHETRD=2): the whole tridiagonalisation in ONE single-CTA launch; meant for the
        // pooled per-block paths (batched svd / eigh of mid-size blocks), where 2n launches per block bound the rate
        const size_t smem = bhetrd_smem_elems(n) * sizeof(T);
        rc = bhetrd_configure<T>(h, smem);
        if (rc) return rc;
        bhetrd_one_kernel<T><<<1, BHETRD_THREADS, smem, s>>>(BhetrdDesc<T>{n, A, lda, x.d, x.e, x.tau}, 0);
        count_launch();
        MAK_LAUNCH_CHECK(h, "bhetrd_one_kernel");
        pt.mark("bhetrd");
    } else {
        rc = hetrd<T>(h, x);
        if (rc) return rc;
        pt.mark("hetrd");
    }
    if (V == nullptr) {
        // values only (job 'N'): Sturm-count K-section on the tridiagonal, one thread per eigenvalue;
        // no eigenvector GEMMs, no back-transformation
        sturm_eigvals_kernel<<<(n + 31) / 32, 32, 0, s>>>(n, x.d, x.e, W);
        count_launch();
        MAK_LAUNCH_CHECK(h, "sturm_eigvals_kernel");
        pt.mark("sturm");
        pt.report("eigh_vals");
        return 0;
    }
    double* Z;
    int ldz;
    if constexpr (is_cplx<T>::value) { Z = Zreal; ldz = n; }
    else { Z = V; ldz = ldv; }
    rc = stedc(h, n, x.d, x.e, W, Z, ldz, sub, sb, info_dev);
    if (rc) return rc;
    pt.mark("stedc");
    if constexpr (is_cplx<T>::value) {
        size_t total = (size_t)n * n;
        int blocks = (int)((total + 255) / 256 < (size_t)h->num_sms * 16 ? (total + 255) / 256 : (size_t)h->num_sms * 16);
        real_to_T_kernel<<<blocks, 256, 0, s>>>(n, Zreal, n, V, ldv);
        MAK_LAUNCH_CHECK(h, "real_to_T_kernel");
    }
    if (two_stage) {
        // X = Q1 Q2 Z: chase reflectors in diamond blocks, then the stage-1 block reflectors
        // (QR-type columns of A[b:, 0:n-b])
        rc = sbr_apply_q2_t<T>(h, n, b2, eigh_twostage_g(b2), ts.V2, n, ts.tau2, ts.ldt, V + (size_t)c0 * ldv, ldv, nc, sub, sb);
        if (rc) return rc;
        pt.mark("q2");
        rc = ormqr_left_t<T>(h, n - b2, n - b2, A + b2, lda, ts.tau1, V + (size_t)c0 * ldv + b2, ldv, nc, sub, sb);
        if (rc) return rc;
        pt.mark("q1");
    } else if (n > 1) {
        // V[1:, :] <- H_0 ... H_{n-2} V[1:, :]; reflectors = QR-type columns of B = A[1:, 0:n-1]
        rc = ormqr_left_t<T>(h, n - 1, n - 1, A + 1, lda, x.tau, V + (size_t)c0 * ldv + 1, ldv, nc, sub, sb);
        if (rc) return rc;
    }
    pt.mark("backtransform");
    count_launch(3);  // mirror, last_d, gauge
    if (fixgauge) rc = gauge_columns<T>(h, n, nc, V + (size_t)c0 * ldv, ldv, (T*)nullptr, 0, 0);
    pt.mark("gauge");
    pt.report("eigh");
    return rc;
}

template size_t eigh_worksize_t<double>(makb200_handle*, int);
template size_t eigh_worksize_t<cplx>(makb200_handle*, int);
template int eigh_t<double>(makb200_handle*, int, double*, int, double*, double*, int, int, void*, size_t, int*, int,
                            const TrdPre<double>*);
template int eigh_t<cplx>(makb200_handle*, int, cplx*, int, double*, cplx*, int, int, void*, size_t, int*, int,
                          const TrdPre<cplx>*);

template <typename T>
int bhetrd_batched_t(makb200_handle* h, int nblk, const void* descs_dev, int nmax) {
    if (nblk <= 0) return 0;
    if (nmax > BHETRD_MAX_N) return -1;
    const size_t smem = bhetrd_smem_elems(nmax) * sizeof(T);
    int rc = bhetrd_configure<T>(h, smem);
    if (rc) return rc;
    bhetrd_kernel<T><<<nblk, BHETRD_THREADS, smem, h->stream>>>((const BhetrdDesc<T>*)descs_dev, nmax, 1);
    count_launch();
    MAK_LAUNCH_CHECK(h, "bhetrd_kernel");
    return 0;
}
template int bhetrd_batched_t<double>(makb200_handle*, int, const void*, int);
template int bhetrd_batched_t<cplx>(makb200_handle*, int, const void*, int);

}  // namespace mak
