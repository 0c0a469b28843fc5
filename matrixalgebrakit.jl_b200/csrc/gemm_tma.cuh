// TMA-fed, warp-specialised FP64 DMMA GEMM (Float64; included by gemm.cu).
//
// The cp.async kernel (gemm.cu) runs the tensor pipe at 82 % of its active cycles on 4096^3 (ncu,
// profiles/r2_ncu_gemm_details.txt): every k-tile ends in a block-wide barrier, so all 16 warps drain and
// refill the DMMA queue together.  Here nobody waits for the block:
//   * one elected lane issues cp.async.bulk.tensor (TMA) for the A and B tiles of a stage - no registers, no
//     address arithmetic or bounds predicates in the math path (out-of-range rows/columns/K are zero-filled by
//     the tensor map), and every operand byte comes in as a bulk 128-byte-swizzled box.  The lane belongs to
//     math warp 0 (a 17th warp would round the CTA up to 640 threads' worth of registers: 96 per thread and
//     spills); it refills the stage the block released LAST iteration, ST - 1 stages ahead of the math;
//   * the 16 math warps wait on that stage's "full" mbarrier, issue their 64 DMMAs and arrive on its "empty"
//     mbarrier one by one - a warp that is ahead starts the next stage as soon as its bytes have landed.
//
// Shared-memory layout (SWIZZLE_128B: the 16-byte chunk index inside every 128-byte row is XORed with row % 8):
//   operand stored with the tile's M (or N) index contiguous in global memory ("MN-major", op = N for A, T for B):
//       8 boxes of 16 (mn) x 16 (k) doubles, box b = mn / 16, row = k           (2 KB per box)
//   operand stored with K contiguous ("K-major", op = T/C for A, N for B):
//       1 box of 16 (k) x 128 (mn) doubles, row = mn                             (16 KB)
// DMMA m8n8k4 fragments: lane (lr = lane / 4, lc = lane % 4) holds A[lr][k(lc)] and B[k(lc)][lr].  Which four k of
// the 16 in a stage form one DMMA step is free (a DMMA sums over its k; A and B only have to agree), and the table
//       step 0: k = 0, 3,12,15   step 1: 2, 5,14, 9   step 2: 4, 7, 8,11   step 3: 6, 1,10,13
// makes every half-warp of a fragment load hit 16 distinct 8-byte banks in BOTH layouts (K-major needs two even and
// two odd k from different halves of the row; MN-major needs k % 8 from four different pairs).
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>

namespace mak {

constexpr int TG_BM = 128, TG_BN = 128, TG_BK = 16, TG_ST = 5;
constexpr int TG_MATH_WARPS = 16, TG_THREADS = TG_MATH_WARPS * 32;
constexpr unsigned TG_TILE_BYTES = TG_BM * TG_BK * 8;          // 16 KB per operand per stage
constexpr unsigned TG_STAGE_BYTES = 2 * TG_TILE_BYTES;

__device__ __forceinline__ unsigned tg_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tg_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(tg_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tg_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(tg_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tg_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(tg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tg_mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TG_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TG_WAIT_DONE;\n"
        "bra TG_WAIT_LOOP;\n"
        "TG_WAIT_DONE:\n"
        "}\n" ::"r"(tg_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tg_tma_2d(unsigned dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
            "r"(dst), "l"(tmap), "r"(tg_smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ double tg_lds(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
    return v;
}

// k of DMMA step s for lane column lc (see the header)
__device__ __forceinline__ int tg_k(int s, int lc) {
    // packed 4-bit entries, row s = [k(lc=0), k(lc=1), k(lc=2), k(lc=3)]
    const unsigned tab[4] = {0xFC30u, 0x9E52u, 0xB874u, 0xDA16u};
    return (tab[s] >> (4 * lc)) & 15;
}

// byte offset (inside one operand tile of a stage) of the fragment element (mn = w0 + 8 i + lr, k = tg_k(s, lc)) is
// split into a per-(s) part that depends on the lane and a per-(i) part that is a compile-time constant.
template <bool KMAJOR>
__device__ __forceinline__ unsigned tg_lane_off(int s, int lr, int lc, int half) {
    const int k = tg_k(s, lc);
    if (KMAJOR) {
        // row = mn, mn % 8 = lr: lr*128 + (((k/2) ^ lr) * 16) + (k%2)*8
        return (unsigned)(lr * 128 + (((k >> 1) ^ lr) << 4) + ((k & 1) << 3));
    } else {
        // box = mn / 16, row = k, chunk = ((mn % 16) / 2) ^ (k % 8) with (mn % 16) = 8 half + lr
        const int k8 = k & 7;
        const int chunk = (((half ^ (k8 >> 2)) << 2) | ((lr >> 1) ^ (k8 & 3)));
        return (unsigned)(k * 128 + (chunk << 4) + ((lr & 1) << 3));
    }
}

// CST: the C tile goes through shared memory by TMA in both directions (rank-k updates on a large C, K <= 1024): the
// load of the old tile (beta != 0) is issued before the main loop and lands under it; the epilogue reads and rewrites
// the tile in shared memory and ONE lane stores it with cp.async.bulk.tensor (SASS UTMASTG) - full-line coalesced
// writes, no per-thread global round trips.  With the 128 KB tile the ring has ST = 3 stages (enough for <= 64 k-tiles).
template <bool TA, bool TB, int ST, bool CST>
__global__ void __launch_bounds__(TG_THREADS, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const GemmProblem<double> p, int splitk, double* __restrict__ ws) {
    constexpr int BM = TG_BM, BN = TG_BN, BK = TG_BK;
    constexpr int WM = 32, WN = 32, MT = 4, NT = 4, WARPS_M = BM / WM;
    extern __shared__ __align__(1024) unsigned char tg_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[ST], empty_bar[ST];
    __shared__ __align__(8) uint64_t c_bar;
    const unsigned ring = (tg_smem_u32(tg_smem_raw) + 1023u) & ~1023u;
    const unsigned sC = ring + (unsigned)ST * TG_STAGE_BYTES;      // CST: 128 x 128 doubles, [n][m] dense

    const int M = p.m, N = p.n, K = p.k;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (m0 >= M || n0 >= N) return;
    if (p.lower && n0 >= m0 + BM) return;

    const int ktiles = (K + BK - 1) / BK;
    int kt_beg = 0, kt_end = ktiles;
    if (!CST && splitk > 1) {
        const int per = (ktiles + splitk - 1) / splitk;
        kt_beg = blockIdx.z * per;
        kt_end = min(ktiles, kt_beg + per);
        if (kt_end < kt_beg) kt_end = kt_beg;
    }
    const int nkt = kt_end - kt_beg;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const bool beta0 = (p.beta == 0.0);
    if (tid == 0) {
        for (int st = 0; st < ST; ++st) { tg_mbar_init(&full_bar[st], 1); tg_mbar_init(&empty_bar[st], TG_MATH_WARPS); }
        if (CST) tg_mbar_init(&c_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (CST && tid == 0 && !beta0) {
        // old C tile: four boxes of 128 (m) x 32 (n), under the main loop
        tg_mbar_expect_tx(&c_bar, BM * BN * 8);
#pragma unroll
        for (int b = 0; b < 4; ++b) tg_tma_2d(sC + b * (BM * 32 * 8), &tmC, &c_bar, m0, n0 + 32 * b);
    }

    // ---- loads: stage `it` of this CTA's k range (issued by thread 0) ----
    auto issue = [&](int it) {
        const int st = it % ST;
        tg_mbar_expect_tx(&full_bar[st], TG_STAGE_BYTES);
        const int k0 = (kt_beg + it) * BK;
        const unsigned sa = ring + (unsigned)st * TG_STAGE_BYTES, sb = sa + TG_TILE_BYTES;
        if (TA) tg_tma_2d(sa, &tmA, &full_bar[st], k0, m0);                       // K-major: {16 k, 128 m}
        else {
#pragma unroll
            for (int b = 0; b < BM / 16; ++b) tg_tma_2d(sa + b * 2048, &tmA, &full_bar[st], m0 + 16 * b, k0);
        }
        if (!TB) tg_tma_2d(sb, &tmB, &full_bar[st], k0, n0);                      // K-major: {16 k, 128 n}
        else {
#pragma unroll
            for (int b = 0; b < BN / 16; ++b) tg_tma_2d(sb + b * 2048, &tmB, &full_bar[st], n0 + 16 * b, k0);
        }
    };
    if (tid == 0) {
        for (int it = 0; it < ST - 1 && it < nkt; ++it) issue(it);
    }

    // ---- math warps ----
    const int wm = (warp % WARPS_M) * WM, wn = (warp / WARPS_M) * WN;
    const int lr = lane >> 2, lc = lane & 3;
    // lane-dependent offsets per DMMA step; A is K-major iff TA, B is K-major iff !TB
    unsigned offA[4][2], offB[4][2];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            offA[s][hf] = tg_lane_off<TA>(s, lr, lc, hf) + (TA ? (unsigned)(wm * 128) : (unsigned)((wm / 16) * 2048));
            offB[s][hf] = tg_lane_off<!TB>(s, lr, lc, hf) + (!TB ? (unsigned)(wn * 128) : (unsigned)((wn / 16) * 2048));
        }
    }
    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int it = 0; it < nkt; ++it) {
        const int st = it % ST, use = it / ST;
        if (tid == 0) {
            // refill the stage everybody finished with in iteration it - 1
            const int nx = it + ST - 1;
            if (nx < nkt) {
                if (nx >= ST) tg_mbar_wait(&empty_bar[nx % ST], (unsigned)(((nx / ST) - 1) & 1));
                issue(nx);
            }
        }
        __syncwarp();
        tg_mbar_wait(&full_bar[st], (unsigned)(use & 1));
        const unsigned sa = ring + (unsigned)st * TG_STAGE_BYTES, sb = sa + TG_TILE_BYTES;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            double af[MT], bf[NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                // K-major: rows (wm + 8 i + lr): + i * 8 * 128.  MN-major: box (wm + 8 i) / 16, half = i & 1
                const unsigned o = TA ? offA[s][0] + (unsigned)(i * 1024) : offA[s][i & 1] + (unsigned)((i >> 1) * 2048);
                af[i] = tg_lds(sa + o);
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const unsigned o = !TB ? offB[s][0] + (unsigned)(j * 1024) : offB[s][j & 1] + (unsigned)((j >> 1) * 2048);
                bf[j] = tg_lds(sb + o);
            }
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncwarp();
        if (lane == 0) tg_mbar_arrive(&empty_bar[st]);
    }

    if constexpr (CST) {
        // ---- staged epilogue: combine with the old tile in shared memory, one bulk store of the tile ----
        if (!beta0) tg_mbar_wait(&c_bar, 0);
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int rl = wm + i * 8 + lr;
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int cl = wn + j * 8 + lc * 2 + e;
                    const unsigned a = sC + (unsigned)((cl * BM + rl) * 8);
                    double v = p.alpha * acc[i][j][e];
                    if (!beta0) v = fma(p.beta, tg_lds(a), v);
                    asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(a), "d"(v) : "memory");
                }
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> visible to the TMA store
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int b = 0; b < 4; ++b)
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(&tmC),
                             "r"(sC + b * (BM * 32 * 8)), "r"(m0), "r"(n0 + 32 * b)
                             : "memory");
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // the tile must outlive the reads of the store
        }
        return;
    }
    // ---- epilogue: thread holds C[row = lr][cols = 2*lc, 2*lc+1] of each 8x8 tile; the C values of one row block
    // are loaded together before any store (beta != 0: the loads of a read-modify-write chain would otherwise
    // serialise behind the stores, 32 L2 round trips per thread) ----
    const bool partial = splitk > 1;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int r = m0 + wm + i * 8 + lr;
        if (r >= M) continue;
        if (partial) {
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = n0 + wn + j * 8 + lc * 2 + e;
                    if (c < N) ws[(size_t)blockIdx.z * M * N + (size_t)c * M + r] = acc[i][j][e];
                }
            continue;
        }
        double old[NT][2];
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = n0 + wn + j * 8 + lc * 2 + e;
                old[j][e] = (!beta0 && c < N) ? p.C[(size_t)c * p.ldc + r] : 0.0;
            }
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = n0 + wn + j * 8 + lc * 2 + e;
                if (c < N) p.C[(size_t)c * p.ldc + r] = beta0 ? p.alpha * acc[i][j][e] : fma(p.alpha, acc[i][j][e], p.beta * old[j][e]);
            }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// ComplexF64: the same pipeline on interleaved (re, im) operands.  One 16-byte swizzle chunk = one complex number, a
// 128-byte row = 8 of them = BK.  64 x 128 x 8 tiles, 8 math warps (32 x 32 each, four real DMMAs per complex product).
//   MN-major operand: boxes of 8 (mn) x 8 (k), box b = mn / 8, row = k, chunk = (mn % 8) ^ k
//   K-major operand : one box of 8 (k) x BM|BN (mn), row = mn, chunk = k ^ (mn % 8)
// The two DMMA steps of a stage take k = 2 lc + s: a quarter-warp (8 lanes, one 128-byte wavefront of LDS.128) then
// reads 8 distinct chunks in both layouts.  The tensor maps describe the operands as Float64 with a doubled inner
// dimension (there is no complex element type).
constexpr int TGC_BM = 64, TGC_BN = 128, TGC_BK = 8, TGC_ST = 6;
constexpr int TGC_MATH_WARPS = 8, TGC_THREADS = TGC_MATH_WARPS * 32;
constexpr unsigned TGC_A_BYTES = TGC_BM * TGC_BK * 16, TGC_B_BYTES = TGC_BN * TGC_BK * 16;
constexpr unsigned TGC_STAGE_BYTES = TGC_A_BYTES + TGC_B_BYTES;

__device__ __forceinline__ cplx tg_lds_c(unsigned addr) {
    cplx v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.re), "=d"(v.im) : "r"(addr));
    return v;
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(TGC_THREADS, 1)
gemm_tma_kernel_c(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const GemmProblem<cplx> p, int splitk, cplx* __restrict__ ws) {
    constexpr int BM = TGC_BM, BN = TGC_BN, BK = TGC_BK, ST = TGC_ST;
    constexpr int WM = 32, MT = 4, NT = 4, WARPS_M = BM / WM;
    extern __shared__ __align__(1024) unsigned char tg_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[ST], empty_bar[ST];
    const unsigned ring = (tg_smem_u32(tg_smem_raw) + 1023u) & ~1023u;

    const int M = p.m, N = p.n, K = p.k;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (m0 >= M || n0 >= N) return;
    if (p.lower && n0 >= m0 + BM) return;
    const int ktiles = (K + BK - 1) / BK;
    int kt_beg = 0, kt_end = ktiles;
    if (splitk > 1) {
        const int per = (ktiles + splitk - 1) / splitk;
        kt_beg = blockIdx.z * per;
        kt_end = min(ktiles, kt_beg + per);
        if (kt_end < kt_beg) kt_end = kt_beg;
    }
    const int nkt = kt_end - kt_beg;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int st = 0; st < ST; ++st) { tg_mbar_init(&full_bar[st], 1); tg_mbar_init(&empty_bar[st], TGC_MATH_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int it) {
        const int st = it % ST;
        tg_mbar_expect_tx(&full_bar[st], TGC_STAGE_BYTES);
        const int k0 = (kt_beg + it) * BK;
        const unsigned sa = ring + (unsigned)st * TGC_STAGE_BYTES, sb = sa + TGC_A_BYTES;
        if (TA) tg_tma_2d(sa, &tmA, &full_bar[st], 2 * k0, m0);                    // K-major: {8 k, 64 m}
        else {
#pragma unroll
            for (int b = 0; b < BM / 8; ++b) tg_tma_2d(sa + b * 1024, &tmA, &full_bar[st], 2 * (m0 + 8 * b), k0);
        }
        if (!TB) tg_tma_2d(sb, &tmB, &full_bar[st], 2 * k0, n0);                   // K-major: {8 k, 128 n}
        else {
#pragma unroll
            for (int b = 0; b < BN / 8; ++b) tg_tma_2d(sb + b * 1024, &tmB, &full_bar[st], 2 * (n0 + 8 * b), k0);
        }
    };
    if (tid == 0) {
        for (int it = 0; it < ST - 1 && it < nkt; ++it) issue(it);
    }

    const int wm = (warp % WARPS_M) * WM, wn = (warp / WARPS_M) * 32;
    const int lr = lane >> 2, lc = lane & 3;
    unsigned offA[2], offB[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int k = 2 * lc + s;
        offA[s] = TA ? (unsigned)((wm + lr) * 128 + ((k ^ lr) << 4)) : (unsigned)((wm / 8) * 1024 + k * 128 + ((lr ^ k) << 4));
        offB[s] = !TB ? (unsigned)((wn + lr) * 128 + ((k ^ lr) << 4)) : (unsigned)((wn / 8) * 1024 + k * 128 + ((lr ^ k) << 4));
    }
    double acc[MT][NT][4];   // r0, r1, i0, i1
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.0;
    const double sgnA = p.conja ? -1.0 : 1.0, sgnB = p.conjb ? -1.0 : 1.0;

    for (int it = 0; it < nkt; ++it) {
        const int st = it % ST, use = it / ST;
        if (tid == 0) {
            const int nx = it + ST - 1;
            if (nx < nkt) {
                if (nx >= ST) tg_mbar_wait(&empty_bar[nx % ST], (unsigned)(((nx / ST) - 1) & 1));
                issue(nx);
            }
        }
        __syncwarp();
        tg_mbar_wait(&full_bar[st], (unsigned)(use & 1));
        const unsigned sa = ring + (unsigned)st * TGC_STAGE_BYTES, sb = sa + TGC_A_BYTES;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            cplx af[MT], bf[NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) af[i] = tg_lds_c(sa + offA[s] + (unsigned)(i * 1024));   // 8 rows or one box further
#pragma unroll
            for (int j = 0; j < NT; ++j) bf[j] = tg_lds_c(sb + offB[s] + (unsigned)(j * 1024));
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const double ar = af[i].re, ai = af[i].im * sgnA, nai = -ai;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const double br = bf[j].re, bi = bf[j].im * sgnB;
                    dmma(acc[i][j][0], acc[i][j][1], ar, br);
                    dmma(acc[i][j][0], acc[i][j][1], nai, bi);
                    dmma(acc[i][j][2], acc[i][j][3], ar, bi);
                    dmma(acc[i][j][2], acc[i][j][3], ai, br);
                }
            }
        }
        __syncwarp();
        if (lane == 0) tg_mbar_arrive(&empty_bar[st]);
    }

    const bool partial = splitk > 1;
    const bool beta0 = is_zero(p.beta);
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int r = m0 + wm + i * 8 + lr;
        if (r >= M) continue;
        if (partial) {
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = n0 + wn + j * 8 + lc * 2 + e;
                    if (c < N) ws[(size_t)blockIdx.z * M * N + (size_t)c * M + r] = cplx{acc[i][j][e], acc[i][j][2 + e]};
                }
            continue;
        }
        cplx old[NT][2];
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = n0 + wn + j * 8 + lc * 2 + e;
                old[j][e] = (!beta0 && c < N) ? p.C[(size_t)c * p.ldc + r] : cplx{0.0, 0.0};
            }
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = n0 + wn + j * 8 + lc * 2 + e;
                if (c >= N) continue;
                cplx out = mul_(p.alpha, cplx{acc[i][j][e], acc[i][j][2 + e]});
                if (!beta0) out = add_(out, mul_(p.beta, old[j][e]));
                p.C[(size_t)c * p.ldc + r] = out;
            }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 tg_encoder() {
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
    return enc;
}
// Tensor map over a column-major operand stored as (rows x cols, leading dimension ld): dim 0 = rows (contiguous).
// box0 x box1 elements per TMA box.
static bool tg_make_map(CUtensorMap* tm, const double* base, int rows, int cols, int ld, int box0, int box1, bool swizzle = true) {
    PFN_cuTensorMapEncodeTiled_v12000 enc = tg_encoder();
    if (!enc) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(double)};
    cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool tg_enabled() {
    static const bool v = []() { const char* e = getenv("MAKB200_GEMM_TMA"); return !(e && e[0] == '0'); }();
    return v;
}
// operands a tensor map can describe: 16-byte aligned base, leading dimension a multiple of 16 bytes
static bool tg_operand_ok(const double* p, int ld) {
    return ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && ((ld & 1) == 0) && ld > 0;
}

template <bool TA, bool TB, int ST, bool CST>
static cudaError_t tg_launch(cudaStream_t stream, dim3 grid, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                             const GemmProblem<double>& p, int splitk, double* ws) {
    constexpr size_t smem = (size_t)ST * TG_STAGE_BYTES + (CST ? (size_t)TG_BM * TG_BN * 8 : 0) + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tma_kernel<TA, TB, ST, CST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    g_clock_gemm.begin(stream);
    gemm_tma_kernel<TA, TB, ST, CST><<<grid, TG_THREADS, smem, stream>>>(tmA, tmB, tmC, p, splitk, ws);
    g_clock_gemm.end(stream);
    count_launch();
    return cudaGetLastError();
}
static int tg_cstage_maxk() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MAKB200_GEMM_CSTAGE_MAXK"); v = e ? atoi(e) : 1024; if (v < 0) v = 0; }
    return v;
}

// returns false when the TMA path does not apply (the caller then takes the cp.async kernel)
static bool gemm_tma_try(cudaStream_t stream, bool ta, bool tb, dim3 grid, const GemmProblem<double>& p, int splitk,
                         double* ws, cudaError_t* err) {
    if (!tg_enabled() || !tg_operand_ok(p.A, p.lda) || !tg_operand_ok(p.B, p.ldb)) return false;
    CUtensorMap tmA, tmB, tmC;
    // A as stored: op N -> M x K (M contiguous, MN-major boxes 16 x 16); op T/C -> K x M (K contiguous, box 16 x 128)
    const bool okA = ta ? tg_make_map(&tmA, p.A, p.k, p.m, p.lda, TG_BK, TG_BM) : tg_make_map(&tmA, p.A, p.m, p.k, p.lda, 16, TG_BK);
    // B as stored: op N -> K x N (K contiguous, box 16 x 128); op T/C -> N x K (N contiguous, boxes 16 x 16)
    const bool okB = tb ? tg_make_map(&tmB, p.B, p.n, p.k, p.ldb, 16, TG_BK) : tg_make_map(&tmB, p.B, p.k, p.n, p.ldb, TG_BK, TG_BN);
    if (!okA || !okB) return false;
    // staged C tile: a rank-k update (short K, no split-K) whose C a tensor map can describe
    const bool cst = splitk == 1 && p.k <= tg_cstage_maxk() && tg_operand_ok(p.C, p.ldc) &&
                     tg_make_map(&tmC, p.C, p.m, p.n, p.ldc, TG_BM, 32, false);
    if (!cst) tmC = tmA;   // unused
#define TG_DISPATCH(TA_, TB_)                                                                                  \
    (cst ? tg_launch<TA_, TB_, 3, true>(stream, grid, tmA, tmB, tmC, p, splitk, ws)                            \
         : tg_launch<TA_, TB_, TG_ST, false>(stream, grid, tmA, tmB, tmC, p, splitk, ws))
    if (!ta && !tb) *err = TG_DISPATCH(false, false);
    else if (ta && !tb) *err = TG_DISPATCH(true, false);
    else if (!ta && tb) *err = TG_DISPATCH(false, true);
    else *err = TG_DISPATCH(true, true);
#undef TG_DISPATCH
    return true;
}

template <bool TA, bool TB>
static cudaError_t tgc_launch(cudaStream_t stream, dim3 grid, const CUtensorMap& tmA, const CUtensorMap& tmB,
                              const GemmProblem<cplx>& p, int splitk, cplx* ws) {
    constexpr size_t smem = (size_t)TGC_ST * TGC_STAGE_BYTES + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tma_kernel_c<TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    g_clock_gemm.begin(stream);
    gemm_tma_kernel_c<TA, TB><<<grid, TGC_THREADS, smem, stream>>>(tmA, tmB, p, splitk, ws);
    g_clock_gemm.end(stream);
    count_launch();
    return cudaGetLastError();
}

// ComplexF64 operands are always 16-byte aligned with a 16-byte multiple leading dimension
static bool gemm_tma_try(cudaStream_t stream, bool ta, bool tb, dim3 grid, const GemmProblem<cplx>& p, int splitk,
                         cplx* ws, cudaError_t* err) {
    if (!tg_enabled() || p.lda <= 0 || p.ldb <= 0) return false;
    if ((reinterpret_cast<uintptr_t>(p.A) & 15) || (reinterpret_cast<uintptr_t>(p.B) & 15)) return false;
    CUtensorMap tmA, tmB;
    // as Float64 with a doubled inner dimension; leading dimensions in doubles = 2 ld
    const bool okA = ta ? tg_make_map(&tmA, (const double*)p.A, 2 * p.k, p.m, 2 * p.lda, 2 * TGC_BK, TGC_BM)
                        : tg_make_map(&tmA, (const double*)p.A, 2 * p.m, p.k, 2 * p.lda, 16, TGC_BK);
    const bool okB = tb ? tg_make_map(&tmB, (const double*)p.B, 2 * p.n, p.k, 2 * p.ldb, 16, TGC_BK)
                        : tg_make_map(&tmB, (const double*)p.B, 2 * p.k, p.n, 2 * p.ldb, 2 * TGC_BK, TGC_BN);
    if (!okA || !okB) return false;
    if (!ta && !tb) *err = tgc_launch<false, false>(stream, grid, tmA, tmB, p, splitk, ws);
    else if (ta && !tb) *err = tgc_launch<true, false>(stream, grid, tmA, tmB, p, splitk, ws);
    else if (!ta && tb) *err = tgc_launch<false, true>(stream, grid, tmA, tmB, p, splitk, ws);
    else *err = tgc_launch<true, true>(stream, grid, tmA, tmB, p, splitk, ws);
    return true;
}

}  // namespace mak
