// Persistent, TMA-fed column kernels of the Hermitian tridiagonalisation (included by eigh.cu).
//
// Round 1 streamed the trailing matrix with ~2700 short-lived CTAs per column (128 KB each between a
// reflector prologue and a reduction epilogue): 0.36 of the measured HBM rate, 24 % warps active (ncu,
// profiles/r1_ncu_summaries.txt).  Here ONE wave of 2 CTAs per SM runs per column.  Every CTA owns a
// contiguous chunk of the band-major tile list (trd_tiles.h) and keeps a 3-stage shared-memory ring full
// for the whole launch: a producer lane issues cp.async.bulk.tensor (one 256-row x CW-column tile = 32 KB
// per stage, no register cost, ~190 KB of loads in flight per SM), eight consumer warps do the two FMAs
// per element out of shared memory (lane = row).  The ring is filled before the reflector scalars are
// known (the loads do not depend on them), so the prologue latency is hidden.
//
//   per tile    column parts  ycol[J][k] = sum_{r in band J} conj(A[r,k]) v[r]   (butterfly reduction, one
//               named barrier per tile), row part accumulated in one register per thread
//   per band    row parts     yrow[g][r] = sum_{k in chunk g} A[r,k] v[k]  (+ the real diagonal term)
//   per CTA     v^H A22 v partial; V(:, i) over its row slice; CTA g < 2i finishes panel dot product g
//               (t1 = W^H v, t2 = V^H v) from the per-CTA partials the previous w kernel left behind
//
// trd_w2_kernel then forms w, writes the reflector back, updates the next column and its partial norms; per row
// it sums <= (chunks per band) + (bands) partials instead of the n/16 strips of round 1.  The panel rows it has
// in hand anyway (V[r,p], W[r,p]) also give the NEXT column's panel dot products up to the reflector scale,
// which is not known yet:  W^H v' = conj(W[row0',:]) + scale' * W^H a'[row0'+1:]  - so they cost no extra pass
// over the panel and no extra launch; no atomics anywhere (bit-reproducible).
#pragma once
#include "trd_tiles.h"

namespace mak {

constexpr int TRD2_BH = 256;
constexpr int TRD2_NST = 3;
template <typename T> struct Trd2SV { static constexpr int value = 1024; };     // entries of v staged per run
template <> struct Trd2SV<cplx> { static constexpr int value = 512; };

__device__ __forceinline__ double shfl_xor_(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ cplx shfl_xor_(cplx v, int o) {
    return cplx{__shfl_xor_sync(0xffffffffu, v.re, o), __shfl_xor_sync(0xffffffffu, v.im, o)};
}

// Sum c[0..CNT) over the 32 lanes with 2*(CNT-1)+... shuffles instead of 5*CNT: after the call lane l holds in c[0]
// the total of entry trd2_owned<CNT>(l) (pairs/quads of lanes hold the same entry).
template <typename T, int CNT, int O>
struct MultiReduce {
    static __device__ __forceinline__ void run(T* c, int lane) {
        if constexpr (CNT > 1) {
            constexpr int H = CNT / 2;
            const bool up = (lane & O) != 0;
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const T send = up ? c[k] : c[k + H];
                const T keep = up ? c[k + H] : c[k];
                c[k] = add_(keep, shfl_xor_(send, O));
            }
            MultiReduce<T, H, O / 2>::run(c, lane);
        } else if constexpr (O >= 1) {
            c[0] = add_(c[0], shfl_xor_(c[0], O));
            MultiReduce<T, 1, O / 2>::run(c, lane);
        }
    }
};
template <int CNT>
__device__ __forceinline__ int trd2_owned(int lane) {
    int k = 0, o = 16;
#pragma unroll
    for (int h = CNT / 2; h >= 1; h >>= 1, o >>= 1)
        if (lane & o) k += h;
    return k;
}
// first lane of the group that holds entry k: the bits used by trd2_owned, the rest zero
template <int CNT>
__device__ __forceinline__ bool trd2_owner_lane(int lane) {
    int used = 0, o = 16;
#pragma unroll
    for (int h = CNT / 2; h >= 1; h >>= 1, o >>= 1) used |= o;
    return (lane & ~used) == 0;
}

__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// v[g] of the current reflector (global row g): 1 at row0, scaled column below, 0 outside the trailing block
template <typename T>
__device__ __forceinline__ T trd2_v(const T* __restrict__ Acol, int g, int row0, int n, T scale) {
    if (g < row0 || g >= n) return zero<T>();
    if (g == row0) return one<T>();
    return mul_(Acol[g], scale);
}

template <typename T>
__global__ void __launch_bounds__(288, 2)
trd_symv2_kernel(const __grid_constant__ CUtensorMap tmap, TrdCtx<T> x, int c, int i, int npn, int early) {
    constexpr int CW = SymvCW<T>::value;
    constexpr int BH = TRD2_BH;
    constexpr int NBOX = sizeof(T) / 8;
    constexpr int RB = BH / NBOX;                  // matrix rows per TMA box
    constexpr unsigned STAGE_BYTES = BH * CW * sizeof(T);
    constexpr int SV = Trd2SV<T>::value;
    constexpr int SVT = SV / CW;                   // tiles per staged run of v
    // `early` (set for every column but the first of a panel): the previous kernel is the w kernel of column c - 1, which
    // writes only column c of A and the panel buffers - not the trailing block the tiles cover (the one stored column of
    // a boundary strip that it does touch is masked out of the sums).  The producer may then fill the ring BEFORE the
    // programmatic dependency resolves: ~190 KB per SM stream in under the w kernel, and the consumers find their first
    // tiles waiting.  The consumers - who read what the w kernel wrote - wait as before.
    if (!early) pdl_wait_then_trigger();
    extern __shared__ __align__(128) unsigned char trd2_smem_raw[];
    unsigned char* ring = trd2_smem_raw + ((128u - (smem_u32(trd2_smem_raw) & 127u)) & 127u);
    __shared__ T s_v[SV];
    __shared__ T s_col[2][8][CW];
    __shared__ T s_vb[TRD2_BH];                    // v over the rows of the current band (column pass)
    __shared__ double s_q[8];
    __shared__ __align__(8) uint64_t full_bar[TRD2_NST], empty_bar[TRD2_NST];

    const int n = x.n, row0 = c + 1, mt = n - row0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x, G = gridDim.x;
    const TrdTiling tl = trd_tiling(n, row0, BH, CW, G);
    const int t_beg = min(tl.NT, g * tl.q), t_end = min(tl.NT, t_beg + tl.q);

    if (tid == 0) {
        for (int st = 0; st < TRD2_NST; ++st) { mbar_init(&full_bar[st], 1); mbar_init(&empty_bar[st], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp == 8) {
        // ---- producer: keep the ring full with this chunk's tiles ----
        if (lane == 0 && t_beg < t_end) {
            int J = trd_tile_band(tl, t_beg);
            int S = tl.S0 + (t_beg - trd_band_first_tile(tl, J));
            int Slast = trd_band_last_strip(tl, J);
            for (int t = t_beg, it = 0; t < t_end; ++t, ++it) {
                const int st = it % TRD2_NST, use = it / TRD2_NST;
                if (use > 0) mbar_wait(&empty_bar[st], (unsigned)((use - 1) & 1));
                mbar_expect_tx(&full_bar[st], STAGE_BYTES);
                unsigned char* dst = ring + (size_t)st * STAGE_BYTES;
#pragma unroll
                for (int bx = 0; bx < NBOX; ++bx)
                    tma_load_2d(dst + (size_t)bx * (STAGE_BYTES / NBOX), &tmap, &full_bar[st], (BH * J) * NBOX + bx * 256,
                                CW * S);
                if (++S > Slast) { ++J; S = tl.S0; Slast = (J < tl.JB) ? trd_band_last_strip(tl, J) : 0; }
            }
        }
        return;   // the producer warp takes no part in the consumer barriers (bar 1) below
    }

    // ---- consumers ----
    if (early) pdl_wait_then_trigger();
    // The launch is a chain of L2 latencies before the first tile is touched (ncu: 12 % of the stall samples sat in this
    // prologue), so every load whose ADDRESS is known is put in flight first: the partial norms, alpha, the raw column
    // entries this thread will scale into v (its row of the first band, its entries of the first staged run of strips,
    // its row of the V(:, i) slice) and the panel-dot partials; the scale-dependent arithmetic follows.
    const T* Acol = x.A + (size_t)c * x.lda;
    {
        const int J1 = (t_beg < t_end) ? trd_tile_band(tl, t_beg) : tl.J0;
        const int S1 = (t_beg < t_end) ? tl.S0 + (t_beg - trd_band_first_tile(tl, J1)) : tl.S0;
        const int g1 = BH * J1 + tid;
        if (g1 > row0 && g1 < n) asm volatile("prefetch.global.L1 [%0];" ::"l"(Acol + g1));
        for (int e = tid; e < SV; e += 256) {
            const int gcol = CW * S1 + e;
            if (gcol > row0 && gcol < n) asm volatile("prefetch.global.L1 [%0];" ::"l"(Acol + gcol));
        }
    }
    // reflector scalars: every warp sums the partial norms itself in the same fixed order (loads first, then the sum)
    double sigma = 0.0;
    {
        constexpr int PU = 8;
        for (int q0 = 0; q0 < npn; q0 += 32 * PU) {
            double pv[PU];
#pragma unroll
            for (int u = 0; u < PU; ++u) {
                const int q = q0 + lane + 32 * u;
                pv[u] = q < npn ? x.pn[q] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < PU; ++u) sigma += pv[u];
        }
    }
    const T alpha = Acol[row0];
    // panel dot product number g: the partials of the previous w kernel's CTAs (rows >= row0 + 1 of the unscaled column)
    const bool tdot = g < 2 * i;
    const bool isw = g < i;
    const int tp = isw ? g : g - i;
    T tacc = zero<T>(), tfirst = zero<T>();
    if (tdot) {
        const T* src = x.tpart + (size_t)(isw ? tp : TRD_NB + tp) * x.tpld;
        for (int b = tid; b < npn; b += 256) tacc = add_(tacc, src[b]);
        if (tid == 0) tfirst = conj_(x.P[(size_t)(isw ? x.pw + tp : tp) * x.ldp + row0]);
    }
    int r_lo, r_hi;
    trd_slice(tl, g, r_lo, r_hi);
    const int slen = r_hi - r_lo;                  // <= 64 (G >= mt / 64 is checked on the host)
    const int sgr = row0 + r_lo + tid;
    T sraw = zero<T>();
    if (tid < slen && sgr > row0 && sgr < n) sraw = Acol[sgr];
    sigma = warp_sum(sigma);
    double beta; T tau, scale;
    larfgp_scalars<T>(alpha, sigma, beta, tau, scale);

    // V(:, i) over this CTA's row slice (two copies in the panel buffer)
    if (tid < slen) {
        const T vq = (sgr == row0) ? one<T>() : ((sgr > row0 && sgr < n) ? mul_(sraw, scale) : zero<T>());
        x.P[(size_t)i * x.ldp + sgr] = vq;
        x.P[(size_t)(2 * x.pw + i) * x.ldp + sgr] = vq;
    }
    // ... times scale, plus the row-row0 term (v[row0] = 1)
    if (tdot) {
        tacc = warp_sum(tacc);
        if (lane == 0) s_col[0][warp][0] = tacc;
        bar_consumers();
        if (tid == 0) {
            T tot = s_col[0][0][0];
#pragma unroll
            for (int w = 1; w < 8; ++w) tot = add_(tot, s_col[0][w][0]);
            x.t[isw ? tp : TRD_NB + tp] = add_(tfirst, mul_(scale, tot));
        }
    }

    double qacc = 0.0;
    if (t_beg < t_end) {
        constexpr int CPW = CW / 8;                  // tile columns owned by a warp in the column pass
        constexpr int RPL = BH / 32;                 // rows per lane in the column pass
        int J = trd_tile_band(tl, t_beg);
        int S = tl.S0 + (t_beg - trd_band_first_tile(tl, J));
        int Slast = trd_band_last_strip(tl, J);
        const int rr = tid;                          // row pass: thread = row of the tile
        T vr = zero<T>(), rowacc = zero<T>();
        T vcol[RPL];                                 // column pass: v at rows lane + 32 i of the band
        double adiag = 0.0;
        bool band_open = false;
        int run_left = 0, sv_base = 0;               // staged strips of v: global columns [sv_base, sv_base + CW*run)
        for (int t = t_beg, it = 0; t < t_end; ++t, ++it) {
            const int gr = BH * J + rr;
            if (!band_open) {
                vr = trd2_v<T>(Acol, gr, row0, n, scale);
                rowacc = zero<T>();
                adiag = 0.0;
                band_open = true;
                run_left = 0;
                bar_consumers();                     // previous band's s_vb has been read by everybody
                s_vb[tid] = vr;
                bar_consumers();
#pragma unroll
                for (int q = 0; q < RPL; ++q) vcol[q] = s_vb[lane + 32 * q];
            }
            if (run_left == 0) {
                // stage v for the next strips of this band inside the chunk
                int cnt = min(min(Slast - S + 1, t_end - t), SVT);
                bar_consumers();                     // everybody is done with the previous contents of s_v
                sv_base = CW * S;
                for (int e = tid; e < cnt * CW; e += 256) s_v[e] = trd2_v<T>(Acol, sv_base + e, row0, n, scale);
                bar_consumers();
                run_left = cnt;
            }
            const int st = it % TRD2_NST, use = it / TRD2_NST;
            mbar_wait(&full_bar[st], (unsigned)(use & 1));
            const T* tile = reinterpret_cast<const T*>(ring + (size_t)st * STAGE_BYTES);
            const T* src = tile + (size_t)(rr / RB) * (CW * RB) + (rr % RB);
            const T* sv = s_v + (CW * S - sv_base);
            const int gc0 = CW * S;
            const bool interior = gc0 + CW - 1 < BH * J;   // every column of the strip is left of every row of the band
            // ---- row pass: y_row[r] += sum_k A[r,k] v[k]  (two accumulators: half the dependent chain) ----
            T racc2 = zero<T>();
            if (interior) {
#pragma unroll
                for (int k = 0; k < CW; k += 2) {
                    fma_(rowacc, src[k * RB], sv[k]);
                    fma_(racc2, src[(k + 1) * RB], sv[k + 1]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < CW; ++k) {
                    const int gc = gc0 + k;
                    if (gr > gc) fma_(rowacc, src[k * RB], sv[k]);
                    else if (gr == gc) adiag = real_(src[k * RB]);   // Hermitian: real diagonal
                }
            }
            rowacc = add_(rowacc, racc2);
            // ---- column pass: warp w owns tile columns [w CPW, (w+1) CPW): y_col[k] = sum_r conj(A[r,k]) v[r] over the
            // band's 256 rows, 8 per lane, one warp reduction per column and NO cross-warp step: the warps never meet
            // inside a run of strips, each releases the stage when it is done with it ----
#pragma unroll
            for (int kk = 0; kk < CPW; ++kk) {
                const int k = warp * CPW + kk, gc = gc0 + k;
                T cacc = zero<T>();
#pragma unroll
                for (int q = 0; q < RPL; ++q) {
                    const int r2 = lane + 32 * q;
                    const T a = tile[(size_t)(r2 / RB) * (CW * RB) + k * RB + (r2 % RB)];
                    if (interior || BH * J + r2 > gc) fmac_(cacc, a, vcol[q]);
                }
                cacc = warp_sum(cacc);
                if (lane == 0 && gc < n) x.ycolp[(size_t)J * x.ldy + gc] = cacc;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
            --run_left;
            const bool band_end = (S == Slast) || (t + 1 == t_end);
            if (band_end) {
                // row part of this band from this chunk (+ the diagonal term, which lives in exactly one chunk)
                const T yr = add_(rowacc, scale_(vr, adiag));
                if (gr < n) x.yrowp[(size_t)g * x.ldy + gr] = yr;
                T cv = zero<T>();
                fmac_(cv, vr, rowacc);               // conj(v_r) * (strictly-lower row part)
                qacc += 2.0 * real_(cv) + adiag * abs2_(vr);
                band_open = false;
            }
            if (S == Slast) { ++J; S = tl.S0; Slast = (J < tl.JB) ? trd_band_last_strip(tl, J) : 0; }
            else ++S;
        }
    }
    qacc = warp_sum(qacc);
    if (lane == 0) s_q[warp] = qacc;
    bar_consumers();
    if (tid == 0) {
        double qs = 0.0;
        for (int w = 0; w < 8; ++w) qs += s_q[w];
        x.pyv[g] = mk<T>(qs);                        // v^H A22 v is real
        if (g == 0) {
            x.tau[c] = tau;
            x.e[c] = beta;
            x.d[c] = real_(x.A[(size_t)c * x.lda + c]);
        }
    }
}

// y[r] for global row r: row parts of band(r) from its chunks + column parts from bands >= band(r)
template <typename T>
__device__ __forceinline__ T trd2_ysum(const TrdCtx<T>& x, const TrdTiling& tl, int r, int part, int nparts) {
    const int J = r / tl.BH;
    int g_lo, g_hi;
    trd_band_chunks(tl, J, g_lo, g_hi);
    T s0 = zero<T>(), s1 = zero<T>();
    for (int gg = g_lo + part; gg <= g_hi; gg += nparts) s0 = add_(s0, x.yrowp[(size_t)gg * x.ldy + r]);
    for (int JJ = J + part; JJ < tl.JB; JJ += nparts) s1 = add_(s1, x.ycolp[(size_t)JJ * x.ldy + r]);
    return add_(s0, s1);
}

// 288 threads: warps 0..7 = 32 rows x 8 groups make ONE pass over the panel rows (V[r,p], W[r,p] feed the w update,
// the left-looking update of the next column and - second use of the same rows - the next column's panel dot
// partials) and share the partial sums of y; warp 8 computes the step's scalars concurrently.
template <typename T>
__global__ void __launch_bounds__(288)
trd_w2_kernel(TrdCtx<T> x, int c, int i, int G, int do_next) {
    pdl_wait_then_trigger();
    __shared__ T sm[TRD_K2_NG][TRD_K2_ROWS];
    __shared__ T sm2[TRD_K2_NG][TRD_K2_ROWS];
    __shared__ T st[2 * TRD_NB];       // t1, t2
    __shared__ T srow[2 * TRD_NB + 2]; // conj(W[c1,p]), conj(V[c1,p])
    __shared__ T sa[TRD_K2_ROWS];      // updated next column over this CTA's rows (0 outside its tail)
    __shared__ T sscal[2];
    __shared__ double red[32];
    const int n = x.n, row0 = c + 1, c1 = c + 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool scalar_warp = (warp == 8);
    const int tx = tid % TRD_K2_ROWS, ty = (tid / TRD_K2_ROWS) % TRD_K2_NG;
    const int r = row0 + blockIdx.x * TRD_K2_ROWS + tx;
    const bool live = !scalar_warp && r < n;
    const TrdTiling tl = trd_tiling(n, row0, TRD2_BH, SymvCW<T>::value, G);
    // every global load of this thread is issued before the first barrier (the launch is a chain of L2 latencies, not
    // bytes): the panel rows stay in registers for all three uses
    constexpr int PU = TRD_NB / TRD_K2_NG;            // panel columns per thread
    T vpr[PU], wpr[PU];
#pragma unroll
    for (int u = 0; u < PU; ++u) {
        const int p = ty + u * TRD_K2_NG;
        const bool on = live && p < i;
        vpr[u] = on ? x.P[(size_t)p * x.ldp + r] : zero<T>();
        wpr[u] = on ? x.P[(size_t)(x.pw + p) * x.ldp + r] : zero<T>();
    }
    const T ysum = live ? trd2_ysum<T>(x, tl, r, ty, TRD_K2_NG) : zero<T>();
    const bool fin = live && ty == 0;
    const T vr = fin ? x.P[(size_t)i * x.ldp + r] : zero<T>();
    const T anext = (fin && do_next) ? x.A[(size_t)c1 * x.lda + r] : zero<T>();
    const T tauc = x.tau[c];
    const double ec = x.e[c];
    for (int p = tid; p < i; p += blockDim.x) {
        st[p] = x.t[p];
        st[TRD_NB + p] = x.t[TRD_NB + p];
        srow[p] = conj_(x.P[(size_t)(x.pw + p) * x.ldp + c1]);          // conj(W[c1,p])
        srow[TRD_NB + 1 + p] = conj_(x.P[(size_t)p * x.ldp + c1]);      // conj(V[c1,p])
    }
    T yhv = zero<T>(), y0 = zero<T>();
    if (scalar_warp) {
        // (loads first, then the sums: a fused loop is one L2 round trip per iteration, and this warp's chain - G / 32
        // of them - is the critical path of the launch)
        constexpr int PU = 8;
        y0 = trd2_ysum<T>(x, tl, row0, lane, 32);
        for (int q0 = 0; q0 < G; q0 += 32 * PU) {
            T pv[PU];
#pragma unroll
            for (int u = 0; u < PU; ++u) {
                const int q = q0 + lane + 32 * u;
                pv[u] = q < G ? x.pyv[q] : zero<T>();
            }
#pragma unroll
            for (int u = 0; u < PU; ++u) yhv = add_(yhv, pv[u]);
        }
    }
    __syncthreads();
    T part = zero<T>(), part2 = zero<T>();
    if (scalar_warp) {
        // alpha2 = -(tau/2) * w^H v,  w^H v = conj(tau) * (y^H v - t1^H t2 - t2^H t1)
        T s12 = zero<T>();
        for (int p = lane; p < i; p += 32) {
            fmac_(s12, st[p], st[TRD_NB + p]);
            fmac_(s12, st[TRD_NB + p], st[p]);
        }
        // first row of w (row c+1): y - V t1 - W t2
        T sf = zero<T>();
        for (int p = lane; p < i; p += 32) {
            fma_(sf, x.P[(size_t)p * x.ldp + row0], st[p]);
            fma_(sf, x.P[(size_t)(x.pw + p) * x.ldp + row0], st[TRD_NB + p]);
        }
        yhv = warp_sum(yhv);
        s12 = warp_sum(s12);
        sf = warp_sum(sf);
        y0 = warp_sum(y0);
        if (lane == 0) {
            const T whv = mul_(conj_(tauc), sub_(yhv, s12));
            const T alpha2 = neg_(scale_(mul_(tauc, whv), 0.5));
            sscal[0] = alpha2;
            sscal[1] = add_(mul_(tauc, sub_(y0, sf)), alpha2);          // w[row0], v[row0] = 1
        }
    } else if (live) {
#pragma unroll
        for (int u = 0; u < PU; ++u) {
            const int p = ty + u * TRD_K2_NG;
            if (p < i) {
                fma_(part, vpr[u], st[p]);
                fma_(part, wpr[u], st[TRD_NB + p]);
                fma_(part2, vpr[u], srow[p]);
                fma_(part2, wpr[u], srow[TRD_NB + 1 + p]);
            }
        }
        part = sub_(part, ysum);                                        // w = tau (y - part)
    }
    if (!scalar_warp) { sm[ty][tx] = part; sm2[ty][tx] = part2; }
    if (tid < TRD_K2_ROWS) sa[tid] = zero<T>();
    __syncthreads();
    const T alpha2 = sscal[0], wfirst = sscal[1];
    double nrm = 0.0;
    T wr = zero<T>();
    if (fin) {
        T s = sm[0][tx];
#pragma unroll
        for (int gq = 1; gq < TRD_K2_NG; ++gq) s = add_(s, sm[gq][tx]);
        wr = add_(mul_(tauc, neg_(s)), mul_(alpha2, vr));
        x.P[(size_t)(x.pw + i) * x.ldp + r] = wr;                              // W(:, i)
        x.A[(size_t)c * x.lda + r] = (r == row0) ? mk<T>(ec) : vr;              // reflector storage
        if (do_next) {
            // left-looking update of column c1 = c+1: previous panel columns + the new one
            T s2 = sm2[0][tx];
#pragma unroll
            for (int gq = 1; gq < TRD_K2_NG; ++gq) s2 = add_(s2, sm2[gq][tx]);
            fma_(s2, vr, conj_(wfirst));   // V[r,i] conj(W[c1,i])
            s2 = add_(s2, wr);             // W[r,i] conj(V[c1,i]), V[c1,i] = 1
            T a = sub_(anext, s2);
            if (r == c1) a = mk<T>(real_(a));
            x.A[(size_t)c1 * x.lda + r] = a;
            if (r >= c1 + 2) { nrm = abs2_(a); sa[tx] = a; }
        }
    }
    if (!do_next) return;
    double tot = block_sum<double>(nrm, red);    // (contains the barriers that publish sa)
    if (tid == 0) x.pn[blockIdx.x] = tot;
    // panel dot partials of the next column (rows >= c1 + 2 of its unscaled tail): sum_r conj(W[r,p]) a[r],
    // sum_r conj(V[r,p]) a[r] for p <= i; warp ty reduces over its 32 rows
    if (scalar_warp) return;
    const T ar = sa[tx];
    const size_t blk = blockIdx.x;
#pragma unroll
    for (int u = 0; u < PU; ++u) {
        const int p = ty + u * TRD_K2_NG;
        if (p < i) {                              // uniform per warp
            T dw = zero<T>(), dv = zero<T>();
            fmac_(dw, wpr[u], ar);
            fmac_(dv, vpr[u], ar);
            dw = warp_sum(dw);
            dv = warp_sum(dv);
            if (lane == 0) {
                x.tpart[(size_t)p * x.tpld + blk] = dw;
                x.tpart[(size_t)(TRD_NB + p) * x.tpld + blk] = dv;
            }
        }
    }
    if (ty == 0) {
        T dw = zero<T>(), dv = zero<T>();
        fmac_(dw, wr, ar);                        // the column of W and V this launch produced (p = i)
        fmac_(dv, vr, ar);
        dw = warp_sum(dw);
        dv = warp_sum(dv);
        if (lane == 0) {
            x.tpart[(size_t)i * x.tpld + blk] = dw;
            x.tpart[(size_t)(TRD_NB + i) * x.tpld + blk] = dv;
        }
    }
}

}  // namespace mak
