// Warp-per-block QR kernels for tiny blocks (m, n <= 32).  Device code only, written against the
// subset of CUDA that tests/cpu_harness/cuda_emu.h emulates, so the same header is compiled by nvcc
// into libmakb200 and by g++ into the CPU logic tests (tests/test_emu_kernels_cpu.py).
#pragma once
#include "devutil.cuh"
#include "batched_desc.h"

namespace mak {

// ---------------------------------------------------------------------------------------
// warp-per-block QR for tiny blocks (m, n <= 32): four blocks per 128-thread CTA, each warp owns
// one block in its own shared-memory region.  lane = column for the reflector application (every
// lane accumulates ITS column's dot product serially over the rows: no cross-lane reduction, the
// reflector entries are shared-memory broadcasts), lane = row for loads/stores and the scaling of
// v.  No block barrier at all.  This is the HBM end of the batched config: at n ~ 24 ComplexF64
// the arithmetic intensity (2/9 n flop/B) equals the machine balance, so FP64 issue and HBM bound
// the kernel together.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
batched_qr_warp_kernel(const QrBlockDesc<T>* __restrict__ descs, int batch, int cap_elems) {
    MAK_DYN_SMEM(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int blk = blockIdx.x * 4 + warp;
    if (blk >= batch) return;
    const QrBlockDesc<T> d = descs[blk];
    const int m = d.m, n = d.n, k = m < n ? m : n;
    const int lds = m | 1;
    T* S = reinterpret_cast<T*>(smem_raw) + (size_t)warp * cap_elems;
    // load (lane = row), tail norms of every column (lane = column)
    for (int c = 0; c < n; ++c)
        if (lane < m) S[c * lds + lane] = d.A[(size_t)c * d.lda + lane];
    __syncwarp();
    T* mycol = S + (lane < n ? lane : 0) * lds;
    double mynrm = 0.0;
    for (int r = 1; r < m; ++r) mynrm += abs2_(mycol[r]);
    T mytau = zero<T>();
    // ---- factorization ----
    for (int j = 0; j < k; ++j) {
        T* cj = S + j * lds;
        const double sg = __shfl_sync(0xffffffffu, mynrm, j);
        double beta; T tau, scale;
        larfgp_scalars<T>(cj[j], sg, beta, tau, scale);
        __syncwarp();
        if (lane > j && lane < m) cj[lane] = mul_(cj[lane], scale);   // v in place
        if (lane == j) { cj[j] = mk<T>(beta); mytau = tau; }
        __syncwarp();
        if (lane > j && lane < n) {
            T s0 = mycol[j], s1 = zero<T>();
            int r = j + 1;
            for (; r + 1 < m; r += 2) { fmac_(s0, cj[r], mycol[r]); fmac_(s1, cj[r + 1], mycol[r + 1]); }
            if (r < m) fmac_(s0, cj[r], mycol[r]);
            const T f = mul_(conj_(tau), add_(s0, s1));
            mycol[j] = sub_(mycol[j], f);
            double nrm = 0.0;
            for (r = j + 1; r < m; ++r) {
                const T x = sub_(mycol[r], mul_(f, cj[r]));
                mycol[r] = x;
                if (r > j + 1) nrm += abs2_(x);
            }
            mynrm = nrm;   // tail norm below row j+1: consumed when this lane's column is the pivot
        }
        __syncwarp();
    }
    // ---- R out (lane = row) ----
    if (d.R) {
        for (int c = 0; c < n; ++c)
            if (lane < k) d.R[(size_t)c * d.ldr + lane] = (lane <= c) ? S[c * lds + lane] : zero<T>();
    }
    __syncwarp();
    // ---- Q in place (columns 0..k-1), backward accumulation ----
    for (int j = k - 1; j >= 0; --j) {
        T* cj = S + j * lds;
        const T tau = shfl_(mytau, j);
        if (lane > j && lane < k) {
            T s0 = mycol[j], s1 = zero<T>();
            int r = j + 1;
            for (; r + 1 < m; r += 2) { fmac_(s0, cj[r], mycol[r]); fmac_(s1, cj[r + 1], mycol[r + 1]); }
            if (r < m) fmac_(s0, cj[r], mycol[r]);
            const T f = mul_(tau, add_(s0, s1));
            mycol[j] = sub_(mycol[j], f);
            for (r = j + 1; r < m; ++r) mycol[r] = sub_(mycol[r], mul_(f, cj[r]));
        }
        __syncwarp();
        if (lane < m) {
            T x;
            if (lane < j) x = zero<T>();
            else if (lane == j) x = sub_(one<T>(), tau);
            else x = neg_(mul_(tau, cj[lane]));
            cj[lane] = x;
        }
        __syncwarp();
    }
    for (int c = 0; c < k; ++c)
        if (lane < m) d.Q[(size_t)c * d.ldq + lane] = S[c * lds + lane];
}

// Register-resident variant (opt-in, MAKB200_BQR_WARP_REG=1): lane = column and the column LIVES in
// registers (M = compile-time row capacity); the pivot column is published raw to a per-warp
// shared-memory vector (one STS by the pivot lane, broadcast LDS by the others) and the reflector's
// `scale` is folded into the two scalars of the step, so a (row, step) costs 10 DFMA + 3 LSU
// wavefronts instead of the 14 LSU wavefronts that bound the shared-memory version.
template <typename T, int M>
__global__ void __launch_bounds__(128)
batched_qr_warp_reg_kernel(const QrBlockDesc<T>* __restrict__ descs, int batch) {
    __shared__ T vsm[4][M];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int blk = blockIdx.x * 4 + warp;
    if (blk >= batch) return;
    const QrBlockDesc<T> d = descs[blk];
    const int m = d.m, n = d.n, k = m < n ? m : n;
    T* vs = vsm[warp];
    T col[M];
    {
        const T* src = d.A + (size_t)(lane < n ? lane : 0) * d.lda;
#pragma unroll
        for (int r = 0; r < M; ++r) col[r] = (lane < n && r < m) ? src[r] : zero<T>();
    }
    T mytau = zero<T>(), myscale = zero<T>();
    // ---- factorization ----
    for (int j = 0; j < k; ++j) {
        double sig = 0.0;
        T piv = zero<T>();
#pragma unroll
        for (int r = 0; r < M; ++r) {
            if (r > j) sig += abs2_(col[r]);
            if (r == j) piv = col[r];
        }
        sig = __shfl_sync(0xffffffffu, sig, j);
        const T alpha = shfl_(piv, j);
        double beta; T tau, scale;
        larfgp_scalars<T>(alpha, sig, beta, tau, scale);
        if (lane == j) {
#pragma unroll
            for (int r = 0; r < M; ++r) {
                if (r > j) vs[r] = col[r];
                if (r == j) col[r] = mk<T>(beta);
            }
            mytau = tau; myscale = scale;
        }
        __syncwarp();
        if (lane > j && lane < n) {
            T s0 = zero<T>(), s1 = zero<T>();
#pragma unroll
            for (int r = 0; r < M; ++r)
                if (r > j) { if (r & 1) fmac_(s1, vs[r], col[r]); else fmac_(s0, vs[r], col[r]); }
            // s = c_j + conj(scale) * sum conj(a_r) c_r ;  f = conj(tau) s ;  g = f * scale
            T sdot = add_(s0, s1), st = piv;
            fmac_(st, scale, sdot);
            const T f = mul_(conj_(tau), st), g = mul_(f, scale);
#pragma unroll
            for (int r = 0; r < M; ++r) {
                if (r == j) col[r] = sub_(col[r], f);
                if (r > j) col[r] = sub_(col[r], mul_(g, vs[r]));
            }
        }
        __syncwarp();
    }
    // ---- R out: lane = column, rows 0..k-1 ----
    if (d.R && lane < n) {
        T* dst = d.R + (size_t)lane * d.ldr;
#pragma unroll
        for (int r = 0; r < M; ++r)
            if (r < k) dst[r] = (r <= lane) ? col[r] : zero<T>();
    }
    // ---- Q in place (columns 0..k-1), backward accumulation; v_j = scale_j * (raw column below j) ----
    for (int j = k - 1; j >= 0; --j) {
        const T tau = shfl_(mytau, j), scale = shfl_(myscale, j);
        if (lane == j) {
#pragma unroll
            for (int r = 0; r < M; ++r)
                if (r > j) vs[r] = col[r];
        }
        __syncwarp();
        if (lane > j && lane < k) {
            T s0 = zero<T>(), s1 = zero<T>(), qj = zero<T>();
#pragma unroll
            for (int r = 0; r < M; ++r) {
                if (r == j) qj = col[r];
                if (r > j) { if (r & 1) fmac_(s1, vs[r], col[r]); else fmac_(s0, vs[r], col[r]); }
            }
            T sdot = add_(s0, s1), st = qj;
            fmac_(st, scale, sdot);
            const T f = mul_(tau, st), g = mul_(f, scale);
#pragma unroll
            for (int r = 0; r < M; ++r) {
                if (r == j) col[r] = sub_(col[r], f);
                if (r > j) col[r] = sub_(col[r], mul_(g, vs[r]));
            }
        } else if (lane == j) {
            const T ts = neg_(mul_(tau, scale));
#pragma unroll
            for (int r = 0; r < M; ++r) {
                if (r < j) col[r] = zero<T>();
                else if (r == j) col[r] = sub_(one<T>(), tau);
                else col[r] = mul_(ts, col[r]);
            }
        }
        __syncwarp();
    }
    if (lane < k) {
        T* dst = d.Q + (size_t)lane * d.ldq;
#pragma unroll
        for (int r = 0; r < M; ++r)
            if (r < m) dst[r] = col[r];
    }
}

// Panel-blocked variant (opt-in, MAKB200_BQR_WARP_BLK=1, round-2 bring-up): the shared-memory kernel above
// moves 14 LSU wavefronts per (row, step) - two passes over the lane's column per reflector - against 5 clk
// of DFMA.  Here reflectors are generated four at a time with lane = ROW (the 4-column panel lives in
// registers, norms and dots are warp reductions, the 4x4 compact-WY factor Tf is built on the fly), and the
// trailing columns (lane = COLUMN) take the whole block reflector in TWO passes:
//   s = V^H c (4 dots in one sweep),  w = Tf^H s,  c -= V w
// i.e. (8 + 12)/4 = 5 wavefronts per (row, step).  Q is accumulated backwards panel by panel the same way
// (c -= V (Tf (V^H c))), the panel's own columns are E - V (Tf V_top^H) with lane = row.
// Layout: per warp the padded block S (lds = m|1) followed by 8 x 16 entries for the Tf of every panel.
constexpr int BQW_NB = 4;
constexpr int BQW_TF_ELEMS = 8 * BQW_NB * BQW_NB;   // 32 / NB panels at most

template <typename T>
__global__ void __launch_bounds__(128)
batched_qr_warp_blk_kernel(const QrBlockDesc<T>* __restrict__ descs, int batch, int cap_elems) {
    MAK_DYN_SMEM(smem_raw);
    constexpr int NB = BQW_NB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int blk = blockIdx.x * 4 + warp;
    if (blk >= batch) return;
    const QrBlockDesc<T> d = descs[blk];
    const int m = d.m, n = d.n, k = m < n ? m : n;
    const int lds = m | 1;
    T* S = reinterpret_cast<T*>(smem_raw) + (size_t)warp * (cap_elems + BQW_TF_ELEMS);
    T* Tst = S + cap_elems;
    for (int c = 0; c < n; ++c)
        if (lane < m) S[c * lds + lane] = d.A[(size_t)c * d.lda + lane];
    __syncwarp();
    T* mycol = S + (lane < n ? lane : 0) * lds;
    const int npanels = (k + NB - 1) / NB;

    // ---------------- factorization ----------------
    for (int p = 0; p < npanels; ++p) {
        const int j0 = p * NB, pb = (k - j0 < NB) ? (k - j0) : NB;
        // ---- panel (lane = row): generate pb reflectors and Tf ----
        T a[NB], Tf[NB][NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            a[i] = (i < pb && lane >= j0 && lane < m) ? S[(j0 + i) * lds + lane] : zero<T>();
#pragma unroll
            for (int q = 0; q < NB; ++q) Tf[q][i] = zero<T>();
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            if (i < pb) {                                   // uniform across the warp
                const int jc = j0 + i;
                const double sigma = warp_sum((lane > jc && lane < m) ? abs2_(a[i]) : 0.0);
                const T alpha = shfl_(a[i], jc);
                double beta; T tau, scale;
                larfgp_scalars<T>(alpha, sigma, beta, tau, scale);
                if (lane > jc && lane < m) a[i] = mul_(a[i], scale);
                if (lane == jc) a[i] = mk<T>(beta);
                // v_i as seen by this lane: 0 above the pivot row, 1 on it, the scaled entry below
                const T vi = (lane < jc || lane >= m) ? zero<T>() : (lane == jc ? one<T>() : a[i]);
                // apply H_i^H to the rest of the panel
#pragma unroll
                for (int l = 0; l < NB; ++l) {
                    if (l > i && l < pb) {
                        T sl = zero<T>();
                        fmac_(sl, vi, a[l]);
                        sl = warp_sum(sl);
                        const T f = mul_(conj_(tau), sl);
                        a[l] = sub_(a[l], mul_(f, vi));
                    }
                }
                // Tf(0:i, i) = -tau Tf(0:i, 0:i) (V(:, 0:i)^H v_i),  Tf(i, i) = tau
                T z[NB];
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    z[q] = zero<T>();
                    if (q < i) {
                        // V(:, q) at this lane: rows above j0+q are R entries (not V): only rows >= jc matter, all below j0+q
                        const T vq = (lane < jc || lane >= m) ? zero<T>() : a[q];
                        T zz = zero<T>();
                        fmac_(zz, vq, vi);
                        z[q] = warp_sum(zz);
                    }
                }
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    if (q < i) {
                        T acc = zero<T>();
#pragma unroll
                        for (int l = 0; l < NB; ++l)
                            if (l >= q && l < i) fma_(acc, Tf[q][l], z[l]);
                        Tf[q][i] = neg_(mul_(tau, acc));
                    }
                }
                Tf[i][i] = tau;
            }
        }
        // write the panel back (rows >= j0: R entries of the panel's upper triangle, beta, v) and Tf
#pragma unroll
        for (int i = 0; i < NB; ++i)
            if (i < pb && lane >= j0 && lane < m) S[(j0 + i) * lds + lane] = a[i];
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < NB; ++q)
#pragma unroll
                for (int i = 0; i < NB; ++i) Tst[p * NB * NB + q * NB + i] = Tf[q][i];
        }
        __syncwarp();
        // ---- trailing columns (lane = column): c <- (I - V Tf^H V^H) c ----
        const int c0 = j0 + pb;
        if (lane >= c0 && lane < n) {
            T sdot[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) sdot[i] = zero<T>();
#pragma unroll
            for (int tt = 0; tt < NB; ++tt) {               // triangle rows j0 .. j0+pb-1 (unit lower part of V)
                if (tt < pb) {
                    const T x = mycol[j0 + tt];
#pragma unroll
                    for (int i = 0; i < NB; ++i) {
                        if (i == tt) sdot[i] = add_(sdot[i], x);
                        else if (i < tt) fmac_(sdot[i], S[(j0 + i) * lds + j0 + tt], x);
                    }
                }
            }
            for (int r = c0; r < m; ++r) {
                const T x = mycol[r];
#pragma unroll
                for (int i = 0; i < NB; ++i)
                    if (i < pb) fmac_(sdot[i], S[(j0 + i) * lds + r], x);
            }
            T w[NB];                                        // w = Tf^H s
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                w[i] = zero<T>();
#pragma unroll
                for (int q = 0; q < NB; ++q)
                    if (q <= i && i < pb) fmac_(w[i], Tst[p * NB * NB + q * NB + i], sdot[q]);
            }
#pragma unroll
            for (int tt = 0; tt < NB; ++tt) {
                if (tt < pb) {
                    T x = mycol[j0 + tt];
#pragma unroll
                    for (int i = 0; i < NB; ++i) {
                        if (i == tt) x = sub_(x, w[i]);
                        else if (i < tt) x = sub_(x, mul_(S[(j0 + i) * lds + j0 + tt], w[i]));
                    }
                    mycol[j0 + tt] = x;
                }
            }
            for (int r = c0; r < m; ++r) {
                T x = mycol[r];
#pragma unroll
                for (int i = 0; i < NB; ++i)
                    if (i < pb) x = sub_(x, mul_(S[(j0 + i) * lds + r], w[i]));
                mycol[r] = x;
            }
        }
        __syncwarp();
    }
    // ---------------- R out (lane = row) ----------------
    if (d.R) {
        for (int c = 0; c < n; ++c)
            if (lane < k) d.R[(size_t)c * d.ldr + lane] = (lane <= c) ? S[c * lds + lane] : zero<T>();
    }
    __syncwarp();
    // ---------------- Q in place (columns 0..k-1), backward over the panels ----------------
    for (int p = npanels - 1; p >= 0; --p) {
        const int j0 = p * NB, pb = (k - j0 < NB) ? (k - j0) : NB;
        const int c0 = j0 + pb;
        // ---- already formed columns (lane = column): c <- (I - V Tf V^H) c; their rows < c0 are zero ----
        if (lane >= c0 && lane < k) {
            T sdot[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) sdot[i] = zero<T>();
            for (int r = c0; r < m; ++r) {
                const T x = mycol[r];
#pragma unroll
                for (int i = 0; i < NB; ++i)
                    if (i < pb) fmac_(sdot[i], S[(j0 + i) * lds + r], x);
            }
            T w[NB];                                        // w = Tf s
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                w[i] = zero<T>();
#pragma unroll
                for (int q = 0; q < NB; ++q)
                    if (q >= i && q < pb) fma_(w[i], Tst[p * NB * NB + i * NB + q], sdot[q]);
            }
#pragma unroll
            for (int tt = 0; tt < NB; ++tt) {
                if (tt < pb) {
                    T x = zero<T>();                        // rows j0..c0-1 of a formed column are zero on entry
#pragma unroll
                    for (int i = 0; i < NB; ++i) {
                        if (i == tt) x = sub_(x, w[i]);
                        else if (i < tt) x = sub_(x, mul_(S[(j0 + i) * lds + j0 + tt], w[i]));
                    }
                    mycol[j0 + tt] = x;
                }
            }
            for (int r = c0; r < m; ++r) {
                T x = mycol[r];
#pragma unroll
                for (int i = 0; i < NB; ++i)
                    if (i < pb) x = sub_(x, mul_(S[(j0 + i) * lds + r], w[i]));
                mycol[r] = x;
            }
        }
        __syncwarp();
        // ---- the panel's own columns (lane = row): Q(:, j0+c) = e_{j0+c} - V (Tf V_top^H)(:, c) ----
        T vrow[NB], Mx[NB][NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            T x = zero<T>();
            if (i < pb && lane < m) {
                if (lane == j0 + i) x = one<T>();
                else if (lane > j0 + i) x = S[(j0 + i) * lds + lane];
            }
            vrow[i] = x;
        }
        // u_c = V_top^H e_c = conj(row c of the unit lower triangle);  Mx(:, c) = Tf u_c
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            T u[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                u[i] = zero<T>();
                if (c < pb) {
                    if (i == c) u[i] = one<T>();
                    else if (i < c) u[i] = conj_(S[(j0 + i) * lds + j0 + c]);
                }
            }
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                T acc = zero<T>();
#pragma unroll
                for (int i = 0; i < NB; ++i)
                    if (i >= q && i <= c && c < pb) fma_(acc, Tst[p * NB * NB + q * NB + i], u[i]);
                Mx[q][c] = acc;
            }
        }
        __syncwarp();                                       // every lane has read V and V_top before the overwrite
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            if (c < pb && lane < m) {
                T x = (lane == j0 + c) ? one<T>() : zero<T>();
#pragma unroll
                for (int i = 0; i < NB; ++i) x = sub_(x, mul_(vrow[i], Mx[i][c]));
                S[(j0 + c) * lds + lane] = x;
            }
        }
        __syncwarp();
    }
    for (int c = 0; c < k; ++c)
        if (lane < m) d.Q[(size_t)c * d.ldq + lane] = S[c * lds + lane];
}

}  // namespace mak
