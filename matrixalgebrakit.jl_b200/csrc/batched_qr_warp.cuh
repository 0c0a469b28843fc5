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

}  // namespace mak
