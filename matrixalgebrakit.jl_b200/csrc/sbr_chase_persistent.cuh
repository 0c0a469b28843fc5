// Persistent form of the bulge-chasing stage (band -> tridiagonal, geometry in sbr_core.h).
//
// chase_wave_kernel (sbr.cu) spends one launch per wavefront 2s+k: 2n launches of ~12 us each at
// n = 8192 although a task is a few microseconds of work.  Here ONE cooperative launch runs the whole
// stage: CTA c owns sweeps c, c+G, c+2G, ... and walks the tasks k = 0, 1, ... of a sweep in order; the
// only inter-CTA dependency, (s, k) after (s-1, k+1), is a per-sweep progress counter in global memory
// (prog[s] = tasks of sweep s completed; release store by the producer, acquire poll by thread 0 of the
// consumer).  Sweep s-1 is owned by CTA c-1 (mod G) and was started before sweep s, so with all G CTAs
// resident (cooperative launch) the wait graph has no cycle.  Inside a sweep the reflector of task k is
// handed to task k+1 in shared memory.  The band is read with L2-only loads (ld.global.cg): every
// element is produced by another SM of the same launch.
//
// Device code only, written against the CUDA subset of tests/cpu_harness/cuda_emu.h: the same header is
// compiled by g++ with all CTAs as co-resident fibers (emu::launch_coresident) and compared with the
// sequential chase of sbr_core.h (tests/test_emu_kernels_cpu.py).
#pragma once
#include "devutil.cuh"
#include "sbr_core.h"

namespace mak {

constexpr int SBRP_THREADS = 256;
constexpr int SBRP_SPIN_LIMIT = 1 << 24;   // polls before a consumer gives up and raises prog[n] (no hang on a logic error)

// shared memory of one CTA in units of T: G, D (b x (b+1) each), two reflectors, pw, two partial-sum
// planes, the block_sum scratch of T and of double (32 doubles fit 32 T)
__host__ __device__ __forceinline__ size_t chase_persistent_smem_elems(int b) {
    return (size_t)2 * b * (b + 1) + 3 * (size_t)b + 2 * SBRP_THREADS + 80;
}

template <typename T>
__global__ void __launch_bounds__(SBRP_THREADS)
chase_persistent_kernel(int n, int b, T* __restrict__ AB, int ldab, T* __restrict__ V2, int ldv, T* __restrict__ tau2,
                        int ldt, int* __restrict__ prog) {
    MAK_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x;
    const int ldg = b + 1;                       // odd leading dimension: conflict-free row and column walks
    T* G = reinterpret_cast<T*>(smem_raw);       // [b][ldg]  (column-major: G[i + j*ldg])
    T* D = G + (size_t)b * ldg;                  // [b][ldg]  full Hermitian
    T* v = D + (size_t)b * ldg;                  // [b] reflector of this task
    T* vp = v + b;                               // [b] reflector of the previous task of the sweep
    T* pw = vp + b;                              // [b] p / w of the two-sided update
    T* ps = pw + b;                              // [256] partial sums
    T* ps2 = ps + SBRP_THREADS;                  // [256] partial sums
    T* red = ps2 + SBRP_THREADS;                 // [32]
    double* redd = reinterpret_cast<double*>(red + 32);   // [32]
    int* sflag = reinterpret_cast<int*>(redd + 32);       // [1] abort broadcast

    // Work split: the 256 threads form NQ groups of RG (= 32 or 64 >= b) threads; thread (i, q) owns row
    // (or column) i and the q-th slice of the other index, partial sums meet in shared memory.
    const int RG = (b <= 32) ? 32 : 64, NQ = SBRP_THREADS / RG;
    const int i = tid % RG, q = tid / RG;

    for (int s = blockIdx.x; s <= n - 2; s += gridDim.x) {
        const int nt = sbr::sweep_ntasks(n, b, s);
        const int ntp = s > 0 ? sbr::sweep_ntasks(n, b, s - 1) : 0;
        T taup = zero<T>();
        for (int k = 0; k < nt; ++k) {
            // ---- wait for (s-1, k+1) ----
            if (tid == 0) {
                int bad = 0;
                if (s > 0) {
                    const int need = (k + 2 < ntp) ? k + 2 : ntp;
                    int spins = 0;
                    while (ld_acquire_gpu(prog + (s - 1)) < need) {
                        MAK_SPIN_PAUSE();
                        if ((++spins & 1023) == 0 && (spins >= SBRP_SPIN_LIMIT || ld_acquire_gpu(prog + n) != 0)) {
                            st_release_gpu(prog + n, 1);
                            bad = 1;
                            break;
                        }
                    }
                }
                *sflag = bad;
            }
            __syncthreads();   // also: every thread is done with the shared memory of the previous task
            if (*sflag) return;

            const sbr::Task tk = sbr::task_geometry(n, b, s, k);
            const int L = tk.L, Lp = tk.Lp, r0 = tk.r0, c0 = tk.c0;

            // ---- load: thread (i, q) owns row i and columns q, q+NQ, ...; for a fixed column the rows are
            // contiguous in the band storage (coalesced); only the lower triangle of D is read, the upper
            // one is mirrored in shared memory ----
            if (k > 0) {
                if (i < L) {
#pragma unroll 4
                    for (int j = q; j < Lp; j += NQ) G[i + j * ldg] = ld_cg(AB + (size_t)(c0 + j) * ldab + (r0 + i - c0 - j));
                }
            } else {
                if (tid < L) G[tid] = ld_cg(AB + (size_t)s * ldab + (r0 + tid - s));
            }
            if (i < L) {
#pragma unroll 4
                for (int j = q; j <= i; j += NQ) {
                    const T x = ld_cg(AB + (size_t)(r0 + j) * ldab + (i - j));
                    D[i + j * ldg] = x;
                    if (j != i) D[j + i * ldg] = conj_(x);
                }
            }
            __syncthreads();

            // ---- G <- G H_prev (k >= 1), H_prev handed over in shared memory ----
            if (k > 0) {
                if (!is_zero(taup)) {   // uniform
                    const int jw = (Lp + NQ - 1) / NQ, j0 = q * jw, j1 = (Lp < j0 + jw) ? Lp : j0 + jw;
                    T g = zero<T>();
                    if (i < L)
                        for (int j = j0; j < j1; ++j) fma_(g, G[i + j * ldg], vp[j]);
                    ps[q * RG + i] = g;
                    __syncthreads();
                    if (i < L) {
                        T gs = zero<T>();
                        for (int qq = 0; qq < NQ; ++qq) gs = add_(gs, ps[qq * RG + i]);
                        gs = mul_(taup, gs);
                        for (int j = j0; j < j1; ++j) G[i + j * ldg] = sub_(G[i + j * ldg], mul_(gs, conj_(vp[j])));
                    }
                }
                __syncthreads();
            }

            // ---- reflector from column 0 of G ----
            double part = 0.0;
            if (tid >= 1 && tid < L) part = abs2_(G[tid]);
            const double sigma = block_sum<double>(part, redd);
            double beta; T tau, scale;
            larfgp_scalars<T>(G[0], sigma, beta, tau, scale);
            __syncthreads();
            if (tid < L) {
                v[tid] = (tid == 0) ? one<T>() : mul_(G[tid], scale);
                G[tid] = (tid == 0) ? mk<T>(beta) : zero<T>();
            } else if (tid < b) {
                v[tid] = zero<T>();   // a short last block hands a zero-padded reflector on (never read: it is the last)
            }
            __syncthreads();

            if (!is_zero(tau)) {   // uniform
                // ---- G <- H^H G on columns 1 .. Lp-1 (k >= 1): thread (j, q) owns column j, row slice q ----
                const int iw = (L + NQ - 1) / NQ, i0 = q * iw, i1 = (L < i0 + iw) ? L : i0 + iw;
                if (k > 0) {
                    T dsum = zero<T>();
                    if (i >= 1 && i < Lp)
                        for (int r = i0; r < i1; ++r) fmac_(dsum, v[r], G[r + i * ldg]);
                    ps[q * RG + i] = dsum;
                }
                // ---- p = D v (row i, column slice q) ----
                {
                    T p = zero<T>();
                    if (i < L)
                        for (int j = i0; j < i1; ++j) fma_(p, D[i + j * ldg], v[j]);
                    ps2[q * RG + i] = p;
                }
                __syncthreads();
                if (k > 0 && i >= 1 && i < Lp) {
                    T dsum = zero<T>();
                    for (int qq = 0; qq < NQ; ++qq) dsum = add_(dsum, ps[qq * RG + i]);
                    dsum = mul_(conj_(tau), dsum);
                    for (int r = i0; r < i1; ++r) G[r + i * ldg] = sub_(G[r + i * ldg], mul_(v[r], dsum));
                }
                T p = zero<T>();
                if (tid < L)
                    for (int qq = 0; qq < NQ; ++qq) p = add_(p, ps2[qq * RG + tid]);
                T a = zero<T>();
                if (tid < L) fmac_(a, v[tid], p);
                const T alpha = block_sum<T>(a, red);     // v^H D v (real)
                const double half = 0.5 * abs2_(tau) * real_(alpha);
                if (tid < L) pw[tid] = sub_(mul_(tau, p), scale_(v[tid], half));
                __syncthreads();
                // ---- D <- D - v w^H - w v^H (row i, column slice q) ----
                if (i < L) {
                    const T vi = v[i], wi = pw[i];
                    for (int j = i0; j < i1; ++j) {
                        T x = D[i + j * ldg];
                        x = sub_(x, mul_(vi, conj_(pw[j])));
                        x = sub_(x, mul_(wi, conj_(v[j])));
                        if (j == i) x = mk<T>(real_(x));
                        D[i + j * ldg] = x;
                    }
                }
            }
            __syncthreads();

            // ---- store (same thread map as the load) ----
            if (k > 0) {
                if (i < L)
                    for (int j = q; j < Lp; j += NQ) AB[(size_t)(c0 + j) * ldab + (r0 + i - c0 - j)] = G[i + j * ldg];
            } else {
                if (tid < L) AB[(size_t)s * ldab + (r0 + tid - s)] = G[tid];
            }
            if (i < L)
                for (int j = q; j <= i; j += NQ) AB[(size_t)(r0 + j) * ldab + (i - j)] = D[i + j * ldg];
            for (int ii = tid; ii < L; ii += SBRP_THREADS) V2[(size_t)s * ldv + r0 + ii] = v[ii];
            if (tid == 0) tau2[(size_t)s * ldt + k] = tau;
            // ---- publish: every thread's stores are ordered before the counter ----
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release_gpu(prog + s, k + 1);
            // hand the reflector to the next task of the sweep
            T* tmp = v; v = vp; vp = tmp;
            taup = tau;
        }
    }
}

}  // namespace mak
