// EXPERIMENTAL (round-2 bring-up, opt-in MAKB200_BHETRD=1): Hermitian tridiagonalisation of MANY mid-size
// blocks in ONE launch — one CTA per block, the whole reduction of a block inside its CTA.
//
// Why: for the 65..512 blocks of the batched config the per-block path runs the single-matrix hetrd,
// i.e. two launches per COLUMN per block (229..2234 blocks/s, launch
// bound).  A 512 x 512 ComplexF64 block is 4 MB: it lives in L2 while its CTA works on it, so a CTA
// can stream the trailing matrix from L2 once per pass and 148 blocks reduce concurrently.
//
// Algorithm (zhetd2 'L' with the package's non-negative-beta reflectors, scalar.h larfgp_scalars; same
// output convention as hetrd() in eigh.cu so stedc and ormqr_left_t follow unchanged):
//   for j = 0 .. n-2:   x = A[j+1:, j];  H_j = I - tau v v^H,  H_j^H x = beta e_1,  v = [1; scale x[1:]]
//       A[j+2:, j] <- v[1:],  e[j] = beta,  d[j] = real(A[j,j]),  tau[j] = tau
//       y = tau A22 v            (lower triangle only: entry (r,c) serves y[r] and y[c])
//       w = y - (tau/2)(y^H v) v
//       A22 <- A22 - v w^H - w v^H   (lower triangle) -- applied LAZILY, fused into the symv pass of step j+1:
//                                    one read + one write of the trailing matrix per column (16 n^3/3 B per c128 block)
// Columns of A22 are dealt to the warps round-robin, lanes run down the rows (coalesced); the row
// contributions of the symv go to a per-warp shared-memory vector (no atomics, deterministic).
//
// Device code only, written against the CUDA subset of tests/cpu_harness/cuda_emu.h
// (tests/test_emu_kernels_cpu.py validates it with g++ against LAPACK's eigenvalues and by
// reconstructing A from the reflectors).
#pragma once
#include "devutil.cuh"

namespace mak {

constexpr int BHETRD_THREADS = 256;
constexpr int BHETRD_NW = BHETRD_THREADS / 32;

template <typename T>
struct BhetrdDesc {
    int n;
    T* A; int lda;      // in: Hermitian, lower triangle read (upper when mirror_upper); out: reflectors below the sub-diagonal
    double* d;          // n
    double* e;          // n-1
    T* tau;             // n-1
};

// dynamic shared memory in elements of T for blocks up to nmax: v, w, NW partial-y vectors, 64 scratch, vn
inline size_t bhetrd_smem_elems(int nmax) { return (size_t)(3 + BHETRD_NW) * (size_t)nmax + 64; }

template <typename T>
__device__ __forceinline__ void bhetrd_body(const BhetrdDesc<T>& D, int nmax, int mirror_upper, unsigned char* smem_raw) {
    T* v = reinterpret_cast<T*>(smem_raw);
    T* w = v + nmax;
    T* yp = w + nmax;                          // [NW][nmax]
    T* scratch = yp + (size_t)BHETRD_NW * nmax;  // 64
    const int n = D.n, lda = D.lda;
    T* A = D.A;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n <= 0) return;
    if (mirror_upper) {
        // the reference calls LAPACK with uplo = 'U' (yalapack.jl:994,1286): lower <- conj(upper), one time
        for (int c = warp; c < n; c += BHETRD_NW)
            for (int r = c + 1 + lane; r < n; r += 32) A[(size_t)c * lda + r] = conj_(A[(size_t)r * lda + c]);
        __syncthreads();
    }

    // The rank-2 update of step j-1 is applied lazily: `pend` says that the stored trailing matrix still
    // lacks  - vp wp^H - wp vp^H  (vp, wp = v, w of the previous step, first entry = row/column j).  Step j
    // first makes column j true, generates its reflector, and then makes ONE pass over the stored A22 that
    // writes the true values back and feeds them to the symv of the new reflector: one read + one write of
    // the trailing matrix per column instead of a symv pass plus a separate update pass.
    bool pend = false;
    for (int j = 0; j + 1 < n; ++j) {
        const int mt = n - j - 1;              // order of A22 = A[j+1:, j+1:]
        T* x = A + (size_t)j * lda + (j + 1);  // column j below the diagonal, length mt
        // ---- column j (rows j..n-1) of the true matrix: v, w still hold the pending pair, index 0 = row j ----
        if (pend) {
            const T w0 = conj_(w[0]), v0 = conj_(v[0]);
            for (int r = tid; r <= mt; r += BHETRD_THREADS) {
                T* p = A + (size_t)j * lda + j + r;
                T a = sub_(sub_(*p, mul_(v[r], w0)), mul_(w[r], v0));
                if (r == 0) a = mk<T>(real_(a));
                *p = a;
            }
            __syncthreads();
        }
        // ---- reflector ----
        double part = 0.0;
        for (int r = 1 + tid; r < mt; r += BHETRD_THREADS) part += abs2_(x[r]);
        const double sigma = block_sum<double>(part, reinterpret_cast<double*>(scratch));
        double beta; T tau, scale;
        larfgp_scalars<T>(x[0], sigma, beta, tau, scale);
        const bool have = !is_zero(tau);       // uniform across the CTA
        // v, w keep the pending pair during the pass; the NEW reflector goes to the spare vector vn and is copied
        // into v when w is formed
        T* vn = scratch + 64;                  // nmax entries behind the scratch (see bhetrd_smem_elems)
        __syncthreads();
        for (int r = tid; r < mt; r += BHETRD_THREADS) {
            const T vr = (r == 0) ? one<T>() : mul_(x[r], scale);
            vn[r] = vr;
            if (r > 0) x[r] = vr;              // reflector storage (geqrf layout of A[1:, 0:n-1])
        }
        if (tid == 0) {
            D.e[j] = beta;
            D.d[j] = real_(A[(size_t)j * lda + j]);
            D.tau[j] = tau;
        }
        for (int i = tid; i < BHETRD_NW * mt; i += BHETRD_THREADS) yp[(size_t)(i / mt) * nmax + (i % mt)] = zero<T>();
        __syncthreads();
        if (!have && !pend) continue;          // nothing to apply, nothing to compute

        T* A22 = A + (size_t)(j + 1) * lda + (j + 1);
        // ---- one pass over the lower triangle of A22: apply the pending update, y = A22 vn ----
        {
            T* myp = yp + (size_t)warp * nmax;
            for (int c = warp; c < mt; c += BHETRD_NW) {
                T* col = A22 + (size_t)c * lda;
                const T vc = vn[c];
                T wc = zero<T>(), vpc = zero<T>();
                if (pend) { wc = conj_(w[c + 1]); vpc = conj_(v[c + 1]); }
                T acc = zero<T>();
                for (int r = c + lane; r < mt; r += 32) {
                    T a = col[r];
                    if (pend) {
                        a = sub_(sub_(a, mul_(v[r + 1], wc)), mul_(w[r + 1], vpc));
                        if (r == c) a = mk<T>(real_(a));
                        col[r] = a;
                    }
                    if (have) {
                        if (r == c) {
                            acc = add_(acc, scale_(vc, real_(a)));          // diagonal: real, serves y[c] once
                        } else {
                            fmac_(acc, a, vn[r]);                           // conj(A[r,c]) v[r]  -> y[c]
                            myp[r] = add_(myp[r], mul_(a, vc));             // A[r,c] v[c]        -> y[r]
                        }
                    }
                }
                if (have) {
                    acc = warp_sum(acc);
                    if (lane == 0) myp[c] = add_(myp[c], acc);
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (!have) { pend = false; continue; }
        // ---- x = tau y,  g = (tau/2) x^H v,  w = x - g v;  (v, w) <- (vn, w) becomes the pending pair ----
        T xv_part = zero<T>();
        for (int r = tid; r < mt; r += BHETRD_THREADS) {
            T sacc = zero<T>();
            for (int q = 0; q < BHETRD_NW; ++q) sacc = add_(sacc, yp[(size_t)q * nmax + r]);
            sacc = mul_(tau, sacc);
            w[r] = sacc;
            v[r] = vn[r];
            fmac_(xv_part, sacc, vn[r]);       // conj(x_r) v_r
        }
        const T xv = block_sum<T>(xv_part, scratch);
        const T g = scale_(mul_(tau, xv), 0.5);
        __syncthreads();
        for (int r = tid; r < mt; r += BHETRD_THREADS) w[r] = sub_(w[r], mul_(g, v[r]));
        __syncthreads();
        pend = true;
    }
    if (tid == 0) {
        T a = A[(size_t)(n - 1) * lda + (n - 1)];
        if (pend) {   // v, w have one entry left: row n-1
            a = sub_(sub_(a, mul_(v[0], conj_(w[0]))), mul_(w[0], conj_(v[0])));
        }
        D.d[n - 1] = real_(a);
    }
}

// one CTA per block of a batch (descriptors in device memory)
template <typename T>
__global__ void __launch_bounds__(BHETRD_THREADS) bhetrd_kernel(const BhetrdDesc<T>* __restrict__ descs, int nmax,
                                                                int mirror_upper) {
    MAK_DYN_SMEM(smem_raw);
    const BhetrdDesc<T> D = descs[blockIdx.x];
    bhetrd_body<T>(D, nmax, mirror_upper, smem_raw);
}
// a single block, descriptor by value (per-block path: one launch instead of two per column)
template <typename T>
__global__ void __launch_bounds__(BHETRD_THREADS) bhetrd_one_kernel(BhetrdDesc<T> D, int mirror_upper) {
    MAK_DYN_SMEM(smem_raw);
    bhetrd_body<T>(D, D.n, mirror_upper, smem_raw);
}

}  // namespace mak
