// One-CTA Cholesky of a diagonal block with its explicit inverse: the serial link of the blocked Cholesky
// (polar.cu: potrf_blocked) and, one CTA per matrix, of the lock-step batched Cholesky (polar_lockstep.cuh).
#pragma once
#include "common.cuh"
#include "devutil.cuh"
namespace mak {
template <typename T> struct CholNB { static constexpr int value = 128; };
template <> struct CholNB<cplx> { static constexpr int value = 64; };

__device__ __forceinline__ double shfl_xor_any(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ cplx shfl_xor_any(cplx v, int o) {
    return cplx{__shfl_xor_sync(0xffffffffu, v.re, o), __shfl_xor_sync(0xffffffffu, v.im, o)};
}
// ---------------------------------------------------------------------------------------
// Cholesky building blocks
// ---------------------------------------------------------------------------------------
// one CTA: L = chol(Zblk) (nb x nb, lower) and Linv = L^-1; both written with zeros above the
// diagonal.  info[0] set to 1 if a pivot is not positive.
// The block is held in REGISTERS, 2-D cyclic over a 16 x 16 thread grid (thread (ti,tj) owns rows
// ti+16a, columns tj+16b); per column k: pivot -> scaled column to shared memory -> rank-1 update
// of the register tile.  Two block barriers per column, no shared-memory traffic for the matrix.
template <typename T, int NB>
__device__ __forceinline__ void potf2_inv_body(int nb, const T* __restrict__ Zb, int ldz, T* __restrict__ Lb, int ldl,
                                               T* __restrict__ Linv, int ldi, int* info) {
    constexpr int E = NB / 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* S = reinterpret_cast<T*>(smem_raw);  // [NB][NB+1] column-major: S[c*(NB+1)+r] (L for the inverse)
    __shared__ T s_col[NB];
    __shared__ double s_piv;
    const int lds = NB + 1, tid = threadIdx.x, ti = tid & 15, tj = tid >> 4;
    T a[E][E];
#pragma unroll
    for (int ia = 0; ia < E; ++ia)
#pragma unroll
        for (int ib = 0; ib < E; ++ib) {
            const int r = ti + 16 * ia, c = tj + 16 * ib;
            a[ia][ib] = (r < nb && c < nb && r >= c) ? Zb[(size_t)c * ldz + r] : zero<T>();
        }
#pragma unroll
    for (int kb = 0; kb < E; ++kb) {
        for (int kk = 0; kk < 16; ++kk) {
            const int k = kb * 16 + kk;
            if (k >= nb) break;
            if (ti == kk && tj == kk) {
                double akk = real_(a[kb][kb]);
                if (!(akk > 0.0)) { atomicExch(info, 1); akk = 1.0; }
                s_piv = sqrt(akk);
            }
            __syncthreads();
            const double piv = s_piv, inv = 1.0 / piv;
            if (tj == kk) {
#pragma unroll
                for (int ia = 0; ia < E; ++ia) {
                    const int r = ti + 16 * ia;
                    if (r > k) {
                        T v = scale_(a[ia][kb], inv);
                        a[ia][kb] = v;
                        s_col[r] = v;
                    } else if (r == k) {
                        a[ia][kb] = mk<T>(piv);
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int ia = 0; ia < E; ++ia) {
                const int r = ti + 16 * ia;
                if (r > k && r < nb) {
                    const T lr = s_col[r];
#pragma unroll
                    for (int ib = 0; ib < E; ++ib) {
                        const int c = tj + 16 * ib;
                        if (c > k && c <= r) a[ia][ib] = sub_(a[ia][ib], mul_(lr, conj_(s_col[c])));
                    }
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int ia = 0; ia < E; ++ia)
#pragma unroll
        for (int ib = 0; ib < E; ++ib) {
            const int r = ti + 16 * ia, c = tj + 16 * ib;
            if (r < nb && c < nb) {
                T v = (r >= c) ? a[ia][ib] : zero<T>();
                S[c * lds + r] = v;
                Lb[(size_t)c * ldl + r] = v;
            }
        }
    __syncthreads();
    // inverse, in place in shared memory (S is a private copy of L), one column per step from the last to the first:
    //   X[j,j] = 1 / L[j,j],   X[j+1:, j] = -(X[j+1:, j+1:] L[j+1:, j]) X[j,j]
    // Row-parallel: thread pair (r, half) takes half of row r's dot product (S is column-major with an odd leading
    // dimension: consecutive rows hit consecutive banks, L[p,j] is a broadcast).  The r1 version gave every thread a whole
    // forward substitution through GLOBAL memory (8 k dependent steps for column 0): ~170 us of the ~200 us this kernel
    // took, and n/128 of them are the serial chain of the blocked Cholesky.
    {
        const int r = tid >> 1, hf = tid & 1;
        for (int j = nb - 1; j >= 0; --j) {
            T acc = zero<T>();
            const double dinv = 1.0 / real_(S[j * lds + j]);   // (read before the barrier: row j's thread overwrites it below)
            if (r > j && r < nb) {
                // p in (j, r]: X[r,p] * L[p,j]; the two halves interleave p
                for (int pp = j + 1 + hf; pp <= r; pp += 2) fma_(acc, S[pp * lds + r], S[j * lds + pp]);
            }
            // combine the halves (lanes 2q and 2q+1 of the same warp)
            acc = add_(acc, shfl_xor_any(acc, 1));
            __syncthreads();                       // everybody has read column j of L
            if (hf == 0) {
                if (r == j) S[j * lds + j] = mk<T>(dinv);
                else if (r > j && r < nb) S[j * lds + r] = scale_(neg_(acc), dinv);
            }
            __syncthreads();
        }
        for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
            const int c = idx / nb, rr = idx - c * nb;
            Linv[(size_t)c * ldi + rr] = (rr >= c) ? S[c * lds + rr] : zero<T>();
        }
    }
}


template <typename T, int NB>
__global__ void __launch_bounds__(256)
potf2_inv_kernel(int nb, const T* __restrict__ Zb, int ldz, T* __restrict__ Lb, int ldl, T* __restrict__ Linv,
                 int ldi, int* info) {
    potf2_inv_body<T, NB>(nb, Zb, ldz, Lb, ldl, Linv, ldi, info);
}
}  // namespace mak
