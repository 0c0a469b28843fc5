// Fused application of the chase reflectors, Z <- Q2 Z (second back-transformation of the two-stage
// eigh_full!), one CTA per COLUMN SLAB of Z.
//
// The grouped-GEMM form (sbr.cu: sbr_apply_q2_t) walks the diamond wavefronts with three launches each
// and streams Z through HBM for every block.  Columns of Z are independent under a left multiplication,
// so here a CTA owns `cw` columns and walks ALL diamond blocks in a fixed serial order that satisfies
// every dependency of sbr_core.h (groups descending, chase position k ascending inside a group).  For
// one group the rows a block touches slide down by b per step, so the slab is held in shared memory as
// a 2b-row ring: each step loads b new rows and stores the b finished ones -- per group the slab is
// read and written once (n^3 sz / g bytes in total, ~10 ms at n = 8192, g = 64).  Per step, on the FP64
// tensor cores (DMMA m8n8k4 fragments straight from padded shared-memory tiles):
//     W  = V^H Zw      (g x cw)   only the non-zero rows of the parallelogram V: b+8 of b+g per 8 sweeps
//     W <- T W         (T upper triangular: half the products), staged in registers, in place
//     Zw -= V W        only the non-zero sweeps per 8 rows
// V (ld b+g, zero outside the reflectors) and T (ld g) come from q2_build_kernel's pools, addressed
// through blkmap[grp * kmax + k] -> Q2BlockDesc.
//
// Requirements: b % 8 == 0, g % 8 == 0, g <= b, g <= 64, cw in {32, 64}.  Device code only, written
// against the CUDA subset of tests/cpu_harness/cuda_emu.h (CPU logic test: test_emu_kernels_cpu.py).
#pragma once
#include "devutil.cuh"
#include "sbr_core.h"

namespace mak {

struct Q2BlockDesc {
    int s0, ns, base, rows, k, pad;
    size_t voff, toff;   // element offsets into the V / T pools
};

constexpr int Q2S_THREADS = 256;
constexpr int Q2S_WARPS = Q2S_THREADS / 32;
constexpr int Q2S_NT = 4;      // a warp item is an 8 x 32 block of the output: one A fragment feeds four DMMAs
constexpr int Q2S_MAXI = 2;    // items per warp of the W phases: (g/8)(cw/32)/8 <= 2 for g, cw <= 64

// leading dimension == 4 (mod 16): the 4 x 4 lane pattern of a DMMA fragment load (index lr*ld + lc or
// lc*ld + lr) touches 16 distinct 8-byte banks per half warp
__host__ __device__ __forceinline__ int q2s_pad_ld(int rows) { return rows + ((rows % 16 == 0) ? 4 : (rows % 16 == 8 ? 12 : 20 - rows % 16)); }

struct Q2SlabSmem {
    int ldzs, ldvs, ldts, ldws;
    size_t zs, vs, ts, ws, total;   // element offsets / total elements
};
__host__ __device__ __forceinline__ Q2SlabSmem q2_slab_smem(int b, int g, int cw) {
    Q2SlabSmem m;
    m.ldzs = q2s_pad_ld(2 * b);
    m.ldvs = q2s_pad_ld(b + g);
    m.ldts = q2s_pad_ld(g);
    m.ldws = q2s_pad_ld(g);
    m.zs = 0;
    m.vs = m.zs + (size_t)cw * m.ldzs;
    m.ts = m.vs + (size_t)g * m.ldvs;
    m.ws = m.ts + (size_t)g * m.ldts;
    m.total = m.ws + (size_t)cw * m.ldws;
    return m;
}

template <typename T>
__global__ void __launch_bounds__(Q2S_THREADS)
q2_slab_kernel(int n, int b, int g, int cw, int ngroups, int kmax, const int* __restrict__ blkmap,
               const Q2BlockDesc* __restrict__ descs, const T* __restrict__ Vpool, const T* __restrict__ Tpool,
               T* __restrict__ Z, int ldz, int ncols) {
    MAK_DYN_SMEM(smem_raw);
    const Q2SlabSmem sm = q2_slab_smem(b, g, cw);
    T* Zs = reinterpret_cast<T*>(smem_raw) + sm.zs;   // ring of 2b rows x cw columns, column-major
    T* Vs = reinterpret_cast<T*>(smem_raw) + sm.vs;   // (b+g) x g
    T* Ts = reinterpret_cast<T*>(smem_raw) + sm.ts;   // g x g upper
    T* Ws = reinterpret_cast<T*>(smem_raw) + sm.ws;   // g x cw, column-major
    const int ldzs = sm.ldzs, ldvs = sm.ldvs, ldts = sm.ldts, ldws = sm.ldws;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lr = lane >> 2, lc = lane & 3;
    const int col0 = blockIdx.x * cw;
    const int ncw = (ncols - col0 < cw) ? ncols - col0 : cw;
    if (ncw <= 0) return;
    const int ring = 2 * b, ldvb = b + g;
    const int g8 = g / 8, cgn = cw / (8 * Q2S_NT);
    const int nitemsW = g8 * cgn, nitemsZ = (ldvb / 8) * cgn;

    // rows [row0, row0 + b) of the slab <-> ring rows [half*b, half*b + b); rows >= n and columns >= ncw are zero
    auto load_half = [&](int half, int row0) {
        for (int c = warp; c < cw; c += Q2S_WARPS) {
            T* dst = Zs + (size_t)c * ldzs + half * b;
            const T* src = Z + (size_t)(col0 + c) * ldz + row0;
            for (int r = lane; r < b; r += 32) dst[r] = (c < ncw && row0 + r < n) ? src[r] : zero<T>();
        }
    };
    auto store_half = [&](int half, int row0) {
        for (int c = warp; c < ncw; c += Q2S_WARPS) {
            const T* src = Zs + (size_t)c * ldzs + half * b;
            T* dst = Z + (size_t)(col0 + c) * ldz + row0;
            for (int r = lane; r < b; r += 32)
                if (row0 + r < n) dst[r] = src[r];
        }
    };

    for (int grp = ngroups - 1; grp >= 0; --grp) {
        int klast = -1, base_last = 0;
        for (int k = 0; k < kmax; ++k) {
            const int bi = blkmap[grp * kmax + k];
            if (bi < 0) break;                       // blocks of a group are a prefix in k
            const Q2BlockDesc d = descs[bi];
            const int top = k & 1, zoff = top * b;   // ring half that holds rows [base, base + b)
            if (k == 0) load_half(0, d.base);
            load_half(top ^ 1, d.base + b);
            {
                const T* Vg = Vpool + d.voff;
                const T* Tg = Tpool + d.toff;
                for (int j = warp; j < g; j += Q2S_WARPS) {
                    for (int r = lane; r < ldvb; r += 32) Vs[(size_t)j * ldvs + r] = Vg[(size_t)j * ldvb + r];
                    for (int r = lane; r < g; r += 32) Ts[(size_t)j * ldts + r] = Tg[(size_t)j * g + r];
                }
            }
            __syncthreads();

            // ---- W = V^H Zw ----
            for (int item = warp; item < nitemsW; item += Q2S_WARPS) {
                const int jt = item % g8, cg = item / g8;
                TileAcc<T> acc[Q2S_NT];
#pragma unroll
                for (int t = 0; t < Q2S_NT; ++t) tile_zero(acc[t]);
                const int rlo = 8 * jt, rhi = (8 * jt + 8 + b < ldvb) ? 8 * jt + 8 + b : ldvb;
                const T* va = Vs + (size_t)(8 * jt + lr) * ldvs + lc;
#pragma unroll 2
                for (int r = rlo; r < rhi; r += 4) {
                    int zr = zoff + r;
                    if (zr >= ring) zr -= ring;
                    const T a = conj_(va[r]);
#pragma unroll
                    for (int t = 0; t < Q2S_NT; ++t)
                        tile_mma(acc[t], a, Zs[(size_t)(32 * cg + 8 * t + lr) * ldzs + zr + lc]);
                }
#pragma unroll
                for (int t = 0; t < Q2S_NT; ++t) {
                    T* w = Ws + (size_t)(32 * cg + 8 * t + 2 * lc) * ldws + 8 * jt + lr;
                    w[0] = tile_get0(acc[t]);
                    w[ldws] = tile_get1(acc[t]);
                }
            }
            __syncthreads();

            // ---- W <- T W (upper triangular T: products with j >= i only), register staged ----
            {
                TileAcc<T> acc2[Q2S_MAXI][Q2S_NT];
#pragma unroll
                for (int ii = 0; ii < Q2S_MAXI; ++ii) {
                    const int item = warp + ii * Q2S_WARPS;
#pragma unroll
                    for (int t = 0; t < Q2S_NT; ++t) tile_zero(acc2[ii][t]);
                    if (item < nitemsW) {
                        const int it = item % g8, cg = item / g8;
#pragma unroll 2
                        for (int j = 8 * it; j < g; j += 4) {
                            const T a = Ts[(size_t)(j + lc) * ldts + 8 * it + lr];
#pragma unroll
                            for (int t = 0; t < Q2S_NT; ++t)
                                tile_mma(acc2[ii][t], a, Ws[(size_t)(32 * cg + 8 * t + lr) * ldws + j + lc]);
                        }
                    }
                }
                __syncthreads();
#pragma unroll
                for (int ii = 0; ii < Q2S_MAXI; ++ii) {
                    const int item = warp + ii * Q2S_WARPS;
                    if (item < nitemsW) {
                        const int it = item % g8, cg = item / g8;
#pragma unroll
                        for (int t = 0; t < Q2S_NT; ++t) {
                            T* w = Ws + (size_t)(32 * cg + 8 * t + 2 * lc) * ldws + 8 * it + lr;
                            w[0] = tile_get0(acc2[ii][t]);
                            w[ldws] = tile_get1(acc2[ii][t]);
                        }
                    }
                }
            }
            __syncthreads();

            // ---- Zw -= V W (rows 8rt .. 8rt+7 see sweeps j in (r - b, r] only) ----
            for (int item = warp; item < nitemsZ; item += Q2S_WARPS) {
                const int rt = item % (ldvb / 8), cg = item / (ldvb / 8);
                const int r0l = 8 * rt;
                int zr = zoff + r0l;
                if (zr >= ring) zr -= ring;
                TileAcc<T> acc[Q2S_NT];
#pragma unroll
                for (int t = 0; t < Q2S_NT; ++t) {
                    const T* z = Zs + (size_t)(32 * cg + 8 * t + 2 * lc) * ldzs + zr + lr;
                    tile_set(acc[t], z[0], z[ldzs]);
                }
                const int jlo = (r0l >= b) ? r0l - b : 0, jhi = (r0l + 8 < g) ? r0l + 8 : g;
#pragma unroll 2
                for (int j = jlo; j < jhi; j += 4) {
                    const T a = neg_(Vs[(size_t)(j + lc) * ldvs + r0l + lr]);
#pragma unroll
                    for (int t = 0; t < Q2S_NT; ++t)
                        tile_mma(acc[t], a, Ws[(size_t)(32 * cg + 8 * t + lr) * ldws + j + lc]);
                }
#pragma unroll
                for (int t = 0; t < Q2S_NT; ++t) {
                    T* z = Zs + (size_t)(32 * cg + 8 * t + 2 * lc) * ldzs + zr + lr;
                    z[0] = tile_get0(acc[t]);
                    z[ldzs] = tile_get1(acc[t]);
                }
            }
            __syncthreads();

            // rows [base, base + b) are final for this group
            store_half(top, d.base);
            klast = k;
            base_last = d.base;
            __syncthreads();   // the next step refills this half
        }
        if (klast >= 0) {
            store_half((klast & 1) ^ 1, base_last + b);
            __syncthreads();
        }
    }
}

}  // namespace mak
