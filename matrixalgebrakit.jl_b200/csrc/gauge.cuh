// Gauge of one column (common/gauge.jl:12-14,38-45): V[:, j] *= conj(sign(first entry of maximal modulus)); the row j
// of an optional second matrix (svd: V^H) is scaled by the sign.  Called by the whole CTA; shared by the single-matrix
// kernel (eigh.cu) and the lock-step batched one (eigh_lockstep.cuh).
#pragma once
#include "common.cuh"
#include "devutil.cuh"
namespace mak {
template <typename T>
__device__ __forceinline__ void gauge_column_body(int m, int j, T* __restrict__ V, int ldv, T* __restrict__ other, int ldo,
                                                  int other_rows_n) {
    __shared__ double bv[32];
    __shared__ int bi[32];
    __shared__ T sfac;
    T* col = V + (size_t)j * ldv;
    double best = -1.0;
    int besti = 0x7fffffff;
    for (int r = threadIdx.x; r < m; r += blockDim.x) {
        double a = is_cplx<T>::value ? abs2_(col[r]) : fabs(real_(col[r]));
        if (a > best) { best = a; besti = r; }  // strict: first maximum wins within a thread
    }
    // warp reduce (value desc, index asc)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = besti; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (bv[w] > best || (bv[w] == best && bi[w] < besti)) { best = bv[w]; besti = bi[w]; }
        T piv = (m > 0 && besti < m) ? col[besti] : zero<T>();
        double a = sqrt(abs2_(piv));
        // Julia sign(): 0 at 0 -> the column (all zeros) is zeroed either way; keep it unchanged
        sfac = (a == 0.0) ? one<T>() : scale_(piv, 1.0 / a);
    }
    __syncthreads();
    const T sg = sfac, csg = conj_(sfac);
    for (int r = threadIdx.x; r < m; r += blockDim.x) col[r] = mul_(col[r], csg);
    if (other)
        for (int q = threadIdx.x; q < other_rows_n; q += blockDim.x) {
            T* p = other + (size_t)q * ldo + j;
            *p = mul_(*p, sg);
        }
    __syncthreads();
}
}  // namespace mak
