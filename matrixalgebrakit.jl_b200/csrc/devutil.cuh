// Device-side helpers shared by every kernel: warp/block reductions, complex shuffles and the
// dynamic-shared-memory declaration.  No runtime-API dependency: compiles under nvcc and, through
// tests/cpu_harness/cuda_emu.h (MAK_EMU), under plain g++ for the CPU logic tests.
#pragma once
#include "scalar.h"
#ifndef MAK_EMU
#define MAK_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace mak {

// ---------------------------------------------------------------------------------------
// warp / block reductions (deterministic order)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ cplx warp_sum(cplx v) {
    v.re = warp_sum(v.re);
    v.im = warp_sum(v.im);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum; `scratch` must hold >= 32 T; result broadcast to all threads.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    T r = zero<T>();
    for (int i = 0; i < nw; ++i) r = add_(r, scratch[i]);
    return r;
}

__device__ __forceinline__ double shfl_(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ cplx shfl_(cplx v, int src) {
    return cplx{__shfl_sync(0xffffffffu, v.re, src), __shfl_sync(0xffffffffu, v.im, src)};
}

}  // namespace mak

// ---------------------------------------------------------------------------------------
// inter-CTA flags of the persistent kernels (gpu-scope release/acquire, L2-only loads of data that
// other CTAs of the SAME launch produce).  Under MAK_EMU memory is sequentially consistent and a
// polling loop must yield to the other fibers (MAK_SPIN_PAUSE comes from cuda_emu.h).
// ---------------------------------------------------------------------------------------
namespace mak {
#ifdef MAK_EMU
inline int ld_acquire_gpu(const int* p) { return *(const volatile int*)p; }
inline void st_release_gpu(int* p, int v) { *(volatile int*)p = v; }
inline double ld_cg(const double* p) { return *p; }
inline cplx ld_cg(const cplx* p) { return *p; }
#else
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_cg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ cplx ld_cg(const cplx* p) {
    const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
    return cplx{v.x, v.y};
}
#define MAK_SPIN_PAUSE() __nanosleep(32)
#endif
}  // namespace mak

// ---------------------------------------------------------------------------------------
// FP64 tensor-core tile product D(8x8) += A(8x4) B(4x8) (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4).
// Lane l holds a = A[l/4][l%4], b = B[l%4][l/4], d0/d1 = D[l/4][2(l%4) + {0,1}].
// ---------------------------------------------------------------------------------------
namespace mak {
#ifdef MAK_EMU
inline void dmma_f64(double& d0, double& d1, double a, double b) { emu::dmma(d0, d1, a, b); }
#else
__device__ __forceinline__ void dmma_f64(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
#endif
// accumulator of one 8x8 tile per lane and acc += a*b on the fragments (complex: four real products)
template <typename T> struct TileAcc;
template <> struct TileAcc<double> { double c0, c1; };
template <> struct TileAcc<cplx> { double r0, r1, i0, i1; };
__device__ __forceinline__ void tile_zero(TileAcc<double>& a) { a.c0 = a.c1 = 0.0; }
__device__ __forceinline__ void tile_zero(TileAcc<cplx>& a) { a.r0 = a.r1 = a.i0 = a.i1 = 0.0; }
__device__ __forceinline__ void tile_set(TileAcc<double>& a, double x0, double x1) { a.c0 = x0; a.c1 = x1; }
__device__ __forceinline__ void tile_set(TileAcc<cplx>& a, cplx x0, cplx x1) { a.r0 = x0.re; a.r1 = x1.re; a.i0 = x0.im; a.i1 = x1.im; }
__device__ __forceinline__ double tile_get0(const TileAcc<double>& a) { return a.c0; }
__device__ __forceinline__ double tile_get1(const TileAcc<double>& a) { return a.c1; }
__device__ __forceinline__ cplx tile_get0(const TileAcc<cplx>& a) { return cplx{a.r0, a.i0}; }
__device__ __forceinline__ cplx tile_get1(const TileAcc<cplx>& a) { return cplx{a.r1, a.i1}; }
__device__ __forceinline__ void tile_mma(TileAcc<double>& acc, double a, double b) { dmma_f64(acc.c0, acc.c1, a, b); }
__device__ __forceinline__ void tile_mma(TileAcc<cplx>& acc, cplx a, cplx b) {
    dmma_f64(acc.r0, acc.r1, a.re, b.re);
    dmma_f64(acc.r0, acc.r1, -a.im, b.im);
    dmma_f64(acc.i0, acc.i1, a.re, b.im);
    dmma_f64(acc.i0, acc.i1, a.im, b.re);
}
}  // namespace mak
