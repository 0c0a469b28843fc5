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
