// Scalar layer shared by the CUDA kernels and the CPU harnesses (tests/cpu_harness): double and an
// interleaved complex with the arithmetic helpers every kernel uses, and the reflector scalars.
// Compiles with nvcc (as part of common.cuh) and with plain g++ (the qualifiers become no-ops).
#pragma once
#include <math.h>
#include <stdint.h>
#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#ifndef __align__
#define __align__(x) alignas(x)
#endif
#endif

namespace mak {

// ---------------------------------------------------------------------------------------
// scalar types: double and an interleaved complex (layout == Julia ComplexF64)
// ---------------------------------------------------------------------------------------
struct __align__(16) cplx {
    double re, im;
};

template <typename T> struct is_cplx { static constexpr bool value = false; };
template <> struct is_cplx<cplx> { static constexpr bool value = true; };

__host__ __device__ __forceinline__ double zero_of(double) { return 0.0; }
__host__ __device__ __forceinline__ cplx zero_of(cplx) { return cplx{0.0, 0.0}; }
template <typename T> __host__ __device__ __forceinline__ T zero() { return zero_of(T{}); }
__host__ __device__ __forceinline__ double one_of(double) { return 1.0; }
__host__ __device__ __forceinline__ cplx one_of(cplx) { return cplx{1.0, 0.0}; }
template <typename T> __host__ __device__ __forceinline__ T one() { return one_of(T{}); }

__host__ __device__ __forceinline__ double conj_(double a) { return a; }
__host__ __device__ __forceinline__ cplx conj_(cplx a) { return cplx{a.re, -a.im}; }
__host__ __device__ __forceinline__ double real_(double a) { return a; }
__host__ __device__ __forceinline__ double real_(cplx a) { return a.re; }
__host__ __device__ __forceinline__ double imag_(double) { return 0.0; }
__host__ __device__ __forceinline__ double imag_(cplx a) { return a.im; }
__host__ __device__ __forceinline__ double abs2_(double a) { return a * a; }
__host__ __device__ __forceinline__ double abs2_(cplx a) { return a.re * a.re + a.im * a.im; }
// |a| without forming the square (no underflow / overflow for |a| outside [1e-154, 1e154])
__host__ __device__ __forceinline__ double abs_(double a) { return fabs(a); }
__host__ __device__ __forceinline__ double abs_(cplx a) { return hypot(a.re, a.im); }
__host__ __device__ __forceinline__ double add_(double a, double b) { return a + b; }
__host__ __device__ __forceinline__ cplx add_(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
__host__ __device__ __forceinline__ double sub_(double a, double b) { return a - b; }
__host__ __device__ __forceinline__ cplx sub_(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
__host__ __device__ __forceinline__ double mul_(double a, double b) { return a * b; }
__host__ __device__ __forceinline__ cplx mul_(cplx a, cplx b) {
    return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__host__ __device__ __forceinline__ double neg_(double a) { return -a; }
__host__ __device__ __forceinline__ cplx neg_(cplx a) { return cplx{-a.re, -a.im}; }
__host__ __device__ __forceinline__ double scale_(double a, double s) { return a * s; }
__host__ __device__ __forceinline__ cplx scale_(cplx a, double s) { return cplx{a.re * s, a.im * s}; }
// a += b*c
__host__ __device__ __forceinline__ void fma_(double& a, double b, double c) { a = fma(b, c, a); }
__host__ __device__ __forceinline__ void fma_(cplx& a, cplx b, cplx c) {
    a.re = fma(b.re, c.re, a.re);
    a.re = fma(-b.im, c.im, a.re);
    a.im = fma(b.re, c.im, a.im);
    a.im = fma(b.im, c.re, a.im);
}
// a += conj(b)*c
__host__ __device__ __forceinline__ void fmac_(double& a, double b, double c) { a = fma(b, c, a); }
__host__ __device__ __forceinline__ void fmac_(cplx& a, cplx b, cplx c) {
    a.re = fma(b.re, c.re, a.re);
    a.re = fma(b.im, c.im, a.re);
    a.im = fma(b.re, c.im, a.im);
    a.im = fma(-b.im, c.re, a.im);
}
__host__ __device__ __forceinline__ double div_(double a, double b) { return a / b; }
__host__ __device__ __forceinline__ cplx div_(cplx a, cplx b) {
    // Smith's algorithm
    if (fabs(b.re) >= fabs(b.im)) {
        double r = b.im / b.re, d = b.re + b.im * r;
        return cplx{(a.re + a.im * r) / d, (a.im - a.re * r) / d};
    } else {
        double r = b.re / b.im, d = b.re * r + b.im;
        return cplx{(a.re * r + a.im) / d, (a.im * r - a.re) / d};
    }
}
__host__ __device__ __forceinline__ double from_real(double r, double) { return r; }
__host__ __device__ __forceinline__ cplx from_real(double r, cplx) { return cplx{r, 0.0}; }
template <typename T> __host__ __device__ __forceinline__ T mk(double r) { return from_real(r, T{}); }
__host__ __device__ __forceinline__ bool is_zero(double a) { return a == 0.0; }
__host__ __device__ __forceinline__ bool is_zero(cplx a) { return a.re == 0.0 && a.im == 0.0; }
__host__ __device__ __forceinline__ bool is_one(double a) { return a == 1.0; }
__host__ __device__ __forceinline__ bool is_one(cplx a) { return a.re == 1.0 && a.im == 0.0; }

// stable non-negative-beta reflector scalars: given alpha = x[0], sigma = ||x[1:]||^2
// returns beta >= 0, tau, scale with v = x[1:]*scale, H = I - tau*[1;v][1;v]^H, H^H x = beta e1.
template <typename T>
__host__ __device__ __forceinline__ void larfgp_scalars(T alpha, double sigma, double& beta, T& tau, T& scale) {
    if (sigma == 0.0 && imag_(alpha) == 0.0 && real_(alpha) >= 0.0) {
        beta = real_(alpha);
        tau = zero<T>();
        scale = zero<T>();
        return;
    }
    beta = sqrt(abs2_(alpha) + sigma);
    T d;
    if (real_(alpha) < 0.0) {
        d = sub_(alpha, mk<T>(beta));
    } else {
        // alpha - beta = ((alpha - conj(alpha))*beta - sigma) / (conj(alpha) + beta)
        T num = sub_(scale_(sub_(alpha, conj_(alpha)), beta), mk<T>(sigma));
        T den = add_(conj_(alpha), mk<T>(beta));
        d = div_(num, den);
    }
    tau = scale_(neg_(d), 1.0 / beta);
    scale = div_(one<T>(), d);
}

}  // namespace mak
