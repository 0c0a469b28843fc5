// Eigenvalues-only solver for a real symmetric tridiagonal matrix: Sturm-sequence counts + K-section
// (the dstebz/dlaebz class; Demmel, Dhillon & Ren 1995 is the published statement of the count's
// robustness).  Used by the values-only paths (eigh_vals! / svd_vals!, reference job 'N':
// yalapack.jl:1164-1279 with V of length 0, :2105-2127), where the divide-and-conquer solver's
// eigenvector GEMMs and the back-transformation are not needed.
//
// Every function is `__host__ __device__`: sturm_eigvals_kernel (eigh.cu) runs one thread per
// eigenvalue through kth_eigenvalue(), and tests/cpu_harness/sturm_host.cpp compiles the same header
// with g++ (tests/test_sturm_core_cpu.py).
#pragma once
#include <math.h>

#ifndef MAK_HD
#ifdef __CUDACC__
#define MAK_HD __host__ __device__ __forceinline__
#else
#define MAK_HD inline
#endif
#endif

namespace mak {
namespace sturm {

constexpr double ST_EPS = 1.1102230246251565e-16;      // 2^-53
constexpr double ST_PIVMIN = 2.2250738585072014e-308;  // smallest normal; the scaled matrix has max|e| <= 1
constexpr int ST_K = 7;                                // interior points per pass (3 bits per pass)
constexpr int ST_MAXPASS = 64;

// Scale and Gershgorin interval of the SCALED matrix T/scale (scale = max(|d|,|e|), 1 if T = 0):
// every eigenvalue of T/scale lies in [gl, gu] (widened by the dstebz slack).
struct Bounds {
    double scale, inv, gl, gu;
};

MAK_HD Bounds bounds(int n, const double* d, const double* e) {
    Bounds b;
    double mx = 0.0;
    for (int i = 0; i < n; ++i) {
        mx = fmax(mx, fabs(d[i]));
        if (i + 1 < n) mx = fmax(mx, fabs(e[i]));
    }
    b.scale = mx > 0.0 ? mx : 1.0;
    b.inv = 1.0 / b.scale;
    double gl = 0.0, gu = 0.0;
    for (int i = 0; i < n; ++i) {
        const double r = (i > 0 ? fabs(e[i - 1] * b.inv) : 0.0) + (i + 1 < n ? fabs(e[i] * b.inv) : 0.0);
        const double c = d[i] * b.inv;
        if (i == 0 || c - r < gl) gl = c - r;
        if (i == 0 || c + r > gu) gu = c + r;
    }
    const double tn = fmax(fabs(gl), fabs(gu));
    const double slack = 2.0 * tn * ST_EPS * n + 2.0 * ST_PIVMIN;
    b.gl = gl - slack;
    b.gu = gu + slack;
    return b;
}

// cnt[j] = number of eigenvalues of T/scale below x[j], j < K: K independent Sturm recurrences
//   q_0 = d_0 - x,  q_i = d_i - x - e_{i-1}^2 / q_{i-1}   (|q| < pivmin -> -pivmin)
// advanced together so their divisions overlap.
template <int K>
MAK_HD void count_below(int n, const double* d, const double* e, double inv, const double* x, int* cnt) {
    double q[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        q[j] = 1.0;
        cnt[j] = 0;
    }
    double e2 = 0.0;
    for (int i = 0; i < n; ++i) {
        const double di = d[i] * inv;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            double t = (di - x[j]) - e2 / q[j];
            if (fabs(t) < ST_PIVMIN) t = -ST_PIVMIN;
            cnt[j] += t < 0.0 ? 1 : 0;
            q[j] = t;
        }
        if (i + 1 < n) {
            const double ei = e[i] * inv;
            e2 = ei * ei;
        }
    }
}

// k-th smallest eigenvalue (0-based) of T.  All threads start from the same interval and split it at
// the same points, so two indices share their bracket until the counts separate them: the results
// come out in ascending order without a sort.
MAK_HD double kth_eigenvalue(int n, const double* d, const double* e, const Bounds& b, int k) {
    double lo = b.gl, hi = b.gu;   // invariant: count(lo) <= k < count(hi)
    const double atol = 0.25 * ST_EPS;
    for (int pass = 0; pass < ST_MAXPASS; ++pass) {
        const double w = hi - lo;
        if (w <= atol + 4.0 * ST_EPS * fmax(fabs(lo), fabs(hi))) break;
        double x[ST_K];
        int c[ST_K];
        const double h = w * (1.0 / (ST_K + 1));
#pragma unroll
        for (int j = 0; j < ST_K; ++j) x[j] = lo + (j + 1) * h;
        count_below<ST_K>(n, d, e, b.inv, x, c);
        double nlo = lo, nhi = hi;
        bool open = true;
#pragma unroll
        for (int j = 0; j < ST_K; ++j) {
            if (open) {
                if (c[j] <= k) nlo = x[j];
                else { nhi = x[j]; open = false; }
            }
        }
        if (!(nhi - nlo < w)) break;   // the interval no longer shrinks (points collapsed in floating point)
        lo = nlo;
        hi = nhi;
    }
    return 0.5 * (lo + hi) * b.scale;
}

}  // namespace sturm
}  // namespace mak
