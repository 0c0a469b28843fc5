// Truncation search on a sorted spectrum: findtruncated_svd + truncation_error!
// (src/implementations/truncation.jl:54-58 rank, :69-79 tolerance, :86-102 error, :168-174 error norm;
// strategy composition src/interface/truncation.jl:37-66).  For singular values (sorted descending)
// every supported strategy keeps a PREFIX, so an intersection is the minimum of the ranks and a union
// the maximum; the whole decision is one sequential pass per block.
//
// `__host__ __device__`: trunc_select_kernel (truncation.cu) runs one thread per block of a batch;
// tests/cpu_harness/trunc_host.cpp compiles the same function with g++ (tests/test_trunc_core_cpu.py).
#pragma once
#include <math.h>
#include "../../include/makb200.h"

#ifndef MAK_HD
#ifdef __CUDACC__
#define MAK_HD __host__ __device__ __forceinline__
#else
#define MAK_HD inline
#endif
#endif

namespace mak {
namespace trunc {

MAK_HD double powp(double v, double p) {
    v = fabs(v);
    if (p == 2.0) return v * v;
    if (p == 1.0) return v;
    return pow(v, p);
}

// number of leading values kept by trunctol(atol, rtol, p): |v| >= max(atol, rtol * ||values||_p)
MAK_HD int rank_by_value(int k, const double* S, double atol, double rtol, double p) {
    double thr = atol;
    if (rtol > 0.0) {
        double np_ = 0.0;
        for (int j = 0; j < k; ++j) np_ += powp(S[j], p);
        const double nrm = p == 2.0 ? sqrt(np_) : (p == 1.0 ? np_ : pow(np_, 1.0 / p));
        thr = fmax(atol, rtol * nrm);
    }
    int r = 0;
    for (int j = 0; j < k; ++j) r += fabs(S[j]) >= thr ? 1 : 0;
    return r;
}

// truncerror(atol, rtol, p): drop the smallest values while the p-norm of what is dropped stays below
// max(atol, rtol * ||values||_p)
MAK_HD int rank_by_error(int k, const double* S, double atol, double rtol, double p) {
    double np_ = 0.0;
    for (int j = 0; j < k; ++j) np_ += powp(S[j], p);
    const double ep = fmax(powp(atol, p), powp(rtol, p) * np_);
    if (ep >= np_) return 0;
    double cs = 0.0;
    for (int j = k - 1; j >= 0; --j) {
        cs += powp(S[j], p);
        if (cs >= ep) return j + 1;
    }
    return 0;
}

// rank kept by the composed strategy and the 2-norm of the discarded values
MAK_HD void select(int k, const double* S, const makb200_trunc_spec& sp, int* rank, double* eps) {
    int r = k;
    bool any = false;
    if (sp.maxrank >= 0) { r = sp.maxrank < r ? sp.maxrank : r; any = true; }
    if (sp.by_value) { const int rv = rank_by_value(k, S, sp.vatol, sp.vrtol, sp.vp); r = rv < r ? rv : r; any = true; }
    if (sp.by_error) { const int re = rank_by_error(k, S, sp.eatol, sp.ertol, sp.ep); r = re < r ? re : r; any = true; }
    if (sp.minrank >= 0) {
        const int rm = sp.minrank < k ? sp.minrank : k;
        r = any ? (rm > r ? rm : r) : rm;
    }
    double t = 0.0;
    for (int j = r; j < k; ++j) t += S[j] * S[j];
    *rank = r;
    *eps = sqrt(t);
}

}  // namespace trunc
}  // namespace mak
