// band -> tridiagonal bulge chasing (second stage of the two-stage tridiagonalisation); experimental
#pragma once
#include "common.cuh"
namespace mak {
template <typename T> size_t sbr_chase_worksize_t(int n, int b);
// A: n x n Hermitian, only the lower band of width b is read (not modified).  d[n], e[n-1]: the
// tridiagonal; V2 (ldv x n), tau2 (ldt x n, ldt >= ceil(n/b)+1): the chase reflectors (sbr_core.h layout)
template <typename T>
int sbr_chase_t(makb200_handle* h, int n, int b, const T* A, int lda, double* d, double* e, T* V2, int ldv, T* tau2,
                int ldt, void* work, size_t lwork);
// Z (n x ncols) <- Q2 Z with diamond blocks of g sweeps (grouped DMMA GEMMs per diamond wavefront)
template <typename T> size_t sbr_apply_q2_worksize_t(int n, int b, int g, int ncols);
template <typename T>
int sbr_apply_q2_t(makb200_handle* h, int n, int b, int g, const T* V2, int ldv, const T* tau2, int ldt, T* Z, int ldz,
                   int ncols, void* work, size_t lwork);
}  // namespace mak
