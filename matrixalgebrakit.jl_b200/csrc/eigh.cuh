#pragma once
#include "common.cuh"
namespace mak {
template <typename T> int herm_defect_t(makb200_handle* h, int n, const T* A, int lda, double* out2);
// B = (A +- A^H)/2 (B may be A); out4 = {||vanishing part||_F^2, max|A_ij|, ||A||_F^2, #exact mismatches};
// out2 = {||P||_F^2, ||P - I||_F^2}
template <typename T> int project_herm_t(makb200_handle* h, int anti, int n, const T* A, int lda, T* B, int ldb);
template <typename T> int herm_props_t(makb200_handle* h, int anti, int n, const T* A, int lda, double* out4);
template <typename T> int gram_defect_t(makb200_handle* h, int n, const T* P, int ldp, double* out2);
// mode 0: A = I, 1: zero below the diagonal (uppertriangular!), 2: zero above it (lowertriangular!)
template <typename T> int tri_init_t(makb200_handle* h, int mode, int m, int n, T* A, int lda);
// out1[0] = ||A||_F^2
template <typename T> int fro2_t(makb200_handle* h, int m, int n, const T* A, int lda, double* out1);
template <typename T> size_t eigh_worksize_t(makb200_handle* h, int n);
template <typename T>
struct TrdPre { double* d; double* e; T* tau; };   // a block already tridiagonalised in place (bhetrd_batched_t)
template <typename T>
int eigh_t(makb200_handle* h, int n, T* A, int lda, double* W, T* V, int ldv, int fixgauge, void* work, size_t lwork,
           int* info_dev, int top = 0, const TrdPre<T>* pre = nullptr);
// V == nullptr: values only; 0 < top < n: only the last `top` eigenvectors; pre: skip mirror + hetrd
// EXPERIMENTAL (MAKB200_BHETRD=1): tridiagonalise `nblk` blocks (n <= BHETRD_MAX_N) in ONE launch, one CTA per block;
// descs_dev: DEVICE array of BhetrdDesc<T> (bhetrd.cuh); reads the upper triangle (uplo = 'U')
constexpr int BHETRD_MAX_N = 512;
template <typename T> int bhetrd_batched_t(makb200_handle* h, int nblk, const void* descs_dev, int nmax);
// V[:, j] *= conj(sign(pivot_j)); optional `other` (k x other_n, row j scaled by sign(pivot_j))
template <typename T>
int gauge_columns(makb200_handle* h, int m, int ncols, T* V, int ldv, T* other, int ldo, int other_n);
}  // namespace mak
