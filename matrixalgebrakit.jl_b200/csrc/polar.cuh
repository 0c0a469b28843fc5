#pragma once
#include "common.cuh"
#include "nccl_dl.h"
#include "qdwh_schedule.h"
namespace mak {
int polar_init(makb200_handle* h);
template <typename T> size_t polar_worksize_t(makb200_handle* h, int m, int n);
// l0: lower bound on sigma_min(A / ||A||_F); iters_host (optional, HOST) receives the step count
template <typename T>
int polar_qdwh_t(makb200_handle* h, int m, int n, T* A, int lda, T* W, int ldw, T* P, int ldp, double l0,
                 int maxiter, void* work, size_t lwork, int* iters_host, int* info_dev);
template <typename T> size_t svd_worksize_t(makb200_handle* h, int m, int n);
template <typename T>
int svd_t(makb200_handle* h, int m, int n, T* A, int lda, double* S, T* U, int ldu, T* Vh, int ldvh, int fixgauge,
          double l0, void* work, size_t lwork, int* info_dev, int r = 0);   // 0 < r < min(m,n): U m x r, Vh r x n
// phased SVD of one block (batched path, m >= n): see polar.cu
template <typename T> struct TrdPre;
template <typename T> size_t svd_phase_scratch_t(makb200_handle* h, int m, int n);
template <typename T>
int svd_phase1_t(makb200_handle* h, int m, int n, T* A, int lda, T* Wp, T* P, double l0, void* scratch, size_t lscratch);
template <typename T>
int svd_phase2_t(makb200_handle* h, int m, int n, T* Wp, T* P, T* V, double* wv, double* flag, const TrdPre<T>* pre,
                 double* S, T* U, int ldu, T* Vh, int ldvh, int fixgauge, void* scratch, size_t lscratch);
// tall-skinny local QR (CholeskyQR2; nshift > 0: shifted CholeskyQR with that many preconditioning
// passes) used by TSQR; A is overwritten, diag(R) > 0
template <typename T> size_t cholqr2_worksize_t(makb200_handle* h, int m, int n);
template <typename T>
int cholqr2_t(makb200_handle* h, int m, int n, T* A, int lda, T* Q, int ldq, T* R, int ldr, void* work, size_t lwork,
              int* info_dev, int nshift = 0);
// multi-GPU TSQR: local CholeskyQR2 + binary-tree reduction of R over NCCL, tree factors folded into the last solve
template <typename T> size_t tsqr_worksize_t(makb200_handle* h, int m, int n, int nranks);
template <typename T>
int tsqr_t(makb200_handle* h, const NcclApi* api, ncclComm_t comm, int m, int n, T* A, int lda, T* Q, int ldq, T* R,
           int ldr, void* work, size_t lwork, int* info_dev);
// B (n x m) = A^H for A (m x n); tiled, coalesced on both sides
template <typename T> int adjoint_t(makb200_handle* h, int m, int n, const T* A, int lda, T* B, int ldb);
}  // namespace mak
