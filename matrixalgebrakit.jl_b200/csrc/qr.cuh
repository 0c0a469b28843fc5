#pragma once
#include "common.cuh"
namespace mak {
int qr_init(makb200_handle* h);
template <typename T> size_t qr_worksize_t(makb200_handle* h, int m, int n, int ncols_q);
template <typename T>
int qr_fused_t(makb200_handle* h, int mode, int m, int n, T* A, int lda, T* Q, int ldq, T* R, int ldr, void* work,
               size_t lwork);
template <typename T> int geqrf_t(makb200_handle* h, int m, int n, T* A, int lda, T* tau, void* work, size_t lwork);
template <typename T>
int orgqr_t(makb200_handle* h, int m, int ncols, int k, const T* A, int lda, const T* tau, T* Q, int ldq, void* work,
            size_t lwork);
template <typename T> size_t ormqr_worksize_t(makb200_handle* h, int m, int k, int nc);
template <typename T>
int ormqr_left_t(makb200_handle* h, int m, int k, const T* A, int lda, const T* tau, T* C, int ldc, int nc, void* work,
                 size_t lwork);
// stable non-negative-beta reflector (shared with eigh.cu)
}  // namespace mak
