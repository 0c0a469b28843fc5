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
                 size_t lwork, bool adjoint = false);   // adjoint: C <- Q^H C
// lock-step batched C_i <- H_0 ... H_{k_i-1} C_i (reflectors below the diagonal of A_i, m_i x k_i): see qr.cu.
// blks: HOST array sorted by k descending with the work buffers carved; tables: device region of ormqr_batched_table_bytes
constexpr int BORM_NB = 64;
template <typename T>
struct OrmqrBatchBlk { int m, k, nc; const T* A; int lda; const T* tau; T* C; int ldc; T *Vw, *G, *Tb, *W, *W2; };
template <typename T> size_t ormqr_batched_block_elems(int m, int nc);
template <typename T> void ormqr_batched_carve(OrmqrBatchBlk<T>& b, T*& p);
template <typename T> size_t ormqr_batched_table_bytes(int count, int kmax);
template <typename T>
int ormqr_left_batched(makb200_handle* h, int count, const OrmqrBatchBlk<T>* blks, char* tables, size_t tables_bytes);
// EXPERIMENTAL: dense -> band (first stage of the two-stage tridiagonalisation); A full Hermitian, in place
template <typename T> size_t sy2sb_worksize_t(makb200_handle* h, int n, int b);
template <typename T> int sy2sb_t(makb200_handle* h, int n, int b, T* A, int lda, T* tau1, void* work, size_t lwork);
}  // namespace mak
