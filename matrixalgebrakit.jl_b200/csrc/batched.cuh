#pragma once
#include "common.cuh"
namespace mak {
template <typename T>
struct QrBlockDesc {
    int m, n;
    T* A; int lda;
    T* Q; int ldq;
    T* R; int ldr;   // R == nullptr -> not requested
};
// shared-memory elements the one-CTA kernel needs for an m x n block, and the largest it accepts
size_t batched_qr_smem_elems(int m, int n);
template <typename T> size_t batched_qr_max_smem_elems();
int batched_init(makb200_handle* h);
// descs: DEVICE array; max_smem_elems: max of batched_qr_smem_elems over the batch
template <typename T>
int batched_qr_smem(makb200_handle* h, int batch, size_t max_smem_elems, const QrBlockDesc<T>* descs, int* info);
// tiny blocks (m, n <= 32): one warp per block, rmax in {16, 24, 32} = row capacity of the variant
template <typename T> int batched_qr_warp(makb200_handle* h, int batch, int rmax, const QrBlockDesc<T>* descs);
template <typename T>
struct SvdBlockDesc {
    int m, n, fixgauge;
    const T* A; int lda;
    double* S;
    T* U; int ldu;     // U == nullptr -> values only
    T* Vh; int ldvh;
};
size_t batched_svd_smem_bytes(int m, int n, size_t elem);
size_t batched_svd_max_smem_bytes();
template <typename T>
int batched_svd_smem(makb200_handle* h, int batch, size_t max_smem_bytes, const SvdBlockDesc<T>* descs, int* info);
}  // namespace mak
