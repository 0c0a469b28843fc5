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
// tiny blocks (m, n <= 32): one warp per block; cap_elems = max over the class of (m|1)*n (per-warp smem)
template <typename T> int batched_qr_warp(makb200_handle* h, int batch, int cap_elems, const QrBlockDesc<T>* descs);
template <typename T>
struct SvdBlockDesc {
    int m, n, fixgauge;
    const T* A; int lda;
    double* S;
    T* U; int ldu;     // U == nullptr -> values only
    T* Vh; int ldvh;
};
template <typename T>
struct EighBlockDesc {
    int n, fixgauge;
    const T* A; int lda;
    double* W;
    T* V; int ldv;     // V == nullptr -> values only
};
size_t batched_eigh_smem_bytes(int n, size_t elem);
size_t batched_eigh_max_smem_bytes();
template <typename T>
int batched_eigh_smem(makb200_handle* h, int batch, size_t max_smem_bytes, const EighBlockDesc<T>* descs, int* info);
size_t batched_svd_smem_bytes(int m, int n, size_t elem);
size_t batched_svd_max_smem_bytes();
template <typename T>
int batched_svd_smem(makb200_handle* h, int batch, size_t max_smem_bytes, const SvdBlockDesc<T>* descs, int* info);
}  // namespace mak

// ---- lock-step blocked QR over many mid-size blocks (batched_blocked.cu) ----------------------
#include <vector>
namespace mak {
template <typename T> struct GemmProblem;
constexpr int BQR_NB = 32;   // widest column step
template <typename T>
struct BqrBlock {
    int m, n, k;
    T* A; int lda;
    T* Q; int ldq;
    T* R; int ldr;   // R == nullptr -> not requested
    T* Vw;           // m x NB   explicit V of the current step (ld m)
    T* W;            // NB x max(n,k)
    T* W2;           // NB x max(n,k)
    T* Tf;           // NB x NB per step: compact-WY T factors
};
struct BqrStep { int j0, jb, active, max_rows, max_nc, max_ncq; };
template <typename T> bool bqr_fits(int m, int n);
template <typename T>
std::vector<BqrStep> bqr_steps(const std::vector<int>& ms, const std::vector<int>& ns, const std::vector<int>& ks);
template <typename T> size_t bqr_block_work_elems(int m, int n, int k, int nsteps);
// blocks_dev: DEVICE array sorted by k descending (the blocks active at a step are a prefix);
// probs_dev: DEVICE scratch for 3*count GEMM descriptors
template <typename T>
int batched_qr_blocked(makb200_handle* h, int count, const BqrBlock<T>* blocks_dev, const std::vector<BqrStep>& steps,
                       GemmProblem<T>* probs_dev);
int batched_blocked_init(makb200_handle* h);
}  // namespace mak
