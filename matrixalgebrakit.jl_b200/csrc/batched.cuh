#pragma once
#include "common.cuh"
#include "batched_desc.h"
namespace mak {
// shared-memory elements the one-CTA kernel needs for an m x n block, and the largest it accepts
size_t batched_qr_smem_elems(int m, int n);
template <typename T> size_t batched_qr_max_smem_elems();
int batched_init(makb200_handle* h);
// descs: DEVICE array; max_smem_elems: max of batched_qr_smem_elems over the batch
template <typename T>
int batched_qr_smem(makb200_handle* h, int batch, size_t max_smem_elems, const QrBlockDesc<T>* descs, int* info);
// tiny blocks (m, n <= 32): one warp per block; cap_elems = max over the class of (m|1)*n (per-warp smem)
// rmax = row/column capacity of the class (16, 24 or 32): selects the register-resident variant's instantiation
template <typename T> int batched_qr_warp(makb200_handle* h, int batch, int cap_elems, const QrBlockDesc<T>* descs, int rmax);
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
constexpr int BQR_NB = 32;        // widest inner column step (panel width)
constexpr int BQR_NBO_MAX = 128;  // widest outer block (K of the trailing-update GEMMs)
template <typename T>
struct BqrBlock {
    int m, n, k;
    T* A; int lda;
    T* Q; int ldq;
    T* R; int ldr;   // R == nullptr -> not requested
    T* Vw;           // m x nbo   explicit V of the current outer block (ld m, rows relative to J0)
    T* W;            // nbo x max(n,k)
    T* W2;           // nbo x max(n,k)
    T* Tin;          // NB x NB: compact-WY T of the current inner step
    T* G;            // nbo x nbo: V^H V of the current outer block
    T* Tout;         // nbo x nbo per outer block: compact-WY T factors (kept for the Q phase)
    T* tau;          // k
};
struct BqrStep { int j0, jb, active, max_rows, max_nc, J0; };                            // inner step
struct BqrOuter { int J0, active, max_rows, max_ne, max_nc, max_ncq, s_begin, s_end; };  // outer block
struct BqrSchedule {
    int nbo = 64;
    std::vector<BqrStep> steps;
    std::vector<BqrOuter> outer;
};
template <typename T> bool bqr_fits(int m, int n);
template <typename T>
BqrSchedule bqr_schedule(const std::vector<int>& ms, const std::vector<int>& ns, const std::vector<int>& ks);
template <typename T> size_t bqr_block_work_elems(const BqrSchedule& sc, int m, int n, int k);
template <typename T> void bqr_carve_block(const BqrSchedule& sc, BqrBlock<T>& b, T*& p);
// blocks_dev: DEVICE array sorted by k descending (the blocks active at a step are a prefix);
// probs_dev: DEVICE scratch for 4*count GEMM descriptors
template <typename T>
int batched_qr_blocked(makb200_handle* h, int count, const BqrBlock<T>* blocks_dev, const BqrSchedule& sc,
                       GemmProblem<T>* probs_dev);
int batched_blocked_init(makb200_handle* h);
}  // namespace mak
