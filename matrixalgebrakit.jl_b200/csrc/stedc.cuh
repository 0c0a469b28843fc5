#pragma once
#include "common.cuh"
namespace mak {
size_t stedc_worksize(int n);
// eigen-decomposition of the real symmetric tridiagonal (d, e): w ascending, Z (n x n) orthonormal.
// All pointers device; info_dev (optional) receives 0 / 1 (leaf QL failed to converge).
int stedc(makb200_handle* h, int n, const double* d, const double* e, double* w, double* Z, int ldz, void* work,
          size_t lwork, int* info_dev);
// The same solver for MANY tridiagonals in one pass of the kernels (lock-step batched eigh / svd): see stedc.cu.
// blks: HOST array; d, e (n, n-1), w (n) and V (n x n of T, leading dimension ldv) are device pointers.
struct StedcBlk { int n; const double* d; const double* e; double* w; void* V; int ldv; };
size_t stedc_batched_worksize(int nblk, const int* n);
template <typename T>
int stedc_batched(makb200_handle* h, int nblk, const StedcBlk* blks, void* work, size_t lwork, int* info_dev);
}  // namespace mak
