#pragma once
#include "common.cuh"
namespace mak {
size_t stedc_worksize(int n);
// eigen-decomposition of the real symmetric tridiagonal (d, e): w ascending, Z (n x n) orthonormal.
// All pointers device; info_dev (optional) receives 0 / 1 (leaf QL failed to converge).
int stedc(makb200_handle* h, int n, const double* d, const double* e, double* w, double* Z, int ldz, void* work,
          size_t lwork, int* info_dev);
}  // namespace mak
