// Per-block descriptors of the batched entry points (plain structs: shared by the CUDA kernels, the
// C-ABI layer and the CPU logic tests).
#pragma once
#include "scalar.h"
namespace mak {
template <typename T>
struct QrBlockDesc {
    int m, n;
    T* A; int lda;
    T* Q; int ldq;
    T* R; int ldr;   // R == nullptr -> not requested
};
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
}  // namespace mak
