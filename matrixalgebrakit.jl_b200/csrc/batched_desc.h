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

// one problem of a (grouped) GEMM launch: C = alpha * op(A) * op(B) + beta * C, column-major
template <typename T>
struct GemmProblem {
    int m, n, k;
    const T* A;
    int lda;
    const T* B;
    int ldb;
    T* C;
    int ldc;
    T alpha, beta;
    int conja, conjb;
    int lower;  // 1: C is only needed on/below the diagonal (tiles above are skipped)
};

// ---- lock-step batched polar decomposition (polar_lockstep_plan.h / polar_lockstep.cuh) ----
// One block (m >= n) with its work buffers.  Every work matrix is stored with its own row count as leading dimension.
template <typename T>
struct LsBlk {
    int m, n;
    const T* A; int lda;     // input, m x n (not modified)
    T* W;                    // out: isometry, m x n, ld m
    T* P;                    // out: Hermitian factor, n x n, ld n
    const T* S; int lds;     // the n x n matrix the iteration runs on: A itself (m == n) or R0 of A = Q0 R0
    T* X;                    // n x n   iterate
    T* B;                    // max(2n, m) x n   [sqrt(c) X; I]  /  X Z^-1  /  copy of a tall A
    T* Q;                    // 2n x n  orthonormal basis of B  /  X L^-H
    T* T2;                   // 2n x n  right-hand side of the triangular solves
    T* Z;                    // n x n   I + c X^H X  /  X^H S
    T* L;                    // n x n   Cholesky factor
    T* Linv;                 // nb x nb per diagonal block: inverses
    T* Q0;                   // m x n   (m > n only)
    T* R0;                   // n x n   (m > n only)
};
constexpr int LS_NPROBE = 8;   // Hutchinson probes of the sigma_min estimate
enum LsBuf { LS_X = 0, LS_B, LS_Q, LS_T2, LS_Z, LS_L, LS_W, LS_A };
enum LsKind { LS_GEMM = 0, LS_PREP, LS_STACK, LS_ADDDIAG, LS_AXPBY, LS_COPY, LS_SYMM, LS_POTF2, LS_QR_STACK, LS_QR_TALL, LS_PROBE, LS_FRO };
// one launch of the lock-step sequence.  `count` blocks take part (a prefix of the n-descending order).
struct LsAct {
    int kind;
    int opa, opb, max_m, max_n;   // LS_GEMM (ops: 0 = N, 2 = C as MAKB200_OP_*)
    int count;
    size_t off;                   // LS_GEMM: first descriptor
    int a0, a1, a2;               // LS_COPY: src buffer, dst buffer, row multiple (1: n rows, 2: 2n rows); LS_POTF2: j0, block index
    double p0, p1;                // LS_STACK: sqrt(c); LS_AXPBY: X = p0 X + p1 B
                                  // LS_PROBE: T2 (LS_NPROBE x n, ld LS_NPROBE) = +-1 entries; LS_FRO: est[i] = ||Q (LS_NPROBE x n)||_F^2
};
}  // namespace mak
