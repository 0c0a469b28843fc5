// FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) GEMM for Float64 / ComplexF64, column-major.
//   C = alpha * op(A) * op(B) + beta * C,  op in {N, T, C}
// tcgen05 has no f64 kind on sm_100a, so the FP64 tensor path is the warp-level DMMA
// (SASS DMMA.8x8x4); operands are staged global -> shared with a multi-stage cp.async
// pipeline into padded (bank-conflict-free) tiles and read as per-lane fragments.
// Complex products are 4 real DMMAs on the interleaved (re,im) tiles.
#pragma once
#include "common.cuh"
#include "batched_desc.h"   // GemmProblem<T>

namespace mak {


// Host-side launcher.  opa/opb: MAKB200_OP_{N,T,C}.  `ws`/`ws_bytes`: optional split-K
// scratch (may be null -> no split-K).  Asynchronous on `stream`.
template <typename T>
cudaError_t gemm(cudaStream_t stream, int num_sms, int opa, int opb, int m, int n, int k, T alpha,
                 const T* A, int lda, const T* B, int ldb, T beta, T* C, int ldc,
                 void* ws = nullptr, size_t ws_bytes = 0, bool lower = false);

// Grouped launch: `count` problems described in DEVICE memory (dims may be produced on the
// device by an earlier kernel); max_m/max_n are host upper bounds used to size the grid.
// All problems share opa/opb. Problems with m, n or k <= 0 are skipped (k<=0: C = beta*C).
template <typename T>
cudaError_t gemm_grouped(cudaStream_t stream, int opa, int opb, int count, int max_m, int max_n,
                         const GemmProblem<T>* problems_dev);

}  // namespace mak
