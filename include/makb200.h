/* makb200 — C ABI of the B200-native dense-factorization library (sm_100a).
 *
 * Drop-in boundary for MatrixAlgebraKit.jl v0.6.9: every entry point below replaces a
 * LAPACK (`src/yalapack.jl`, ccall into libblastrampoline) or cuSOLVER
 * (`ext/MatrixAlgebraKitCUDAExt/yacusolver.jl`) call site on the path
 * qr_compact!/qr_full!, svd_compact!/svd_trunc!, eigh_full!, left_polar!.
 * The Julia extension `ext/MatrixAlgebraKitB200Ext` `ccall`s exactly these symbols;
 * the Python host layer (`matrixalgebrakit.jl_b200/`) binds the same symbols via ctypes.
 *
 * Conventions (all from the reference's shims):
 *  - column-major, unit stride in dim 1, explicit leading dimensions in ELEMENTS
 *    (`lda = stride(A,2)`, yalapack.jl:173,179; yacusolver.jl:72-74);
 *  - all matrix/vector pointers are DEVICE pointers; plain sizes are host ints;
 *  - return value: 0 ok, -i = i-th argument invalid (LAPACK `info<0`, yalapack.jl:193),
 *    >0 numerical failure / CUDA error (MAKB200_ERR_*). Never throws, never allocates
 *    device memory: scratch comes from the caller (`*_worksize` query + `work`,`lwork`),
 *    mirroring CUDA.jl's `with_workspace` (yacusolver.jl:76-90);
 *  - asynchronous on the handle's stream; no host sync unless documented;
 *  - dtype: MAKB200_F64 = Float64, MAKB200_C128 = ComplexF64 (interleaved re,im).
 */
#ifndef MAKB200_H
#define MAKB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct makb200_handle makb200_handle_t;

enum { MAKB200_F64 = 0, MAKB200_C128 = 1 };
enum { MAKB200_QR_COMPACT = 0, MAKB200_QR_FULL = 1 };
enum { MAKB200_OP_N = 0, MAKB200_OP_T = 1, MAKB200_OP_C = 2 };
enum { MAKB200_ERR_CUDA = 1000, MAKB200_ERR_WORKSPACE = 1001, MAKB200_ERR_NOCONV = 1002, MAKB200_ERR_NCCL = 1003 };

/* -- handle: plays the role of cuSOLVER.dense_handle() (yacusolver.jl:76) ------------ */
int makb200_create(makb200_handle_t** h, int device);
int makb200_destroy(makb200_handle_t* h);
int makb200_set_stream(makb200_handle_t* h, void* cuda_stream);
int makb200_version(void);
/* last CUDA error string recorded on this handle (host memory, owned by the library) */
const char* makb200_last_error(makb200_handle_t* h);

/* -- instrumentation for bench.py (no effect on results) -------------------------------------
 * makb200_launch_count: kernels launched by this library in this process so far.
 * makb200_kernel_timing(1): bracket the dominant kernels with CUDA events on the launching
 *   stream; makb200_kernel_time(which, &ms, &launches) collects and resets
 *   (which = 0: tridiagonalisation column-dot kernel, 1: DMMA GEMM). */
unsigned long long makb200_launch_count(void);
int makb200_kernel_timing(int enable);
int makb200_kernel_time(int which, double* ms, int* launches);
/* real flops (2mnk, x4 for complex) issued through the DMMA GEMM since makb200_kernel_timing() */
double makb200_gemm_flops(void);

/* -- GEMM building block: C = alpha*op(A)*op(B) + beta*C  (FP64 DMMA tiles) ----------
 * replaces `mul!` on CuArray -> cuBLAS gemm (implementations/polar.jl:63,88).
 * alpha/beta: HOST pointers to one scalar of `dtype`. */
int makb200_gemm(makb200_handle_t* h, int dtype, int opa, int opb, int m, int n, int k,
                 const void* alpha, const void* A, int lda, const void* B, int ldb,
                 const void* beta, void* C, int ldc);

/* -- L1 (LAPACK-shaped) QR shims -----------------------------------------------------
 * makb200_geqrf  replaces geqrf!/geqrt! (yalapack.jl:168-200,303-333; yacusolver.jl:12).
 *   On exit A holds R (upper) and the Householder vectors (below the diagonal, unit
 *   diagonal implicit), `tau` the k=min(m,n) scalar factors.  Reflector convention:
 *   H_j = I - tau_j v v^H with NON-NEGATIVE real R[j,j] (src/common/householder.jl:35-67),
 *   so the reference's QR gauge (common/gauge.jl:16-25) is already satisfied.
 * makb200_orgqr   replaces ungqr!/gemqrt!-on-identity (yalapack.jl:550-583,882-928):
 *   Q (m x ncols, ncols<=m) = first ncols columns of H_1...H_k. */
size_t makb200_geqrf_worksize(makb200_handle_t* h, int dtype, int m, int n);
int makb200_geqrf(makb200_handle_t* h, int dtype, int m, int n, void* A, int lda, void* tau,
                  void* work, size_t lwork);
size_t makb200_orgqr_worksize(makb200_handle_t* h, int dtype, int m, int ncols, int k);
int makb200_orgqr(makb200_handle_t* h, int dtype, int m, int ncols, int k, const void* A,
                  int lda, const void* tau, void* Q, int ldq, void* work, size_t lwork);

/* makb200_ormqr  replaces unmqr!/ormqr (yalapack.jl:688-735; yacusolver.jl:14; MatrixAlgebraKitCUDAExt.jl:32-34),
 *   the consumer being qr_null_householder! (implementations/qr.jl:236-262: unmqr!(driver,'L','N',A,tau,N)):
 *   C (m x n) <- Q C (trans = MAKB200_OP_N) or Q^H C (MAKB200_OP_C; MAKB200_OP_T for Float64), Q = H_1...H_k from
 *   makb200_geqrf (A m x k below the diagonal, tau).  side: 0 = left; the right side is not provided (-3).
 *   Compact-WY blocks of 128 reflectors, three DMMA GEMMs per block. */
size_t makb200_ormqr_worksize(makb200_handle_t* h, int dtype, int m, int n, int k);
int makb200_ormqr(makb200_handle_t* h, int dtype, int side, int trans, int m, int n, int k, const void* A,
                  int lda, const void* tau, void* C, int ldc, void* work, size_t lwork);

/* -- L2 fused QR: qr_householder!(driver, A, Q, R; positive) (implementations/qr.jl:132-188)
 *   mode COMPACT: Q m x k, R k x n; FULL: Q m x m, R m x n (k = min(m,n)).
 *   R == NULL or ldr == 0  => "R not requested" (zero-length R, qr.jl:149).
 *   A is destroyed.  `positive` is accepted for signature parity; the factorization always
 *   has diag(R) >= 0, which is a valid result for positive=false as well. */
size_t makb200_qr_worksize(makb200_handle_t* h, int dtype, int mode, int m, int n);
int makb200_qr(makb200_handle_t* h, int dtype, int mode, int positive, int m, int n, void* A,
               int lda, void* Q, int ldq, void* R, int ldr, void* work, size_t lwork);

/* -- batched QR of many small blocks (one CTA per block, block resident in shared memory) ---
 * New capability: semantics = qr_compact!(A_i,(Q_i,R_i)) applied per block
 * (downstream TensorKit-style block loops call the single-matrix op once per block).
 * m,n,lda,ldq,ldr: HOST int arrays (Julia knows the block sizes on the host);
 * A,Q,R: HOST arrays of DEVICE pointers; R may be NULL (or R[i] NULL) = not requested.
 * Blocks that do not fit one CTA's shared memory are routed through the blocked DMMA path.
 * info: DEVICE int[batch] or NULL. */
size_t makb200_qr_batched_worksize(makb200_handle_t* h, int dtype, int batch, const int* m,
                                   const int* n);
int makb200_qr_batched(makb200_handle_t* h, int dtype, int batch, const int* m, const int* n,
                       void* const* A, const int* lda, void* const* Q, const int* ldq,
                       void* const* R, const int* ldr, int* info, void* work, size_t lwork);

/* Plan form of makb200_qr_batched for block structures that are reused across calls (a block-sparse
 * tensor keeps its sectors): create() classifies the blocks, carves `work` and uploads the
 * descriptors once (asynchronously on the handle's stream); run() only launches kernels.
 * `work` (makb200_qr_batched_worksize bytes) and all block pointers must stay valid while the
 * plan lives.  info: DEVICE int[batch] or NULL. */
typedef struct makb200_qr_batched_plan makb200_qr_batched_plan_t;
int makb200_qr_batched_plan_create(makb200_handle_t* h, int dtype, int batch, const int* m, const int* n,
                                   void* const* A, const int* lda, void* const* Q, const int* ldq,
                                   void* const* R, const int* ldr, void* work, size_t lwork,
                                   makb200_qr_batched_plan_t** plan);
int makb200_qr_batched_plan_run(makb200_handle_t* h, makb200_qr_batched_plan_t* plan, int* info);
int makb200_qr_batched_plan_destroy(makb200_qr_batched_plan_t* plan);

/* -- eigh_full! ------------------------------------------------------------------------
 * makb200_hermitian_defect: the device half of check_hermitian (implementations/eigh.jl:11-18,
 *   matrixproperties.jl:150-172; MatrixAlgebraKitCUDAExt.jl:147-152): out2_dev[0] =
 *   ||(A - A^H)/2||_F^2, out2_dev[1] = max |A_ij| (for default_hermitian_tol, defaults.jl:44).
 * makb200_eigh: replaces heevd!/heevr! + gaugefix!(eigh_full!) (yalapack.jl:1164-1362,
 *   yacusolver.jl:766-810, common/gauge.jl:38-45).  Only the upper triangle of A is read
 *   (uplo='U', yalapack.jl:994,1286); A is destroyed.  W: n real eigenvalues ascending;
 *   V: n x n eigenvectors (must not alias A); V == NULL: values only (eigh_vals!, eigh.jl:157-161,
 *   LAPACK job 'N'): after the tridiagonalisation the eigenvalues come from a Sturm-count K-section
 *   kernel (one thread per eigenvalue) and neither the D&C eigenvector GEMMs nor the
 *   back-transformation run.  info_dev: optional DEVICE int, >0 if the tridiagonal solver failed
 *   to converge.
 *   Pipeline: blocked Householder tridiagonalisation (HBM-bound column-dot kernel + DMMA her2k),
 *   tridiagonal divide & conquer (makb200_stedc), compact-WY back-transformation, fused gauge. */
int makb200_hermitian_defect(makb200_handle_t* h, int dtype, int n, const void* A, int lda,
                             double* out2_dev);
/* project_hermitian! / project_antihermitian! (implementations/projections.jl:60-139; on a CuArray the
 * reference's NativeBlocked loop is O((n/32)^2) broadcasts): B = (A + A^H)/2 (anti = 0) or
 * (A - A^H)/2 (anti != 0) in ONE launch, entry for entry the reference's arithmetic.  B may be A
 * (in place, the reference's default output, projections.jl:38-43) or a distinct n x n matrix. */
int makb200_project_hermitian(makb200_handle_t* h, int dtype, int anti, int n, const void* A, int lda,
                              void* B, int ldb);
/* ishermitian / isantihermitian (common/matrixproperties.jl:77-195), every ingredient in one pass:
 * out4_dev (DEVICE double[4]) = { ||part that must vanish||_F^2  (anti = 0: (A - A^H)/2),
 * max |A_ij|, ||A||_F^2, number of entries i <= j with A_ij != +-conj(A_ji) (the exact test) }. */
int makb200_hermitian_props(makb200_handle_t* h, int dtype, int anti, int n, const void* A, int lda,
                            double* out4_dev);
/* is_left_isometric (common/matrixproperties.jl:53-58) on the Gram matrix P = A^H A (makb200_gemm):
 * out2_dev (DEVICE double[2]) = { ||P||_F^2, ||P - I||_F^2 }. */
int makb200_gram_defect(makb200_handle_t* h, int dtype, int n, const void* P, int ldp, double* out2_dev);
/* out1_dev[0] = ||A||_F^2 (DEVICE double).  Used after makb200_polar_qdwh: ||W||_F^2 = n for an isometric polar
 * factor, = rank(A) for the partial isometry QDWH converges to on singular input (the host layer then
 * takes the PolarViaSVD recipe, implementations/polar.jl:59-70, so that W is isometric as with LAPACK). */
int makb200_fro2(makb200_handle_t* h, int dtype, int m, int n, const void* A, int lda, double* out1_dev);
/* one! / uppertriangular! / lowertriangular! (src/common/initialization.jl:11-36; on a CuArray the
 * reference's uppertriangular! is one zero! launch per column): one launch.
 * mode 0: A = I (rectangular identity), 1: zero strictly below the diagonal, 2: zero strictly above it. */
int makb200_tri_init(makb200_handle_t* h, int dtype, int mode, int m, int n, void* A, int lda);
size_t makb200_eigh_worksize(makb200_handle_t* h, int dtype, int n);
int makb200_eigh(makb200_handle_t* h, int dtype, int fixgauge, int n, void* A, int lda, double* W,
                 void* V, int ldv, void* work, size_t lwork, int* info_dev);
/* symmetric tridiagonal divide & conquer (stedc class): d[n], e[n-1] -> W ascending, Z n x n */
size_t makb200_stedc_worksize(makb200_handle_t* h, int n);
int makb200_stedc(makb200_handle_t* h, int n, const double* d, const double* e, double* W, double* Z,
                  int ldz, void* work, size_t lwork, int* info_dev);

/* -- left_polar! (QDWH) ---------------------------------------------------------------------
 * New algorithm `B200_QDWH` (the reference has PolarViaSVD / PolarNewton, implementations/
 * polar.jl:59-166; contract = its result: A = W P, W isometric m x n, P Hermitian PSD n x n).
 * m >= n required (polar.jl:9-10).  P == NULL or ldp == 0: P not requested (zero-length P,
 * polar.jl:14,64,102).  l0: lower bound for sigma_min(A/||A||_F), <= 0 selects eps (valid for
 * every kappa <= 1e16).  A is destroyed.  iters_host: optional HOST int (QDWH steps scheduled).
 * Host synchronisation: none for n < 1024 or a caller-supplied l0 > 0 (the schedule is a function of l0 alone and is
 * computed on the host before the first launch).  For n >= 1024 with l0 <= 0 the sigma_max / sigma_min estimates are
 * formed on the device and ONE 24-byte device-to-host read (the reference's `info`-style read, yacusolver.jl:101-181)
 * picks the schedule; MAKB200_QDWH_ESTIMATE=0 or l0 > 0 avoids it. */
size_t makb200_polar_worksize(makb200_handle_t* h, int dtype, int m, int n);
int makb200_polar_qdwh(makb200_handle_t* h, int dtype, int m, int n, void* A, int lda, void* W,
                       int ldw, void* P, int ldp, double l0, int maxiter, void* work, size_t lwork,
                       int* iters_host, int* info_dev);

/* -- svd_compact! / svd_vals! ----------------------------------------------------------------
 * replaces gesdd!/gesvd!/gesvdp! + gaugefix!(svd_compact!) (yalapack.jl:1959-2202,
 * yacusolver.jl:101-181, common/gauge.jl:69-77).  Algorithm: QDWH polar + Hermitian D&C
 * (the reference's `SVDViaPolar` tag).  U m x k, S k (descending, >= 0), Vh k x n, k = min(m,n);
 * U == Vh == NULL: values only (job 'N').  m < n handled through A^H (svd.jl:134-142). */
size_t makb200_svd_worksize(makb200_handle_t* h, int dtype, int m, int n);
int makb200_svd(makb200_handle_t* h, int dtype, int fixgauge, int m, int n, void* A, int lda, double* S,
                void* U, int ldu, void* Vh, int ldvh, double l0, void* work, size_t lwork,
                int* info_dev);

/* svd_trunc! with the rank known before the decomposition (truncrank(r): svd.jl:226-237,
 * truncation.jl:54-58; BASELINE config 5 keeps 1024 of 16384 triplets).  The reference computes the
 * full compact SVD and slices; here S still receives ALL k = min(m,n) singular values (the truncation
 * error needs the discarded ones) but only the r leading triplets' vectors are formed: the
 * back-transformation runs on r columns and U is an m x r x n product.  U: m x r, Vh: r x n
 * (ldvh >= r), 1 <= r <= k; same workspace as makb200_svd; gauge as in makb200_svd. */
int makb200_svd_leading(makb200_handle_t* h, int dtype, int fixgauge, int m, int n, int r, void* A, int lda,
                        double* S, void* U, int ldu, void* Vh, int ldvh, double l0, void* work,
                        size_t lwork, int* info_dev);

/* -- TSQR building block: local tall-skinny QR of one row shard ------------------------------
 * New capability: qr_compact! of a row-sharded m x n matrix (m >> n) =
 * local factorization per rank + binary-tree reduction of the n x n R factors over NCCL (host
 * layer: matrixalgebrakit.jl_b200/tsqr.py).  The local step is CholeskyQR2 on the DMMA GEMM
 * (Gram matrix, blocked Cholesky, blocked TRSM, twice); requires kappa(A) <~ 1e7, reports a
 * Cholesky breakdown in info_dev.  A is overwritten; Q m x n; R n x n upper, diag(R) > 0. */
size_t makb200_tsqr_local_worksize(makb200_handle_t* h, int dtype, int m, int n);
int makb200_tsqr_local(makb200_handle_t* h, int dtype, int m, int n, void* A, int lda, void* Q,
                       int ldq, void* R, int ldr, void* work, size_t lwork, int* info_dev);
/* Same with `nshift` (0..3) shifted-CholeskyQR preconditioning passes in front (Fukaya et al. 2020:
 * G + sI, s = 11(mn + n(n+1)) u ||A||_F^2): every pass divides kappa by ~1/sqrt(11(mn+n^2)u), so
 * nshift = 1 covers kappa <~ 1e10 and nshift = 2 any numerically full-rank input, at (2+nshift)/2
 * the cost.  Same workspace size as makb200_tsqr_local. */
int makb200_tsqr_local_ex(makb200_handle_t* h, int dtype, int m, int n, void* A, int lda, void* Q,
                          int ldq, void* R, int ldr, int nshift, void* work, size_t lwork,
                          int* info_dev);

/* -- multi-GPU TSQR: qr_compact! of a row-sharded tall-skinny matrix (BASELINE config 4) -------------
 * New capability (SURVEY.md 8b "multi-GPU: tsqr(h, ncclComm_t, ...)", 8e).  One process per GPU; rank p
 * passes its m_local x n shard (m_local may differ per rank, n <= sum of the m_local).  Result = the
 * single-GPU qr_compact!(vcat(A_0, ..., A_{P-1})): Q_p (m_local x n) on rank p, the same R (n x n upper,
 * diag(R) > 0: common/gauge.jl:16-25 is satisfied) on every rank.
 *   local step     CholeskyQR2 on the DMMA GEMM (as makb200_tsqr_local; kappa(A) <~ 1e7, breakdown -> info_dev)
 *   exchange step  binary tree over ranks on the n x n R factors: ncclSend/ncclRecv issued on the handle's
 *                  stream (no host synchronisation), Householder QR of each stacked pair, the path product of
 *                  the tree factors sent back down and folded into the last local triangular solve
 *                  (Q_p = Q1_p (L2^-H T_p): no tall-matrix work beyond the single-GPU algorithm), ncclBroadcast of R.
 * `nccl_comm` is an ncclComm_t (NCCL.jl: `comm.handle`; Python: makb200_comm_create below) whose rank
 * order is the row order of the shards; NULL = single rank.  A is overwritten.  NCCL is resolved with
 * dlopen("libnccl.so.2") at first use (the copy already mapped by the process wins; env MAKB200_NCCL_LIB
 * overrides), so the single-GPU entry points have no NCCL dependency.
 * makb200_nccl_unique_id / makb200_comm_create / _destroy: thin wrappers of ncclGetUniqueId /
 * ncclCommInitRank / ncclCommDestroy for hosts without their own NCCL binding; id128: HOST, 128 bytes. */
int makb200_nccl_unique_id(void* id128);
int makb200_comm_create(void** nccl_comm, int nranks, int rank, const void* id128);
int makb200_comm_destroy(void* nccl_comm);
size_t makb200_tsqr_worksize(makb200_handle_t* h, int dtype, int m_local, int n, int nranks);
int makb200_tsqr(makb200_handle_t* h, void* nccl_comm, int dtype, int m_local, int n, void* A, int lda,
                 void* Q, int ldq, void* R, int ldr, void* work, size_t lwork, int* info_dev);

/* -- batched svd_compact! of many small blocks -------------------------------------------------
 * New capability (reference: commented-out gesvdjBatched stubs yacusolver.jl:506-569).
 * Semantics = svd_compact!(A_i,(U_i,S_i,Vh_i)) per block incl. the SVD gauge.  One CTA per block,
 * one-sided Jacobi in shared memory; blocks that do not fit are routed through makb200_svd.
 * m,n,lda,ldu,ldvh: HOST int arrays; A,S,U,Vh: HOST arrays of DEVICE pointers (U == Vh == NULL:
 * values only).  Small blocks are NOT destroyed; routed large blocks are.  info: DEVICE int[batch]. */
size_t makb200_svd_batched_worksize(makb200_handle_t* h, int dtype, int batch, const int* m,
                                    const int* n);
int makb200_svd_batched(makb200_handle_t* h, int dtype, int fixgauge, int batch, const int* m,
                        const int* n, void* const* A, const int* lda, void* const* S, void* const* U,
                        const int* ldu, void* const* Vh, const int* ldvh, int* info, void* work,
                        size_t lwork);

/* -- batched eigh_full! / eigh_vals! of many small Hermitian blocks ---------------------------
 * New capability (reference: commented-out heevjBatched stubs, yacusolver.jl:571-649; per-block
 * semantics = eigh_full!(A_i,(D_i,V_i)) incl. the eigh gauge, implementations/eigh.jl:123-156).
 * One CTA per block, two-sided Jacobi in shared memory (uplo='U' is read, A is NOT destroyed);
 * blocks that do not fit are routed through makb200_eigh (those ARE destroyed).
 * n,lda,ldv: HOST int arrays; A,W,V: HOST arrays of DEVICE pointers (V == NULL: values only).  info: DEVICE int[batch] or NULL. */
size_t makb200_eigh_batched_worksize(makb200_handle_t* h, int dtype, int batch, const int* n);
int makb200_eigh_batched(makb200_handle_t* h, int dtype, int fixgauge, int batch, const int* n,
                         void* const* A, const int* lda, void* const* W, void* const* V,
                         const int* ldv, int* info, void* work, size_t lwork);

/* -- truncation search for svd_trunc! over a batch of sorted spectra -----------------------------
 * findtruncated_svd + truncation_error! (implementations/truncation.jl:54-102,168-174) for the strategy
 * TruncationStrategy(; atol, rtol, maxrank, maxerror, minrank) builds (interface/truncation.jl:37-66):
 *   ( truncrank(maxrank) & trunctol(vatol, vrtol, vp) & truncerror(eatol, ertol, ep) ) | truncrank(minrank)
 * with every absent component switched off (maxrank / minrank < 0, by_value / by_error = 0).  On sorted
 * singular values each component keeps a prefix, so the result is a rank.  The reference's GPU path
 * copies every values vector to the host (MatrixAlgebraKitCUDAExt.jl:64-66); here rank_dev[i] and
 * eps_dev[i] = ||S_i[rank_i:]||_2 for ALL blocks come from one launch (one thread per block).
 * k: HOST int[batch]; S: HOST array of DEVICE pointers (descending values); rank_dev / eps_dev: DEVICE. */
typedef struct makb200_trunc_spec {
    int maxrank, minrank;      /* < 0: not set */
    int by_value, by_error;    /* 0: not set */
    double vatol, vrtol, vp;   /* trunctol:   keep |s| >= max(vatol, vrtol ||S||_vp) */
    double eatol, ertol, ep;   /* truncerror: discarded ||.||_ep < max(eatol, ertol ||S||_ep) */
} makb200_trunc_spec;
size_t makb200_trunc_select_batched_worksize(makb200_handle_t* h, int batch);
/* maxrank_blk: optional HOST int[batch] of per-block rank caps (BASELINE config 3 truncates block i at
 * truncrank(n_i / 2)); entry < 0 or NULL: no per-block cap.  It intersects with spec->maxrank. */
int makb200_trunc_select_batched(makb200_handle_t* h, int batch, const int* k, double* const* S,
                                 const makb200_trunc_spec* spec, const int* maxrank_blk, int* rank_dev,
                                 double* eps_dev, void* work, size_t lwork);

/* -- gaugefix!(eigh_full!, V) on its own (common/gauge.jl:38-45; also the per-column rule of the svd_full! gauge,
 * gauge.jl:47-67, for the columns / rows beyond min(m,n)): every column of V (m x ncols) is multiplied by
 * conj(sign(first entry of maximal modulus)).  One launch (the reference loops over columns with a host scalar). */
int makb200_gauge_columns(makb200_handle_t* h, int dtype, int m, int ncols, void* V, int ldv);

/* -- adjoint: B (n x m) = A^H.  Used by the LQ family, which every GPU driver of the reference
 * routes through QR of the adjoint (lq_via_qr!, implementations/lq.jl:130-131,303-327), and by
 * svd_via_adjoint! (implementations/svd.jl:134-142). */
int makb200_adjoint(makb200_handle_t* h, int dtype, int m, int n, const void* A, int lda, void* B,
                    int ldb);

/* -- EXPERIMENTAL (round 1): second stage of a two-stage Hermitian tridiagonalisation ------------
 * Band (lower bandwidth b <= 64) -> real symmetric tridiagonal by bulge chasing, T = Q2^H B Q2
 * (what LAPACK's ?hbtrd / the sb2st stage of ?hetrd_2stage do; replaces nothing in the reference yet:
 * eigh_full! still runs the one-stage reduction).  A: n x n, only the lower band is read.  d[n],
 * e[n-1]: DEVICE doubles.  V2 (ldv x n), tau2 (ldt x n, ldt >= ceil(n/b)+1): the chase reflectors in
 * the layout of csrc/sbr_core.h (sweep s in column s).  All pointers DEVICE. */
size_t makb200_sbr_chase_worksize(makb200_handle_t* h, int dtype, int n, int b);
int makb200_sbr_chase(makb200_handle_t* h, int dtype, int n, int b, const void* A, int lda, double* d,
                      double* e, void* V2, int ldv, void* tau2, int ldt, void* work, size_t lwork);

/* EXPERIMENTAL: first stage, dense -> band (what ?hetrd_he2hb does).  A: FULL n x n Hermitian (both
 * triangles), overwritten: lower band (width b) = Q1^H A Q1, reflectors below the band (QR-type columns
 * of A[b:, 0:n-b]); tau1: DEVICE, n entries. */
size_t makb200_sy2sb_worksize(makb200_handle_t* h, int dtype, int n, int b);
int makb200_sy2sb(makb200_handle_t* h, int dtype, int n, int b, void* A, int lda, void* tau1, void* work,
                  size_t lwork);
/* EXPERIMENTAL: Z (n x ncols) <- Q2 Z with the reflectors of makb200_sbr_chase, diamond blocks of g sweeps. */
size_t makb200_sbr_apply_q2_worksize(makb200_handle_t* h, int dtype, int n, int b, int g, int ncols);
int makb200_sbr_apply_q2(makb200_handle_t* h, int dtype, int n, int b, int g, const void* V2, int ldv,
                         const void* tau2, int ldt, void* Z, int ldz, int ncols, void* work, size_t lwork);

#ifdef __cplusplus
}
#endif
#endif /* MAKB200_H */
