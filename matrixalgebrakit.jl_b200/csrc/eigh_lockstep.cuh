// Lock-step eigensolve of MANY Hermitian blocks that were tridiagonalised by bhetrd_batched_t (one launch for all), and
// the lock-step tail of the batched SVD.  Before this the per-block chain (D&C + back-transformation + gauge, ~100
// launches per block, 0.2-0.6 ms per block on a 32-stream pool) was 97 % / 90 % / 70 % of the batched eigh time of
// the 65-128 / 129-256 / 257-512 buckets (profiles/r2_bhetrd_share.log); now a chunk of blocks costs
//   stedc_batched        ~ 12 launches per tree level for the whole chunk (stedc.cu)
//   ormqr_left_batched   7 launches per panel of 64 reflectors (qr.cu)
//   one gauge launch, and for the SVD: one reorder launch, one grouped GEMM U = W V, one defect launch + ONE
//   device-to-host read for the chunk (rank-deficient blocks are then repaired one by one, as before).
// Included by capi.cu.  Reference semantics: /root/reference/src/implementations/eigh.jl (eigh_full! per block),
// /root/reference/src/implementations/svd.jl:196-237.
#pragma once
#include "common.cuh"
#include "devutil.cuh"
#include "gemm.cuh"
#include "qr.cuh"
#include "stedc.cuh"
#include "gauge.cuh"
#include <algorithm>
#include <vector>

namespace mak {

template <typename T>
struct EighLsBlk {
    int n;
    T* A; int lda;        // reflectors below the sub-diagonal (bhetrd layout)
    double* d; double* e; T* tau;
    double* w;            // out: eigenvalues ascending
    T* V; int ldv;        // out: eigenvectors
};

template <typename T>
struct GaugeDesc { int m, ncols; T* V; int ldv; T* other; int ldo, other_n; };
template <typename T>
__global__ void ls_gauge_kernel(const GaugeDesc<T>* __restrict__ descs) {
    const GaugeDesc<T> d = descs[blockIdx.y];
    for (int j = blockIdx.x; j < d.ncols; j += gridDim.x) gauge_column_body<T>(d.m, j, d.V, d.ldv, d.other, d.ldo, d.other_n);
}

// ---- workspace: [tables | per-block ormqr buffers | stedc_batched workspace] ----
template <typename T>
struct EighLsLayout { size_t tables, blocks, stedc, total; };
template <typename T>
inline EighLsLayout<T> eigh_lockstep_layout(int count, const int* n) {
    EighLsLayout<T> L{};
    int nmax = 0;
    size_t be = 0;
    for (int i = 0; i < count; ++i) {
        nmax = std::max(nmax, n[i]);
        be += align_up(ormqr_batched_block_elems<T>(n[i] - 1, n[i]) * sizeof(T), 256);
    }
    L.tables = align_up(ormqr_batched_table_bytes<T>(count, nmax), 256) + align_up(sizeof(GaugeDesc<T>) * (size_t)count, 256);
    L.blocks = be;
    L.stedc = align_up(stedc_batched_worksize(count, n), 256);
    L.total = L.tables + L.blocks + L.stedc + 1024;
    return L;
}
// bound for `count` blocks no larger than nmax
template <typename T>
inline size_t eigh_lockstep_bytes(int count, int nmax) {
    std::vector<int> n((size_t)count, nmax);
    return eigh_lockstep_layout<T>(count, n.data()).total;
}

// blks: HOST array sorted by n descending, n >= 3.  fixgauge: gauge the eigenvector columns (eigh); the SVD tail
// gauges U / Vh itself.
template <typename T>
int eigh_lockstep_run(makb200_handle* h, int count, const EighLsBlk<T>* blks, int fixgauge, char* work, size_t lwork) {
    if (count <= 0) return 0;
    cudaStream_t s = h->stream;
    std::vector<int> n(count);
    for (int i = 0; i < count; ++i) n[i] = blks[i].n;
    const EighLsLayout<T> L = eigh_lockstep_layout<T>(count, n.data());
    if (L.total > lwork) return MAKB200_ERR_WORKSPACE;
    char* base = (char*)align_up((size_t)(uintptr_t)work, 256);
    char* tables = base;
    GaugeDesc<T>* gdev = (GaugeDesc<T>*)(tables + align_up(ormqr_batched_table_bytes<T>(count, blks[0].n), 256));
    char* bbase = tables + L.tables;
    char* swork = bbase + L.blocks;
    // 1. tridiagonal D&C of every block: w_i, V_i = Z_i
    std::vector<StedcBlk> sb(count);
    for (int i = 0; i < count; ++i) sb[i] = StedcBlk{blks[i].n, blks[i].d, blks[i].e, blks[i].w, (void*)blks[i].V, blks[i].ldv};
    int rc = stedc_batched<T>(h, count, sb.data(), swork, L.stedc, nullptr);
    if (rc) return rc;
    // 2. V_i[1:, :] <- H_0 ... H_{n-2} V_i[1:, :]   (reflectors = QR-type columns of A_i[1:, 0:n-1])
    std::vector<OrmqrBatchBlk<T>> ob(count);
    T* p = (T*)bbase;
    for (int i = 0; i < count; ++i) {
        OrmqrBatchBlk<T>& o = ob[i];
        o.m = blks[i].n - 1; o.k = blks[i].n - 1; o.nc = blks[i].n;
        o.A = blks[i].A + 1; o.lda = blks[i].lda; o.tau = blks[i].tau;
        o.C = blks[i].V + 1; o.ldc = blks[i].ldv;
        ormqr_batched_carve<T>(o, p);
        p = (T*)align_up((size_t)(uintptr_t)p, 256);
    }
    if ((char*)p > swork) return MAKB200_ERR_WORKSPACE;
    rc = ormqr_left_batched<T>(h, count, ob.data(), tables, align_up(ormqr_batched_table_bytes<T>(count, blks[0].n), 256));
    if (rc) return rc;
    if (fixgauge) {
        std::vector<GaugeDesc<T>> gd(count);
        int nmax = 0;
        for (int i = 0; i < count; ++i) {
            gd[i] = GaugeDesc<T>{blks[i].n, blks[i].n, blks[i].V, blks[i].ldv, nullptr, 0, 0};
            nmax = std::max(nmax, blks[i].n);
        }
        {
            Stager st(h, sizeof(GaugeDesc<T>) * (size_t)count + 512);
            MAK_CUDA(h, st.put(gdev, gd.data(), sizeof(GaugeDesc<T>) * (size_t)count, s));
        }
        ls_gauge_kernel<T><<<dim3(nmax, count), 128, 0, s>>>(gdev);
        count_launch();
        MAK_LAUNCH_CHECK(h, "ls_gauge_kernel");
    }
    return 0;
}

// ---- SVD tail ---------------------------------------------------------------------------------------------
template <typename T>
struct SvdTailBlk {
    int m, n;
    const T* Wp;          // m x n polar isometry, ld m
    const double* w;      // eigenvalues of P ascending
    const T* V; int ldv;  // eigenvectors of P
    double* S;            // out: n, descending
    T* U; int ldu;        // out: m x n
    T* Vh; int ldvh;      // out: n x n
    double* flag;         // out: max_j | ||U(:, j)||^2 - 1 |
};
// S[j] = max(w[n-1-j], 0), Vh[j, :] = conj(V[:, n-1-j]); flag = 0
template <typename T>
__global__ void ls_svd_reorder_kernel(const SvdTailBlk<T>* __restrict__ blks) {
    const SvdTailBlk<T> b = blks[blockIdx.y];
    const int n = b.n;
    const size_t start = blockIdx.x * (size_t)blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    if (start == 0) b.flag[0] = 0.0;
    for (size_t j = start; j < (size_t)n; j += step) b.S[j] = fmax(b.w[n - 1 - j], 0.0);
    const size_t total = (size_t)n * n;
    for (size_t idx = start; idx < total; idx += step) {
        const int i = (int)(idx % n), j = (int)(idx / n);       // read V[i, n-1-j] (coalesced), write Vh[j, i]
        b.Vh[(size_t)i * b.ldvh + j] = conj_(b.V[(size_t)(n - 1 - j) * b.ldv + i]);
    }
}
template <typename T>
__global__ void ls_svd_defect_kernel(const SvdTailBlk<T>* __restrict__ blks) {
    const SvdTailBlk<T> b = blks[blockIdx.y];
    const int lane = threadIdx.x & 31;
    for (int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < b.n; j += gridDim.x * (blockDim.x >> 5)) {
        const T* u = b.U + (size_t)j * b.ldu;
        double s = 0.0;
        for (int r = lane; r < b.m; r += 32) s += abs2_(u[r]);
        s = warp_sum(s);
        if (lane == 0) {
            double dft = fabs(s - 1.0);
            if (!(dft == dft)) dft = 1e300;   // NaN counts as a defect
            atomicMax((unsigned long long*)b.flag, (unsigned long long)__double_as_longlong(dft));
        }
    }
}
template <typename T>
inline size_t svd_tail_lockstep_bytes(int count) {
    return align_up(sizeof(SvdTailBlk<T>) * (size_t)count, 256) + align_up(sizeof(GemmProblem<T>) * (size_t)count, 256) +
           align_up(sizeof(GaugeDesc<T>) * (size_t)count, 256) + align_up(sizeof(double) * (size_t)count, 256) + 1024;
}
// blks: HOST array.  repair(i): called (after a stream sync) for every block whose U is not isometric (rank-deficient
// input): must replace U_i by an orthonormal completion, as the single-matrix svd_tail does.
template <typename T, typename REPAIR>
int svd_tail_lockstep_run(makb200_handle* h, int count, const SvdTailBlk<T>* blks, int fixgauge, char* work, size_t lwork,
                          REPAIR repair) {
    if (count <= 0) return 0;
    if (svd_tail_lockstep_bytes<T>(count) > lwork) return MAKB200_ERR_WORKSPACE;
    cudaStream_t s = h->stream;
    char* base = (char*)align_up((size_t)(uintptr_t)work, 256);
    SvdTailBlk<T>* bdev = (SvdTailBlk<T>*)base;
    GemmProblem<T>* pdev = (GemmProblem<T>*)(base + align_up(sizeof(SvdTailBlk<T>) * (size_t)count, 256));
    GaugeDesc<T>* gdev = (GaugeDesc<T>*)((char*)pdev + align_up(sizeof(GemmProblem<T>) * (size_t)count, 256));
    double* flags = (double*)((char*)gdev + align_up(sizeof(GaugeDesc<T>) * (size_t)count, 256));
    std::vector<SvdTailBlk<T>> tb(blks, blks + count);
    std::vector<GemmProblem<T>> gp(count);
    std::vector<GaugeDesc<T>> gd(count);
    int mmax = 0, nmax = 0;
    for (int i = 0; i < count; ++i) {
        tb[i].flag = flags + i;
        GemmProblem<T>& p = gp[i];
        p.m = blks[i].m; p.n = blks[i].n; p.k = blks[i].n;
        p.A = blks[i].Wp; p.lda = blks[i].m; p.B = blks[i].Vh; p.ldb = blks[i].ldvh; p.C = blks[i].U; p.ldc = blks[i].ldu;
        p.alpha = one<T>(); p.beta = zero<T>(); p.conja = 0; p.conjb = 1; p.lower = 0;
        gd[i] = GaugeDesc<T>{blks[i].m, blks[i].n, blks[i].U, blks[i].ldu, blks[i].Vh, blks[i].ldvh, blks[i].n};
        mmax = std::max(mmax, blks[i].m); nmax = std::max(nmax, blks[i].n);
    }
    {
        Stager st(h, (sizeof(SvdTailBlk<T>) + sizeof(GemmProblem<T>) + sizeof(GaugeDesc<T>)) * (size_t)count + 2048);
        MAK_CUDA(h, st.put(bdev, tb.data(), sizeof(SvdTailBlk<T>) * (size_t)count, s));
        MAK_CUDA(h, st.put(pdev, gp.data(), sizeof(GemmProblem<T>) * (size_t)count, s));
        MAK_CUDA(h, st.put(gdev, gd.data(), sizeof(GaugeDesc<T>) * (size_t)count, s));
    }
    ls_svd_reorder_kernel<T><<<dim3(32, count), 256, 0, s>>>(bdev);
    count_launch();
    MAK_LAUNCH_CHECK(h, "ls_svd_reorder_kernel");
    // U = W V_desc = W Vh^H
    cudaError_t e = gemm_grouped<T>(s, MAKB200_OP_N, MAKB200_OP_C, count, mmax, nmax, pdev);
    if (e != cudaSuccess) return cuda_fail(h, e, "svd tail: grouped gemm");
    ls_svd_defect_kernel<T><<<dim3(std::max(1, (nmax + 7) / 8), count), 256, 0, s>>>(bdev);
    count_launch();
    MAK_LAUNCH_CHECK(h, "ls_svd_defect_kernel");
    std::vector<double> hf(count);
    MAK_CUDA(h, cudaMemcpyAsync(hf.data(), flags, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, s));
    MAK_CUDA(h, cudaStreamSynchronize(s));
    for (int i = 0; i < count; ++i)
        if (hf[i] > 1e-6) {
            int rc = repair(i);
            if (rc) return rc;
        }
    if (fixgauge) {
        ls_gauge_kernel<T><<<dim3(nmax, count), 128, 0, s>>>(gdev);
        count_launch();
        MAK_LAUNCH_CHECK(h, "ls_gauge_kernel");
    }
    return 0;
}

}  // namespace mak
