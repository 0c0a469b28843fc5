// Lock-step blocked Householder QR over MANY mid-size blocks (the 65-512 end of the block-sparse
// batch, BASELINE configs[2]).  Blocks that do not fit one CTA's shared memory are factorized
// together: the column range is cut into steps (j0, jb <= 32) common to all blocks, and every step
// is a handful of launches that each cover ALL blocks still active at that column:
//   1. bqr_panel_kernel      one CTA per block: the (m_i-j0) x jb panel is factorized in shared
//                            memory (non-negative-beta reflectors, compact-WY T accumulated on the
//                            fly), V\R written back, explicit V and T written for the GEMMs
//   2. bqr_problems_kernel   builds the three grouped-GEMM descriptor arrays on the device
//   3. three grouped DMMA GEMMs   W = V^H C,  W2 = T^H W,  C -= V W2   (gemm.cu, one launch each)
// Q is formed the same way backwards (orgqr with the stored T factors).  Launch count is
// O(max_k / 32), independent of the number of blocks; no host round trip anywhere.
// Replaces the per-block loop a TensorKit-style caller runs over qr_compact! (SURVEY.md §8b
// "What calls it"; the reference has only commented-out batched stubs, yacusolver.jl:506-649).
#include "batched.cuh"
#include "gemm.cuh"
#include <vector>
#include <algorithm>

namespace mak {

constexpr int BP_THREADS = 256;
constexpr size_t BP_SMEM_BYTES = 200 * 1024;

template <typename T>
__host__ __device__ inline size_t bqr_panel_smem_bytes(int rows, int jb) {
    return ((size_t)(rows | 1) * jb + 3 * BQR_NB + BQR_NB * BQR_NB + 4) * sizeof(T);
}

// ---------------------------------------------------------------------------------------
// panel factorization, one CTA per block
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(BP_THREADS)
bqr_panel_kernel(const BqrBlock<T>* __restrict__ blocks, int j0, int jb, int step) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const BqrBlock<T> b = blocks[blockIdx.x];
    const int rows = b.m - j0;
    const int cols = min(jb, b.k - j0);
    if (cols <= 0 || rows <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = BP_THREADS / 32;
    const int lds = rows | 1;
    T* slab = reinterpret_cast<T*>(smem_raw);   // [cols][lds]
    T* dots = slab + (size_t)lds * jb;          // [NB]
    T* coef = dots + BQR_NB;                    // [NB]
    T* zbuf = coef + BQR_NB;                    // [NB]
    T* Tsm = zbuf + BQR_NB;                     // [NB][NB] column-major
    T* scal = Tsm + BQR_NB * BQR_NB;            // scale, tau, beta

    T* Ap = b.A + (size_t)j0 * b.lda + j0;
    for (int c = warp; c < cols; c += NW)
        for (int r = lane; r < rows; r += 32) slab[(size_t)c * lds + r] = Ap[(size_t)c * b.lda + r];
    for (int idx = tid; idx < BQR_NB * BQR_NB; idx += BP_THREADS) Tsm[idx] = zero<T>();
    __syncthreads();

    for (int j = 0; j < cols; ++j) {
        const T* cj = slab + (size_t)j * lds;
        // dot products over rows > j: l >= j: conj(a_j) a_l ; l < j: conj(v_l) a_j
        for (int l = warp; l < cols; l += NW) {
            const T* cl = slab + (size_t)l * lds;
            T s = zero<T>();
            if (l >= j) { for (int r = j + 1 + lane; r < rows; r += 32) fmac_(s, cj[r], cl[r]); }
            else        { for (int r = j + 1 + lane; r < rows; r += 32) fmac_(s, cl[r], cj[r]); }
            s = warp_sum(s);
            if (lane == 0) dots[l] = s;
        }
        __syncthreads();
        if (tid < cols) {
            const int l = tid;
            const T topj = slab[(size_t)j * lds + j], topl = slab[(size_t)l * lds + j];
            double beta; T tau, scale;
            larfgp_scalars<T>(topj, real_(dots[j]), beta, tau, scale);
            if (l > j) coef[l] = mul_(conj_(tau), add_(topl, mul_(conj_(scale), dots[l])));
            else if (l < j) zbuf[l] = add_(conj_(topl), mul_(scale, dots[l]));   // v_l^H v_j
            else { scal[0] = scale; scal[1] = tau; scal[2] = mk<T>(beta); }
        }
        __syncthreads();
        const T scale = scal[0], tau = scal[1];
        if (tid < j) {   // T[0:j, j] = -tau T[0:j,0:j] z
            T s = zero<T>();
            for (int p = tid; p < j; ++p) fma_(s, Tsm[p * BQR_NB + tid], zbuf[p]);
            Tsm[j * BQR_NB + tid] = neg_(mul_(tau, s));
        } else if (tid == j) {
            Tsm[j * BQR_NB + j] = tau;
        }
        T* cjw = slab + (size_t)j * lds;
        for (int r = j + 1 + tid; r < rows; r += BP_THREADS) {
            const T v = mul_(cjw[r], scale);
            cjw[r] = v;
            for (int l = j + 1; l < cols; ++l) {
                T* p = slab + (size_t)l * lds + r;
                *p = sub_(*p, mul_(coef[l], v));
            }
        }
        if (tid < cols) {
            if (tid > j) {
                T* p = slab + (size_t)tid * lds + j;
                *p = sub_(*p, coef[tid]);
            } else if (tid == j) {
                slab[(size_t)j * lds + j] = scal[2];
            }
        }
        __syncthreads();
    }

    T* Vw = b.Vw;
    for (int c = warp; c < cols; c += NW) {
        for (int r = lane; r < rows; r += 32) {
            const T v = slab[(size_t)c * lds + r];
            Ap[(size_t)c * b.lda + r] = v;
            Vw[(size_t)c * b.m + r] = (r < c) ? zero<T>() : (r == c ? one<T>() : v);
        }
    }
    T* Tf = b.Tf + (size_t)step * BQR_NB * BQR_NB;
    for (int idx = tid; idx < BQR_NB * BQR_NB; idx += BP_THREADS) Tf[idx] = Tsm[idx];
}

// explicit V of one step from the factored A (orgqr phase)
template <typename T>
__global__ void __launch_bounds__(256)
bqr_copy_v_kernel(const BqrBlock<T>* __restrict__ blocks, int j0, int jb) {
    const BqrBlock<T> b = blocks[blockIdx.x];
    const int rows = b.m - j0, cols = min(jb, b.k - j0);
    if (cols <= 0 || rows <= 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T* Ap = b.A + (size_t)j0 * b.lda + j0;
    for (int c = warp; c < cols; c += 8)
        for (int r = lane; r < rows; r += 32) {
            T v = (r < c) ? zero<T>() : (r == c ? one<T>() : Ap[(size_t)c * b.lda + r]);
            b.Vw[(size_t)c * b.m + r] = v;
        }
}

// R = triu(A[0:k, 0:n]) and Q = [I; 0]  (one CTA per block)
template <typename T>
__global__ void __launch_bounds__(256)
bqr_extract_kernel(const BqrBlock<T>* __restrict__ blocks) {
    const BqrBlock<T> b = blocks[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (b.R) {
        for (int c = warp; c < b.n; c += 8)
            for (int r = lane; r < b.k; r += 32)
                b.R[(size_t)c * b.ldr + r] = (r <= c) ? b.A[(size_t)c * b.lda + r] : zero<T>();
    }
    for (int c = warp; c < b.k; c += 8)
        for (int r = lane; r < b.m; r += 32) b.Q[(size_t)c * b.ldq + r] = (r == c) ? one<T>() : zero<T>();
}

// phase 0: trailing update of A (H^H from the left);  phase 1: Q accumulation (H from the left)
template <typename T>
__global__ void bqr_problems_kernel(const BqrBlock<T>* __restrict__ blocks, int count, int j0, int jb, int step,
                                    int phase, GemmProblem<T>* __restrict__ P1, GemmProblem<T>* __restrict__ P2,
                                    GemmProblem<T>* __restrict__ P3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const BqrBlock<T> b = blocks[i];
    const int rows = b.m - j0, cols = min(jb, b.k - j0);
    int nc;
    T* C;
    int ldc;
    if (phase == 0) { nc = b.n - j0 - cols; C = b.A + (size_t)(j0 + cols) * b.lda + j0; ldc = b.lda; }
    else            { nc = b.k - j0;        C = b.Q + (size_t)j0 * b.ldq + j0;          ldc = b.ldq; }
    const bool live = cols > 0 && rows > 0 && nc > 0;
    GemmProblem<T> p;
    p.lower = 0;
    // W = V^H C
    p.m = live ? cols : 0; p.n = nc; p.k = rows;
    p.A = b.Vw; p.lda = b.m; p.B = C; p.ldb = ldc; p.C = b.W; p.ldc = BQR_NB;
    p.alpha = one<T>(); p.beta = zero<T>(); p.conja = 1; p.conjb = 0;
    P1[i] = p;
    // W2 = op(T) W
    p.k = cols;
    p.A = b.Tf + (size_t)step * BQR_NB * BQR_NB; p.lda = BQR_NB; p.B = b.W; p.ldb = BQR_NB; p.C = b.W2; p.ldc = BQR_NB;
    p.conja = (phase == 0) ? 1 : 0;
    P2[i] = p;
    // C -= V W2
    p.m = live ? rows : 0; p.k = cols;
    p.A = b.Vw; p.lda = b.m; p.B = b.W2; p.ldb = BQR_NB; p.C = C; p.ldc = ldc;
    p.alpha = neg_(one<T>()); p.beta = one<T>(); p.conja = 0;
    P3[i] = p;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <typename T>
bool bqr_fits(int m, int n) {
    // the narrowest panel (8 columns) of the tallest step must fit one CTA's shared memory
    return m > 0 && n > 0 && bqr_panel_smem_bytes<T>(m, 8) <= BP_SMEM_BYTES;
}

// column steps common to all blocks; `ms`/`ks` sorted by k descending
template <typename T>
std::vector<BqrStep> bqr_steps(const std::vector<int>& ms, const std::vector<int>& ns, const std::vector<int>& ks) {
    std::vector<BqrStep> steps;
    const int count = (int)ks.size();
    if (count == 0) return steps;
    const int kmax = ks[0];
    int j0 = 0, active = count;
    while (j0 < kmax) {
        while (active > 0 && ks[active - 1] <= j0) --active;
        int max_rows = 0, max_nc = 0, max_ncq = 0;
        for (int i = 0; i < active; ++i) {
            max_rows = std::max(max_rows, ms[i] - j0);
            max_ncq = std::max(max_ncq, ks[i] - j0);
        }
        int jb = BQR_NB;
        while (jb > 8 && bqr_panel_smem_bytes<T>(max_rows, jb) > BP_SMEM_BYTES) jb /= 2;
        for (int i = 0; i < active; ++i) max_nc = std::max(max_nc, ns[i] - j0 - std::min(jb, ks[i] - j0));
        steps.push_back(BqrStep{j0, jb, active, max_rows, max_nc, max_ncq});
        j0 += jb;
    }
    return steps;
}

template <typename T>
size_t bqr_block_work_elems(int m, int n, int k, int nsteps) {
    const size_t wc = (size_t)std::max(n, k);
    auto up = [](size_t e) { return (e + 15) / 16 * 16; };
    return up((size_t)m * BQR_NB) + 2 * up((size_t)BQR_NB * wc) + up((size_t)BQR_NB * BQR_NB * nsteps);
}

template <typename T>
int batched_qr_blocked(makb200_handle* h, int count, const BqrBlock<T>* blocks_dev, const std::vector<BqrStep>& steps,
                       GemmProblem<T>* probs_dev) {
    if (count <= 0 || steps.empty()) return 0;
    cudaStream_t s = h->stream;
    GemmProblem<T>*P1 = probs_dev, *P2 = probs_dev + count, *P3 = probs_dev + 2 * (size_t)count;
    constexpr int ZMAX = 32768;   // gridDim.z limit of the grouped launch
    auto grouped3 = [&](int phase, const BqrStep& st, int max_nc) -> int {
        for (int z0 = 0; z0 < st.active; z0 += ZMAX) {
            const int zc = std::min(ZMAX, st.active - z0);
            cudaError_t e = gemm_grouped<T>(s, MAKB200_OP_C, MAKB200_OP_N, zc, st.jb, max_nc, P1 + z0);
            if (e == cudaSuccess)
                e = gemm_grouped<T>(s, phase == 0 ? MAKB200_OP_C : MAKB200_OP_N, MAKB200_OP_N, zc, st.jb, max_nc, P2 + z0);
            if (e == cudaSuccess) e = gemm_grouped<T>(s, MAKB200_OP_N, MAKB200_OP_N, zc, st.max_rows, max_nc, P3 + z0);
            if (e != cudaSuccess) return cuda_fail(h, e, "gemm_grouped");
        }
        return 0;
    };
    // ---- factorization ----
    for (size_t si = 0; si < steps.size(); ++si) {
        const BqrStep& st = steps[si];
        const size_t smem = bqr_panel_smem_bytes<T>(st.max_rows, st.jb);
        bqr_panel_kernel<T><<<st.active, BP_THREADS, smem, s>>>(blocks_dev, st.j0, st.jb, (int)si);
        count_launch();
        MAK_LAUNCH_CHECK(h, "bqr_panel_kernel");
        if (st.max_nc > 0) {
            bqr_problems_kernel<T><<<(st.active + 127) / 128, 128, 0, s>>>(blocks_dev, st.active, st.j0, st.jb, (int)si, 0,
                                                                          P1, P2, P3);
            count_launch();
            MAK_LAUNCH_CHECK(h, "bqr_problems_kernel");
            int rc = grouped3(0, st, st.max_nc);
            if (rc) return rc;
        }
    }
    // ---- R out, Q = I ----
    bqr_extract_kernel<T><<<count, 256, 0, s>>>(blocks_dev);
    count_launch();
    MAK_LAUNCH_CHECK(h, "bqr_extract_kernel");
    // ---- Q = H_1 ... H_k [I; 0], backwards over the steps ----
    for (int si = (int)steps.size() - 1; si >= 0; --si) {
        const BqrStep& st = steps[si];
        bqr_copy_v_kernel<T><<<st.active, 256, 0, s>>>(blocks_dev, st.j0, st.jb);
        bqr_problems_kernel<T><<<(st.active + 127) / 128, 128, 0, s>>>(blocks_dev, st.active, st.j0, st.jb, si, 1, P1, P2,
                                                                      P3);
        count_launch(2);
        MAK_LAUNCH_CHECK(h, "bqr_copy_v_kernel");
        int rc = grouped3(1, st, st.max_ncq);
        if (rc) return rc;
    }
    return 0;
}

int batched_blocked_init(makb200_handle* h) {
    MAK_CUDA(h, cudaFuncSetAttribute(bqr_panel_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BP_SMEM_BYTES));
    MAK_CUDA(h, cudaFuncSetAttribute(bqr_panel_kernel<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BP_SMEM_BYTES));
    return 0;
}

template bool bqr_fits<double>(int, int);
template bool bqr_fits<cplx>(int, int);
template std::vector<BqrStep> bqr_steps<double>(const std::vector<int>&, const std::vector<int>&, const std::vector<int>&);
template std::vector<BqrStep> bqr_steps<cplx>(const std::vector<int>&, const std::vector<int>&, const std::vector<int>&);
template size_t bqr_block_work_elems<double>(int, int, int, int);
template size_t bqr_block_work_elems<cplx>(int, int, int, int);
template int batched_qr_blocked<double>(makb200_handle*, int, const BqrBlock<double>*, const std::vector<BqrStep>&,
                                        GemmProblem<double>*);
template int batched_qr_blocked<cplx>(makb200_handle*, int, const BqrBlock<cplx>*, const std::vector<BqrStep>&,
                                      GemmProblem<cplx>*);

}  // namespace mak
