// Lock-step blocked Householder QR over MANY mid-size blocks (the 65-512 end of the block-sparse
// batch, BASELINE configs[2]).  Blocks that do not fit one CTA's shared memory are factorized
// together with TWO-LEVEL blocking common to all blocks: outer column blocks (J0, nbo = 64..128) cut
// into inner steps (j0, jb <= 32).  Every step is a handful of launches that each cover ALL blocks
// still active at that column:
//   inner step
//   1. bqr_panel_kernel      one CTA per block: the (m_i-j0) x jb panel is factorized in shared
//                            memory (non-negative-beta reflectors, compact-WY T accumulated on the
//                            fly), V\R written back, explicit V (into the outer block's V), T, tau
//   2. bqr_problems_kernel + three grouped DMMA GEMMs (W = V^H C, W2 = T^H W, C -= V W2) on the
//                            columns that remain INSIDE the outer block only (short, L2-resident)
//   outer step
//   3. G = V^H V (one grouped GEMM), bqr_tout_kernel: the nbo x nbo compact-WY T of the whole outer
//      block from G and tau (larft recurrence), then the three grouped GEMMs with K = nbo on the
//      trailing matrix -- so the trailing matrix is read and written once per nbo columns, not once
//      per 16-32 (the K = 16..32 updates of the one-level version were HBM-bound at 5 TF/s).
// Q is formed backwards over the OUTER blocks with the stored T factors (three K = nbo GEMMs each).
// Launch count is O(max_k / 32), independent of the number of blocks; no host round trip anywhere.
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
bqr_panel_kernel(const BqrBlock<T>* __restrict__ blocks, int j0, int jb, int J0) {
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

    // explicit V of this step inside the outer block's V (rows relative to J0, zero above the step)
    const int off = j0 - J0;
    T* Vw = b.Vw + (size_t)off * b.m;
    for (int c = warp; c < cols; c += NW) {
        T* vc = Vw + (size_t)c * b.m;
        for (int r = lane; r < off; r += 32) vc[r] = zero<T>();
        for (int r = lane; r < rows; r += 32) {
            const T v = slab[(size_t)c * lds + r];
            Ap[(size_t)c * b.lda + r] = v;
            vc[off + r] = (r < c) ? zero<T>() : (r == c ? one<T>() : v);
        }
    }
    for (int idx = tid; idx < BQR_NB * BQR_NB; idx += BP_THREADS) b.Tin[idx] = Tsm[idx];
    if (tid < cols) b.tau[j0 + tid] = Tsm[tid * BQR_NB + tid];
}

// compact-WY T (ne x ne, upper triangular, zero below) of one OUTER block from G = V^H V and tau:
//   T[0:j, j] = -tau_j T[0:j,0:j] G[0:j, j],  T[j,j] = tau_j        (larft, forward columnwise)
template <typename T>
__global__ void __launch_bounds__(BQR_NBO_MAX)
bqr_tout_kernel(const BqrBlock<T>* __restrict__ blocks, int J0, int nbo, int outer_idx) {
    const BqrBlock<T> b = blocks[blockIdx.x];
    const int ne = min(nbo, b.k - J0);
    if (ne <= 0) return;
    const T* G = b.G;
    T* To = b.Tout + (size_t)outer_idx * nbo * nbo;
    const T* tau = b.tau + J0;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < nbo * nbo; idx += BQR_NBO_MAX) To[idx] = zero<T>();
    __syncthreads();
    for (int j = 0; j < ne; ++j) {
        const T tj = tau[j];
        if (tid < j) {
            T s = zero<T>();
            for (int p = tid; p < j; ++p) fma_(s, To[(size_t)p * nbo + tid], G[(size_t)j * nbo + p]);
            To[(size_t)j * nbo + tid] = neg_(mul_(tj, s));
        } else if (tid == j) {
            To[(size_t)j * nbo + j] = tj;
        }
        __syncthreads();
    }
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

// Grouped-GEMM descriptors of one step, built on the device.
//   kind 0: inner step (j0, jb) of the outer block starting at J0: update of the columns that remain
//           inside the outer block, with the step's own V (a sub-block of Vw) and T (Tin)
//   kind 1: outer step, factorization: trailing columns beyond the outer block, V = Vw, T = Tout^H;
//           P4 = the Gram problem G = V^H V
//   kind 2: outer step, Q accumulation (H from the left): C = Q[J0:, J0:k], T = Tout
template <typename T>
__global__ void bqr_problems_kernel(const BqrBlock<T>* __restrict__ blocks, int count, int j0, int jb, int J0, int nbo,
                                    int outer_idx, int kind, GemmProblem<T>* __restrict__ P1,
                                    GemmProblem<T>* __restrict__ P2, GemmProblem<T>* __restrict__ P3,
                                    GemmProblem<T>* __restrict__ P4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const BqrBlock<T> b = blocks[i];
    int rows, cols, nc, ldc;
    T* C;
    const T* V;
    const T* Tm;
    int ldt;
    if (kind == 0) {
        rows = b.m - j0; cols = min(jb, b.k - j0);
        const int cend = min(b.k, J0 + nbo);   // reflector columns only: [J0+ne, n) belongs to the outer update
        nc = cend - j0 - cols; C = b.A + (size_t)(j0 + cols) * b.lda + j0; ldc = b.lda;
        V = b.Vw + (size_t)(j0 - J0) * b.m + (j0 - J0);
        Tm = b.Tin; ldt = BQR_NB;
    } else {
        rows = b.m - J0; cols = min(nbo, b.k - J0);
        V = b.Vw;
        Tm = b.Tout + (size_t)outer_idx * nbo * nbo; ldt = nbo;
        if (kind == 1) { nc = b.n - J0 - cols; C = b.A + (size_t)(J0 + cols) * b.lda + J0; ldc = b.lda; }
        else           { nc = b.k - J0;        C = b.Q + (size_t)J0 * b.ldq + J0;          ldc = b.ldq; }
    }
    const bool live = cols > 0 && rows > 0 && nc > 0;
    GemmProblem<T> p;
    p.lower = 0;
    // W = V^H C
    p.m = live ? cols : 0; p.n = nc; p.k = rows;
    p.A = V; p.lda = b.m; p.B = C; p.ldb = ldc; p.C = b.W; p.ldc = nbo;
    p.alpha = one<T>(); p.beta = zero<T>(); p.conja = 1; p.conjb = 0;
    P1[i] = p;
    // W2 = op(T) W
    p.k = cols;
    p.A = Tm; p.lda = ldt; p.B = b.W; p.ldb = nbo; p.C = b.W2; p.ldc = nbo;
    p.conja = (kind == 2) ? 0 : 1;
    P2[i] = p;
    // C -= V W2
    p.m = live ? rows : 0; p.k = cols;
    p.A = V; p.lda = b.m; p.B = b.W2; p.ldb = nbo; p.C = C; p.ldc = ldc;
    p.alpha = neg_(one<T>()); p.beta = one<T>(); p.conja = 0;
    P3[i] = p;
    if (kind == 1) {   // G = V^H V
        const bool lg = cols > 0 && rows > 0;
        p.m = lg ? cols : 0; p.n = cols; p.k = rows;
        p.A = V; p.lda = b.m; p.B = V; p.ldb = b.m; p.C = b.G; p.ldc = nbo;
        p.alpha = one<T>(); p.beta = zero<T>(); p.conja = 1; p.conjb = 0;
        P4[i] = p;
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <typename T>
bool bqr_fits(int m, int n) {
    // the narrowest panel (8 columns) of the tallest step must fit one CTA's shared memory
    return m > 0 && n > 0 && bqr_panel_smem_bytes<T>(m, 8) <= BP_SMEM_BYTES;
}

int bqr_outer_width() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MAKB200_BQR_NBO");
        v = e ? atoi(e) : 64;
        v = (v + 31) / 32 * 32;
        if (v < 32) v = 32;
        if (v > BQR_NBO_MAX) v = BQR_NBO_MAX;
    }
    return v;
}

// two-level column schedule common to all blocks; `ms`/`ns`/`ks` sorted by k descending
template <typename T>
BqrSchedule bqr_schedule(const std::vector<int>& ms, const std::vector<int>& ns, const std::vector<int>& ks) {
    BqrSchedule sc;
    sc.nbo = bqr_outer_width();
    const int count = (int)ks.size();
    if (count == 0) return sc;
    const int kmax = ks[0];
    int active = count;
    for (int J0 = 0; J0 < kmax; J0 += sc.nbo) {
        while (active > 0 && ks[active - 1] <= J0) --active;
        BqrOuter o{};
        o.J0 = J0; o.active = active; o.s_begin = (int)sc.steps.size();
        for (int i = 0; i < active; ++i) {
            const int ne = std::min(sc.nbo, ks[i] - J0);
            o.max_rows = std::max(o.max_rows, ms[i] - J0);
            o.max_ne = std::max(o.max_ne, ne);
            o.max_nc = std::max(o.max_nc, ns[i] - J0 - ne);
            o.max_ncq = std::max(o.max_ncq, ks[i] - J0);
        }
        const int jend = std::min(J0 + sc.nbo, kmax);
        int ia = active;
        for (int j0 = J0; j0 < jend;) {
            while (ia > 0 && ks[ia - 1] <= j0) --ia;
            BqrStep st{};
            st.j0 = j0; st.J0 = J0; st.active = ia;
            for (int i = 0; i < ia; ++i) st.max_rows = std::max(st.max_rows, ms[i] - j0);
            int jb = BQR_NB;
            while (jb > 8 && bqr_panel_smem_bytes<T>(st.max_rows, jb) > BP_SMEM_BYTES) jb /= 2;
            jb = std::min(jb, jend - j0);
            st.jb = jb;
            for (int i = 0; i < ia; ++i)
                st.max_nc = std::max(st.max_nc, std::min(ks[i], J0 + sc.nbo) - j0 - std::min(jb, ks[i] - j0));
            sc.steps.push_back(st);
            j0 += jb;
        }
        o.s_end = (int)sc.steps.size();
        sc.outer.push_back(o);
    }
    return sc;
}

static inline size_t bqr_up(size_t e) { return (e + 15) / 16 * 16; }

// outer steps a block with k reflectors takes part in
static inline int bqr_nouter(const BqrSchedule& sc, int k) { return k <= 0 ? 0 : (k + sc.nbo - 1) / sc.nbo; }

template <typename T>
size_t bqr_block_work_elems(const BqrSchedule& sc, int m, int n, int k) {
    const size_t wc = (size_t)std::max(n, k), nbo = (size_t)sc.nbo;
    return bqr_up((size_t)m * nbo) + 2 * bqr_up(nbo * wc) + bqr_up((size_t)BQR_NB * BQR_NB) + bqr_up(nbo * nbo) +
           bqr_up(nbo * nbo * (size_t)bqr_nouter(sc, k)) + bqr_up((size_t)std::max(k, 1));
}

// carve the per-block scratch out of `p` (advanced); the layout matches bqr_block_work_elems
template <typename T>
void bqr_carve_block(const BqrSchedule& sc, BqrBlock<T>& b, T*& p) {
    const size_t wc = (size_t)std::max(b.n, b.k), nbo = (size_t)sc.nbo;
    b.Vw = p;   p += bqr_up((size_t)b.m * nbo);
    b.W = p;    p += bqr_up(nbo * wc);
    b.W2 = p;   p += bqr_up(nbo * wc);
    b.Tin = p;  p += bqr_up((size_t)BQR_NB * BQR_NB);
    b.G = p;    p += bqr_up(nbo * nbo);
    b.Tout = p; p += bqr_up(nbo * nbo * (size_t)bqr_nouter(sc, b.k));
    b.tau = p;  p += bqr_up((size_t)std::max(b.k, 1));
}

template <typename T>
int batched_qr_blocked(makb200_handle* h, int count, const BqrBlock<T>* blocks_dev, const BqrSchedule& sc,
                       GemmProblem<T>* probs_dev) {
    if (count <= 0 || sc.outer.empty()) return 0;
    cudaStream_t s = h->stream;
    const int nbo = sc.nbo;
    GemmProblem<T>*P1 = probs_dev, *P2 = probs_dev + count, *P3 = probs_dev + 2 * (size_t)count,
                  *P4 = probs_dev + 3 * (size_t)count;
    constexpr int ZMAX = 32768;   // gridDim.z limit of the grouped launch
    // W = V^H C (cols x nc, K = rows), W2 = op(T) W, C -= V W2 (rows x nc, K = cols)
    auto grouped3 = [&](bool tconj, int active, int max_cols, int max_rows, int max_nc) -> int {
        for (int z0 = 0; z0 < active; z0 += ZMAX) {
            const int zc = std::min(ZMAX, active - z0);
            cudaError_t e = gemm_grouped<T>(s, MAKB200_OP_C, MAKB200_OP_N, zc, max_cols, max_nc, P1 + z0);
            if (e == cudaSuccess)
                e = gemm_grouped<T>(s, tconj ? MAKB200_OP_C : MAKB200_OP_N, MAKB200_OP_N, zc, max_cols, max_nc, P2 + z0);
            if (e == cudaSuccess) e = gemm_grouped<T>(s, MAKB200_OP_N, MAKB200_OP_N, zc, max_rows, max_nc, P3 + z0);
            if (e != cudaSuccess) return cuda_fail(h, e, "gemm_grouped");
        }
        return 0;
    };
    // ---- factorization ----
    for (size_t oi = 0; oi < sc.outer.size(); ++oi) {
        const BqrOuter& o = sc.outer[oi];
        for (int si = o.s_begin; si < o.s_end; ++si) {
            const BqrStep& st = sc.steps[si];
            if (st.active <= 0) continue;
            const size_t smem = bqr_panel_smem_bytes<T>(st.max_rows, st.jb);
            bqr_panel_kernel<T><<<st.active, BP_THREADS, smem, s>>>(blocks_dev, st.j0, st.jb, o.J0);
            count_launch();
            MAK_LAUNCH_CHECK(h, "bqr_panel_kernel");
            if (st.max_nc > 0) {
                bqr_problems_kernel<T><<<(st.active + 127) / 128, 128, 0, s>>>(blocks_dev, st.active, st.j0, st.jb, o.J0, nbo,
                                                                              (int)oi, 0, P1, P2, P3, P4);
                count_launch();
                MAK_LAUNCH_CHECK(h, "bqr_problems_kernel");
                int rc = grouped3(true, st.active, st.jb, st.max_rows, st.max_nc);
                if (rc) return rc;
            }
        }
        // T of the whole outer block (also needed by the Q phase), then the K = nbo trailing update
        bqr_problems_kernel<T><<<(o.active + 127) / 128, 128, 0, s>>>(blocks_dev, o.active, 0, 0, o.J0, nbo, (int)oi, 1, P1,
                                                                     P2, P3, P4);
        count_launch();
        MAK_LAUNCH_CHECK(h, "bqr_problems_kernel");
        for (int z0 = 0; z0 < o.active; z0 += ZMAX) {
            cudaError_t e = gemm_grouped<T>(s, MAKB200_OP_C, MAKB200_OP_N, std::min(ZMAX, o.active - z0), o.max_ne, o.max_ne,
                                            P4 + z0);
            if (e != cudaSuccess) return cuda_fail(h, e, "gemm_grouped");
        }
        bqr_tout_kernel<T><<<o.active, BQR_NBO_MAX, 0, s>>>(blocks_dev, o.J0, nbo, (int)oi);
        count_launch();
        MAK_LAUNCH_CHECK(h, "bqr_tout_kernel");
        if (o.max_nc > 0) {
            int rc = grouped3(true, o.active, o.max_ne, o.max_rows, o.max_nc);
            if (rc) return rc;
        }
    }
    // ---- R out, Q = I ----
    bqr_extract_kernel<T><<<count, 256, 0, s>>>(blocks_dev);
    count_launch();
    MAK_LAUNCH_CHECK(h, "bqr_extract_kernel");
    // ---- Q = H_1 ... H_k [I; 0], backwards over the outer blocks ----
    for (int oi = (int)sc.outer.size() - 1; oi >= 0; --oi) {
        const BqrOuter& o = sc.outer[oi];
        bqr_copy_v_kernel<T><<<o.active, 256, 0, s>>>(blocks_dev, o.J0, nbo);
        bqr_problems_kernel<T><<<(o.active + 127) / 128, 128, 0, s>>>(blocks_dev, o.active, 0, 0, o.J0, nbo, oi, 2, P1, P2,
                                                                     P3, P4);
        count_launch(2);
        MAK_LAUNCH_CHECK(h, "bqr_copy_v_kernel");
        int rc = grouped3(false, o.active, o.max_ne, o.max_rows, o.max_ncq);
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
template BqrSchedule bqr_schedule<double>(const std::vector<int>&, const std::vector<int>&, const std::vector<int>&);
template BqrSchedule bqr_schedule<cplx>(const std::vector<int>&, const std::vector<int>&, const std::vector<int>&);
template size_t bqr_block_work_elems<double>(const BqrSchedule&, int, int, int);
template size_t bqr_block_work_elems<cplx>(const BqrSchedule&, int, int, int);
template void bqr_carve_block<double>(const BqrSchedule&, BqrBlock<double>&, double*&);
template void bqr_carve_block<cplx>(const BqrSchedule&, BqrBlock<cplx>&, cplx*&);
template int batched_qr_blocked<double>(makb200_handle*, int, const BqrBlock<double>*, const BqrSchedule&,
                                        GemmProblem<double>*);
template int batched_qr_blocked<cplx>(makb200_handle*, int, const BqrBlock<cplx>*, const BqrSchedule&,
                                      GemmProblem<cplx>*);

}  // namespace mak
