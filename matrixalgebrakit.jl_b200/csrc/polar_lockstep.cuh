// Lock-step QDWH polar decomposition of many mid-size blocks: the batched kernels and the replay of the plan built by
// polar_lockstep_plan.h (see there for the algorithm).  Included by capi.cu, which supplies the batched QR.
#pragma once
#include "common.cuh"
#include "devutil.cuh"
#include "gemm.cuh"
#include "potf2.cuh"
#include "polar_lockstep_plan.h"

namespace mak {

template <typename T>
__device__ __forceinline__ T* ls_buf_dev(const LsBlk<T>& b, int id) {
    switch (id) {
        case LS_X: return b.X;
        case LS_B: return b.B;
        case LS_Q: return b.Q;
        case LS_T2: return b.T2;
        case LS_Z: return b.Z;
        case LS_L: return b.L;
        case LS_W: return b.W;
        default: return nullptr;
    }
}

// X0 = S / ||S||_F of every block (one CTA per block), entries pre-scaled by 1 / max|s_ij| so that the sum of squares
// neither underflows nor overflows (same rule as the single-matrix driver, polar.cu: absmax / fro2_pre / scale_copy_pre)
template <typename T>
__global__ void __launch_bounds__(1024) ls_prep_kernel(const LsBlk<T>* __restrict__ blks) {
    __shared__ double red[32];
    __shared__ double s_f;
    const LsBlk<T> b = blks[blockIdx.x];
    const int n = b.n, tid = threadIdx.x;
    const size_t total = (size_t)n * n;
    double mx = 0.0;
    for (size_t idx = tid; idx < total; idx += blockDim.x) {
        const int r = (int)(idx % n), c = (int)(idx / n);
        const T a = b.S[(size_t)c * b.lds + r];
        const double v = fmax(fabs(real_(a)), fabs(imag_(a)));
        mx = (v > mx || v != v) ? v : mx;     // NaN propagates
    }
    mx = warp_max(mx);
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t = fmax(t, red[i]);
        s_f = (t > 0.0 && isfinite(t)) ? 1.0 / t : 1.0;
    }
    __syncthreads();
    const double f = s_f;
    double s = 0.0;
    for (size_t idx = tid; idx < total; idx += blockDim.x) {
        const int r = (int)(idx % n), c = (int)(idx / n);
        s += abs2_(scale_(b.S[(size_t)c * b.lds + r], f));
    }
    const double nn = block_sum<double>(s, red);
    const double inv = nn > 0.0 ? 1.0 / sqrt(nn) : 1.0;
    for (size_t idx = tid; idx < total; idx += blockDim.x) {
        const int r = (int)(idx % n), c = (int)(idx / n);
        b.X[idx] = scale_(scale_(b.S[(size_t)c * b.lds + r], f), inv);
    }
}

// element-wise steps, blockIdx.y = block
template <typename T>
__global__ void ls_ew_kernel(const LsBlk<T>* __restrict__ blks, int kind, int a0, int a1, int a2, double p0, double p1) {
    const LsBlk<T> b = blks[blockIdx.y];
    const int n = b.n;
    const size_t start = blockIdx.x * (size_t)blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    switch (kind) {
        case LS_STACK: {   // B = [p0 X; I]  (2n x n)
            const int mb = 2 * n;
            const size_t total = (size_t)mb * n;
            for (size_t idx = start; idx < total; idx += step) {
                const int r = (int)(idx % mb), c = (int)(idx / mb);
                T v;
                if (r < n) v = scale_(b.X[(size_t)c * n + r], p0);
                else v = (r - n == c) ? one<T>() : zero<T>();
                b.B[idx] = v;
            }
            break;
        }
        case LS_ADDDIAG:
            for (size_t i = start; i < (size_t)n; i += step) b.Z[i * n + i] = add_(b.Z[i * n + i], one<T>());
            break;
        case LS_AXPBY: {   // X = p0 X + p1 B
            const size_t total = (size_t)n * n;
            for (size_t idx = start; idx < total; idx += step) b.X[idx] = add_(scale_(b.X[idx], p0), scale_(b.B[idx], p1));
            break;
        }
        case LS_COPY: {
            if (a0 == LS_A) {   // tall blocks only: B (m x n, ld m) = A
                if (b.m <= n) return;
                const size_t total = (size_t)b.m * n;
                for (size_t idx = start; idx < total; idx += step) {
                    const int r = (int)(idx % b.m), c = (int)(idx / b.m);
                    b.B[idx] = b.A[(size_t)c * b.lda + r];
                }
                return;
            }
            if (a1 == LS_W && b.m != n) return;   // a tall block's W is Q0 X (grouped GEMM)
            const T* __restrict__ src = ls_buf_dev(b, a0);
            T* __restrict__ dst = ls_buf_dev(b, a1);
            const size_t total = (size_t)a2 * n * n;
            for (size_t idx = start; idx < total; idx += step) dst[idx] = src[idx];
            break;
        }
        case LS_PROBE: {   // T2 (LS_NPROBE x n, ld LS_NPROBE) = Rademacher probes of the sigma_min estimate
            const size_t total = (size_t)LS_NPROBE * n;
            for (size_t idx = start; idx < total; idx += step)
                b.T2[idx] = mk<T>(ls_probe_entry((int)(idx % LS_NPROBE), (int)(idx / LS_NPROBE)));
            break;
        }
        case LS_SYMM: {   // P = (Z + Z^H) / 2, real diagonal
            const size_t total = (size_t)n * n;
            for (size_t idx = start; idx < total; idx += step) {
                const int r = (int)(idx % n), c = (int)(idx / n);
                T v = scale_(add_(b.Z[idx], conj_(b.Z[(size_t)r * n + c])), 0.5);
                if (r == c) v = mk<T>(real_(v));
                b.P[idx] = v;
            }
            break;
        }
        default: break;
    }
}

// est[i] = ||Q_i (LS_NPROBE x n)||_F^2  (the solved probes of the sigma_min estimate), one CTA per block
template <typename T>
__global__ void __launch_bounds__(256) ls_fro_kernel(const LsBlk<T>* __restrict__ blks, double* __restrict__ est) {
    __shared__ double red[32];
    const LsBlk<T> b = blks[blockIdx.x];
    const size_t total = (size_t)LS_NPROBE * b.n;
    double s = 0.0;
    for (size_t idx = threadIdx.x; idx < total; idx += blockDim.x) s += abs2_(b.Q[idx]);
    const double t = block_sum<double>(s, red);
    if (threadIdx.x == 0) est[blockIdx.x] = t;
}

// diagonal block (column j0, index bi) of every active block: L_jj and its inverse, one CTA per block
template <typename T, int NB>
__global__ void __launch_bounds__(256) ls_potf2_kernel(const LsBlk<T>* __restrict__ blks, int j0, int bi, int* info) {
    const LsBlk<T> b = blks[blockIdx.x];
    const int n = b.n, jb = (n - j0 < NB) ? (n - j0) : NB;
    if (jb <= 0) return;
    potf2_inv_body<T, NB>(jb, b.Z + (size_t)j0 * n + j0, n, b.L + (size_t)j0 * n + j0, n, b.Linv + (size_t)bi * NB * NB, NB,
                          info + blockIdx.x);
}

template <typename T>
inline int ls_init(makb200_handle* h) {
    constexpr int nb = CholNB<T>::value;
    static bool done = false;   // (the lock-step driver runs on the calling thread only)
    if (done) return 0;
    MAK_CUDA(h, cudaFuncSetAttribute(ls_potf2_kernel<T, nb>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(sizeof(T) * nb * (nb + 1))));
    done = true;
    return 0;
}

// bytes of the device-side tables of a chunk of `count` blocks with sizes up to nmax
template <typename T>
inline size_t ls_tables_bytes(int count, int nmax, bool any_tall) {
    constexpr int nb = CholNB<T>::value;
    const std::vector<QdwhStep> sched = qdwh_schedule(2.2e-16, 12, 100.0);
    const size_t launches = (size_t)ls_gemm_launch_bound<T>(nmax, nb, any_tall, sched);
    return align_up(sizeof(LsBlk<T>) * (size_t)count, 256) + align_up(sizeof(int) * (size_t)count, 256) +
           align_up(sizeof(double) * (size_t)count, 256) + align_up(sizeof(GemmProblem<T>) * launches * (size_t)count, 256) + 1024;
}

// sigma_min estimate per chunk (default on): MAKB200_LS_ESTIMATE=0 runs the fixed l0 = eps schedule
inline bool ls_estimate_enabled() {
    const char* e = getenv("MAKB200_LS_ESTIMATE");   // read per call: tests run both
    return !(e && e[0] == '0');
}

// Uploads one plan (blocks + descriptors) and replays it.  qr(batch, m, n, A, lda, Q, ldq, R, ldr): the batched QR of
// capi.cu (A overwritten, R may be null).
template <typename T, typename QRF>
int polar_lockstep_replay(makb200_handle* h, const std::vector<LsBlk<T>>& blk, const LsPlan<T>& pl, LsBlk<T>* bdev, int* info,
                          double* est, GemmProblem<T>* pdev, size_t pcap, QRF& qr) {
    constexpr int nb = CholNB<T>::value;
    const int count = (int)blk.size();
    cudaStream_t s = h->stream;
    if (pl.probs.size() > pcap) return MAKB200_ERR_WORKSPACE;
    {
        Stager st(h, sizeof(LsBlk<T>) * (size_t)count + sizeof(GemmProblem<T>) * pl.probs.size() + 1024);
        MAK_CUDA(h, st.put(bdev, blk.data(), sizeof(LsBlk<T>) * (size_t)count, s));
        MAK_CUDA(h, st.put(pdev, pl.probs.data(), sizeof(GemmProblem<T>) * pl.probs.size(), s));
    }
    MAK_CUDA(h, cudaMemsetAsync(info, 0, sizeof(int) * (size_t)count, s));
    auto ew = [&](const LsAct& a, int gx) {
        ls_ew_kernel<T><<<dim3(gx, a.count), 256, 0, s>>>(bdev, a.kind, a.a0, a.a1, a.a2, a.p0, a.p1);
        count_launch();
    };
    int rc = 0;
    for (const LsAct& a : pl.acts) {
        switch (a.kind) {
            case LS_GEMM: {
                cudaError_t e = gemm_grouped<T>(s, a.opa, a.opb, a.count, a.max_m, a.max_n, pdev + a.off);
                if (e != cudaSuccess) return cuda_fail(h, e, "lock-step grouped gemm");
                break;
            }
            case LS_PREP:
                ls_prep_kernel<T><<<a.count, 1024, 0, s>>>(bdev);
                count_launch();
                break;
            case LS_STACK: case LS_AXPBY: case LS_COPY: case LS_SYMM:
                ew(a, 32);
                break;
            case LS_ADDDIAG: case LS_PROBE:
                ew(a, 1);
                break;
            case LS_FRO:
                ls_fro_kernel<T><<<a.count, 256, 0, s>>>(bdev, est);
                count_launch();
                break;
            case LS_POTF2:
                ls_potf2_kernel<T, nb><<<a.count, 256, sizeof(T) * nb * (nb + 1), s>>>(bdev, a.a0, a.a1, info);
                count_launch();
                break;
            case LS_QR_TALL: {
                LsAct c{};
                c.kind = LS_COPY; c.count = count; c.a0 = LS_A; c.a1 = LS_B; c.a2 = 1;
                ew(c, 32);
                std::vector<int> m, n, lda, ldq, ldr;
                std::vector<void*> A, Q, R;
                for (const auto& b : blk) {
                    if (b.m <= b.n) continue;
                    m.push_back(b.m); n.push_back(b.n);
                    A.push_back(b.B); lda.push_back(b.m);
                    Q.push_back(b.Q0); ldq.push_back(b.m);
                    R.push_back(b.R0); ldr.push_back(b.n);
                }
                rc = qr((int)m.size(), m.data(), n.data(), A.data(), lda.data(), Q.data(), ldq.data(), R.data(), ldr.data());
                if (rc) return rc;
                break;
            }
            case LS_QR_STACK: {
                std::vector<int> m, n, lda;
                std::vector<void*> A, Q;
                for (const auto& b : blk) {
                    m.push_back(2 * b.n); n.push_back(b.n);
                    A.push_back(b.B); lda.push_back(2 * b.n);
                    Q.push_back(b.Q);
                }
                rc = qr(count, m.data(), n.data(), A.data(), lda.data(), Q.data(), lda.data(), (void* const*)nullptr, (const int*)nullptr);
                if (rc) return rc;
                break;
            }
            default: break;
        }
    }
    MAK_LAUNCH_CHECK(h, "polar_lockstep_replay");
    return 0;
}

// The lock-step polar decomposition of a chunk.  `blk`: host copy of the blocks (n descending, buffers carved); tables:
// device region of ls_tables_bytes.
//   1. X0 and, unless MAKB200_LS_ESTIMATE=0, the sigma_min estimate of every block (one plan, ONE device-to-host read)
//   2. blocks with a usable estimate run the QDWH schedule of the SMALLEST l0 among them (a lower bound for each: the
//      schedule converges for all; for well-conditioned blocks it has no Householder step and one step fewer);
//      the others (Cholesky of X0^H X0 broke down, l0 <= 1e-7: rank-deficient or kappa > ~1e6) run the l0 = eps schedule
template <typename T, typename QRF>
int polar_lockstep_run(makb200_handle* h, const std::vector<LsBlk<T>>& blk, char* tables, size_t tables_bytes, QRF qr) {
    constexpr int nb = CholNB<T>::value;
    const int count = (int)blk.size();
    if (count == 0) return 0;
    int rc = ls_init<T>(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    size_t off = 0;
    LsBlk<T>* bdev = (LsBlk<T>*)(tables + off); off += align_up(sizeof(LsBlk<T>) * (size_t)count, 256);
    int* info = (int*)(tables + off); off += align_up(sizeof(int) * (size_t)count, 256);
    double* est = (double*)(tables + off); off += align_up(sizeof(double) * (size_t)count, 256);
    if (off + 1024 > tables_bytes) return MAKB200_ERR_WORKSPACE;
    GemmProblem<T>* pdev = (GemmProblem<T>*)(tables + off);
    const size_t pcap = (tables_bytes - off) / sizeof(GemmProblem<T>);
    const bool estimate = ls_estimate_enabled();
    {
        LsPlan<T> pl;
        LsPlanner<T> planner(blk, nb, pl);
        planner.build_prepare(estimate);
        rc = polar_lockstep_replay<T>(h, blk, pl, bdev, info, est, pdev, pcap, qr);
        if (rc) return rc;
    }
    std::vector<LsBlk<T>> fast, slow;
    double l0_fast = 0.9;
    if (estimate) {
        std::vector<double> hest(count);
        std::vector<int> hinfo(count);
        MAK_CUDA(h, cudaMemcpyAsync(hest.data(), est, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, s));
        MAK_CUDA(h, cudaMemcpyAsync(hinfo.data(), info, sizeof(int) * (size_t)count, cudaMemcpyDeviceToHost, s));
        MAK_CUDA(h, cudaStreamSynchronize(s));
        for (int i = 0; i < count; ++i) {
            const double l0 = ls_l0_from_estimate(hest[i], hinfo[i]);
            if (l0 > 1e-7) { fast.push_back(blk[i]); l0_fast = std::min(l0_fast, l0); }
            else slow.push_back(blk[i]);
        }
    } else {
        slow = blk;
    }
    if (getenv("MAKB200_LOCKSTEP_VERBOSE"))
        fprintf(stderr, "[makb200] lock-step QDWH: %zu blocks on the l0 = %.2e schedule, %zu on l0 = eps\n", fast.size(), l0_fast, slow.size());
    if (!fast.empty()) {
        LsPlan<T> pl;
        LsPlanner<T> planner(fast, nb, pl);
        planner.build_iterate(qdwh_schedule(l0_fast, 12, 100.0));
        rc = polar_lockstep_replay<T>(h, fast, pl, bdev, info, est, pdev, pcap, qr);
        if (rc) return rc;
    }
    if (!slow.empty()) {
        LsPlan<T> pl;
        LsPlanner<T> planner(slow, nb, pl);
        planner.build_iterate(qdwh_schedule(2.2e-16, 12, 100.0));
        rc = polar_lockstep_replay<T>(h, slow, pl, bdev, info, est, pdev, pcap, qr);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace mak
