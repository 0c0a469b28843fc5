// Blocked Householder QR for sm_100a (geqrf/geqrt + orgqr/gemqrt equivalents).
//
//  * panel_kernel: one thread-block CLUSTER factorizes an (mp x ib) panel that stays resident in
//    the shared memory of the cluster's CTAs (row slabs).  Per column there is exactly one
//    cluster-wide reduction (partial dot products exchanged through distributed shared memory)
//    and one cluster barrier; the compact-WY T factor is accumulated on the fly.
//    Reflector convention: beta = +||x|| (src/common/householder.jl:35-67 of the reference), so
//    diag(R) >= 0 and the reference's QR gauge (common/gauge.jl:16-25) needs no extra pass.
//  * everything GEMM-shaped (compact-WY trailing updates, T-factor coupling, Q formation) goes
//    through the DMMA GEMM in gemm.cu.
#include <cooperative_groups.h>
#include <type_traits>
#include <vector>
#include <algorithm>
#include "qr.cuh"
#include "projections.cuh"
#include "gemm.cuh"

namespace cg = cooperative_groups;

namespace mak {

constexpr int QR_NB_MAX = 128;    // outer block width (K of the trailing-update GEMMs)
static int qr_nb() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MAKB200_QR_NB");
        v = e ? atoi(e) : QR_NB_MAX;
        if (v != 32 && v != 64 && v != 128) v = QR_NB_MAX;
    }
    return v;
}
constexpr int IBMAX = 32;         // inner panel width (max)
constexpr int PANEL_THREADS = 512;
constexpr size_t PANEL_SLAB_BYTES = 160 * 1024;  // shared-memory budget for the row slab

// ---------------------------------------------------------------------------------------
// cluster panel factorization
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(PANEL_THREADS, 1)
panel_kernel(int mp, int ib, T* __restrict__ A, int lda, T* __restrict__ tau_out, T* __restrict__ Tout,
             int ldt, int rpc) {
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = PANEL_THREADS / 32;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* slab = reinterpret_cast<T*>(smem_raw);              // [ib][rpc]
    T* xch = slab + (size_t)ib * rpc;                       // [2][CS+1][IBMAX]
    T* coef = xch + 2 * (CS + 1) * IBMAX;                   // [IBMAX]
    T* zbuf = coef + IBMAX;                                 // [IBMAX]
    T* Tsm = zbuf + IBMAX;                                  // [IBMAX][IBMAX] (col-major), rank 0 only
    T* scal = Tsm + IBMAX * IBMAX;                          // [4]: scale, tau, beta

    const int r0 = rank * rpc;
    const int nr = max(0, min(rpc, mp - r0));

    for (int idx = tid; idx < nr * ib; idx += PANEL_THREADS) {
        int c = idx / nr, r = idx - c * nr;
        slab[(size_t)c * rpc + r] = A[(size_t)c * lda + r0 + r];
    }
    for (int idx = tid; idx < IBMAX * IBMAX; idx += PANEL_THREADS) Tsm[idx] = zero<T>();
    __syncthreads();
    cluster.sync();  // all CTAs resident & initialised before any DSMEM traffic

    for (int j = 0; j < ib; ++j) {
        const int par = j & 1;
        const int ls = max(0, j + 1 - r0);  // first local row strictly below the diagonal
        // ---- partial dot products over own rows: one warp per pair of columns (shared a_j loads) ----
        const T* cj = slab + (size_t)j * rpc;
        for (int l0 = 2 * warp; l0 < ib; l0 += 2 * NW) {
            const int l1 = l0 + 1;
            const bool has1 = l1 < ib;
            const T* c0 = slab + (size_t)l0 * rpc;
            const T* c1 = slab + (size_t)(has1 ? l1 : l0) * rpc;
            T s0 = zero<T>(), s1 = zero<T>();
            // l >= j: conj(a_j) * a_l ; l < j: conj(v_l) * a_j
            const bool f0 = l0 >= j, f1 = l1 >= j;
            for (int r = ls + lane; r < nr; r += 32) {
                const T aj = cj[r], x0 = c0[r], x1 = c1[r];
                if (f0) fmac_(s0, aj, x0); else fmac_(s0, x0, aj);
                if (f1) fmac_(s1, aj, x1); else fmac_(s1, x1, aj);
            }
            s0 = warp_sum(s0);
            s1 = warp_sum(s1);
            if (lane < CS) {
                T* dst = cluster.map_shared_rank(xch, lane);
                dst[(par * (CS + 1) + rank) * IBMAX + l0] = s0;
                if (has1) dst[(par * (CS + 1) + rank) * IBMAX + l1] = s1;
            }
        }
        if (rank == 0) {
            // broadcast row j of the panel (v_l[j] for l<j, a_jl for l>=j)
            for (int idx = tid; idx < ib * CS; idx += PANEL_THREADS) {
                int l = idx % ib, d = idx / ib;
                T* dst = cluster.map_shared_rank(xch, d);
                dst[(par * (CS + 1) + CS) * IBMAX + l] = slab[(size_t)l * rpc + j];
            }
        }
        cluster.sync();
        // ---- reduce, reflector scalars, update coefficients ----
        if (tid < ib) {
            const T* xp = xch + par * (CS + 1) * IBMAX;
            const T* top = xp + CS * IBMAX;
            T accl = zero<T>(), accj = zero<T>();
            for (int r = 0; r < CS; ++r) {
                accl = add_(accl, xp[r * IBMAX + tid]);
                accj = add_(accj, xp[r * IBMAX + j]);
            }
            double beta;
            T tau, scale;
            larfgp_scalars<T>(top[j], real_(accj), beta, tau, scale);
            const int l = tid;
            if (l > j) {
                T f = add_(top[l], mul_(conj_(scale), accl));
                coef[l] = mul_(conj_(tau), f);
            } else if (l < j) {
                zbuf[l] = add_(conj_(top[l]), mul_(scale, accl));  // v_l^H v_j
            } else {
                scal[0] = scale;
                scal[1] = tau;
                scal[2] = mk<T>(beta);
            }
        }
        __syncthreads();
        const T scale = scal[0], tau = scal[1];
        // ---- T column j (rank 0): T[0:j, j] = -tau * T[0:j,0:j] * z ----
        if (rank == 0 && tid < ib) {
            if (tid < j) {
                T s = zero<T>();
                for (int p = tid; p < j; ++p) fma_(s, Tsm[p * IBMAX + tid], zbuf[p]);
                Tsm[j * IBMAX + tid] = neg_(mul_(tau, s));
            } else if (tid == j) {
                Tsm[j * IBMAX + j] = tau;
                tau_out[j] = tau;
            }
        }
        // ---- apply H_j^H to the remaining columns; scale column j into v ----
        T* cjw = slab + (size_t)j * rpc;
        for (int r = ls + tid; r < nr; r += PANEL_THREADS) {
            T v = mul_(cjw[r], scale);
            cjw[r] = v;
            for (int l = j + 1; l < ib; ++l) {
                T* p = slab + (size_t)l * rpc + r;
                *p = sub_(*p, mul_(coef[l], v));
            }
        }
        if (rank == 0 && tid < ib) {
            if (tid > j) {
                T* p = slab + (size_t)tid * rpc + j;
                *p = sub_(*p, coef[tid]);
            } else if (tid == j) {
                slab[(size_t)j * rpc + j] = scal[2];
            }
        }
        __syncthreads();
    }

    for (int idx = tid; idx < nr * ib; idx += PANEL_THREADS) {
        int c = idx / nr, r = idx - c * nr;
        A[(size_t)c * lda + r0 + r] = slab[(size_t)c * rpc + r];
    }
    if (rank == 0 && Tout) {
        for (int idx = tid; idx < ib * ib; idx += PANEL_THREADS) {
            int c = idx / ib, r = idx - c * ib;
            Tout[(size_t)c * ldt + r] = Tsm[c * IBMAX + r];
        }
    }
    cluster.sync();  // no CTA may exit while peers can still address its shared memory
}

template <typename T>
static size_t panel_smem_bytes(int ib, int rpc, int cs) {
    return ((size_t)ib * rpc + 2 * (size_t)(cs + 1) * IBMAX + 2 * IBMAX + IBMAX * IBMAX + 4) * sizeof(T);
}

template <typename T>
static cudaError_t launch_panel(makb200_handle* h, int mp, int ib, T* A, int lda, T* tau, T* Tout, int ldt) {
    // cluster size: ~512 rows per CTA, every CTA must own >= ib rows so that the panel's
    // top triangle lives in rank 0, and the slab must fit in shared memory.
    int cs = 1;
    auto rows_per = [&](int c) { return ((mp + c - 1) / c + 31) / 32 * 32; };
    while (cs < h->max_cluster && mp / (cs * 2) >= ib &&
           ((mp + cs - 1) / cs > 256 || (size_t)rows_per(cs) * ib * sizeof(T) > PANEL_SLAB_BYTES))
        cs *= 2;
    int rpc = rows_per(cs);
    if ((size_t)rpc * ib * sizeof(T) > PANEL_SLAB_BYTES) return cudaErrorInvalidValue;  // caller sizes ib
    size_t smem = panel_smem_bytes<T>(ib, rpc, cs);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs, 1, 1);
    cfg.blockDim = dim3(PANEL_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    count_launch();
    return cudaLaunchKernelEx(&cfg, panel_kernel<T>, mp, ib, A, lda, tau, Tout, ldt, rpc);
}

// inner panel width such that the slab of the tallest panel fits the cluster's shared memory
template <typename T>
static int choose_ib(const makb200_handle* h, int m) {
    int ib = IBMAX;
    while (ib > 4) {
        int rpc = ((m + h->max_cluster - 1) / h->max_cluster + 31) / 32 * 32;
        if ((size_t)rpc * ib * sizeof(T) <= PANEL_SLAB_BYTES) break;
        ib /= 2;
    }
    return ib;
}

template <typename T>
static int panel_init(makb200_handle* h) {
    MAK_CUDA(h, cudaFuncSetAttribute(panel_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    MAK_CUDA(h, cudaFuncSetAttribute(panel_kernel<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    return 0;
}

int qr_init(makb200_handle* h) {
    int rc = panel_init<double>(h);
    if (rc) return rc;
    rc = panel_init<cplx>(h);
    if (rc) return rc;
    // largest cluster that can be co-scheduled with a full slab
    int best = 1;
    for (int cs = 16; cs >= 2; cs /= 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs, 1, 1);
        cfg.blockDim = dim3(PANEL_THREADS, 1, 1);
        cfg.dynamicSmemBytes = PANEL_SLAB_BYTES + 40 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int ncl = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, panel_kernel<cplx>, &cfg);
        if (e == cudaSuccess && ncl > 0) { best = cs; break; }
        cudaGetLastError();
    }
    h->max_cluster = best;
    return 0;
}

// ---------------------------------------------------------------------------------------
// small helper kernels
// ---------------------------------------------------------------------------------------
// Vw (mp x jb, ld ldv) = unit-lower-trapezoidal part of the panel A (explicit 1 / 0)
// (column c has its diagonal at row c + doff)
template <typename T>
__global__ void copy_v_kernel(int mp, int jb, int doff, const T* __restrict__ A, int lda, T* __restrict__ V,
                              int ldv) {
    size_t total = (size_t)mp * jb;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % mp), c = (int)(idx / mp);
        T v = A[(size_t)c * lda + r];
        if (r < c + doff) v = zero<T>();
        else if (r == c + doff) v = one<T>();
        V[(size_t)c * ldv + r] = v;
    }
}

template <typename T>
__global__ void set_identity_kernel(int m, int n, T* __restrict__ Q, int ldq) {
    size_t total = (size_t)m * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % m), c = (int)(idx / m);
        Q[(size_t)c * ldq + r] = (r == c) ? one<T>() : zero<T>();
    }
}

// R (rr x n) = upper-triangular part of A[0:rr, 0:n]  (uppertriangular! + copyto!, qr.jl:177-181)
template <typename T>
__global__ void extract_r_kernel(int rr, int n, int ma, const T* __restrict__ A, int lda, T* __restrict__ R,
                                 int ldr) {
    size_t total = (size_t)rr * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % rr), c = (int)(idx / rr);
        T v = zero<T>();
        if (r <= c && r < ma) v = A[(size_t)c * lda + r];
        R[(size_t)c * ldr + r] = v;
    }
}

// T[0:i0, i0:i0+ib] = -T[0:i0,0:i0] * G(i0 x ib) * T[i0:i0+ib, i0:i0+ib]; one CTA per column
template <typename T>
__global__ void t_couple_kernel(int i0, int ib, T* __restrict__ Tm, int ldt, const T* __restrict__ G, int ldg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* x = reinterpret_cast<T*>(smem_raw);  // [i0]
    const int c = blockIdx.x;
    const T* Tb = Tm + (size_t)i0 * ldt + i0;
    for (int r = threadIdx.x; r < i0; r += blockDim.x) {
        T s = zero<T>();
        for (int p = 0; p <= c; ++p) fma_(s, G[(size_t)p * ldg + r], Tb[(size_t)c * ldt + p]);
        x[r] = s;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < i0; r += blockDim.x) {
        T s = zero<T>();
        for (int p = r; p < i0; ++p) fma_(s, Tm[(size_t)p * ldt + r], x[p]);
        Tm[(size_t)(i0 + c) * ldt + r] = neg_(s);
    }
}

static inline int grid_for(size_t total, int num_sms) {
    size_t b = (total + 255) / 256;
    size_t cap = (size_t)num_sms * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------------------------------
// blocked driver
// ---------------------------------------------------------------------------------------
template <typename T>
struct QrWork {
    T* Vw;     // m x nb
    T* W;      // nb x ncmax
    T* W2;     // nb x ncmax
    T* G;      // nb x nb
    T* Tall;   // nb x nb per outer block
    T* tau;    // k
    void* ws;  // split-K scratch
    size_t ws_bytes;
    int nb, ib;
    // second set for the look-ahead stream (trailing update of block b overlaps panel b+1)
    T* Vw2;
    T* W_b;
    T* W2_b;
    void* ws_b;
};

template <typename T>
static size_t splitk_ws_bytes(const makb200_handle* h) {
    return (size_t)h->num_sms * 128 * 128 * sizeof(double) * (is_cplx<T>::value ? 1 : 1);
}

template <typename T, typename AR>
static void qr_carve(const makb200_handle* h, AR& ar, int m, int n, int ncols_q, QrWork<T>* w) {
    const int k = m < n ? m : n;
    const int nb = qr_nb();
    const int ncmax = (n > ncols_q ? n : ncols_q);
    const int nblk = (k + nb - 1) / nb;
    w->Vw = ar.template get<T>((size_t)(m > 0 ? m : 1) * nb);
    w->W = ar.template get<T>((size_t)nb * (ncmax > 0 ? ncmax : 1));
    w->W2 = ar.template get<T>((size_t)nb * (ncmax > 0 ? ncmax : 1));
    w->G = ar.template get<T>((size_t)nb * nb);
    w->Tall = ar.template get<T>((size_t)nb * nb * (nblk > 0 ? nblk : 1));
    w->tau = ar.template get<T>((size_t)(k > 0 ? k : 1));
    w->ws_bytes = splitk_ws_bytes<T>(h);
    w->ws = ar.template get<char>(w->ws_bytes);
    w->nb = nb;
    w->Vw2 = ar.template get<T>((size_t)(m > 0 ? m : 1) * nb);
    w->W_b = ar.template get<T>((size_t)nb * (ncmax > 0 ? ncmax : 1));
    w->W2_b = ar.template get<T>((size_t)nb * (ncmax > 0 ? ncmax : 1));
    w->ws_b = ar.template get<char>(w->ws_bytes);
}

#define MAK_GEMM(h, ...)                                                \
    do {                                                                \
        cudaError_t _e = gemm<T>(__VA_ARGS__);                          \
        if (_e != cudaSuccess) return cuda_fail(h, _e, "gemm");         \
    } while (0)

// Apply the block reflector H^H = I - V T^H V^H (trans=true) or H = I - V T V^H (trans=false) to
// C (mc x nc) from the left.  V: mc x kb (explicit unit-lower-trapezoidal), T: kb x kb upper.
template <typename T>
static int apply_block_reflector(makb200_handle* h, bool trans, int mc, int nc, int kb, const T* V, int ldv,
                                 const T* Tm, int ldt, T* C, int ldc, QrWork<T>& w) {
    if (mc <= 0 || nc <= 0 || kb <= 0) return 0;
    cudaStream_t s = h->stream;
    const T one_ = one<T>(), zero_ = zero<T>(), mone = neg_(one<T>());
    // W = V^H C   (kb x nc)
    MAK_GEMM(h, s, h->num_sms, MAKB200_OP_C, MAKB200_OP_N, kb, nc, mc, one_, V, ldv, C, ldc, zero_, w.W, kb, w.ws,
             w.ws_bytes);
    // W2 = op(T) W
    MAK_GEMM(h, s, h->num_sms, trans ? MAKB200_OP_C : MAKB200_OP_N, MAKB200_OP_N, kb, nc, kb, one_, Tm, ldt, w.W, kb,
             zero_, w.W2, kb, nullptr, 0);
    // C -= V W2
    MAK_GEMM(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_N, mc, nc, kb, mone, V, ldv, w.W2, kb, one_, C, ldc, nullptr,
             0);
    return 0;
}

// geqrt-style blocked factorization: A -> (V\R), tau, T factors per outer block in w.Tall.
// Look-ahead: the panel stage of block b+1 (latency-bound cluster kernels on <= 16 SMs) runs on the
// caller's stream while the bulk of block b's trailing update (DMMA GEMMs) runs on the handle's
// auxiliary stream; only the next panel's columns are updated in the panel stream's critical path.
static bool qr_lookahead() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MAKB200_QR_LOOKAHEAD"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

template <typename T>
static int geqrf_blocked(makb200_handle* h, int m, int n, T* A, int lda, QrWork<T>& w) {
    const int k = m < n ? m : n;
    if (k == 0) return 0;
    // panel chain on the high-priority auxiliary stream, bulk updates on the caller's stream
    cudaStream_t sMain = h->stream;
    struct Restore { makb200_handle* h; cudaStream_t s; ~Restore() { h->stream = s; } } restore{h, sMain};
    const bool la_on = qr_lookahead() && !h->no_lookahead && ((k + w.nb - 1) / w.nb) > 1;
    cudaStream_t sP = la_on ? h->aux_stream : sMain, sG = sMain;
    const int nb = w.nb;
    const int ib_max = choose_ib<T>(h, m);
    const int nblk = (k + nb - 1) / nb;
    const bool la = la_on;
    if (la) {  // fork: the auxiliary stream starts after everything already queued on the caller's
        MAK_CUDA(h, cudaEventRecord(h->ev[0], sMain));
        MAK_CUDA(h, cudaStreamWaitEvent(sP, h->ev[0], 0));
    }
    MAK_CUDA(h, cudaMemsetAsync(w.Tall, 0, sizeof(T) * (size_t)nb * nb * nblk, sP));
    // panel-stream and update-stream views of the workspace
    QrWork<T> wP = w, wG = w;
    wG.W = w.W_b; wG.W2 = w.W2_b; wG.ws = w.ws_b;
    int rc = 0;
    for (int b = 0; b < nblk; ++b) {
        const int j0 = b * nb;
        const int jb = (k - j0 < nb) ? (k - j0) : nb;
        const int mp = m - j0;
        T* Ap = A + (size_t)j0 * lda + j0;
        T* Tb = w.Tall + (size_t)b * nb * nb;
        T* Vw = (la && (b & 1)) ? w.Vw2 : w.Vw;
        cudaEvent_t evP = h->ev[1 + (b & 1)], evG = h->ev[3 + (b & 1)], evGprev = h->ev[3 + ((b + 1) & 1)];
        h->stream = sP;
        cudaStream_t s = sP;
        for (int i0 = 0; i0 < jb; i0 += ib_max) {
            const int ib = (jb - i0 < ib_max) ? (jb - i0) : ib_max;
            T* Ai = Ap + (size_t)i0 * lda + i0;
            const int mi = mp - i0;
            cudaError_t e = launch_panel<T>(h, mi, ib, Ai, lda, w.tau + j0 + i0, Tb + (size_t)i0 * nb + i0, nb);
            if (e != cudaSuccess) return cuda_fail(h, e, "panel_kernel");
            // explicit V for this inner block into Vw[:, i0:i0+ib] (zeros above the diagonal)
            copy_v_kernel<T><<<grid_for((size_t)mp * ib, h->num_sms), 256, 0, s>>>(
                mp, ib, i0, Ap + (size_t)i0 * lda, lda, Vw + (size_t)i0 * mp, mp);
            count_launch();
            MAK_LAUNCH_CHECK(h, "copy_v_kernel");
            const T* Vi = Vw + (size_t)i0 * mp + i0;
            // update the rest of the outer panel
            const int nci = jb - i0 - ib;
            if (nci > 0) {
                rc = apply_block_reflector<T>(h, true, mi, nci, ib, Vi, mp, Tb + (size_t)i0 * nb + i0, nb,
                                              Ai + (size_t)ib * lda, lda, wP);
                if (rc) return rc;
            }
            // couple into the outer T: T[0:i0, i0:i0+ib] = -T_a (V_a^H V_i) T_i
            if (i0 > 0) {
                const T one_ = one<T>(), zero_ = zero<T>();
                MAK_GEMM(h, s, h->num_sms, MAKB200_OP_C, MAKB200_OP_N, i0, ib, mp, one_, Vw, mp,
                         Vw + (size_t)i0 * mp, mp, zero_, w.G, nb, wP.ws, wP.ws_bytes);
                t_couple_kernel<T><<<ib, 128, sizeof(T) * i0, s>>>(i0, ib, Tb, nb, w.G, nb);
                count_launch();
                MAK_LAUNCH_CHECK(h, "t_couple_kernel");
            }
        }
        // trailing update with the outer block reflector
        const int nc = n - j0 - jb;
        if (nc <= 0) continue;
        if (!la) {
            rc = apply_block_reflector<T>(h, true, mp, nc, jb, Vw, mp, Tb, nb, Ap + (size_t)jb * lda, lda, wP);
            if (rc) return rc;
            continue;
        }
        MAK_CUDA(h, cudaEventRecord(evP, sP));
        const int n1 = nc < nb ? nc : nb;  // the next panel's columns: critical path, panel stream
        const int n2 = nc - n1;            // the rest: auxiliary stream, overlaps the next panel stage
        if (n2 > 0) {
            MAK_CUDA(h, cudaStreamWaitEvent(sG, evP, 0));
            h->stream = sG;
            rc = apply_block_reflector<T>(h, true, mp, n2, jb, Vw, mp, Tb, nb, Ap + (size_t)(jb + n1) * lda, lda, wG);
            h->stream = sP;
            if (rc) return rc;
            MAK_CUDA(h, cudaEventRecord(evG, sG));
        }
        // the next panel's columns were last written by the previous block's bulk update
        if (b > 0) MAK_CUDA(h, cudaStreamWaitEvent(sP, evGprev, 0));
        rc = apply_block_reflector<T>(h, true, mp, n1, jb, Vw, mp, Tb, nb, Ap + (size_t)jb * lda, lda, wP);
        if (rc) return rc;
    }
    h->stream = sMain;
    if (la) {  // join: the caller's stream continues after the panel chain has drained
        MAK_CUDA(h, cudaEventRecord(h->ev[5], sP));
        MAK_CUDA(h, cudaStreamWaitEvent(sMain, h->ev[5], 0));
    }
    return 0;
}

// Q (m x ncols) = first ncols columns of H_1 ... H_k, from V (in A) and the stored T blocks
template <typename T>
static int orgqr_blocked(makb200_handle* h, int m, int ncols, int k, const T* A, int lda, T* Q, int ldq,
                         QrWork<T>& w) {
    cudaStream_t s = h->stream;
    if (m <= 0 || ncols <= 0) return 0;
    set_identity_kernel<T><<<grid_for((size_t)m * ncols, h->num_sms), 256, 0, s>>>(m, ncols, Q, ldq);
    count_launch();
    MAK_LAUNCH_CHECK(h, "set_identity_kernel");
    if (k == 0) return 0;
    const int nb = w.nb;
    const int nblk = (k + nb - 1) / nb;
    for (int b = nblk - 1; b >= 0; --b) {
        const int j0 = b * nb;
        const int jb = (k - j0 < nb) ? (k - j0) : nb;
        const int mp = m - j0;
        const T* Ap = A + (size_t)j0 * lda + j0;
        const T* Tb = w.Tall + (size_t)b * nb * nb;
        copy_v_kernel<T><<<grid_for((size_t)mp * jb, h->num_sms), 256, 0, s>>>(mp, jb, 0, Ap, lda, w.Vw, mp);
        count_launch();
    MAK_LAUNCH_CHECK(h, "copy_v_kernel");
        // all columns j0 .. ncols-1 (own block columns are still [I;0])
        const int nc = ncols - j0;
        int rc = apply_block_reflector<T>(h, false, mp, nc, jb, w.Vw, mp, Tb, nb, Q + (size_t)j0 * ldq + j0, ldq, w);
        if (rc) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// entry points used by capi.cu
// ---------------------------------------------------------------------------------------
template <typename T>
size_t qr_worksize_t(makb200_handle* h, int m, int n, int ncols_q) {
    ArenaSize ar;
    QrWork<T> w;
    qr_carve<T>(h, ar, m, n, ncols_q, &w);
    return ar.off + 256;
}

template <typename T>
int qr_fused_t(makb200_handle* h, int mode, int m, int n, T* A, int lda, T* Q, int ldq, T* R, int ldr, void* work,
               size_t lwork) {
    const int k = m < n ? m : n;
    const int ncq = (mode == MAKB200_QR_FULL) ? m : k;
    Arena ar(work, lwork);
    QrWork<T> w;
    qr_carve<T>(h, ar, m, n, ncq, &w);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    PhaseTimer pt(h->stream);
    pt.mark("start");
    int rc = geqrf_blocked<T>(h, m, n, A, lda, w);
    if (rc) return rc;
    pt.mark("geqrf");
    if (R && ldr > 0) {
        const int rr = ncq;
        if (rr > 0 && n > 0) {
            extract_r_kernel<T><<<grid_for((size_t)rr * n, h->num_sms), 256, 0, h->stream>>>(rr, n, m, A, lda, R, ldr);
            count_launch();
            MAK_LAUNCH_CHECK(h, "extract_r_kernel");
        }
    }
    rc = orgqr_blocked<T>(h, m, ncq, k, A, lda, Q, ldq, w);
    pt.mark("orgqr");
    pt.report("qr");
    return rc;
}

template <typename T>
int geqrf_t(makb200_handle* h, int m, int n, T* A, int lda, T* tau, void* work, size_t lwork) {
    const int k = m < n ? m : n;
    Arena ar(work, lwork);
    QrWork<T> w;
    qr_carve<T>(h, ar, m, n, k, &w);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    int rc = geqrf_blocked<T>(h, m, n, A, lda, w);
    if (rc) return rc;
    if (k > 0) MAK_CUDA(h, cudaMemcpyAsync(tau, w.tau, sizeof(T) * k, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

// rebuild the T factors of every outer block from V and tau (larft), for the L1 orgqr entry
// T (jb x jb, upper) from G = V^H V and tau:  T^-1 = diag(1/tau) + striu(G)  =>  column j of T by
// back substitution  x_j = tau_j,  x_i = -tau_i * sum_{p=i+1..j} G[i,p] x_p   (no division: tau_i = 0
// gives a zero row, as larft does).  One thread per column, strictly-upper G packed in shared memory.
template <typename T>
__device__ __forceinline__ void larft_diag_body(int jb, const T* __restrict__ tau, T* __restrict__ Tm, int ldt,
                                                const T* __restrict__ G, int ldg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Gs = reinterpret_cast<T*>(smem_raw);          // packed: (i,p), i<p  ->  p(p-1)/2 + i
    T* ts = Gs + (size_t)jb * (jb - 1) / 2;          // tau
    for (int idx = threadIdx.x; idx < jb * jb; idx += blockDim.x) {
        int p = idx / jb, i = idx - p * jb;
        if (i < p) Gs[(size_t)p * (p - 1) / 2 + i] = G[(size_t)p * ldg + i];
    }
    for (int j = threadIdx.x; j < jb; j += blockDim.x) ts[j] = tau[j];
    __syncthreads();
    for (int j = threadIdx.x; j < jb; j += blockDim.x) {
        T* col = Tm + (size_t)j * ldt;
        col[j] = ts[j];
        for (int i = j - 1; i >= 0; --i) {
            T s = zero<T>();
            for (int p = i + 1; p <= j; ++p) fma_(s, Gs[(size_t)p * (p - 1) / 2 + i], col[p]);
            col[i] = neg_(mul_(ts[i], s));
        }
    }
}
template <typename T>
__global__ void larft_diag_kernel(int jb, const T* __restrict__ tau, T* __restrict__ Tm, int ldt,
                                  const T* __restrict__ G, int ldg) {
    larft_diag_body<T>(jb, tau, Tm, ldt, G, ldg);
}
template <typename T> static size_t larft_smem(int jb) { return sizeof(T) * ((size_t)jb * (jb - 1) / 2 + jb + 2); }
template <typename T>
static cudaError_t larft_launch(cudaStream_t s, int jb, const T* tau, T* Tm, int ldt, const T* G, int ldg) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(larft_diag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    larft_diag_kernel<T><<<1, 128, larft_smem<T>(jb), s>>>(jb, tau, Tm, ldt, G, ldg);
    return cudaGetLastError();
}

template <typename T>
int orgqr_t(makb200_handle* h, int m, int ncols, int k, const T* A, int lda, const T* tau, T* Q, int ldq, void* work,
            size_t lwork) {
    Arena ar(work, lwork);
    QrWork<T> w;
    qr_carve<T>(h, ar, m, k, ncols, &w);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    cudaStream_t s = h->stream;
    const int nb = w.nb;
    const int nblk = (k + nb - 1) / nb;
    if (k > 0) {
        MAK_CUDA(h, cudaMemsetAsync(w.Tall, 0, sizeof(T) * (size_t)nb * nb * nblk, s));
        for (int b = 0; b < nblk; ++b) {
            const int j0 = b * nb;
            const int jb = (k - j0 < nb) ? (k - j0) : nb;
            const int mp = m - j0;
            copy_v_kernel<T><<<grid_for((size_t)mp * jb, h->num_sms), 256, 0, s>>>(
                mp, jb, 0, A + (size_t)j0 * lda + j0, lda, w.Vw, mp);
            count_launch();
    MAK_LAUNCH_CHECK(h, "copy_v_kernel");
            MAK_GEMM(h, s, h->num_sms, MAKB200_OP_C, MAKB200_OP_N, jb, jb, mp, one<T>(), w.Vw, mp, w.Vw, mp, zero<T>(),
                     w.G, nb, w.ws, w.ws_bytes);
            {
                // through larft_launch: it opts the kernel in to > 48 KB of dynamic shared memory (jb = 128 needs 66 KB)
                cudaError_t e = larft_launch<T>(s, jb, tau + j0, w.Tall + (size_t)b * nb * nb, nb, w.G, nb);
                if (e != cudaSuccess) return cuda_fail(h, e, "larft_diag_kernel");
            }
            count_launch();
        }
    }
    return orgqr_blocked<T>(h, m, ncols, k, A, lda, Q, ldq, w);
}

// C (m x nc) <- H_0 ... H_{k-1} C  (ormqr/unmqr 'L','N'): V below the diagonal of A (m x k), tau
template <typename T>
size_t ormqr_worksize_t(makb200_handle* h, int m, int k, int nc) {
    ArenaSize ar;
    QrWork<T> w;
    qr_carve<T>(h, ar, m, k, nc, &w);
    return ar.off + 256;
}

template <typename T>
int ormqr_left_t(makb200_handle* h, int m, int k, const T* A, int lda, const T* tau, T* C, int ldc, int nc,
                 void* work, size_t lwork, bool adjoint) {
    if (m <= 0 || nc <= 0 || k <= 0) return 0;
    Arena ar(work, lwork);
    QrWork<T> w;
    qr_carve<T>(h, ar, m, k, nc, &w);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    cudaStream_t s = h->stream;
    const int nb = w.nb;
    const int nblk = (k + nb - 1) / nb;
    T* Tb = w.Tall;  // one block at a time
    // Q C = H_0 (H_1 (... C)): last block first;  Q^H C = H_{k-1}^H (... (H_0^H C)): first block first, T^H
    for (int bb = 0; bb < nblk; ++bb) {
        const int b = adjoint ? bb : nblk - 1 - bb;
        const int j0 = b * nb;
        const int jb = (k - j0 < nb) ? (k - j0) : nb;
        const int mp = m - j0;
        copy_v_kernel<T><<<grid_for((size_t)mp * jb, h->num_sms), 256, 0, s>>>(mp, jb, 0, A + (size_t)j0 * lda + j0,
                                                                                lda, w.Vw, mp);
        count_launch();
    MAK_LAUNCH_CHECK(h, "copy_v_kernel");
        MAK_CUDA(h, cudaMemsetAsync(Tb, 0, sizeof(T) * (size_t)nb * nb, s));
        MAK_GEMM(h, s, h->num_sms, MAKB200_OP_C, MAKB200_OP_N, jb, jb, mp, one<T>(), w.Vw, mp, w.Vw, mp, zero<T>(), w.G,
                 nb, w.ws, w.ws_bytes);
        {
            cudaError_t e = larft_launch<T>(s, jb, tau + j0, Tb, nb, w.G, nb);
            if (e != cudaSuccess) return cuda_fail(h, e, "larft_diag_kernel");
        }
        count_launch();
        int rc = apply_block_reflector<T>(h, adjoint, mp, nc, jb, w.Vw, mp, Tb, nb, C + j0, ldc, w);
        if (rc) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// C_i <- H_0 ... H_{k_i - 1} C_i for MANY blocks in lock-step (back-transformation of the batched eigh / svd of
// mid-size blocks): per panel of BORM_NB reflectors ONE kernel builds the explicit unit-lower V panel of every active
// block, one grouped GEMM forms V^H V, one kernel (a CTA per block) the compact-WY T, three grouped GEMMs apply
// I - V T V^H.  Blocks are sorted by k descending, so the blocks that own panel b are a prefix.  The single-matrix
// routine above costs ~8 launches per panel PER BLOCK.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void bormqr_prep_kernel(const OrmqrBatchBlk<T>* __restrict__ blks, int j0) {
    const OrmqrBatchBlk<T> b = blks[blockIdx.y];
    const int jb = (b.k - j0 < BORM_NB) ? (b.k - j0) : BORM_NB, mp = b.m - j0;
    if (jb <= 0) return;
    const T* A = b.A + (size_t)j0 * b.lda + j0;
    const size_t start = blockIdx.x * (size_t)blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    const size_t total = (size_t)mp * jb;
    for (size_t idx = start; idx < total; idx += step) {
        const int r = (int)(idx % mp), c = (int)(idx / mp);
        T v = A[(size_t)c * b.lda + r];
        if (r < c) v = zero<T>();
        else if (r == c) v = one<T>();
        b.Vw[idx] = v;
    }
    for (size_t idx = start; idx < (size_t)BORM_NB * BORM_NB; idx += step) b.Tb[idx] = zero<T>();
}
template <typename T>
__global__ void __launch_bounds__(128) bormqr_larft_kernel(const OrmqrBatchBlk<T>* __restrict__ blks, int j0) {
    const OrmqrBatchBlk<T> b = blks[blockIdx.x];
    const int jb = (b.k - j0 < BORM_NB) ? (b.k - j0) : BORM_NB;
    if (jb <= 0) return;
    larft_diag_body<T>(jb, b.tau + j0, b.Tb, BORM_NB, b.G, BORM_NB);
}

template <typename T>
size_t ormqr_batched_block_elems(int m, int nc) {
    auto ev = [](size_t e) { return (e + 1) & ~(size_t)1; };
    return ev((size_t)(m > 0 ? m : 1) * BORM_NB) + 2 * ev((size_t)BORM_NB * BORM_NB) + 2 * ev((size_t)BORM_NB * (nc > 0 ? nc : 1));
}
template <typename T>
void ormqr_batched_carve(OrmqrBatchBlk<T>& b, T*& p) {
    auto take = [&](size_t e) { T* r = p; p += (e + 1) & ~(size_t)1; return r; };
    b.Vw = take((size_t)(b.m > 0 ? b.m : 1) * BORM_NB);
    b.G = take((size_t)BORM_NB * BORM_NB);
    b.Tb = take((size_t)BORM_NB * BORM_NB);
    b.W = take((size_t)BORM_NB * (b.nc > 0 ? b.nc : 1));
    b.W2 = take((size_t)BORM_NB * (b.nc > 0 ? b.nc : 1));
}
template <typename T>
size_t ormqr_batched_table_bytes(int count, int kmax) {
    const size_t panels = (size_t)((kmax + BORM_NB - 1) / BORM_NB);
    return align_up(sizeof(OrmqrBatchBlk<T>) * (size_t)count, 256) + align_up(sizeof(GemmProblem<T>) * 4 * panels * (size_t)count, 256) + 512;
}

template <typename T>
int ormqr_left_batched(makb200_handle* h, int count, const OrmqrBatchBlk<T>* blks, char* tables, size_t tables_bytes) {
    if (count <= 0) return 0;
    cudaStream_t s = h->stream;
    int kmax = 0, mmax = 0, ncmax = 0;
    for (int i = 0; i < count; ++i) {
        if (i > 0 && blks[i].k > blks[i - 1].k) return -2;   // k descending
        kmax = std::max(kmax, blks[i].k); mmax = std::max(mmax, blks[i].m); ncmax = std::max(ncmax, blks[i].nc);
    }
    if (kmax <= 0) return 0;
    const int npanel = (kmax + BORM_NB - 1) / BORM_NB;
    struct Launch { int opa, opb, count, max_m, max_n; size_t off; };
    std::vector<GemmProblem<T>> probs;
    std::vector<Launch> launches;   // 4 per panel, in order
    std::vector<int> actives;
    auto prob = [](int m, int n, int k, const T* A, int lda, const T* B, int ldb, T* C, int ldc, double al, double be, int ca) {
        GemmProblem<T> p;
        p.m = m; p.n = n; p.k = k; p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
        p.alpha = mk<T>(al); p.beta = mk<T>(be); p.conja = ca; p.conjb = 0; p.lower = 0;
        return p;
    };
    for (int bb = npanel - 1; bb >= 0; --bb) {
        const int j0 = bb * BORM_NB;
        int na = 0;
        while (na < count && blks[na].k > j0) ++na;
        actives.push_back(na);
        Launch l[4] = {{MAKB200_OP_C, MAKB200_OP_N, na, 0, 0, 0}, {MAKB200_OP_C, MAKB200_OP_N, na, 0, 0, 0},
                       {MAKB200_OP_N, MAKB200_OP_N, na, 0, 0, 0}, {MAKB200_OP_N, MAKB200_OP_N, na, 0, 0, 0}};
        for (int g = 0; g < 4; ++g) {
            l[g].off = probs.size();
            for (int i = 0; i < na; ++i) {
                const OrmqrBatchBlk<T>& b = blks[i];
                const int jb = std::min(BORM_NB, b.k - j0), mp = b.m - j0;
                GemmProblem<T> p;
                if (g == 0) p = prob(jb, jb, mp, b.Vw, mp, b.Vw, mp, b.G, BORM_NB, 1.0, 0.0, 1);                  // G = V^H V
                else if (g == 1) p = prob(jb, b.nc, mp, b.Vw, mp, b.C + j0, b.ldc, b.W, jb, 1.0, 0.0, 1);           // W = V^H C
                else if (g == 2) p = prob(jb, b.nc, jb, b.Tb, BORM_NB, b.W, jb, b.W2, jb, 1.0, 0.0, 0);             // W2 = T W
                else p = prob(mp, b.nc, jb, b.Vw, mp, b.W2, jb, b.C + j0, b.ldc, -1.0, 1.0, 0);                     // C -= V W2
                l[g].max_m = std::max(l[g].max_m, p.m); l[g].max_n = std::max(l[g].max_n, p.n);
                probs.push_back(p);
            }
            launches.push_back(l[g]);
        }
    }
    OrmqrBatchBlk<T>* bdev = (OrmqrBatchBlk<T>*)tables;
    GemmProblem<T>* pdev = (GemmProblem<T>*)(tables + align_up(sizeof(OrmqrBatchBlk<T>) * (size_t)count, 256));
    if (align_up(sizeof(OrmqrBatchBlk<T>) * (size_t)count, 256) + sizeof(GemmProblem<T>) * probs.size() > tables_bytes)
        return MAKB200_ERR_WORKSPACE;
    {
        Stager st(h, sizeof(OrmqrBatchBlk<T>) * (size_t)count + sizeof(GemmProblem<T>) * probs.size() + 1024);
        MAK_CUDA(h, st.put(bdev, blks, sizeof(OrmqrBatchBlk<T>) * (size_t)count, s));
        MAK_CUDA(h, st.put(pdev, probs.data(), sizeof(GemmProblem<T>) * probs.size(), s));
    }
    static_assert(sizeof(cplx) * ((size_t)BORM_NB * (BORM_NB - 1) / 2 + BORM_NB + 2) <= 48 * 1024, "larft panel must fit default smem");
    size_t li = 0;
    for (int bb = npanel - 1, pi = 0; bb >= 0; --bb, ++pi) {
        const int j0 = bb * BORM_NB, na = actives[pi];
        if (na <= 0) { li += 4; continue; }
        const int gx = std::max(1, std::min((mmax * BORM_NB + 255) / 256, 32));
        bormqr_prep_kernel<T><<<dim3(gx, na), 256, 0, s>>>(bdev, j0);
        count_launch();
        for (int g = 0; g < 4; ++g, ++li) {
            if (g == 1) {
                bormqr_larft_kernel<T><<<na, 128, larft_smem<T>(BORM_NB), s>>>(bdev, j0);
                count_launch();
            }
            const Launch& l = launches[li];
            cudaError_t e = gemm_grouped<T>(s, l.opa, l.opb, l.count, l.max_m, l.max_n, pdev + l.off);
            if (e != cudaSuccess) return cuda_fail(h, e, "ormqr_left_batched: grouped gemm");
        }
    }
    MAK_LAUNCH_CHECK(h, "ormqr_left_batched");
    return 0;
}

// ---------------------------------------------------------------------------------------
// EXPERIMENTAL (round 1): first stage of the two-stage Hermitian tridiagonalisation, dense -> band
// (lower bandwidth b), in place on a FULL Hermitian matrix (both triangles valid and kept valid):
// for every column panel the block below the band is QR-factorized (cluster panel kernel), and the
// trailing matrix gets the two-sided update  A22 <- Q^H A22 Q = A22 - W V^H - V W^H  with
//   Y = A22 V,  X = Y T,  W = X - 1/2 V (T^H (V^H X))
// all DMMA GEMMs (the rank-2b update as ONE product [V W][W V]^H).  On exit the lower band of A holds
// B = Q1^H A Q1, the reflectors sit below the band (QR-type columns of A[b:, 0:n-b]) with tau1[n-b].
// ---------------------------------------------------------------------------------------
template <typename T>
struct Sy2sbWork {
    QrWork<T> qr;
    T* PB;     // (n-b) x 3b : [V | W | V]
    T* Y;      // (n-b) x b
    T* M;      // b x b
    T* M2;     // b x b
};
template <typename T, typename AR>
static void sy2sb_carve(const makb200_handle* h, AR& ar, int n, int b, Sy2sbWork<T>* w) {
    const int m = n > b ? n - b : 1;
    qr_carve<T>(h, ar, m, b, b, &w->qr);
    w->PB = ar.template get<T>((size_t)m * 3 * b);
    w->Y = ar.template get<T>((size_t)m * b);
    w->M = ar.template get<T>((size_t)b * b);
    w->M2 = ar.template get<T>((size_t)b * b);
}
template <typename T>
size_t sy2sb_worksize_t(makb200_handle* h, int n, int b) {
    ArenaSize ar;
    Sy2sbWork<T> w;
    sy2sb_carve<T>(h, ar, n, b, &w);
    return ar.off + 256;
}
static bool sy2sb_lookahead() {
    const char* e = getenv("MAKB200_SY2SB_LOOKAHEAD");   // read per call: the bring-up runs toggle it
    return e && e[0] == '1';
}
static bool sy2sb_lower() {
    const char* e = getenv("MAKB200_SY2SB_LOWER");   // read per call: the bring-up runs toggle it
    return e && e[0] == '1';
}

template <typename T>
int sy2sb_t(makb200_handle* h, int n, int b, T* A, int lda, T* tau1, void* work, size_t lwork) {
    if (n <= 0) return 0;
    if (b < 1 || b > qr_nb()) return -3;
    Arena ar(work, lwork);
    Sy2sbWork<T> w;
    sy2sb_carve<T>(h, ar, n, b, &w);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    cudaStream_t s = h->stream;
    const int nb = w.qr.nb;
    const T one_ = one<T>(), zero_ = zero<T>(), mone = neg_(one<T>());
    MAK_CUDA(h, cudaMemsetAsync(tau1, 0, sizeof(T) * (size_t)n, s));
    const bool la_save = h->no_lookahead;
    h->no_lookahead = true;   // single-block panels: keep everything on the caller's stream
    struct Restore { makb200_handle* h; bool v; ~Restore() { h->no_lookahead = v; } } restore{h, la_save};
    // opt-in (MAKB200_SY2SB_LOOKAHEAD=1, round-2 bring-up): the next panel's columns are updated first and its QR
    // runs on the high-priority auxiliary stream while the bulk of the rank-2b update drains on the caller's stream
    // (the panel chain is ~25 % of the stage at n = 8192: 128 panels x 64 columns of cluster-barrier latency)
    const bool lookahead = sy2sb_lookahead() && !la_save;   // the auxiliary stream is shared: not inside a pooled call
    const bool lower = sy2sb_lower();
    cudaStream_t sMain = s, sAux = h->aux_stream;
    bool panel_ready = false;   // the QR of the current panel was already done (on sAux) by the previous iteration
    for (int j0 = 0; j0 + b < n; j0 += b) {
        const int mp = n - j0 - b, kb = mp < b ? mp : b;
        T* P = A + (size_t)j0 * lda + (j0 + b);
        int rc = 0;
        if (!panel_ready) {
            rc = geqrf_blocked<T>(h, mp, b, P, lda, w.qr);       // V -> qr.Vw (mp x kb, ld mp), T -> qr.Tall (ld nb)
            if (rc) return rc;
        } else {
            MAK_CUDA(h, cudaStreamWaitEvent(sMain, h->ev[7], 0));   // join the look-ahead panel
        }
        panel_ready = false;
        MAK_CUDA(h, cudaMemcpyAsync(tau1 + j0, w.qr.tau, sizeof(T) * kb, cudaMemcpyDeviceToDevice, s));
        const T* V = w.qr.Vw;
        const T* Tm = w.qr.Tall;
        T* A22 = A + (size_t)(j0 + b) * lda + (j0 + b);
        T* PB0 = w.PB, *PB1 = w.PB + (size_t)kb * mp, *PB2 = w.PB + (size_t)2 * kb * mp;
        MAK_CUDA(h, cudaMemcpyAsync(PB0, V, sizeof(T) * (size_t)mp * kb, cudaMemcpyDeviceToDevice, s));
        MAK_CUDA(h, cudaMemcpyAsync(PB2, V, sizeof(T) * (size_t)mp * kb, cudaMemcpyDeviceToDevice, s));
        // Y = A22 V ; X = Y T (into the W slot) ; M = V^H X ; M2 = T^H M ; W = X - 1/2 V M2
        MAK_GEMM(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_N, mp, kb, mp, one_, A22, lda, V, mp, zero_, w.Y, mp,
                 w.qr.ws, w.qr.ws_bytes);
        MAK_GEMM(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_N, mp, kb, kb, one_, w.Y, mp, Tm, nb, zero_, PB1, mp, nullptr, 0);
        MAK_GEMM(h, s, h->num_sms, MAKB200_OP_C, MAKB200_OP_N, kb, kb, mp, one_, V, mp, PB1, mp, zero_, w.M, b, w.qr.ws,
                 w.qr.ws_bytes);
        MAK_GEMM(h, s, h->num_sms, MAKB200_OP_C, MAKB200_OP_N, kb, kb, kb, one_, Tm, nb, w.M, b, zero_, w.M2, b, nullptr, 0);
        MAK_GEMM(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_N, mp, kb, kb, mk<T>(-0.5), V, mp, w.M2, b, one_, PB1, mp,
                 nullptr, 0);
        // A22 -= [V W] [W V]^H
        const int j1 = j0 + b;
        const bool next = lookahead && (j1 + b < n) && mp > b;
        // columns [c_lo, mp) of A22 are updated by the bulk product below; with look-ahead the first b columns
        // (the next panel and its band block) go first, on their own
        int c_lo = 0;
        if (next) {
            MAK_GEMM(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_C, mp, b, 2 * kb, mone, PB0, mp, PB1, mp, one_, A22, lda,
                     nullptr, 0);
            MAK_CUDA(h, cudaEventRecord(h->ev[6], sMain));
            MAK_CUDA(h, cudaStreamWaitEvent(sAux, h->ev[6], 0));
            h->stream = sAux;
            rc = geqrf_blocked<T>(h, n - j1 - b, b, A + (size_t)j1 * lda + (j1 + b), lda, w.qr);
            h->stream = sMain;
            if (rc) return rc;
            MAK_CUDA(h, cudaEventRecord(h->ev[7], sAux));
            panel_ready = true;
            c_lo = b;
        }
        const int nc = mp - c_lo;
        if (nc <= 0) continue;
        if (lower) {
            // opt-in (MAKB200_SY2SB_LOWER=1): only the tiles on and below the diagonal of the square block
            // A22[c_lo:, c_lo:] (half the flops of the rank-2b update), then upper <- conj(lower) so that the next
            // Y = A22 V still sees the full Hermitian matrix (one pass per panel: ~3 % of the update at n = 8192).
            // Without look-ahead c_lo = 0 and the block is all of A22; with it the strip A22[0:b, b:] above the
            // block is the mirror of the next panel's columns, which nothing reads again.
            T* C = A22 + (size_t)c_lo * lda + c_lo;
            cudaError_t e_ = gemm<T>(s, h->num_sms, MAKB200_OP_N, MAKB200_OP_C, nc, nc, 2 * kb, mone, PB0 + c_lo, mp,
                                     PB1 + c_lo, mp, one_, C, lda, nullptr, 0, true);
            if (e_ != cudaSuccess) return cuda_fail(h, e_, "gemm");
            const int nbt = (nc + 31) / 32;
            mirror_lower_kernel<T><<<dim3(nbt, nbt), dim3(32, 8), 0, s>>>(nc, C, lda);
            count_launch();
            MAK_LAUNCH_CHECK(h, "mirror_lower_kernel");
        } else {
            MAK_GEMM(h, s, h->num_sms, MAKB200_OP_N, MAKB200_OP_C, mp, nc, 2 * kb, mone, PB0, mp, PB1 + c_lo, mp, one_,
                     A22 + (size_t)c_lo * lda, lda, nullptr, 0);
        }
    }
    if (panel_ready) MAK_CUDA(h, cudaStreamWaitEvent(sMain, h->ev[7], 0));
    return 0;
}

#define INST(T)                                                                                            \
    template size_t sy2sb_worksize_t<T>(makb200_handle*, int, int);                                        \
    template int sy2sb_t<T>(makb200_handle*, int, int, T*, int, T*, void*, size_t);                        \
    template size_t qr_worksize_t<T>(makb200_handle*, int, int, int);                                      \
    template int qr_fused_t<T>(makb200_handle*, int, int, int, T*, int, T*, int, T*, int, void*, size_t);  \
    template int geqrf_t<T>(makb200_handle*, int, int, T*, int, T*, void*, size_t);                        \
    template int orgqr_t<T>(makb200_handle*, int, int, int, const T*, int, const T*, T*, int, void*, size_t); \
    template size_t ormqr_worksize_t<T>(makb200_handle*, int, int, int);                                   \
    template int ormqr_left_t<T>(makb200_handle*, int, int, const T*, int, const T*, T*, int, int, void*, size_t, bool); \
    template size_t ormqr_batched_block_elems<T>(int, int);                                                 \
    template void ormqr_batched_carve<T>(OrmqrBatchBlk<T>&, T*&);                                           \
    template size_t ormqr_batched_table_bytes<T>(int, int);                                                 \
    template int ormqr_left_batched<T>(makb200_handle*, int, const OrmqrBatchBlk<T>*, char*, size_t);
INST(double)
INST(cplx)

}  // namespace mak
