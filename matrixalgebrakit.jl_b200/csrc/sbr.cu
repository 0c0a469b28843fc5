// Second stage of the two-stage Hermitian tridiagonalisation: band (lower bandwidth b) -> tridiagonal
// by bulge chasing.  EXPERIMENTAL in round 1: exposed through makb200_sbr_chase for bring-up and
// timing; eigh_full! still uses the one-stage reduction.
//
// Geometry, task order and the independence rule come from sbr_core.h (validated on the CPU by
// tests/test_sbr_core_cpu.py): task (s, k) of sweep s touches the b x b off-diagonal block at
// (r0, r0 - b) and the diagonal block at r0 = s + 1 + k b; tasks with equal 2 s + k are independent.
// One launch per wavefront, one CTA per task, both blocks resident in shared memory; consecutive
// wavefronts are chained with programmatic dependent launch.  The band (2 b x n) stays L2-resident
// (8 MB at n = 8192, b = 64), so the stage is launch/latency bound, not HBM bound.
#include "sbr.cuh"
#include "sbr_core.h"
#include "sbr_chase_persistent.cuh"
#include "sbr_q2_slab.cuh"
#include "gemm.cuh"
#include <cstdlib>
#include <vector>

namespace mak {

constexpr int SBR_THREADS = 256;
constexpr int SBR_BMAX = 64;   // both blocks of a ComplexF64 task fit shared memory (2 x 64 x 65 x 16 B = 133 KB)

// AB[(i-j) + j*ldab] = A[i, j] for 0 <= i-j <= b, zero for b < i-j < ldab; diagonal made real
template <typename T>
__global__ void band_pack_kernel(int n, int b, const T* __restrict__ A, int lda, T* __restrict__ AB, int ldab) {
    const int j = blockIdx.x;
    for (int d = threadIdx.x; d < ldab; d += blockDim.x) {
        const int i = j + d;
        T v = zero<T>();
        if (d <= b && i < n) v = A[(size_t)j * lda + i];
        if (d == 0) v = mk<T>(real_(v));
        AB[(size_t)j * ldab + d] = v;
    }
}

template <typename T>
__global__ void band_diag_kernel(int n, const T* __restrict__ AB, int ldab, double* __restrict__ d, double* __restrict__ e,
                                 const int* __restrict__ abort_flag) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    // a persistent chase that gave up (progress wait timed out) must not look like a result
    const bool bad = abort_flag != nullptr && *abort_flag != 0;
    d[j] = bad ? __longlong_as_double(0x7ff8000000000000LL) : real_(AB[(size_t)j * ldab]);
    if (j + 1 < n) e[j] = real_(AB[(size_t)j * ldab + 1]);
}

__device__ __forceinline__ void sbr_pdl() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// all tasks of wavefront t: CTA q handles k = (t & 1) + 2 q, s = (t - k) / 2
template <typename T>
__global__ void __launch_bounds__(SBR_THREADS)
chase_wave_kernel(int n, int b, T* __restrict__ AB, int ldab, T* __restrict__ V2, int ldv, T* __restrict__ tau2, int ldt,
                  int t) {
    sbr_pdl();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = (t & 1) + 2 * (int)blockIdx.x;
    const int s = (t - k) / 2;
    if (k > t || s > n - 2 || k >= sbr::sweep_ntasks(n, b, s)) return;
    const sbr::Task tk = sbr::task_geometry(n, b, s, k);
    const int L = tk.L, Lp = tk.Lp, r0 = tk.r0, c0 = tk.c0;
    if (L <= 0) return;
    const int tid = threadIdx.x;
    const int ldg = b + 1;                       // odd leading dimension: conflict-free row and column walks
    T* G = reinterpret_cast<T*>(smem_raw);       // [Lp][ldg]  (column-major: G[i + j*ldg])
    T* D = G + (size_t)b * ldg;                  // [L][ldg]   full Hermitian
    T* v = D + (size_t)b * ldg;                  // [b] new reflector
    T* vp = v + b;                               // [b] previous reflector
    T* pw = vp + b;                              // [b] p / w of the two-sided update
    T* ps = pw + b;                              // [256] partial sums
    T* ps2 = ps + SBR_THREADS;                   // [256] partial sums
    __shared__ T red[32];
    __shared__ double redd[32];

    // ---- load ----
    // column c of the band storage holds rows c .. c+ldab-1 contiguously
    if (k > 0) {
        for (int idx = tid; idx < L * Lp; idx += SBR_THREADS) {
            const int i = idx % L, j = idx / L;
            G[i + j * ldg] = AB[(size_t)(c0 + j) * ldab + (r0 + i - c0 - j)];
        }
        for (int j = tid; j < Lp; j += SBR_THREADS) vp[j] = V2[(size_t)s * ldv + c0 + j];
    } else {
        for (int i = tid; i < L; i += SBR_THREADS) G[i] = AB[(size_t)s * ldab + (r0 + i - s)];   // column s
    }
    for (int idx = tid; idx < L * L; idx += SBR_THREADS) {
        const int i = idx % L, j = idx / L;
        T x;
        if (i >= j) x = AB[(size_t)(r0 + j) * ldab + (i - j)];
        else x = conj_(AB[(size_t)(r0 + i) * ldab + (j - i)]);
        D[i + j * ldg] = x;
    }
    __syncthreads();

    // Work split: the 256 threads form NQ groups of RG (= 32 or 64 >= b) threads; thread (i, q) owns row
    // (or column) i and the q-th slice of the other index, partial sums meet in shared memory.
    const int RG = (b <= 32) ? 32 : 64, NQ = SBR_THREADS / RG;
    const int i = tid % RG, q = tid / RG;

    // ---- G <- G H_prev (k >= 1) ----
    if (k > 0) {
        const T taup = tau2[(size_t)s * ldt + (k - 1)];
        if (!is_zero(taup)) {   // uniform
            const int jw = (Lp + NQ - 1) / NQ, j0 = q * jw, j1 = min(Lp, j0 + jw);
            T g = zero<T>();
            if (i < L)
                for (int j = j0; j < j1; ++j) fma_(g, G[i + j * ldg], vp[j]);
            ps[q * RG + i] = g;
            __syncthreads();
            if (i < L) {
                T gs = zero<T>();
                for (int qq = 0; qq < NQ; ++qq) gs = add_(gs, ps[qq * RG + i]);
                gs = mul_(taup, gs);
                for (int j = j0; j < j1; ++j) G[i + j * ldg] = sub_(G[i + j * ldg], mul_(gs, conj_(vp[j])));
            }
        }
        __syncthreads();
    }

    // ---- reflector from column 0 of G ----
    double part = 0.0;
    if (tid >= 1 && tid < L) part = abs2_(G[tid]);
    const double sigma = block_sum<double>(part, redd);
    double beta; T tau, scale;
    larfgp_scalars<T>(G[0], sigma, beta, tau, scale);
    __syncthreads();
    if (tid < L) {
        v[tid] = (tid == 0) ? one<T>() : mul_(G[tid], scale);
        G[tid] = (tid == 0) ? mk<T>(beta) : zero<T>();
    }
    __syncthreads();

    if (!is_zero(tau)) {   // uniform
        // ---- G <- H^H G on columns 1 .. Lp-1 (k >= 1): thread (j, q) owns column j, row slice q ----
        const int iw = (L + NQ - 1) / NQ, i0 = q * iw, i1 = min(L, i0 + iw);
        if (k > 0) {
            T dsum = zero<T>();
            if (i >= 1 && i < Lp)
                for (int r = i0; r < i1; ++r) fmac_(dsum, v[r], G[r + i * ldg]);
            ps[q * RG + i] = dsum;
        }
        // ---- p = D v (row i, column slice q) ----
        {
            T p = zero<T>();
            if (i < L)
                for (int j = i0; j < i1; ++j) fma_(p, D[i + j * ldg], v[j]);
            ps2[q * RG + i] = p;
        }
        __syncthreads();
        if (k > 0 && i >= 1 && i < Lp) {
            T dsum = zero<T>();
            for (int qq = 0; qq < NQ; ++qq) dsum = add_(dsum, ps[qq * RG + i]);
            dsum = mul_(conj_(tau), dsum);
            for (int r = i0; r < i1; ++r) G[r + i * ldg] = sub_(G[r + i * ldg], mul_(v[r], dsum));
        }
        T p = zero<T>();
        if (tid < L)
            for (int qq = 0; qq < NQ; ++qq) p = add_(p, ps2[qq * RG + tid]);
        T a = zero<T>();
        if (tid < L) fmac_(a, v[tid], p);
        const T alpha = block_sum<T>(a, red);     // v^H D v (real)
        const double half = 0.5 * abs2_(tau) * real_(alpha);
        if (tid < L) pw[tid] = sub_(mul_(tau, p), scale_(v[tid], half));
        __syncthreads();
        // ---- D <- D - v w^H - w v^H (row i, column slice q) ----
        if (i < L) {
            const T vi = v[i], wi = pw[i];
            for (int j = i0; j < i1; ++j) {
                T x = D[i + j * ldg];
                x = sub_(x, mul_(vi, conj_(pw[j])));
                x = sub_(x, mul_(wi, conj_(v[j])));
                if (j == i) x = mk<T>(real_(x));
                D[i + j * ldg] = x;
            }
        }
    }
    __syncthreads();

    // ---- store ----
    if (k > 0) {
        for (int idx = tid; idx < L * Lp; idx += SBR_THREADS) {
            const int i = idx % L, j = idx / L;
            AB[(size_t)(c0 + j) * ldab + (r0 + i - c0 - j)] = G[i + j * ldg];
        }
    } else {
        for (int i = tid; i < L; i += SBR_THREADS) AB[(size_t)s * ldab + (r0 + i - s)] = G[i];
    }
    for (int idx = tid; idx < L * L; idx += SBR_THREADS) {
        const int i = idx % L, j = idx / L;
        if (i >= j) AB[(size_t)(r0 + j) * ldab + (i - j)] = D[i + j * ldg];
    }
    for (int i = tid; i < L; i += SBR_THREADS) V2[(size_t)s * ldv + r0 + i] = v[i];
    if (tid == 0) tau2[(size_t)s * ldt + k] = tau;
}

template <typename T>
static size_t chase_smem_bytes(int b) { return ((size_t)2 * b * (b + 1) + 3 * (size_t)b + 2 * SBR_THREADS + 8) * sizeof(T); }

template <typename T>
size_t sbr_chase_worksize_t(int n, int b) {
    const size_t nn = (size_t)(n > 0 ? n : 1);
    return align_up((size_t)2 * b * nn * sizeof(T), 256) + align_up((nn + 1) * sizeof(int), 256) + 256;   // band + progress counters
}

// opt-in (MAKB200_CHASE_PERSISTENT=1): one cooperative launch instead of one launch per wavefront
static bool chase_persistent_enabled() {
    const char* e = getenv("MAKB200_CHASE_PERSISTENT");
    return e && e[0] == '1';
}

// Round-2 bring-up: logic validated on the CPU emulator (tests/test_emu_kernels_cpu.py), not yet timed on a B200.
template <typename T>
static int chase_persistent_launch(makb200_handle* h, int n, int b, T* AB, int ldab, T* V2, int ldv, T* tau2, int ldt,
                                   int* prog) {
    cudaStream_t s = h->stream;
    const size_t smem = chase_persistent_smem_elems(b) * sizeof(T);
    MAK_CUDA(h, cudaFuncSetAttribute(chase_persistent_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, nsm = 0, occ = 0;
    MAK_CUDA(h, cudaGetDevice(&dev));
    MAK_CUDA(h, cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    MAK_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, chase_persistent_kernel<T>, SBRP_THREADS, smem));
    if (occ < 1) return MAKB200_ERR_WORKSPACE;
    // sweep s+1 trails sweep s by two tasks: at most ntasks(0)/2 + 1 sweeps are ever in flight
    const int inflight = sbr::sweep_ntasks(n, b, 0) / 2 + 2;
    int grid = nsm * occ;
    if (grid > inflight) grid = inflight;
    if (grid > n - 1) grid = n - 1;
    if (const char* e = getenv("MAKB200_CHASE_GRID")) { const int g = atoi(e); if (g >= 1 && g < grid) grid = g; }
    MAK_CUDA(h, cudaMemsetAsync(prog, 0, sizeof(int) * (size_t)(n + 1), s));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(SBRP_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;   // all CTAs resident or the launch fails: the progress waits cannot deadlock
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t err = cudaLaunchKernelEx(&cfg, chase_persistent_kernel<T>, n, b, AB, ldab, V2, ldv, tau2, ldt, prog);
    if (err != cudaSuccess) return cuda_fail(h, err, "chase_persistent_kernel");
    count_launch();
    return 0;
}

template <typename T>
int sbr_chase_t(makb200_handle* h, int n, int b, const T* A, int lda, double* d, double* e, T* V2, int ldv, T* tau2,
                int ldt, void* work, size_t lwork) {
    if (n <= 0) return 0;
    if (b < 1 || b > SBR_BMAX) return -4;
    cudaStream_t s = h->stream;
    Arena ar(work, lwork);
    const int ldab = 2 * b;
    T* AB = ar.get<T>((size_t)ldab * n);
    int* prog = ar.get<int>((size_t)n + 1);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    static bool configured = false;
    if (!configured) {
        MAK_CUDA(h, cudaFuncSetAttribute(chase_wave_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)chase_smem_bytes<double>(SBR_BMAX)));
        MAK_CUDA(h, cudaFuncSetAttribute(chase_wave_kernel<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)chase_smem_bytes<cplx>(SBR_BMAX)));
        configured = true;
    }
    MAK_CUDA(h, cudaMemsetAsync(V2, 0, sizeof(T) * (size_t)ldv * n, s));
    MAK_CUDA(h, cudaMemsetAsync(tau2, 0, sizeof(T) * (size_t)ldt * n, s));
    band_pack_kernel<T><<<n, 128, 0, s>>>(n, b, A, lda, AB, ldab);
    count_launch();
    MAK_LAUNCH_CHECK(h, "band_pack_kernel");
    const bool persistent = n >= 2 && chase_persistent_enabled();
    if (persistent) {
        const int rc = chase_persistent_launch<T>(h, n, b, AB, ldab, V2, ldv, tau2, ldt, prog);
        if (rc != 0) return rc;
    } else if (n >= 2) {
        const int kmax = (n - 1 + b - 1) / b;                 // tasks of sweep 0
        const int grid = kmax / 2 + 2;
        const int tmax = sbr::wavefront(n - 2, 0);             // last sweep has one task
        // the last wavefront that holds any task: sweeps near the end have one task each, so 2(n-2) is the maximum
        const size_t smem = chase_smem_bytes<T>(b);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(SBR_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        for (int t = 0; t <= tmax; ++t) {
            cudaError_t err = cudaLaunchKernelEx(&cfg, chase_wave_kernel<T>, n, b, AB, ldab, V2, ldv, tau2, ldt, t);
            if (err != cudaSuccess) return cuda_fail(h, err, "chase_wave_kernel");
        }
        count_launch(tmax + 1);
    }
    band_diag_kernel<T><<<(n + 255) / 256, 256, 0, s>>>(n, AB, ldab, d, e, persistent ? prog + n : nullptr);
    count_launch();
    MAK_LAUNCH_CHECK(h, "band_diag_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------
// Q2 application: Z <- Q2 Z with the diamond blocking of sbr_core.h.  Every block (group of g sweeps,
// chase position k) becomes a compact-WY pair (V parallelogram, T); blocks of one diamond wavefront
// are independent and go through three grouped DMMA GEMMs (W = V^H Z_rows, W2 = T W, Z_rows -= V W2).
// ---------------------------------------------------------------------------------------
// Q2BlockDesc: sbr_q2_slab.cuh

// opt-in (MAKB200_Q2_FUSED=1): one CTA per column slab of Z walks every diamond block (sbr_q2_slab.cuh)
static bool q2_fused_enabled() {
    const char* e = getenv("MAKB200_Q2_FUSED");
    return e && e[0] == '1';
}
static int q2_fused_cw() {
    const char* e = getenv("MAKB200_Q2_CW");
    return (e && atoi(e) == 32) ? 32 : 64;
}
static bool q2_fused_ok(int b, int g) { return b % 8 == 0 && g % 8 == 0 && g <= b && g <= 64; }

// one CTA per block: explicit parallelogram V (rows x ns, ld = ldvb) and T (ns x ns upper, ld = g)
template <typename T>
__global__ void __launch_bounds__(128)
q2_build_kernel(int n, int b, int g, int ldvb, const T* __restrict__ V2, int ldv, const T* __restrict__ tau2, int ldt,
                const Q2BlockDesc* __restrict__ descs, T* __restrict__ Vpool, T* __restrict__ Tpool) {
    __shared__ T z[128];
    const Q2BlockDesc d = descs[blockIdx.x];
    T* V = Vpool + d.voff;
    T* Tm = Tpool + d.toff;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < ldvb * d.ns; idx += 128) V[idx] = zero<T>();
    for (int idx = tid; idx < g * g; idx += 128) Tm[idx] = zero<T>();
    __syncthreads();
    for (int j = 0; j < d.ns; ++j) {
        const sbr::Task t = sbr::task_geometry(n, b, d.s0 + j, d.k);
        const T* v = V2 + (size_t)(d.s0 + j) * ldv + t.r0;
        for (int i = tid; i < t.L; i += 128) V[(size_t)j * ldvb + (t.r0 - d.base) + i] = v[i];
    }
    __syncthreads();
    for (int j = 0; j < d.ns; ++j) {
        const T tj = tau2[(size_t)(d.s0 + j) * ldt + d.k];
        if (tid < j) {
            T a = zero<T>();
            for (int r = 0; r < d.rows; ++r) fmac_(a, V[(size_t)tid * ldvb + r], V[(size_t)j * ldvb + r]);
            z[tid] = a;
        }
        __syncthreads();
        if (tid < j) {
            T acc = zero<T>();
            for (int p = tid; p < j; ++p) fma_(acc, Tm[(size_t)p * g + tid], z[p]);
            Tm[(size_t)j * g + tid] = neg_(mul_(tj, acc));
        } else if (tid == j) {
            Tm[(size_t)j * g + j] = tj;
        }
        __syncthreads();
    }
}

template <typename T>
size_t sbr_apply_q2_worksize_t(int n, int b, int g, int ncols) {
    if (n < 2) return 256;
    const int ngroups = (n - 1 + g - 1) / g, kmax = (n - 1 + b - 1) / b;
    size_t nblocks = 0, maxwave = 0;
    std::vector<size_t> wave(ngroups + kmax + 1, 0);
    for (int grp = 0; grp < ngroups; ++grp)
        for (int k = 0; k < kmax; ++k)
            if (sbr::dblock_geometry(n, b, g, grp, k).ns > 0) { ++nblocks; ++wave[sbr::diamond_wavefront(ngroups, grp, k)]; }
    for (size_t w : wave) maxwave = w > maxwave ? w : maxwave;
    size_t bytes = align_up(nblocks * sizeof(Q2BlockDesc), 256) + 3 * align_up(nblocks * sizeof(GemmProblem<T>), 256) +
                   align_up(nblocks * (size_t)(b + g) * g * sizeof(T), 256) + align_up(nblocks * (size_t)g * g * sizeof(T), 256) +
                   2 * align_up(maxwave * (size_t)g * (size_t)(ncols > 0 ? ncols : 1) * sizeof(T), 256) +
                   align_up((size_t)ngroups * kmax * sizeof(int), 256);   // blkmap of the fused slab kernel
    return bytes + 1024;
}

template <typename T>
int sbr_apply_q2_t(makb200_handle* h, int n, int b, int g, const T* V2, int ldv, const T* tau2, int ldt, T* Z, int ldz,
                   int ncols, void* work, size_t lwork) {
    if (n < 2 || ncols <= 0) return 0;
    if (g < 1 || g > 128 || b < 1) return -4;
    cudaStream_t s = h->stream;
    const int ngroups = (n - 1 + g - 1) / g, kmax = (n - 1 + b - 1) / b, ldvb = b + g;
    // blocks ordered by diamond wavefront
    std::vector<std::vector<Q2BlockDesc>> waves(ngroups + kmax + 1);
    size_t nblocks = 0, maxwave = 0;
    for (int grp = 0; grp < ngroups; ++grp)
        for (int k = 0; k < kmax; ++k) {
            const sbr::DBlock d = sbr::dblock_geometry(n, b, g, grp, k);
            if (d.ns <= 0) continue;
            Q2BlockDesc q{d.s0, d.ns, d.base, d.rows, k, 0, 0, 0};
            waves[sbr::diamond_wavefront(ngroups, grp, k)].push_back(q);
            ++nblocks;
        }
    for (auto& w : waves) maxwave = w.size() > maxwave ? w.size() : maxwave;
    Arena ar(work, lwork);
    Q2BlockDesc* ddev = ar.get<Q2BlockDesc>(nblocks);
    GemmProblem<T>* P1 = ar.get<GemmProblem<T>>(nblocks);
    GemmProblem<T>* P2 = ar.get<GemmProblem<T>>(nblocks);
    GemmProblem<T>* P3 = ar.get<GemmProblem<T>>(nblocks);
    T* Vpool = ar.get<T>(nblocks * (size_t)ldvb * g);
    T* Tpool = ar.get<T>(nblocks * (size_t)g * g);
    T* Wb = ar.get<T>(maxwave * (size_t)g * ncols);
    T* W2b = ar.get<T>(maxwave * (size_t)g * ncols);
    int* blkmap = ar.get<int>((size_t)ngroups * kmax);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    std::vector<Q2BlockDesc> descs;
    std::vector<GemmProblem<T>> p1, p2, p3;
    descs.reserve(nblocks); p1.reserve(nblocks); p2.reserve(nblocks); p3.reserve(nblocks);
    std::vector<size_t> wstart;
    for (auto& w : waves) {
        wstart.push_back(descs.size());
        for (size_t iw = 0; iw < w.size(); ++iw) {
            Q2BlockDesc q = w[iw];
            q.voff = descs.size() * (size_t)ldvb * g;
            q.toff = descs.size() * (size_t)g * g;
            descs.push_back(q);
            T* Vb = Vpool + q.voff;
            T* Tb = Tpool + q.toff;
            T* Wi = Wb + iw * (size_t)g * ncols;
            T* W2i = W2b + iw * (size_t)g * ncols;
            T* Zr = Z + q.base;
            GemmProblem<T> p;
            p.lower = 0;
            p.m = q.ns; p.n = ncols; p.k = q.rows;                       // W = V^H Z_rows
            p.A = Vb; p.lda = ldvb; p.B = Zr; p.ldb = ldz; p.C = Wi; p.ldc = g;
            p.alpha = one<T>(); p.beta = zero<T>(); p.conja = 1; p.conjb = 0;
            p1.push_back(p);
            p.k = q.ns;                                                  // W2 = T W
            p.A = Tb; p.lda = g; p.B = Wi; p.ldb = g; p.C = W2i; p.ldc = g; p.conja = 0;
            p2.push_back(p);
            p.m = q.rows; p.k = q.ns;                                    // Z_rows -= V W2
            p.A = Vb; p.lda = ldvb; p.B = W2i; p.ldb = g; p.C = Zr; p.ldc = ldz;
            p.alpha = neg_(one<T>()); p.beta = one<T>();
            p3.push_back(p);
        }
    }
    wstart.push_back(descs.size());
    {
        Stager st(h, nblocks * (sizeof(Q2BlockDesc) + 3 * sizeof(GemmProblem<T>)) + 4096);
        MAK_CUDA(h, st.put(ddev, descs.data(), nblocks * sizeof(Q2BlockDesc), s));
        MAK_CUDA(h, st.put(P1, p1.data(), nblocks * sizeof(GemmProblem<T>), s));
        MAK_CUDA(h, st.put(P2, p2.data(), nblocks * sizeof(GemmProblem<T>), s));
        MAK_CUDA(h, st.put(P3, p3.data(), nblocks * sizeof(GemmProblem<T>), s));
    }
    q2_build_kernel<T><<<(unsigned)nblocks, 128, 0, s>>>(n, b, g, ldvb, V2, ldv, tau2, ldt, ddev, Vpool, Tpool);
    count_launch();
    MAK_LAUNCH_CHECK(h, "q2_build_kernel");
    if (q2_fused_enabled() && q2_fused_ok(b, g)) {
        // Round-2 bring-up: logic validated on the CPU emulator, not yet timed on a B200.
        const int cw = q2_fused_cw();
        const Q2SlabSmem sm = q2_slab_smem(b, g, cw);
        const size_t smem = sm.total * sizeof(T);
        if (smem <= 227 * 1024) {
            std::vector<int> map((size_t)ngroups * kmax, -1);
            for (size_t i = 0; i < descs.size(); ++i) map[(size_t)(descs[i].s0 / g) * kmax + descs[i].k] = (int)i;
            {
                Stager st(h, map.size() * sizeof(int) + 1024);
                MAK_CUDA(h, st.put(blkmap, map.data(), map.size() * sizeof(int), s));
            }
            MAK_CUDA(h, cudaFuncSetAttribute(q2_slab_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            q2_slab_kernel<T><<<(ncols + cw - 1) / cw, Q2S_THREADS, smem, s>>>(n, b, g, cw, ngroups, kmax, blkmap, ddev, Vpool,
                                                                              Tpool, Z, ldz, ncols);
            count_launch();
            MAK_LAUNCH_CHECK(h, "q2_slab_kernel");
            return 0;
        }
    }
    for (size_t u = 0; u + 1 < wstart.size(); ++u) {
        const int cnt = (int)(wstart[u + 1] - wstart[u]);
        if (cnt <= 0) continue;
        const size_t o = wstart[u];
        cudaError_t e = gemm_grouped<T>(s, MAKB200_OP_C, MAKB200_OP_N, cnt, g, ncols, P1 + o);
        if (e == cudaSuccess) e = gemm_grouped<T>(s, MAKB200_OP_N, MAKB200_OP_N, cnt, g, ncols, P2 + o);
        if (e == cudaSuccess) e = gemm_grouped<T>(s, MAKB200_OP_N, MAKB200_OP_N, cnt, b + g, ncols, P3 + o);
        if (e != cudaSuccess) return cuda_fail(h, e, "gemm_grouped (Q2)");
    }
    return 0;
}

template size_t sbr_apply_q2_worksize_t<double>(int, int, int, int);
template size_t sbr_apply_q2_worksize_t<cplx>(int, int, int, int);
template int sbr_apply_q2_t<double>(makb200_handle*, int, int, int, const double*, int, const double*, int, double*, int,
                                    int, void*, size_t);
template int sbr_apply_q2_t<cplx>(makb200_handle*, int, int, int, const cplx*, int, const cplx*, int, cplx*, int, int,
                                  void*, size_t);
template size_t sbr_chase_worksize_t<double>(int, int);
template size_t sbr_chase_worksize_t<cplx>(int, int);
template int sbr_chase_t<double>(makb200_handle*, int, int, const double*, int, double*, double*, double*, int, double*,
                                 int, void*, size_t);
template int sbr_chase_t<cplx>(makb200_handle*, int, int, const cplx*, int, double*, double*, cplx*, int, cplx*, int,
                               void*, size_t);

}  // namespace mak
