// One-CTA-per-block batched QR for the many small blocks tensor-network codes emit.
// The whole block lives in shared memory: A is read from HBM once, Q and R are written once
// (algorithmic bytes = sz*(mn + mk + kn)), so small blocks run at HBM speed.
// Same reflector convention as the large path (beta >= 0  =>  gauge-fixed Q, R).
#include "batched.cuh"
#include "batched_qr_warp.cuh"

namespace mak {

constexpr int BQ_THREADS = 256;
constexpr size_t BQ_SMEM_BYTES = 200 * 1024;

// shared-memory elements a block needs: padded tile + tau
size_t batched_qr_smem_elems(int m, int n) { return (size_t)(m | 1) * n + (m < n ? m : n) + 1 + (size_t)n; }
template <typename T> size_t batched_qr_max_smem_elems() { return BQ_SMEM_BYTES / sizeof(T) - 64; }
template size_t batched_qr_max_smem_elems<double>();
template size_t batched_qr_max_smem_elems<cplx>();

// NT threads per block (64 / 128 / 256 by size class, so that small blocks get many resident CTAs).
// One block barrier per column: the reflector scalars are formed redundantly by every warp from the
// tail norm that the warp updating column j+1 produced in step j; column j is scaled into v after
// the barrier by one warp while the others already work on step j+1 (it is not read again before
// the Q phase).  In the Q phase the v -> q conversion of column j+1 is done by the warp that owns it
// at the start of step j, which removes the second barrier there as well.
template <typename T, int NT>
__global__ void __launch_bounds__(NT)
batched_qr_kernel(const QrBlockDesc<T>* __restrict__ descs, int* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const QrBlockDesc<T> d = descs[blockIdx.x];
    const int m = d.m, n = d.n, k = m < n ? m : n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    if (m <= 0 || n < 0) { if (tid == 0 && info) info[blockIdx.x] = 0; return; }
    const int lds = m | 1;  // odd leading dimension
    T* S = reinterpret_cast<T*>(smem_raw);        // lds x n
    T* tau = S + (size_t)lds * n;                 // k
    double* sig = reinterpret_cast<double*>(tau + (k > 0 ? k : 1));   // n tail norms^2

    for (int c = warp; c < n; c += NW) {
        const T* src = d.A + (size_t)c * d.lda;
        T* dst = S + (size_t)c * lds;
        double part = 0.0;
        for (int r = lane; r < m; r += 32) {
            const T v = src[r];
            dst[r] = v;
            if (r > 0) part += abs2_(v);
        }
        if (c == 0) { part = warp_sum(part); if (lane == 0) sig[0] = part; }
    }
    __syncthreads();

    // ---- factorization: S -> V \ R ----
    for (int j = 0; j < k; ++j) {
        T* cj = S + (size_t)j * lds;
        double beta; T tj, scale;
        larfgp_scalars<T>(cj[j], sig[j], beta, tj, scale);
        const T ctau = conj_(tj);
        for (int l = j + 1 + warp; l < n; l += NW) {
            T* cl = S + (size_t)l * lds;
            const T clj = cl[j];
            T s0 = zero<T>(), s1 = zero<T>();
            int r = j + 1 + lane;
            for (; r + 32 < m; r += 64) {
                fmac_(s0, mul_(cj[r], scale), cl[r]);
                fmac_(s1, mul_(cj[r + 32], scale), cl[r + 32]);
            }
            if (r < m) fmac_(s0, mul_(cj[r], scale), cl[r]);
            const T s = warp_sum(add_(s0, s1));
            const T f = mul_(ctau, add_(clj, s));
            double nrm = 0.0;
            for (r = j + 1 + lane; r < m; r += 32) {
                const T x = sub_(cl[r], mul_(f, mul_(cj[r], scale)));
                cl[r] = x;
                if (r > j + 1) nrm += abs2_(x);
            }
            if (l == j + 1) { nrm = warp_sum(nrm); if (lane == 0) sig[j + 1] = nrm; }
            if (lane == 0) cl[j] = sub_(clj, f);
        }
        __syncthreads();
        if (warp == j % NW) {
            for (int r = j + 1 + lane; r < m; r += 32) cj[r] = mul_(cj[r], scale);
            if (lane == 0) { cj[j] = mk<T>(beta); tau[j] = tj; }
        }
    }
    __syncthreads();

    // ---- R out (upper triangle, zeros below) ----
    if (d.R) {
        for (int c = warp; c < n; c += NW)
            for (int r = lane; r < k; r += 32) d.R[(size_t)c * d.ldr + r] = (r <= c) ? S[(size_t)c * lds + r] : zero<T>();
    }
    __syncthreads();

    // ---- form Q (m x k) in place over V, backward accumulation (org2r) ----
    auto to_q = [&](int c) {   // column c: v-form -> H_c e_c = e_c - tau_c [0; 1; v]   (one warp)
        T* cc = S + (size_t)c * lds;
        const T tc = tau[c];
        for (int r = lane; r < m; r += 32) {
            T x;
            if (r < c) x = zero<T>();
            else if (r == c) x = sub_(one<T>(), tc);
            else x = neg_(mul_(tc, cc[r]));
            cc[r] = x;
        }
        __syncwarp();
    };
    for (int j = k - 2; j >= 0; --j) {
        const T* cj = S + (size_t)j * lds;
        const T tj = tau[j];
        if (warp == 0) to_q(j + 1);
        for (int l = j + 1 + warp; l < k; l += NW) {
            T* cl = S + (size_t)l * lds;
            const T clj = cl[j];
            T s0 = zero<T>(), s1 = zero<T>();
            int r = j + 1 + lane;
            for (; r + 32 < m; r += 64) { fmac_(s0, cj[r], cl[r]); fmac_(s1, cj[r + 32], cl[r + 32]); }
            if (r < m) fmac_(s0, cj[r], cl[r]);
            const T s = warp_sum(add_(s0, s1));
            const T f = mul_(tj, add_(clj, s));
            for (r = j + 1 + lane; r < m; r += 32) cl[r] = sub_(cl[r], mul_(f, cj[r]));
            if (lane == 0) cl[j] = sub_(clj, f);
        }
        __syncthreads();
    }
    if (warp == 0 && k > 0) to_q(0);
    __syncthreads();
    for (int c = warp; c < k; c += NW)
        for (int r = lane; r < m; r += 32) d.Q[(size_t)c * d.ldq + r] = S[(size_t)c * lds + r];
    if (tid == 0 && info) info[blockIdx.x] = 0;
}


static bool bqr_warp_reg() {
    static const bool v = []() { const char* e = getenv("MAKB200_BQR_WARP_REG"); return e && e[0] == '1'; }();
    return v;
}

static bool bqr_warp_blk() {
    const char* e = getenv("MAKB200_BQR_WARP_BLK");   // read per call: the bring-up tests toggle it
    return e && e[0] == '1';
}

template <typename T>
int batched_qr_warp(makb200_handle* h, int batch, int cap_elems, const QrBlockDesc<T>* descs, int rmax) {
    if (batch <= 0) return 0;
    if (bqr_warp_blk()) {
        // panel-blocked variant (round-2 bring-up): 5 instead of 14 shared-memory wavefronts per (row, step)
        const int grid = (batch + 3) / 4;
        const size_t smem = 4 * (size_t)(cap_elems + BQW_TF_ELEMS) * sizeof(T);
        batched_qr_warp_blk_kernel<T><<<grid, 128, smem, h->stream>>>(descs, batch, cap_elems);
        count_launch();
        MAK_LAUNCH_CHECK(h, "batched_qr_warp_blk_kernel");
        return 0;
    }
    if (bqr_warp_reg()) {
        const int grid = (batch + 3) / 4;
        if (rmax <= 16) batched_qr_warp_reg_kernel<T, 16><<<grid, 128, 0, h->stream>>>(descs, batch);
        else if (rmax <= 24) batched_qr_warp_reg_kernel<T, 24><<<grid, 128, 0, h->stream>>>(descs, batch);
        else batched_qr_warp_reg_kernel<T, 32><<<grid, 128, 0, h->stream>>>(descs, batch);
        count_launch();
        MAK_LAUNCH_CHECK(h, "batched_qr_warp_reg_kernel");
        return 0;
    }
    const int grid = (batch + 3) / 4;
    const size_t smem = 4 * (size_t)cap_elems * sizeof(T);
    batched_qr_warp_kernel<T><<<grid, 128, smem, h->stream>>>(descs, batch, cap_elems);
    count_launch();
    MAK_LAUNCH_CHECK(h, "batched_qr_warp_kernel");
    return 0;
}
template int batched_qr_warp<double>(makb200_handle*, int, int, const QrBlockDesc<double>*, int);
template int batched_qr_warp<cplx>(makb200_handle*, int, int, const QrBlockDesc<cplx>*, int);

// ---------------------------------------------------------------------------------------
// batched svd_compact!: one CTA per block, one-sided (Hestenes) Jacobi with a round-robin
// parallel ordering; G (= A or A^H, tall orientation) and the accumulated V live in shared
// memory.  Epilogue: sort descending, normalise, reference SVD gauge (common/gauge.jl:69-77).
// ---------------------------------------------------------------------------------------
constexpr int BS_THREADS = 256;

template <typename T, int NT>
__global__ void __launch_bounds__(NT)
batched_svd_kernel(const SvdBlockDesc<T>* __restrict__ descs, int* __restrict__ info, int max_sweeps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SvdBlockDesc<T> d = descs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = NT / 32;
    const bool tr = d.m < d.n;                 // work on A^H when wide
    const int mm = tr ? d.n : d.m, nn = tr ? d.m : d.n;   // G is mm x nn, mm >= nn
    if (nn <= 0) { if (tid == 0 && info) info[blockIdx.x] = 0; return; }
    const int ldg = mm | 1, ldv = nn | 1;
    T* G = reinterpret_cast<T*>(smem_raw);     // ldg x nn
    T* V = G + (size_t)ldg * nn;               // ldv x nn
    double* sig = reinterpret_cast<double*>(V + (size_t)ldv * nn);  // nn
    int* perm = reinterpret_cast<int*>(sig + nn);                   // nn
    __shared__ int s_rot;

    for (int idx = tid; idx < mm * nn; idx += NT) {
        int c = idx / mm, r = idx - c * mm;
        G[(size_t)c * ldg + r] = tr ? conj_(d.A[(size_t)r * d.lda + c]) : d.A[(size_t)c * d.lda + r];
    }
    for (int idx = tid; idx < nn * nn; idx += NT) {
        int c = idx / nn, r = idx - c * nn;
        V[(size_t)c * ldv + r] = (r == c) ? one<T>() : zero<T>();
    }
    __syncthreads();

    const int ne = (nn + 1) & ~1;  // players (padded to even)
    const double tol = 4.0 * 2.220446049250313e-16;
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        if (tid == 0) s_rot = 0;
        __syncthreads();
        for (int step = 0; step < ne - 1; ++step) {
            for (int k = warp; k < ne / 2; k += NW) {
                int p, q;
                if (k == 0) { p = ne - 1; q = step; }
                else { p = (step + k) % (ne - 1); q = (step - k + (ne - 1)) % (ne - 1); }
                if (p > q) { int t = p; p = q; q = t; }
                if (q >= nn) continue;  // padding player
                T* x = G + (size_t)p * ldg;
                T* y = G + (size_t)q * ldg;
                double al = 0.0, be = 0.0;
                T ga = zero<T>();
                for (int r = lane; r < mm; r += 32) {
                    T xv = x[r], yv = y[r];
                    al += abs2_(xv); be += abs2_(yv);
                    fmac_(ga, xv, yv);  // x^H y
                }
                al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
                // |x^H y| against ||x|| ||y||: no product of squared norms and no squared inner product, so blocks scaled
                // by 1e-100 or 1e+100 rotate exactly as the same block at unit scale (al * be and |ga|^2 are 4th-order
                // quantities: they left the double range beyond 1e+-77 and the sweep either never rotated or never ended)
                double ag = abs_(ga);
                if (ag > tol * (sqrt(al) * sqrt(be)) && ag > 0.0) {
                    if (lane == 0) s_rot = 1;
                    double zeta = (be - al) / (2.0 * ag);
                    double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                    T ph = scale_(conj_(ga), 1.0 / ag);  // e^{-i phi}
                    T sph = scale_(ph, s), cph = scale_(ph, c);
                    for (int r = lane; r < mm; r += 32) {
                        T xv = x[r], yv = y[r];
                        x[r] = sub_(scale_(xv, c), mul_(sph, yv));
                        y[r] = add_(scale_(xv, s), mul_(cph, yv));
                    }
                    T* vx = V + (size_t)p * ldv;
                    T* vy = V + (size_t)q * ldv;
                    for (int r = lane; r < nn; r += 32) {
                        T xv = vx[r], yv = vy[r];
                        vx[r] = sub_(scale_(xv, c), mul_(sph, yv));
                        vy[r] = add_(scale_(xv, s), mul_(cph, yv));
                    }
                }
            }
            __syncthreads();
        }
        if (s_rot == 0) break;
        __syncthreads();
    }
    // singular values and descending order (stable rank by counting)
    for (int j = warp; j < nn; j += NW) {
        double a = 0.0;
        for (int r = lane; r < mm; r += 32) a += abs2_(G[(size_t)j * ldg + r]);
        a = warp_sum(a);
        if (lane == 0) sig[j] = sqrt(a);
    }
    __syncthreads();
    for (int j = tid; j < nn; j += NT) {
        int rank = 0;
        double sj = sig[j];
        for (int i = 0; i < nn; ++i) rank += (sig[i] > sj) || (sig[i] == sj && i < j);
        perm[rank] = j;
    }
    __syncthreads();
    // normalise the columns of G in place (left vectors of the tall problem); exactly-zero columns stay zero
    for (int j = warp; j < nn; j += NW) {
        const double sj = sig[j], invj = sj > 0.0 ? 1.0 / sj : 0.0;
        for (int r = lane; r < mm; r += 32) G[(size_t)j * ldg + r] = scale_(G[(size_t)j * ldg + r], invj);
    }
    __syncthreads();
    // Collapsed columns (sigma == 0: a zero block, a zero column, or a zero row of a wide block) would leave U
    // (tall) / Vh (wide) non-isometric, while the reference contract (LAPACK gesdd) is an isometry for ANY input.
    // Complete them: the unit vector e_k with the least mass in the span so far, projected off that span
    // (w = e_k - sum_c u_c conj(u_c[k]) needs no dot products) and normalised; ||w||^2 >= 1 - rank/mm.
    {
        int nz = 0;
        while (nz < nn && sig[perm[nn - 1 - nz]] == 0.0) ++nz;    // sorted: the zeros are last
        double* mass = reinterpret_cast<double*>(perm + nn + (nn & 1));   // mm doubles behind perm
        __shared__ double s_best[NT / 32];
        __shared__ int s_bidx[NT / 32];
        __shared__ double s_red[NT / 32];
        if (nz > 0) {
            for (int r = tid; r < mm; r += NT) {
                double a = 0.0;
                for (int cc = 0; cc < nn; ++cc) a += abs2_(G[(size_t)cc * ldg + r]);
                mass[r] = a;
            }
            __syncthreads();
            for (int t = 0; t < nz; ++t) {
                const int z = perm[nn - nz + t];
                // k = first row of least mass
                double best = 1e300; int bi = 0x7fffffff;
                for (int r = tid; r < mm; r += NT)
                    if (mass[r] < best) { best = mass[r]; bi = r; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    double ob = __shfl_xor_sync(0xffffffffu, best, o);
                    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                if (lane == 0) { s_best[warp] = best; s_bidx[warp] = bi; }
                __syncthreads();
                for (int w = 0; w < NW; ++w)
                    if (s_best[w] < best || (s_best[w] == best && s_bidx[w] < bi)) { best = s_best[w]; bi = s_bidx[w]; }
                const int k = bi;
                double nrm = 0.0;
                for (int r = tid; r < mm; r += NT) {
                    T wv = (r == k) ? one<T>() : zero<T>();
                    for (int cc = 0; cc < nn; ++cc) {
                        if (cc == z) continue;
                        const T uk = G[(size_t)cc * ldg + k];
                        wv = sub_(wv, mul_(G[(size_t)cc * ldg + r], conj_(uk)));
                    }
                    G[(size_t)z * ldg + r] = wv;     // column z is read by nobody in this loop (cc != z)
                    nrm += abs2_(wv);
                }
                nrm = warp_sum(nrm);
                if (lane == 0) s_red[warp] = nrm;
                __syncthreads();
                double tot = 0.0;
                for (int w = 0; w < NW; ++w) tot += s_red[w];
                const double invn = tot > 0.0 ? 1.0 / sqrt(tot) : 0.0;
                for (int r = tid; r < mm; r += NT) {
                    const T u = scale_(G[(size_t)z * ldg + r], invn);
                    G[(size_t)z * ldg + r] = u;
                    mass[r] += abs2_(u);
                }
                __syncthreads();
            }
        }
    }
    // outputs: one warp per singular triplet
    for (int j = warp; j < nn; j += NW) {
        const int src = perm[j];
        const double sj = sig[src];
        const double inv = 1.0;               // G is normalised in place above
        const T* g = G + (size_t)src * ldg;   // left vector of the tall problem
        const T* v = V + (size_t)src * ldv;   // right vector of the tall problem
        // U column of the ORIGINAL problem: tall -> g/sigma (length m); wide -> v (length m)
        const int ulen = d.m;
        // gauge: first entry of maximal modulus of the U column
        double best = -1.0; int bi = 0x7fffffff;
        for (int r = lane; r < ulen; r += 32) {
            T uv = tr ? v[r] : scale_(g[r], inv);
            double a = abs2_(uv);
            if (a > best) { best = a; bi = r; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        T ph = one<T>();
        if (d.fixgauge && best > 0.0) {
            T piv = tr ? v[bi] : scale_(g[bi], inv);
            ph = scale_(piv, 1.0 / sqrt(abs2_(piv)));  // sign(pivot)
        }
        const T cph = conj_(ph);
        if (lane == 0) d.S[j] = sj;
        if (d.U) {
            for (int r = lane; r < d.m; r += 32) {
                T uv = tr ? v[r] : scale_(g[r], inv);
                d.U[(size_t)j * d.ldu + r] = mul_(uv, cph);
            }
            for (int c = lane; c < d.n; c += 32) {
                // Vh[j, c] = conj(right vector of the original problem)[c] * sign
                T rv = tr ? scale_(g[c], inv) : v[c];
                d.Vh[(size_t)c * d.ldvh + j] = mul_(conj_(rv), ph);
            }
        }
    }
    if (tid == 0 && info) info[blockIdx.x] = (sweep >= max_sweeps) ? 1 : 0;
}

size_t batched_svd_smem_bytes(int m, int n, size_t elem) {
    int mm = m < n ? n : m, nn = m < n ? m : n;
    // G, V, sig[nn], perm[nn] (+ pad), mass[mm] (completion of collapsed columns)
    return ((size_t)(mm | 1) * nn + (size_t)(nn | 1) * nn) * elem + (size_t)nn * 12 + (size_t)mm * 8 + 64;
}
size_t batched_svd_max_smem_bytes() { return BQ_SMEM_BYTES; }

template <typename T>
int batched_svd_smem(makb200_handle* h, int batch, size_t max_smem_bytes, const SvdBlockDesc<T>* descs, int* info) {
    if (batch <= 0) return 0;
    if (max_smem_bytes > BQ_SMEM_BYTES) return MAKB200_ERR_WORKSPACE;
    if (max_smem_bytes <= 40 * 1024) batched_svd_kernel<T, 128><<<batch, 128, max_smem_bytes, h->stream>>>(descs, info, 40);
    else batched_svd_kernel<T, 256><<<batch, 256, max_smem_bytes, h->stream>>>(descs, info, 40);
    count_launch();
    MAK_LAUNCH_CHECK(h, "batched_svd_kernel");
    return 0;
}
template int batched_svd_smem<double>(makb200_handle*, int, size_t, const SvdBlockDesc<double>*, int*);
template int batched_svd_smem<cplx>(makb200_handle*, int, size_t, const SvdBlockDesc<cplx>*, int*);

// ---------------------------------------------------------------------------------------
// batched eigh_full!: one CTA per block, two-sided (classical) Jacobi with the same round-robin
// parallel ordering; A (mirrored from its upper triangle, LAPACK uplo='U' semantics of
// eigh.jl:150-156) and the accumulated V live in shared memory.  Per round: rotation parameters of
// all disjoint pairs from the current (a_pp, a_qq, a_pq), column rotations of A and V, barrier, row
// rotations of A, barrier.  Epilogue: ascending sort, reference eigh gauge (common/gauge.jl:38-45).
// ---------------------------------------------------------------------------------------
template <typename T, int NT>
__global__ void __launch_bounds__(NT)
batched_eigh_kernel(const EighBlockDesc<T>* __restrict__ descs, int* __restrict__ info, int max_sweeps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const EighBlockDesc<T> d = descs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = NT / 32;
    const int n = d.n;
    if (n <= 0) return;
    const int ld = n | 1;
    T* S = reinterpret_cast<T*>(smem_raw);      // ld x n
    T* V = S + (size_t)ld * n;                  // ld x n
    const int hp = (n + 1) / 2 + 1;
    T* rph = V + (size_t)ld * n;                // [hp] e^{-i phi} per pair
    double* lam = reinterpret_cast<double*>(rph + hp);   // n
    double* rc = lam + n;                       // [hp] c
    double* rs = rc + hp;                       // [hp] s
    int* perm = reinterpret_cast<int*>(rs + hp);
    __shared__ double s_red[32];
    __shared__ int s_rot, s_big;

    double part = 0.0;
    for (int idx = tid; idx < n * n; idx += NT) {
        int c = idx / n, r = idx - c * n;
        T v;
        if (r < c) v = d.A[(size_t)c * d.lda + r];
        else if (r > c) v = conj_(d.A[(size_t)r * d.lda + c]);
        else v = mk<T>(real_(d.A[(size_t)c * d.lda + c]));
        S[(size_t)c * ld + r] = v;
        V[(size_t)c * ld + r] = (r == c) ? one<T>() : zero<T>();
        part += abs2_(v);
    }
    const double fro = sqrt(block_sum<double>(part, s_red));
    __syncthreads();
    const double EPS = 2.220446049250313e-16;
    const double thr = 0.5 * EPS * fro, thr_big = 4.0 * EPS * fro;

    const int ne = (n + 1) & ~1;
    int sweep = 0;
    bool converged = (fro == 0.0);
    for (; sweep < max_sweeps && !converged; ++sweep) {
        if (tid == 0) { s_rot = 0; s_big = 0; }
        __syncthreads();
        for (int step = 0; step < ne - 1; ++step) {
            // ---- rotation parameters + column rotations ----
            for (int k = warp; k < ne / 2; k += NW) {
                int p, q;
                if (k == 0) { p = ne - 1; q = step; }
                else { p = (step + k) % (ne - 1); q = (step - k + (ne - 1)) % (ne - 1); }
                if (p > q) { int t = p; p = q; q = t; }
                double c = 1.0, s = 0.0;
                T ph = one<T>();
                if (q < n) {
                    const T g = S[(size_t)q * ld + p];           // a_pq
                    const double ag = abs_(g);
                    if (ag > thr) {
                        const double al = real_(S[(size_t)p * ld + p]), be = real_(S[(size_t)q * ld + q]);
                        const double zeta = (be - al) / (2.0 * ag);
                        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        c = 1.0 / sqrt(1.0 + t * t); s = c * t;
                        ph = scale_(conj_(g), 1.0 / ag);         // e^{-i phi}
                        if (lane == 0) { s_rot = 1; if (ag > thr_big) s_big = 1; }
                    }
                }
                __syncwarp();   // every lane has read the pivot entries before any lane rotates them
                if (lane == 0) { rc[k] = c; rs[k] = s; rph[k] = ph; }
                if (s != 0.0) {
                    const T sph = scale_(ph, s), cph = scale_(ph, c);
                    T* x = S + (size_t)p * ld; T* y = S + (size_t)q * ld;
                    T* vx = V + (size_t)p * ld; T* vy = V + (size_t)q * ld;
                    for (int r = lane; r < n; r += 32) {
                        T xv = x[r], yv = y[r];
                        x[r] = sub_(scale_(xv, c), mul_(sph, yv));
                        y[r] = add_(scale_(xv, s), mul_(cph, yv));
                        xv = vx[r]; yv = vy[r];
                        vx[r] = sub_(scale_(xv, c), mul_(sph, yv));
                        vy[r] = add_(scale_(xv, s), mul_(cph, yv));
                    }
                }
            }
            __syncthreads();
            // ---- row rotations: A[p,:] = c A[p,:] - s e^{i phi} A[q,:],  A[q,:] = s A[p,:] + c e^{i phi} A[q,:]
            for (int k = warp; k < ne / 2; k += NW) {
                const double s = rs[k];
                if (s == 0.0) continue;
                int p, q;
                if (k == 0) { p = ne - 1; q = step; }
                else { p = (step + k) % (ne - 1); q = (step - k + (ne - 1)) % (ne - 1); }
                if (p > q) { int t = p; p = q; q = t; }
                const double c = rc[k];
                const T phc = conj_(rph[k]);
                const T sph = scale_(phc, s), cph = scale_(phc, c);
                for (int col = lane; col < n; col += 32) {
                    T* cp = S + (size_t)col * ld;
                    T xv = cp[p], yv = cp[q];
                    cp[p] = sub_(scale_(xv, c), mul_(sph, yv));
                    cp[q] = add_(scale_(xv, s), mul_(cph, yv));
                }
            }
            __syncthreads();
            // the pivots are now (numerically) zero and the diagonal is real: make both exact
            for (int k = tid; k < ne / 2; k += NT) {
                if (rs[k] == 0.0) continue;
                int p, q;
                if (k == 0) { p = ne - 1; q = step; }
                else { p = (step + k) % (ne - 1); q = (step - k + (ne - 1)) % (ne - 1); }
                if (p > q) { int t = p; p = q; q = t; }
                S[(size_t)q * ld + p] = zero<T>();
                S[(size_t)p * ld + q] = zero<T>();
                S[(size_t)p * ld + p] = mk<T>(real_(S[(size_t)p * ld + p]));
                S[(size_t)q * ld + q] = mk<T>(real_(S[(size_t)q * ld + q]));
            }
            __syncthreads();
        }
        // converged when a sweep made no rotation, or (after a few sweeps) only noise-level ones
        if (s_rot == 0 || (sweep >= 5 && s_big == 0)) converged = true;
        __syncthreads();
    }
    for (int j = tid; j < n; j += NT) lam[j] = real_(S[(size_t)j * ld + j]);
    __syncthreads();
    // The accumulated product of ~sweeps*n^2/2 rotations drifts from unitarity by ~eps*sqrt(sweeps*n)
    // per column; one Newton-Schulz step V <- V (3I - V^H V)/2 restores it to rounding level.
    // S is free now: S <- (I - V^H V)/2.
    if (d.V) {
        for (int idx = tid; idx < n * n; idx += NT) {
            const int i = idx % n, j = idx / n;
            if (i > j) continue;
            const T* vi = V + (size_t)i * ld;
            const T* vj = V + (size_t)j * ld;
            T e = zero<T>();
            for (int r = 0; r < n; ++r) fmac_(e, vi[r], vj[r]);
            e = scale_(e, -0.5);
            if (i == j) e = mk<T>(0.5 + real_(e));
            S[(size_t)j * ld + i] = e;
            if (i != j) S[(size_t)i * ld + j] = conj_(e);
        }
        __syncthreads();
        constexpr int KQ = 4;   // n <= 128 columns per row = 4 per lane
        for (int r = warp; r < n; r += NW) {
            T v[KQ], acc[KQ];
#pragma unroll
            for (int qd = 0; qd < KQ; ++qd) {
                const int c = lane + 32 * qd;
                v[qd] = (c < n) ? V[(size_t)c * ld + r] : zero<T>();
                acc[qd] = v[qd];
            }
            for (int k = 0; k < n; ++k) {
                T vk = zero<T>();
#pragma unroll
                for (int qd = 0; qd < KQ; ++qd)
                    if ((k >> 5) == qd) vk = v[qd];
                vk = shfl_(vk, k & 31);
                const T* Dk = S + k;   // row k of D: D[k, c] = S[c*ld + k]
#pragma unroll
                for (int qd = 0; qd < KQ; ++qd) {
                    const int c = lane + 32 * qd;
                    if (c < n) fma_(acc[qd], vk, Dk[(size_t)c * ld]);
                }
            }
#pragma unroll
            for (int qd = 0; qd < KQ; ++qd) {
                const int c = lane + 32 * qd;
                if (c < n) V[(size_t)c * ld + r] = acc[qd];
            }
        }
        __syncthreads();
    }
    for (int j = tid; j < n; j += NT) {
        int rank = 0;
        const double lj = lam[j];
        for (int i = 0; i < n; ++i) rank += (lam[i] < lj) || (lam[i] == lj && i < j);
        perm[rank] = j;
    }
    __syncthreads();
    for (int j = warp; j < n; j += NW) {
        const int src = perm[j];
        const T* v = V + (size_t)src * ld;
        if (lane == 0) d.W[j] = lam[src];
        if (!d.V) continue;
        double best = -1.0; int bi = 0x7fffffff;
        for (int r = lane; r < n; r += 32) {
            double a = abs2_(v[r]);
            if (a > best) { best = a; bi = r; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        T cph = one<T>();
        if (d.fixgauge && best > 0.0) {
            T piv = v[bi];
            cph = conj_(scale_(piv, 1.0 / sqrt(abs2_(piv))));
        }
        for (int r = lane; r < n; r += 32) d.V[(size_t)j * d.ldv + r] = mul_(v[r], cph);
    }
    if (tid == 0 && info) info[blockIdx.x] = converged ? 0 : 1;
}

size_t batched_eigh_smem_bytes(int n, size_t elem) {
    size_t ld = (size_t)(n | 1), hp = (size_t)(n + 1) / 2 + 1;
    return 2 * ld * n * elem + hp * elem + (size_t)n * 8 + 2 * hp * 8 + (size_t)n * 4 + 64;
}
size_t batched_eigh_max_smem_bytes() { return BQ_SMEM_BYTES; }

template <typename T>
int batched_eigh_smem(makb200_handle* h, int batch, size_t max_smem_bytes, const EighBlockDesc<T>* descs, int* info) {
    if (batch <= 0) return 0;
    if (max_smem_bytes > BQ_SMEM_BYTES) return MAKB200_ERR_WORKSPACE;
    if (max_smem_bytes <= 40 * 1024) batched_eigh_kernel<T, 128><<<batch, 128, max_smem_bytes, h->stream>>>(descs, info, 30);
    else batched_eigh_kernel<T, 256><<<batch, 256, max_smem_bytes, h->stream>>>(descs, info, 30);
    count_launch();
    MAK_LAUNCH_CHECK(h, "batched_eigh_kernel");
    return 0;
}
template int batched_eigh_smem<double>(makb200_handle*, int, size_t, const EighBlockDesc<double>*, int*);
template int batched_eigh_smem<cplx>(makb200_handle*, int, size_t, const EighBlockDesc<cplx>*, int*);

int batched_init(makb200_handle* h) {
    MAK_CUDA(h, cudaFuncSetAttribute(batched_qr_warp_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_qr_warp_kernel<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_qr_warp_blk_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_qr_warp_blk_kernel<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_qr_kernel<double, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_qr_kernel<double, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_qr_kernel<double, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_qr_kernel<cplx, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_qr_kernel<cplx, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_qr_kernel<cplx, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_svd_kernel<double, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_svd_kernel<double, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_svd_kernel<cplx, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_svd_kernel<cplx, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_eigh_kernel<double, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_eigh_kernel<double, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_eigh_kernel<cplx, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_eigh_kernel<cplx, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    return 0;
}

template <typename T>
int batched_qr_smem(makb200_handle* h, int batch, size_t max_smem_elems, const QrBlockDesc<T>* descs, int* info) {
    if (batch <= 0) return 0;
    size_t smem = (max_smem_elems + 64) * sizeof(T);
    if (smem > BQ_SMEM_BYTES) return MAKB200_ERR_WORKSPACE;
    // threads by size class: small blocks -> small CTAs, many of them resident per SM
    if (smem <= 24 * 1024) batched_qr_kernel<T, 64><<<batch, 64, smem, h->stream>>>(descs, info);
    else if (smem <= 72 * 1024) batched_qr_kernel<T, 128><<<batch, 128, smem, h->stream>>>(descs, info);
    else batched_qr_kernel<T, 256><<<batch, 256, smem, h->stream>>>(descs, info);
    count_launch();
    MAK_LAUNCH_CHECK(h, "batched_qr_kernel");
    return 0;
}
template int batched_qr_smem<double>(makb200_handle*, int, size_t, const QrBlockDesc<double>*, int*);
template int batched_qr_smem<cplx>(makb200_handle*, int, size_t, const QrBlockDesc<cplx>*, int*);

}  // namespace mak
