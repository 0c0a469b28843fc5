// Tridiagonal divide-and-conquer eigensolver on the device (kernels around stedc_core.h).
// All merges of one tree level run in the same launches (grid.y = merge index); the
// eigenvector updates of a level are one grouped DMMA GEMM whose dimensions (the non-deflated
// counts) are produced on the device, so the whole solver runs without a host round trip.
#include "stedc.cuh"
#include "stedc_core.h"
#include "stedc_batch_tables.h"
#include "gemm.cuh"
#include <vector>

namespace mak {
using namespace dc;

struct DcBuffers {
    Ctx ctx;
    Merge* merges;                 // per level slice
    GemmProblem<double>* gemms;    // 2 per merge
    double* E;
    double* rho_cut;               // per cut position
    double* sgn_cut;
    double* scale;                 // [1]
    int* info;                     // [1]
};

__global__ void dc_scale_kernel(int n, const double* __restrict__ d, const double* __restrict__ e, double* D,
                                double* E, double* scale) {
    __shared__ double red[32];
    double m = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, fabs(d[i]));
    for (int i = threadIdx.x; i < n - 1; i += blockDim.x) m = fmax(m, fabs(e[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t = fmax(t, red[i]);
    double sc = t > 0.0 ? t : 1.0;
    if (threadIdx.x == 0) *scale = sc;
    for (int i = threadIdx.x; i < n; i += blockDim.x) D[i] = d[i] / sc;
    for (int i = threadIdx.x; i < n; i += blockDim.x) E[i] = (i < n - 1) ? e[i] / sc : 0.0;
}

// tear the matrix at every leaf boundary (cuts are >= DC_LEAF/2 apart, so no two touch one entry)
__global__ void dc_tear_kernel(int ncut, const int* __restrict__ cuts, double* D, const double* __restrict__ E,
                               double* rho_cut, double* sgn_cut) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncut) return;
    int c = cuts[i];
    double e = E[c - 1];
    double a = fabs(e);
    rho_cut[c] = a;
    sgn_cut[c] = e < 0.0 ? -1.0 : 1.0;
    D[c - 1] -= a;
    D[c] -= a;
}

// one WARP per leaf: lane 0 runs the implicit-QL scalar recurrence (same arithmetic as
// dc::leaf_ql in stedc_core.h) on d/e held in shared memory and broadcasts each plane rotation;
// lane r applies it to row r of the leaf's eigenvector block, also held in shared memory.
constexpr int LEAF_WARPS = 4;
__global__ void __launch_bounds__(LEAF_WARPS * 32)
dc_leaf_kernel(int nleaf, const int* __restrict__ bnd, double* D, const double* __restrict__ E, double* Z, int ldz,
               int* info) {
    __shared__ double sZ[LEAF_WARPS][DC_LEAF][DC_LEAF + 1];
    __shared__ double sd[LEAF_WARPS][DC_LEAF + 1], se[LEAF_WARPS][DC_LEAF + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int leaf = blockIdx.x * LEAF_WARPS + warp;
    if (leaf >= nleaf) return;
    const int lo = bnd[leaf], n = bnd[leaf + 1] - lo;
    double* d = sd[warp];
    double* e = se[warp];
    double (*z)[DC_LEAF + 1] = sZ[warp];
    if (lane < n) { d[lane] = D[lo + lane]; e[lane] = (lane + 1 < n) ? E[lo + lane] : 0.0; }
    for (int c = 0; c < n; ++c) z[lane][c] = (lane == c) ? 1.0 : 0.0;   // z[row][col]
    __syncwarp();
    int fail = 0;
    for (int l = 0; l < n && !fail; ++l) {
        int iter = 0;
        while (true) {
            // lane 0 decides; everybody follows
            int m = l;
            if (lane == 0) {
                for (m = l; m < n - 1; ++m) {
                    double dd = fabs(d[m]) + fabs(d[m + 1]);
                    if (fabs(e[m]) <= DC_EPS * dd) break;
                }
            }
            m = __shfl_sync(0xffffffffu, m, 0);
            if (m == l) break;
            if (iter++ == 60) { fail = 1; break; }
            double g = 0.0, r = 0.0, s = 1.0, c = 1.0, p = 0.0;
            if (lane == 0) {
                g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                r = hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + copysign(r, g));
            }
            int i, brk = 0;
            for (i = m - 1; i >= l; --i) {
                if (lane == 0) {
                    double f = s * e[i], b = c * e[i];
                    r = hypot(f, g);
                    e[i + 1] = r;
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        brk = 1;
                    } else {
                        s = f / r;
                        c = g / r;
                        g = d[i + 1] - p;
                        r = (d[i] - g) * s + 2.0 * c * b;
                        p = s * r;
                        d[i + 1] = g + p;
                        g = c * r - b;
                    }
                }
                brk = __shfl_sync(0xffffffffu, brk, 0);
                if (brk) break;
                const double sb = __shfl_sync(0xffffffffu, s, 0), cb = __shfl_sync(0xffffffffu, c, 0);
                if (lane < n) {
                    const double f2 = z[lane][i + 1];
                    z[lane][i + 1] = sb * z[lane][i] + cb * f2;
                    z[lane][i] = cb * z[lane][i] - sb * f2;
                }
            }
            if (brk) continue;
            if (lane == 0) { d[l] -= p; e[l] = g; e[m] = 0.0; }
            __syncwarp();
        }
    }
    __syncwarp();
    if (fail && lane == 0) atomicExch(info, 1);
    // ascending order: rank by counting (stable), then write columns in sorted order
    int rank = 0;
    double mine = (lane < n) ? d[lane] : 0.0;
    for (int j = 0; j < n; ++j) rank += (lane < n) && ((d[j] < mine) || (d[j] == mine && j < lane));
    __syncwarp();
    if (lane < n) D[lo + rank] = mine;
    // column `lane` of z goes to column `rank`; lanes write their column (rows contiguous in global)
    if (lane < n) {
        double* dst = Z + (size_t)(lo + rank) * ldz + lo;
        for (int r = 0; r < n; ++r) dst[r] = z[r][lane];
    }
}

__global__ void dc_merge_init_kernel(int nm, Merge* mg, const double* __restrict__ rho_cut,
                                     const double* __restrict__ sgn_cut) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nm) return;
    mg[i].rho = 2.0 * rho_cut[mg[i].mid];
    mg[i].sgn = sgn_cut[mg[i].mid];
    mg[i].K = 0; mg[i].k1 = mg[i].k2 = mg[i].k3 = 0; mg[i].nrot = 0;
}

__global__ void dc_z_rank_kernel(Ctx c, const Merge* __restrict__ mgs, const double* __restrict__ Z, int ldz,
                                 int phase) {
    const Merge mg = mgs[blockIdx.y];
    const int N = mg.hi - mg.lo;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        if (phase == 0) merge_z_item(c, mg, Z, ldz, i);
        else merge_rank_item(c, mg, i);
    }
}

__global__ void dc_deflate_kernel(Ctx c, Merge* mgs, GemmProblem<double>* gp, const double* Pack, int ldp,
                                  const double* S, int lds, double* Tmp, int ldt) {
    if (threadIdx.x != 0) return;
    Merge mg = mgs[blockIdx.x];
    deflate_scan(c, mg);
    mgs[blockIdx.x] = mg;
    const int lo = mg.lo, mid = mg.mid, N1 = mg.mid - mg.lo, N2 = mg.hi - mg.mid, K = mg.K;
    const int k12 = mg.k1 + mg.k2, k23 = mg.k2 + mg.k3;
    GemmProblem<double> p;
    p.alpha = 1.0; p.beta = 0.0; p.conja = 0; p.conjb = 0; p.lower = 0;
    // top rows: Tmp[lo:mid, lo:lo+K] = Pack[lo:mid, lo:lo+k12] * S[lo:lo+k12, lo:lo+K]
    p.m = N1; p.n = K; p.k = k12;
    p.A = Pack + (size_t)lo * ldp + lo; p.lda = ldp;
    p.B = S + (size_t)lo * lds + lo; p.ldb = lds;
    p.C = Tmp + (size_t)lo * ldt + lo; p.ldc = ldt;
    gp[2 * blockIdx.x] = p;
    // bottom rows: Tmp[mid:hi, lo:lo+K] = Pack[mid:hi, lo:lo+k23] * S[lo+k1:lo+K, lo:lo+K]
    p.m = N2; p.n = K; p.k = k23;
    p.A = Pack + (size_t)lo * ldp + mid;
    p.B = S + (size_t)lo * lds + lo + mg.k1;
    p.C = Tmp + (size_t)lo * ldt + mid;
    gp[2 * blockIdx.x + 1] = p;
}

__global__ void dc_rotate_kernel(Ctx c, const Merge* __restrict__ mgs, double* Z, int ldz) {
    const Merge mg = mgs[blockIdx.y];
    if (mg.nrot == 0) return;
    const int N = mg.hi - mg.lo;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x)
        rotate_row_item(c, mg, Z, ldz, r);
}

__global__ void dc_secular_kernel(Ctx c, const Merge* __restrict__ mgs) {
    const Merge mg = mgs[blockIdx.y];
    const int lo = mg.lo, K = mg.K;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < K; j += gridDim.x * blockDim.x)
        secular_root(K, j, c.dl + lo, c.zl + lo, mg.rho, c.tau + lo + j, c.orig + lo + j);
}

__global__ void dc_zhat_pos_kernel(Ctx c, const Merge* __restrict__ mgs) {
    const Merge mg = mgs[blockIdx.y];
    const int lo = mg.lo, K = mg.K, N = mg.hi - mg.lo;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        if (j < K) zhat_item(K, c.dl + lo, c.zl + lo, mg.rho, c.tau + lo, c.orig + lo, c.zhat + lo, j);
        final_pos_item(c, mg, j);
    }
}

// S column j (one CTA per column): S[rowpos[i], j] = zhat_i / (d_i - lambda_j), normalised
__global__ void dc_svec_kernel(Ctx c, const Merge* __restrict__ mgs, double* S, int lds) {
    __shared__ double red[32];
    const Merge mg = mgs[blockIdx.y];
    const int lo = mg.lo, K = mg.K;
    for (int j = blockIdx.x; j < K; j += gridDim.x) {
        double part = 0.0;
        for (int i = threadIdx.x; i < K; i += blockDim.x) {
            double v = c.zhat[lo + i] / sec_delta(c.dl + lo, c.tau + lo, c.orig + lo, i, j);
            part += v * v;
        }
        double tot = block_sum<double>(part, red);
        double inv = 1.0 / sqrt(tot);
        double* col = S + (size_t)(lo + j) * lds + lo;
        for (int i = threadIdx.x; i < K; i += blockDim.x) {
            double v = c.zhat[lo + i] / sec_delta(c.dl + lo, c.tau + lo, c.orig + lo, i, j);
            col[c.rowpos[lo + i]] = v * inv;
        }
        __syncthreads();
    }
}

// pack non-deflated columns (type-grouped) and copy deflated columns to their final place
__global__ void dc_pack_kernel(Ctx c, const Merge* __restrict__ mgs, const double* __restrict__ Zin, int ldi,
                               double* Pack, int ldp, double* Zout, int ldo) {
    const Merge mg = mgs[blockIdx.z];
    const int lo = mg.lo, N = mg.hi - mg.lo, N1 = mg.mid - mg.lo, K = mg.K;
    for (int j = blockIdx.y; j < N; j += gridDim.y) {
        const int sc = c.src[lo + j], t = c.ctype[lo + j];
        const double* col = Zin + (size_t)(lo + sc) * ldi + lo;
        if (j < K) {
            const int p = c.rowpos[lo + j];
            double* ptop = Pack + (size_t)(lo + p) * ldp + lo;
            double* pbot = Pack + (size_t)(lo + p - mg.k1) * ldp + lo;
            for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) {
                if (r < N1) { if (t != 3) ptop[r] = col[r]; }
                else { if (t != 1) pbot[r] = col[r]; }
            }
        } else {
            double* dst = Zout + (size_t)(lo + c.pos[lo + j]) * ldo + lo;
            for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) {
                bool top = r < N1;
                dst[r] = ((t == 1 && !top) || (t == 3 && top)) ? 0.0 : col[r];
            }
        }
    }
}

__global__ void dc_scatter_kernel(Ctx c, const Merge* __restrict__ mgs, const double* __restrict__ Tmp, int ldt,
                                  double* Zout, int ldo) {
    const Merge mg = mgs[blockIdx.z];
    const int lo = mg.lo, N = mg.hi - mg.lo, K = mg.K;
    for (int j = blockIdx.y; j < K; j += gridDim.y) {
        const double* srcc = Tmp + (size_t)(lo + j) * ldt + lo;
        double* dst = Zout + (size_t)(lo + c.pos[lo + j]) * ldo + lo;
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) dst[r] = srcc[r];
    }
}

__global__ void dc_finish_kernel(int n, const double* __restrict__ D, const double* __restrict__ scale, double* w) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = D[i] * (*scale);
}

__global__ void dc_copy_kernel(int n, const double* __restrict__ src, int lds, double* dst, int ldd) {
    size_t total = (size_t)n * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(idx % n), cidx = (int)(idx / n);
        dst[(size_t)cidx * ldd + r] = src[(size_t)cidx * lds + r];
    }
}

static int dc_levels(int n) {
    int L = 0;
    while ((n + (1 << L) - 1) / (1 << L) > DC_LEAF) ++L;
    return L;
}

template <typename AR>
static void stedc_carve(AR& ar, int n, DcBuffers* b, double** Zw, double** Pack, double** S, int** bnd_dev,
                        int** cuts_dev) {
    const int L = dc_levels(n), nleaf = 1 << L;
    size_t nn = (size_t)(n > 0 ? n : 1);
    b->ctx.n = n;
    b->ctx.D = ar.template get<double>(nn);
    b->ctx.Dn = ar.template get<double>(nn);
    b->ctx.z = ar.template get<double>(nn);
    b->ctx.perm = ar.template get<int>(nn);
    b->ctx.dl = ar.template get<double>(nn);
    b->ctx.zl = ar.template get<double>(nn);
    b->ctx.src = ar.template get<int>(nn);
    b->ctx.ctype = ar.template get<int>(nn);
    b->ctx.rowpos = ar.template get<int>(nn);
    b->ctx.rot_p = ar.template get<int>(nn);
    b->ctx.rot_q = ar.template get<int>(nn);
    b->ctx.rot_c = ar.template get<double>(nn);
    b->ctx.rot_s = ar.template get<double>(nn);
    b->ctx.rot_tp = ar.template get<int>(nn);
    b->ctx.rot_tq = ar.template get<int>(nn);
    b->ctx.tau = ar.template get<double>(nn);
    b->ctx.orig = ar.template get<int>(nn);
    b->ctx.zhat = ar.template get<double>(nn);
    b->ctx.pos = ar.template get<int>(nn);
    b->merges = ar.template get<Merge>((size_t)nleaf);
    b->gemms = ar.template get<GemmProblem<double>>((size_t)nleaf);
    b->E = ar.template get<double>(nn);
    b->rho_cut = ar.template get<double>(nn + 1);
    b->sgn_cut = ar.template get<double>(nn + 1);
    b->scale = ar.template get<double>(4);
    b->info = ar.template get<int>(4);
    *bnd_dev = ar.template get<int>((size_t)nleaf + 1);
    *cuts_dev = ar.template get<int>((size_t)nleaf);
    *Zw = ar.template get<double>(nn * nn);
    *Pack = ar.template get<double>(nn * nn);
    *S = ar.template get<double>(nn * nn);
}

// The tree tables (leaf boundaries, cut positions, the merges of a level) depend on n only.  They are written by
// kernels, not uploaded: a host-to-device copy of a stack/vector buffer captured into a CUDA graph would be re-read
// from that (dead) host address at every replay (capi.cu replays this whole sequence per block).
__global__ void dc_tree_kernel(int n, int nleaf, int* __restrict__ bnd, int* __restrict__ cuts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nleaf) return;
    const int b = (int)((long long)i * n / nleaf);
    bnd[i] = b;
    if (i >= 1 && i < nleaf) cuts[i - 1] = b;
}
__global__ void dc_merge_fill_kernel(int nm, int step, const int* __restrict__ bnd, Merge* __restrict__ merges) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nm) return;
    Merge m{};
    m.lo = bnd[i * step];
    m.mid = bnd[i * step + step / 2];
    m.hi = bnd[(i + 1) * step];
    merges[i] = m;
}

size_t stedc_worksize(int n) {
    ArenaSize ar;
    DcBuffers b;
    double *a, *p, *s;
    int *bd, *cd;
    stedc_carve(ar, n, &b, &a, &p, &s, &bd, &cd);
    return ar.off + 256;
}

int stedc(makb200_handle* h, int n, const double* d, const double* e, double* w, double* Z, int ldz, void* work,
          size_t lwork, int* info_dev) {
    if (n <= 0) return 0;
    cudaStream_t st = h->stream;
    Arena ar(work, lwork);
    DcBuffers b;
    double *Zw, *Pack, *S;
    int *bnd_dev, *cuts_dev;
    stedc_carve(ar, n, &b, &Zw, &Pack, &S, &bnd_dev, &cuts_dev);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    const int L = dc_levels(n), nleaf = 1 << L;
    std::vector<int> bnd(nleaf + 1), cuts;
    for (int i = 0; i <= nleaf; ++i) bnd[i] = (int)((long long)i * n / nleaf);
    for (int i = 1; i < nleaf; ++i) cuts.push_back(bnd[i]);
    dc_tree_kernel<<<(nleaf + 1 + 127) / 128, 128, 0, st>>>(n, nleaf, bnd_dev, cuts_dev);
    count_launch();
    MAK_CUDA(h, cudaMemsetAsync(b.info, 0, sizeof(int) * 4, st));

    dc_scale_kernel<<<1, 1024, 0, st>>>(n, d, e, b.ctx.D, b.E, b.scale);
    if (!cuts.empty())
        dc_tear_kernel<<<(int)(cuts.size() + 127) / 128, 128, 0, st>>>((int)cuts.size(), cuts_dev, b.ctx.D, b.E,
                                                                        b.rho_cut, b.sgn_cut);
    // ping-pong so that the last level lands in the caller's Z
    const int ldw = n;
    double* Zin = (L % 2 == 0) ? Z : Zw;
    int ldi = (L % 2 == 0) ? ldz : ldw;
    double* Zout = (L % 2 == 0) ? Zw : Z;
    int ldo = (L % 2 == 0) ? ldw : ldz;
    PhaseTimer pt(st);
    pt.mark("start");
    dc_leaf_kernel<<<(nleaf + LEAF_WARPS - 1) / LEAF_WARPS, LEAF_WARPS * 32, 0, st>>>(nleaf, bnd_dev, b.ctx.D, b.E, Zin, ldi, b.info);
    pt.mark("leaf");
    count_launch(4);  // scale, tear, leaf, finish
    MAK_LAUNCH_CHECK(h, "dc_leaf_kernel");

    std::vector<Merge> hm;
    for (int lev = 1; lev <= L; ++lev) {
        const int step = 1 << lev, nm = nleaf / step;
        hm.assign(nm, Merge{});
        int maxN = 0, maxH = 0;
        for (int i = 0; i < nm; ++i) {
            hm[i].lo = bnd[i * step];
            hm[i].mid = bnd[i * step + step / 2];
            hm[i].hi = bnd[(i + 1) * step];
            maxN = std::max(maxN, hm[i].hi - hm[i].lo);
            maxH = std::max(maxH, std::max(hm[i].mid - hm[i].lo, hm[i].hi - hm[i].mid));
        }
        dc_merge_fill_kernel<<<(nm + 127) / 128, 128, 0, st>>>(nm, step, bnd_dev, b.merges);
        count_launch();
        dc_merge_init_kernel<<<(nm + 127) / 128, 128, 0, st>>>(nm, b.merges, b.rho_cut, b.sgn_cut);
        const int bx = std::max(1, std::min((maxN + 255) / 256, 64));
        dim3 g2(bx, nm);
        dc_z_rank_kernel<<<g2, 256, 0, st>>>(b.ctx, b.merges, Zin, ldi, 0);
        dc_z_rank_kernel<<<g2, 256, 0, st>>>(b.ctx, b.merges, Zin, ldi, 1);
        // Tmp (GEMM output) aliases Zin: its columns have been packed / copied before the GEMM runs
        dc_deflate_kernel<<<nm, 32, 0, st>>>(b.ctx, b.merges, b.gemms, Pack, n, S, n, Zin, ldi);
        dc_rotate_kernel<<<g2, 256, 0, st>>>(b.ctx, b.merges, Zin, ldi);
        dim3 gs(std::max(1, std::min((maxN + 127) / 128, 128)), nm);
        dc_secular_kernel<<<gs, 128, 0, st>>>(b.ctx, b.merges);
        dc_zhat_pos_kernel<<<gs, 128, 0, st>>>(b.ctx, b.merges);
        dim3 gv(std::min(maxN, 4096), nm);
        dc_svec_kernel<<<gv, 256, 0, st>>>(b.ctx, b.merges, S, n);
        dim3 gp(std::max(1, std::min((maxN + 255) / 256, 8)), std::min(maxN, 8192), nm);
        // grid.z is limited to 65535 merges per launch, far above n / DC_LEAF for any n that fits
        dc_pack_kernel<<<gp, 256, 0, st>>>(b.ctx, b.merges, Zin, ldi, Pack, n, Zout, ldo);
        MAK_LAUNCH_CHECK(h, "dc_pack_kernel");
        cudaError_t ge = gemm_grouped<double>(st, MAKB200_OP_N, MAKB200_OP_N, 2 * nm, maxH, maxN, b.gemms);
        if (ge != cudaSuccess) return cuda_fail(h, ge, "dc gemm_grouped");
        dc_scatter_kernel<<<gp, 256, 0, st>>>(b.ctx, b.merges, Zin, ldi, Zout, ldo);
        MAK_LAUNCH_CHECK(h, "dc_scatter_kernel");
        count_launch(10);  // the ten non-GEMM kernels of this level
        std::swap(Zin, Zout);
        std::swap(ldi, ldo);
        std::swap(b.ctx.D, b.ctx.Dn);
        pt.mark("level");
    }
    pt.report("stedc");
    // after the swaps Zin is the caller's Z
    dc_finish_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, b.ctx.D, b.scale, w);
    MAK_LAUNCH_CHECK(h, "dc_finish_kernel");
    if (info_dev) MAK_CUDA(h, cudaMemcpyAsync(info_dev, b.info, sizeof(int), cudaMemcpyDeviceToDevice, st));
    return 0;
}

// ---------------------------------------------------------------------------------------
// Many tridiagonals in ONE pass of the same kernels (lock-step batched eigh / svd of mid-size blocks).
//
// The blocks are laid end to end as one block-diagonal tridiagonal of order Ntot = sum n_i (the couplings between
// blocks are zero), each with its own binary tree, so a level's launches cover the merges of EVERY block that still has
// that level.  The kernels above address the eigenvector matrix as Z[(lo + c) * ld + lo + r] with global positions and
// only ever touch diagonal sub-blocks, so the "Ntot x Ntot" matrices are stored as strips with ld = nmax: block i
// (global offset o_i) owns the addresses o_i (ld + 1) + [0, n_i ld), disjoint from every other block since ld >= n_i.
// A block with L_i levels ends in ping-pong buffer L_i mod 2 (eigenvalues and vectors alike); the finishing kernel
// reads each block from its own buffer, rescales the eigenvalues and writes V_i (Float64 or ComplexF64).
// ---------------------------------------------------------------------------------------
struct StedcBlkDev {
    int n, off, par, ldv;
    const double* d;
    const double* e;
    double* w;
    void* V;
};

__global__ void dc_scale_batched_kernel(const StedcBlkDev* __restrict__ blks, double* D, double* E, double* scale) {
    __shared__ double red[32];
    const StedcBlkDev b = blks[blockIdx.x];
    const int n = b.n;
    double m = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, fabs(b.d[i]));
    for (int i = threadIdx.x; i < n - 1; i += blockDim.x) m = fmax(m, fabs(b.e[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t = fmax(t, red[i]);
    const double sc = t > 0.0 ? t : 1.0;
    if (threadIdx.x == 0) scale[blockIdx.x] = sc;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        D[b.off + i] = b.d[i] / sc;
        E[b.off + i] = (i < n - 1) ? b.e[i] / sc : 0.0;
    }
}

template <typename T>
__global__ void dc_finish_batched_kernel(const StedcBlkDev* __restrict__ blks, const double* __restrict__ D0,
                                         const double* __restrict__ D1, const double* __restrict__ Z0,
                                         const double* __restrict__ Z1, int ld, const double* __restrict__ scale) {
    const StedcBlkDev b = blks[blockIdx.y];
    const int n = b.n;
    const double* D = b.par ? D1 : D0;
    const double* Z = (b.par ? Z1 : Z0) + (size_t)b.off * ((size_t)ld + 1);
    T* V = reinterpret_cast<T*>(b.V);
    const size_t start = blockIdx.x * (size_t)blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    const double sc = scale[blockIdx.y];
    for (size_t i = start; i < (size_t)n; i += step) b.w[i] = D[b.off + i] * sc;
    const size_t total = (size_t)n * n;
    for (size_t idx = start; idx < total; idx += step) {
        const int r = (int)(idx % n), c = (int)(idx / n);
        V[(size_t)c * b.ldv + r] = mk<T>(Z[(size_t)c * ld + r]);
    }
}

struct DcBatchLayout {
    Ctx ctx;
    Merge* merges;
    GemmProblem<double>* gemms;
    double *E, *rho_cut, *sgn_cut, *scale;
    int *info, *bnd, *cuts;
    StedcBlkDev* blks;
    double *Z0, *Z1, *Pack, *S;
};
template <typename AR>
static void stedc_batched_carve(AR& ar, int nblk, size_t ntot, int nmax, size_t nleaves, DcBatchLayout* b) {
    const size_t nn = ntot > 0 ? ntot : 1;
    b->ctx.n = (int)ntot;
    b->ctx.D = ar.template get<double>(nn);
    b->ctx.Dn = ar.template get<double>(nn);
    b->ctx.z = ar.template get<double>(nn);
    b->ctx.perm = ar.template get<int>(nn);
    b->ctx.dl = ar.template get<double>(nn);
    b->ctx.zl = ar.template get<double>(nn);
    b->ctx.src = ar.template get<int>(nn);
    b->ctx.ctype = ar.template get<int>(nn);
    b->ctx.rowpos = ar.template get<int>(nn);
    b->ctx.rot_p = ar.template get<int>(nn);
    b->ctx.rot_q = ar.template get<int>(nn);
    b->ctx.rot_c = ar.template get<double>(nn);
    b->ctx.rot_s = ar.template get<double>(nn);
    b->ctx.rot_tp = ar.template get<int>(nn);
    b->ctx.rot_tq = ar.template get<int>(nn);
    b->ctx.tau = ar.template get<double>(nn);
    b->ctx.orig = ar.template get<int>(nn);
    b->ctx.zhat = ar.template get<double>(nn);
    b->ctx.pos = ar.template get<int>(nn);
    b->merges = ar.template get<Merge>(nleaves + 1);          // all levels: sum over levels of merges < leaves
    b->gemms = ar.template get<GemmProblem<double>>(nleaves + 2);
    b->E = ar.template get<double>(nn);
    b->rho_cut = ar.template get<double>(nn + 1);
    b->sgn_cut = ar.template get<double>(nn + 1);
    b->scale = ar.template get<double>((size_t)nblk + 1);
    b->info = ar.template get<int>(4);
    b->bnd = ar.template get<int>(nleaves + 1);
    b->cuts = ar.template get<int>(nleaves + 1);
    b->blks = ar.template get<StedcBlkDev>((size_t)nblk + 1);
    const size_t strip = nn * ((size_t)nmax + 1);
    b->Z0 = ar.template get<double>(strip);
    b->Z1 = ar.template get<double>(strip);
    b->Pack = ar.template get<double>(strip);
    b->S = ar.template get<double>(strip);
}

size_t stedc_batched_worksize(int nblk, const int* n) {
    size_t ntot = 0, nleaves = 0;
    int nmax = 1;
    for (int i = 0; i < nblk; ++i) {
        ntot += (size_t)n[i];
        nleaves += (size_t)1 << dc_tree_levels(n[i]);
        nmax = std::max(nmax, n[i]);
    }
    ArenaSize ar;
    DcBatchLayout b;
    stedc_batched_carve(ar, nblk, ntot, nmax, nleaves, &b);
    return ar.off + 256;
}

template <typename T>
int stedc_batched(makb200_handle* h, int nblk, const StedcBlk* blks, void* work, size_t lwork, int* info_dev) {
    if (nblk <= 0) return 0;
    cudaStream_t st = h->stream;
    std::vector<int> ns(nblk);
    for (int i = 0; i < nblk; ++i) {
        if (blks[i].n <= 0) return -2;
        ns[i] = blks[i].n;
    }
    // host tables: leaf boundaries (global), cuts, the merges of every level (stedc_batch_tables.h), block descriptors
    const BatchTables tb = dc_batch_tables(nblk, ns.data());
    const size_t ntot = tb.ntot, nleaves = tb.nleaves;
    const int nmax = tb.nmax;
    const std::vector<int>&bnd = tb.bnd, &cuts = tb.cuts;
    const std::vector<Merge>& merges = tb.merges;
    if (ntot * ((size_t)nmax + 1) > (size_t)0x7fffffff * 64) return -3;
    Arena ar(work, lwork);
    DcBatchLayout b;
    stedc_batched_carve(ar, nblk, ntot, nmax, nleaves, &b);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    std::vector<StedcBlkDev> bd(nblk);
    for (int i = 0; i < nblk; ++i)
        bd[i] = StedcBlkDev{blks[i].n, tb.off[i], tb.lev[i] & 1, blks[i].ldv, blks[i].d, blks[i].e, blks[i].w, blks[i].V};
    if (merges.size() > nleaves + 1) return -4;
    {
        Stager sg(h, (bnd.size() + cuts.size()) * sizeof(int) + merges.size() * sizeof(Merge) + bd.size() * sizeof(StedcBlkDev) + 4096);
        MAK_CUDA(h, sg.put(b.bnd, bnd.data(), bnd.size() * sizeof(int), st));
        MAK_CUDA(h, sg.put(b.cuts, cuts.data(), cuts.size() * sizeof(int), st));
        MAK_CUDA(h, sg.put(b.merges, merges.data(), merges.size() * sizeof(Merge), st));
        MAK_CUDA(h, sg.put(b.blks, bd.data(), bd.size() * sizeof(StedcBlkDev), st));
    }
    MAK_CUDA(h, cudaMemsetAsync(b.info, 0, sizeof(int) * 4, st));
    dc_scale_batched_kernel<<<nblk, 256, 0, st>>>(b.blks, b.ctx.D, b.E, b.scale);
    if (!cuts.empty())
        dc_tear_kernel<<<(int)(cuts.size() + 127) / 128, 128, 0, st>>>((int)cuts.size(), b.cuts, b.ctx.D, b.E, b.rho_cut, b.sgn_cut);
    const int ld = nmax;
    double *Zin = b.Z0, *Zout = b.Z1;
    double *D0 = b.ctx.D, *D1 = b.ctx.Dn;
    const int nl_all = (int)nleaves;
    dc_leaf_kernel<<<(nl_all + LEAF_WARPS - 1) / LEAF_WARPS, LEAF_WARPS * 32, 0, st>>>(nl_all, b.bnd, b.ctx.D, b.E, Zin, ld, b.info);
    count_launch(3);
    MAK_LAUNCH_CHECK(h, "dc_leaf_kernel (batched)");
    for (const BatchLevel& li : tb.levels) {
        const int nm = li.nm, maxN = li.maxN, maxH = li.maxH;
        if (nm == 0) continue;
        Merge* mg = b.merges + li.first;
        for (int m0 = 0; m0 < nm; m0 += 32768) {      // grid.y / grid.z limit
            const int nmc = std::min(32768, nm - m0);
            Merge* mgc = mg + m0;
            dc_merge_init_kernel<<<(nmc + 127) / 128, 128, 0, st>>>(nmc, mgc, b.rho_cut, b.sgn_cut);
            const int bx = std::max(1, std::min((maxN + 255) / 256, 64));
            dim3 g2(bx, nmc);
            dc_z_rank_kernel<<<g2, 256, 0, st>>>(b.ctx, mgc, Zin, ld, 0);
            dc_z_rank_kernel<<<g2, 256, 0, st>>>(b.ctx, mgc, Zin, ld, 1);
            dc_deflate_kernel<<<nmc, 32, 0, st>>>(b.ctx, mgc, b.gemms, b.Pack, ld, b.S, ld, Zin, ld);
            dc_rotate_kernel<<<g2, 256, 0, st>>>(b.ctx, mgc, Zin, ld);
            dim3 gs(std::max(1, std::min((maxN + 127) / 128, 128)), nmc);
            dc_secular_kernel<<<gs, 128, 0, st>>>(b.ctx, mgc);
            dc_zhat_pos_kernel<<<gs, 128, 0, st>>>(b.ctx, mgc);
            dim3 gv(std::min(maxN, 4096), nmc);
            dc_svec_kernel<<<gv, 256, 0, st>>>(b.ctx, mgc, b.S, ld);
            dim3 gp(std::max(1, std::min((maxN + 255) / 256, 8)), std::min(maxN, 8192), nmc);
            dc_pack_kernel<<<gp, 256, 0, st>>>(b.ctx, mgc, Zin, ld, b.Pack, ld, Zout, ld);
            MAK_LAUNCH_CHECK(h, "dc_pack_kernel (batched)");
            cudaError_t ge = gemm_grouped<double>(st, MAKB200_OP_N, MAKB200_OP_N, 2 * nmc, maxH, maxN, b.gemms);
            if (ge != cudaSuccess) return cuda_fail(h, ge, "dc gemm_grouped (batched)");
            dc_scatter_kernel<<<gp, 256, 0, st>>>(b.ctx, mgc, Zin, ld, Zout, ld);
            MAK_LAUNCH_CHECK(h, "dc_scatter_kernel (batched)");
            count_launch(9);
        }
        std::swap(Zin, Zout);
        std::swap(b.ctx.D, b.ctx.Dn);
    }
    {
        const int gx = std::max(1, std::min((nmax * nmax + 255) / 256, 64));
        dc_finish_batched_kernel<T><<<dim3(gx, nblk), 256, 0, st>>>(b.blks, D0, D1, b.Z0, b.Z1, ld, b.scale);
        count_launch();
        MAK_LAUNCH_CHECK(h, "dc_finish_batched_kernel");
    }
    if (info_dev) MAK_CUDA(h, cudaMemcpyAsync(info_dev, b.info, sizeof(int), cudaMemcpyDeviceToDevice, st));
    return 0;
}
template int stedc_batched<double>(makb200_handle*, int, const StedcBlk*, void*, size_t, int*);
template int stedc_batched<cplx>(makb200_handle*, int, const StedcBlk*, void*, size_t, int*);

}  // namespace mak
