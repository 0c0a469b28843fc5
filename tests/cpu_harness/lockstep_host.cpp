// CPU replay of the lock-step batched QDWH plan (csrc/polar_lockstep_plan.h): the PRODUCT planner builds the launch
// list and the grouped-GEMM descriptors; this file executes every action with naive host loops, so the test checks
// the schedule, the pointer arithmetic of every descriptor and the blocked Cholesky / triangular-solve sweeps without a
// GPU.  `lower` GEMMs write NaN above the diagonal (the kernel skips those tiles), so any read of an unwritten entry
// poisons the result.  Test infrastructure only (tests/test_lockstep_plan_cpu.py).
#include "../../matrixalgebrakit.jl_b200/csrc/polar_lockstep_plan.h"
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

using namespace mak;

template <typename T>
static void host_gemm(const GemmProblem<T>& p, int opa, int opb) {
    if (p.m <= 0 || p.n <= 0) return;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int j = 0; j < p.n; ++j)
        for (int i = 0; i < p.m; ++i) {
            T* c = p.C + (size_t)j * p.ldc + i;
            if (p.lower && i < j) { *c = mk<T>(nan); continue; }
            T acc = zero<T>();
            for (int l = 0; l < p.k; ++l) {
                T a = opa ? p.A[(size_t)i * p.lda + l] : p.A[(size_t)l * p.lda + i];
                T b = opb ? p.B[(size_t)l * p.ldb + j] : p.B[(size_t)j * p.ldb + l];
                if (opa && p.conja) a = conj_(a);
                if (opb && p.conjb) b = conj_(b);
                fma_(acc, a, b);
            }
            T out = mul_(p.alpha, acc);
            if (!is_zero(p.beta)) out = add_(out, mul_(p.beta, *c));
            *c = out;
        }
}

// Q (m x n) with orthonormal columns spanning A (m x n, ld lda), Gram-Schmidt with re-orthogonalisation
template <typename T>
static void host_orth(int m, int n, const T* A, int lda, T* Q, int ldq) {
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < m; ++i) Q[(size_t)j * ldq + i] = A[(size_t)j * lda + i];
        for (int pass = 0; pass < 2; ++pass)
            for (int k = 0; k < j; ++k) {
                T d = zero<T>();
                for (int i = 0; i < m; ++i) fmac_(d, Q[(size_t)k * ldq + i], Q[(size_t)j * ldq + i]);
                for (int i = 0; i < m; ++i) Q[(size_t)j * ldq + i] = sub_(Q[(size_t)j * ldq + i], mul_(Q[(size_t)k * ldq + i], d));
            }
        double s = 0.0;
        for (int i = 0; i < m; ++i) s += abs2_(Q[(size_t)j * ldq + i]);
        const double inv = s > 0.0 ? 1.0 / sqrt(s) : 0.0;
        for (int i = 0; i < m; ++i) Q[(size_t)j * ldq + i] = scale_(Q[(size_t)j * ldq + i], inv);
    }
}

template <typename T>
static int host_potf2(int jb, const T* Z, int ldz, T* L, int ldl, T* Linv, int nb) {
    int info = 0;
    std::vector<T> a((size_t)jb * jb);
    for (int c = 0; c < jb; ++c)
        for (int r = 0; r < jb; ++r) a[(size_t)c * jb + r] = r >= c ? Z[(size_t)c * ldz + r] : zero<T>();
    for (int k = 0; k < jb; ++k) {
        double akk = real_(a[(size_t)k * jb + k]);
        if (!(akk > 0.0)) { info = 1; akk = 1.0; }
        const double piv = sqrt(akk);
        a[(size_t)k * jb + k] = mk<T>(piv);
        for (int r = k + 1; r < jb; ++r) a[(size_t)k * jb + r] = scale_(a[(size_t)k * jb + r], 1.0 / piv);
        for (int c = k + 1; c < jb; ++c)
            for (int r = c; r < jb; ++r)
                a[(size_t)c * jb + r] = sub_(a[(size_t)c * jb + r], mul_(a[(size_t)k * jb + r], conj_(a[(size_t)k * jb + c])));
    }
    for (int c = 0; c < jb; ++c)
        for (int r = 0; r < jb; ++r) L[(size_t)c * ldl + r] = a[(size_t)c * jb + r];
    // inverse by forward substitution, column by column
    for (int c = 0; c < nb; ++c)
        for (int r = 0; r < nb; ++r) Linv[(size_t)c * nb + r] = zero<T>();
    for (int c = 0; c < jb; ++c) {
        for (int r = c; r < jb; ++r) {
            T s = r == c ? one<T>() : zero<T>();
            for (int k = c; k < r; ++k) s = sub_(s, mul_(a[(size_t)k * jb + r], Linv[(size_t)c * nb + k]));
            Linv[(size_t)c * nb + r] = scale_(s, 1.0 / real_(a[(size_t)r * jb + r]));
        }
    }
    return info;
}

template <typename T>
static T* host_buf(const LsBlk<T>& b, int id) { return LsPlanner<T>::buf(b, id); }

// executes one plan on the host; info[i] / est[i] per block of `blk`
template <typename T>
static int exec_plan(const std::vector<LsBlk<T>>& blk, const LsPlan<T>& pl, int nb, std::vector<int>& info, std::vector<double>& est) {
    info.assign(blk.size(), 0);
    est.assign(blk.size(), 0.0);
    for (const LsAct& a : pl.acts) {
        switch (a.kind) {
            case LS_GEMM:
                for (int i = 0; i < a.count; ++i) {
                    const GemmProblem<T>& g = pl.probs[a.off + i];
                    if (g.m > a.max_m || g.n > a.max_n) return -3;
                    host_gemm<T>(g, a.opa != 0, a.opb != 0);
                }
                break;
            case LS_PREP:
                for (int i = 0; i < a.count; ++i) {
                    const LsBlk<T>& b = blk[i];
                    double mx = 0.0, s = 0.0;
                    for (int c = 0; c < b.n; ++c)
                        for (int r = 0; r < b.n; ++r) {
                            const T v = b.S[(size_t)c * b.lds + r];
                            mx = fmax(mx, fmax(fabs(real_(v)), fabs(imag_(v))));
                        }
                    const double f = (mx > 0.0 && std::isfinite(mx)) ? 1.0 / mx : 1.0;
                    for (int c = 0; c < b.n; ++c)
                        for (int r = 0; r < b.n; ++r) s += abs2_(scale_(b.S[(size_t)c * b.lds + r], f));
                    const double inv = s > 0.0 ? 1.0 / sqrt(s) : 1.0;
                    for (int c = 0; c < b.n; ++c)
                        for (int r = 0; r < b.n; ++r) b.X[(size_t)c * b.n + r] = scale_(scale_(b.S[(size_t)c * b.lds + r], f), inv);
                }
                break;
            case LS_STACK:
                for (int i = 0; i < a.count; ++i) {
                    const LsBlk<T>& b = blk[i];
                    const int nn = b.n;
                    for (int c = 0; c < nn; ++c)
                        for (int r = 0; r < 2 * nn; ++r)
                            b.B[(size_t)c * 2 * nn + r] = r < nn ? scale_(b.X[(size_t)c * nn + r], a.p0) : (r - nn == c ? one<T>() : zero<T>());
                }
                break;
            case LS_ADDDIAG:
                for (int i = 0; i < a.count; ++i)
                    for (int d = 0; d < blk[i].n; ++d) blk[i].Z[(size_t)d * blk[i].n + d] = add_(blk[i].Z[(size_t)d * blk[i].n + d], one<T>());
                break;
            case LS_AXPBY:
                for (int i = 0; i < a.count; ++i) {
                    const LsBlk<T>& b = blk[i];
                    for (size_t e = 0; e < (size_t)b.n * b.n; ++e) b.X[e] = add_(scale_(b.X[e], a.p0), scale_(b.B[e], a.p1));
                }
                break;
            case LS_COPY:
                for (int i = 0; i < a.count; ++i) {
                    const LsBlk<T>& b = blk[i];
                    if (a.a1 == LS_W && b.m != b.n) continue;
                    const T* src = host_buf(b, a.a0);
                    T* dst = host_buf(b, a.a1);
                    for (size_t e = 0; e < (size_t)a.a2 * b.n * b.n; ++e) dst[e] = src[e];
                }
                break;
            case LS_SYMM:
                for (int i = 0; i < a.count; ++i) {
                    const LsBlk<T>& b = blk[i];
                    for (int c = 0; c < b.n; ++c)
                        for (int r = 0; r < b.n; ++r) {
                            T v = scale_(add_(b.Z[(size_t)c * b.n + r], conj_(b.Z[(size_t)r * b.n + c])), 0.5);
                            if (r == c) v = mk<T>(real_(v));
                            b.P[(size_t)c * b.n + r] = v;
                        }
                }
                break;
            case LS_POTF2:
                for (int i = 0; i < a.count; ++i) {
                    const LsBlk<T>& b = blk[i];
                    const int nn = b.n, j0 = a.a0, jb = std::min(nb, nn - j0);
                    if (jb <= 0) return -4;   // inactive block inside the active prefix
                    info[i] |= host_potf2<T>(jb, b.Z + (size_t)j0 * nn + j0, nn, b.L + (size_t)j0 * nn + j0, nn,
                                             b.Linv + (size_t)a.a1 * nb * nb, nb);
                }
                break;
            case LS_PROBE:
                for (int i = 0; i < a.count; ++i)
                    for (int c = 0; c < blk[i].n; ++c)
                        for (int r = 0; r < LS_NPROBE; ++r) blk[i].T2[(size_t)c * LS_NPROBE + r] = mk<T>(ls_probe_entry(r, c));
                break;
            case LS_FRO:
                for (int i = 0; i < a.count; ++i) {
                    double s = 0.0;
                    for (size_t e = 0; e < (size_t)LS_NPROBE * blk[i].n; ++e) s += abs2_(blk[i].Q[e]);
                    est[i] = s;
                }
                break;
            case LS_QR_TALL:
                for (int i = 0; i < a.count; ++i) {
                    const LsBlk<T>& b = blk[i];
                    if (b.m <= b.n) continue;
                    host_orth<T>(b.m, b.n, b.A, b.lda, b.Q0, b.m);
                    for (int c = 0; c < b.n; ++c)
                        for (int r = 0; r < b.n; ++r) {
                            T s = zero<T>();
                            if (r <= c)
                                for (int k = 0; k < b.m; ++k) fmac_(s, b.Q0[(size_t)r * b.m + k], b.A[(size_t)c * b.lda + k]);
                            b.R0[(size_t)c * b.n + r] = s;
                        }
                }
                break;
            case LS_QR_STACK:
                for (int i = 0; i < a.count; ++i) host_orth<T>(2 * blk[i].n, blk[i].n, blk[i].B, 2 * blk[i].n, blk[i].Q, 2 * blk[i].n);
                break;
            default: return -5;
        }
    }
    return 0;
}

// the driver of csrc/polar_lockstep.cuh: polar_lockstep_run, on the host
template <typename T>
static int replay(int count, const int* m, const int* n, void* const* A, void* const* W, void* const* P, int nb, int estimate,
                  int* n_acts, int* n_gemm, int* n_fast, double* l0_out) {
    std::vector<LsBlk<T>> blk(count);
    size_t elems = 0;
    for (int i = 0; i < count; ++i) elems += ls_block_elems<T>(m[i], n[i], nb);
    const double nan = std::numeric_limits<double>::quiet_NaN();
    std::vector<T> store(elems, mk<T>(nan));
    T* p = store.data();
    for (int i = 0; i < count; ++i) {
        if (i > 0 && n[i] > n[i - 1]) return -1;   // the caller sorts by n descending
        LsBlk<T>& b = blk[i];
        b = LsBlk<T>{};
        b.m = m[i]; b.n = n[i];
        b.A = (const T*)A[i]; b.lda = m[i];
        b.W = (T*)W[i]; b.P = (T*)P[i];
        ls_carve_block<T>(b, p, nb);
    }
    if ((size_t)(p - store.data()) > elems) return -2;
    std::vector<int> info;
    std::vector<double> est;
    int acts = 0, gemms = 0;
    {
        LsPlan<T> pl;
        LsPlanner<T> planner(blk, nb, pl);
        planner.build_prepare(estimate != 0);
        int rc = exec_plan<T>(blk, pl, nb, info, est);
        if (rc) return rc;
        acts += (int)pl.acts.size(); gemms += pl.gemm_launches;
    }
    std::vector<LsBlk<T>> fast, slow;
    double l0_fast = 0.9;
    if (estimate) {
        for (int i = 0; i < count; ++i) {
            const double l0 = ls_l0_from_estimate(est[i], info[i]);
            if (l0 > 1e-7) { fast.push_back(blk[i]); l0_fast = std::min(l0_fast, l0); }
            else slow.push_back(blk[i]);
        }
    } else {
        slow = blk;
    }
    int bad = 0;
    for (int g = 0; g < 2; ++g) {
        const std::vector<LsBlk<T>>& grp = g == 0 ? fast : slow;
        if (grp.empty()) continue;
        LsPlan<T> pl;
        LsPlanner<T> planner(grp, nb, pl);
        planner.build_iterate(qdwh_schedule(g == 0 ? l0_fast : 2.2e-16, 12, 100.0));
        std::vector<int> info2;
        std::vector<double> est2;
        int rc = exec_plan<T>(grp, pl, nb, info2, est2);
        if (rc) return rc;
        for (int v : info2) bad |= v;   // the iteration's matrices I + c X^H X are positive definite
        acts += (int)pl.acts.size(); gemms = std::max(gemms, pl.gemm_launches);
    }
    if (n_acts) *n_acts = acts;
    if (n_gemm) *n_gemm = gemms;
    if (n_fast) *n_fast = (int)fast.size();
    if (l0_out) *l0_out = l0_fast;
    return bad;
}

extern "C" int lockstep_replay(int dtype, int count, const int* m, const int* n, void* const* A, void* const* W, void* const* P,
                               int nb, int estimate, int* n_acts, int* n_gemm, int* n_fast, double* l0) {
    return dtype == 0 ? replay<double>(count, m, n, A, W, P, nb, estimate, n_acts, n_gemm, n_fast, l0)
                      : replay<cplx>(count, m, n, A, W, P, nb, estimate, n_acts, n_gemm, n_fast, l0);
}
extern "C" int lockstep_launch_bound(int dtype, int nmax, int nb, int any_tall) {
    const std::vector<QdwhStep> sched = qdwh_schedule(2.2e-16, 12, 100.0);
    return dtype == 0 ? ls_gemm_launch_bound<double>(nmax, nb, any_tall != 0, sched)
                      : ls_gemm_launch_bound<cplx>(nmax, nb, any_tall != 0, sched);
}
