// Lock-step QDWH polar decomposition of MANY mid-size blocks (65 <= n <= 512): the launch plan.
//
// The batched svd_compact!/svd_trunc! of block-sparse tensors (SURVEY config 3) ran one QDWH chain per block on a
// stream pool: ~500 dependent launches per block, device-dispatch bound (profiles/r2_batched_graphs.log).  Here ALL
// blocks of a chunk advance together: every launch of the sequence below is ONE grouped GEMM (device-side problem
// descriptors, csrc/gemm.cu: gemm_grouped) or ONE batched kernel with a CTA (or CTA column) per block, so a chunk of
// 192 blocks costs ~600 launches instead of ~100 000.
//
// One QDWH schedule serves a whole group of blocks:
//   part 1 (build_prepare)  X0 = S / ||S||_F and a sigma_min estimate per block (Gram, Cholesky, LS_NPROBE Rademacher
//                           probes solved against L; the rule of the single-matrix driver, polar.cu); one D2H read
//   part 2 (build_iterate)  the blocks with a usable estimate run the schedule of the smallest l0 among them, e.g.
//                             step 1  (c ~ 1e6)   CholeskyQR2 of [sqrt(c) X; I]   -> Gram, Cholesky, triangular solve, twice
//                             steps 2-5           Z = I + c X^H X = L L^H,  X <- (b/c) X + (a - b/c) (X L^-H) L^-1
//                           the others the l0 = eps schedule, whose first step (c ~ 1e21) is a Householder QR of
//                           [sqrt(c) X; I] -> the lock-step batched QR (batched_blocked.cu)
// and then P = sym(X^H A), W = X (or Q0 X for a tall block, A = Q0 R0 first as the single-matrix driver does).
// Blocks are sorted by n descending, so the blocks still active at column j0 of a blocked sweep are a prefix.
//
// This header is plain C++ (no CUDA): the plan is built on the host, uploaded once per chunk and replayed by
// polar_lockstep.cuh; tests/cpu_harness/lockstep_host.cpp replays the same plan with naive host loops and checks
// W^H W = I, W P = A, P = P^H >= 0 (tests/test_lockstep_plan_cpu.py).
// Reference semantics: /root/reference/src/implementations/svd.jl:196-237 (batched wrappers = loop over blocks),
// /root/reference/src/implementations/polar.jl:13-45.
#pragma once
#include "batched_desc.h"
#include "qdwh_schedule.h"
#include <algorithm>
#include <cstddef>
#include <vector>

namespace mak {

template <typename T>
struct LsPlan {
    std::vector<LsAct> acts;
    std::vector<GemmProblem<T>> probs;
    int gemm_launches = 0;
};

// elements of T one block needs (every buffer rounded to an even count: 16-byte alignment for Float64)
template <typename T>
inline size_t ls_block_elems(int m, int n, int nb) {
    const size_t nn = (size_t)n, mm = (size_t)m, rb = std::max<size_t>(2 * nn, mm);
    auto ev = [](size_t e) { return (e + 1) & ~(size_t)1; };
    size_t e = ev(nn * nn) + ev(rb * nn) + 2 * ev(2 * nn * nn) + 2 * ev(nn * nn) + ev((size_t)nb * nb * ((nn + nb - 1) / nb));
    if (m > n) e += ev(mm * nn) + ev(nn * nn);
    return e;
}
template <typename T>
inline void ls_carve_block(LsBlk<T>& b, T*& p, int nb) {
    const size_t nn = (size_t)b.n, mm = (size_t)b.m, rb = std::max<size_t>(2 * nn, mm);
    auto take = [&](size_t e) { T* r = p; p += (e + 1) & ~(size_t)1; return r; };
    b.X = take(nn * nn);
    b.B = take(rb * nn);
    b.Q = take(2 * nn * nn);
    b.T2 = take(2 * nn * nn);
    b.Z = take(nn * nn);
    b.L = take(nn * nn);
    b.Linv = take((size_t)nb * nb * ((nn + nb - 1) / nb));
    if (b.m > b.n) {
        b.Q0 = take(mm * nn);
        b.R0 = take(nn * nn);
        b.S = b.R0; b.lds = b.n;
    } else {
        b.Q0 = nullptr; b.R0 = nullptr;
        b.S = b.A; b.lds = b.lda;
    }
}

template <typename T>
inline GemmProblem<T> ls_prob(int m, int n, int k, const T* A, int lda, const T* B, int ldb, T* C, int ldc, double alpha,
                              double beta, int ca, int cb, int lower) {
    GemmProblem<T> p;
    p.m = m; p.n = n; p.k = k;
    p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
    p.alpha = mk<T>(alpha); p.beta = mk<T>(beta);
    p.conja = ca; p.conjb = cb; p.lower = lower;
    return p;
}

template <typename T>
struct LsPlanner {
    const std::vector<LsBlk<T>>& blk;   // n descending
    const int nb, count, nmax;
    LsPlan<T>& pl;
    LsPlanner(const std::vector<LsBlk<T>>& b, int nb_, LsPlan<T>& p)
        : blk(b), nb(nb_), count((int)b.size()), nmax(b.empty() ? 0 : b[0].n), pl(p) {}

    int active_for(int j0) const { int a = 0; while (a < count && blk[a].n > j0) ++a; return a; }
    static T* buf(const LsBlk<T>& b, int id) {
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
    void act(int kind, int cnt, int a0 = 0, int a1 = 0, int a2 = 0, double p0 = 0.0, double p1 = 0.0) {
        LsAct a{};
        a.kind = kind; a.count = cnt; a.a0 = a0; a.a1 = a1; a.a2 = a2; a.p0 = p0; a.p1 = p1;
        pl.acts.push_back(a);
    }
    // one grouped launch; problems with nothing to do are kept as m = 0 entries so that entry i is block i
    void gemm(int opa, int opb, std::vector<GemmProblem<T>>& g) {
        int max_m = 0, max_n = 0;
        for (auto& p : g) {
            if (p.m <= 0 || p.n <= 0) { p.m = 0; p.n = 0; continue; }
            max_m = std::max(max_m, p.m); max_n = std::max(max_n, p.n);
        }
        while (!g.empty() && g.back().m == 0) g.pop_back();
        if (g.empty() || max_m == 0) return;
        LsAct a{};
        a.kind = LS_GEMM; a.opa = opa; a.opb = opb; a.max_m = max_m; a.max_n = max_n;
        a.count = (int)g.size(); a.off = pl.probs.size();
        pl.probs.insert(pl.probs.end(), g.begin(), g.end());
        pl.acts.push_back(a);
        ++pl.gemm_launches;
    }
    // Z (lower) = alpha * S^H S, S = buffer `src` with rowmult * n rows
    void gram(int src, int rowmult, double alpha) {
        std::vector<GemmProblem<T>> g;
        for (const auto& b : blk) {
            const int r = rowmult * b.n;
            g.push_back(ls_prob<T>(b.n, b.n, r, buf(b, src), r, buf(b, src), r, b.Z, b.n, alpha, 0.0, 1, 0, 1));
        }
        gemm(2, 0, g);
    }
    // Z = L L^H (lower; left-looking, nb columns per step) with the inverse of every diagonal block
    void potrf() {
        for (int j0 = 0, bi = 0; j0 < nmax; j0 += nb, ++bi) {
            const int na = active_for(j0);
            if (j0 > 0) {
                std::vector<GemmProblem<T>> g;
                for (int i = 0; i < na; ++i) {
                    const auto& b = blk[i];
                    const int n = b.n, jb = std::min(nb, n - j0);
                    g.push_back(ls_prob<T>(n - j0, jb, j0, b.L + j0, n, b.L + j0, n, b.Z + (size_t)j0 * n + j0, n, -1.0, 1.0, 0, 1, 0));
                }
                gemm(0, 2, g);
            }
            act(LS_POTF2, na, j0, bi);
            std::vector<GemmProblem<T>> g;
            for (int i = 0; i < na; ++i) {
                const auto& b = blk[i];
                const int n = b.n, jb = std::min(nb, n - j0), mr = n - j0 - jb;
                g.push_back(ls_prob<T>(mr, jb, jb, b.Z + (size_t)j0 * n + j0 + jb, n, b.Linv + (size_t)bi * nb * nb, nb,
                                       b.L + (size_t)j0 * n + j0 + jb, n, 1.0, 0.0, 0, 1, 0));
            }
            gemm(0, 2, g);
        }
    }
    // dst = src L^-H (conj) or src L^-1; src, dst, T2 hold rowmult * n rows, or `rows` rows when rows > 0 (then the
    // right-hand side is already in T2)
    void trsm(bool conj, int rowmult, int src, int dst, int rows = 0) {
        if (rows <= 0) act(LS_COPY, count, src, LS_T2, rowmult);
        const int nblk = (nmax + nb - 1) / nb;
        for (int bb = 0; bb < nblk; ++bb) {
            const int bi = conj ? bb : nblk - 1 - bb, j0 = bi * nb;
            const int na = active_for(j0);
            std::vector<GemmProblem<T>> g1, g2;
            for (int i = 0; i < na; ++i) {
                const auto& b = blk[i];
                const int n = b.n, mr = rows > 0 ? rows : rowmult * n, jb = std::min(nb, n - j0);
                T* Y = buf(b, dst);
                T* Tj = b.T2 + (size_t)j0 * mr;
                const T* Li = b.Linv + (size_t)bi * nb * nb;
                if (conj) {
                    // (Y L^H)_j = sum_{i<=j} Y_i L_ji^H
                    g1.push_back(ls_prob<T>(j0 > 0 ? mr : 0, jb, j0, Y, mr, b.L + j0, n, Tj, mr, -1.0, 1.0, 0, 1, 0));
                    g2.push_back(ls_prob<T>(mr, jb, jb, Tj, mr, Li, nb, Y + (size_t)j0 * mr, mr, 1.0, 0.0, 0, 1, 0));
                } else {
                    // (Y L)_j = sum_{i>=j} Y_i L_ij
                    const int j1 = j0 + jb, rest = n - j1;
                    g1.push_back(ls_prob<T>(rest > 0 ? mr : 0, jb, rest, Y + (size_t)j1 * mr, mr, b.L + (size_t)j0 * n + j1, n, Tj, mr,
                                            -1.0, 1.0, 0, 0, 0));
                    g2.push_back(ls_prob<T>(mr, jb, jb, Tj, mr, Li, nb, Y + (size_t)j0 * mr, mr, 1.0, 0.0, 0, 0, 0));
                }
            }
            gemm(0, conj ? 2 : 0, g1);
            gemm(0, conj ? 2 : 0, g2);
        }
    }
    // X <- al * Qf[0:n] Qf[n:2n]^H + be * X
    void qq_update(int qf, double al, double be) {
        std::vector<GemmProblem<T>> g;
        for (const auto& b : blk) {
            const int n = b.n;
            g.push_back(ls_prob<T>(n, n, n, buf(b, qf), 2 * n, buf(b, qf) + n, 2 * n, b.X, n, al, be, 0, 1, 0));
        }
        gemm(0, 2, g);
    }
    // part 1: X0 = S / ||S||_F (after A = Q0 R0 for tall blocks) and the sigma_min estimate of every block:
    // Z = X0^H X0 = L L^H, est_i = ||G L^-H||_F^2 for LS_NPROBE Rademacher rows G  (= sum_k g_k^T Z^-1 g_k ~ NPROBE tr(Z^-1),
    // and sigma_min(X0) >= 1 / sqrt(tr(Z^-1))): ls_l0_from_estimate turns (est_i, Cholesky info_i) into l0
    void build_prepare(bool estimate) {
        bool any_tall = false;
        for (const auto& b : blk) any_tall = any_tall || b.m > b.n;
        if (any_tall) act(LS_QR_TALL, count);
        act(LS_PREP, count);
        if (!estimate) return;
        gram(LS_X, 1, 1.0);
        potrf();
        act(LS_PROBE, count);
        trsm(true, 1, LS_T2, LS_Q, LS_NPROBE);
        act(LS_FRO, count);
    }
    // part 2: the QDWH steps of `sched` and W, P
    void build_iterate(const std::vector<QdwhStep>& sched) {
        bool any_tall = false;
        for (const auto& b : blk) any_tall = any_tall || b.m > b.n;
        for (const QdwhStep& st : sched) {
            if (st.qr) {
                act(LS_STACK, count, 0, 0, 0, sqrt(st.c));
                int qf;
                if (st.c > QDWH_CHOLQR_MAX_C) {
                    act(LS_QR_STACK, count);
                    qf = LS_Q;
                } else {
                    // CholeskyQR2 of B = [sqrt(c) X; I]: B^H B = I + c X^H X
                    gram(LS_X, 1, st.c);
                    act(LS_ADDDIAG, count);
                    potrf();
                    trsm(true, 2, LS_B, LS_Q);
                    gram(LS_Q, 2, 1.0);
                    potrf();
                    trsm(true, 2, LS_Q, LS_B);
                    qf = LS_B;
                }
                qq_update(qf, (st.a - st.b / st.c) / sqrt(st.c), st.b / st.c);
            } else {
                gram(LS_X, 1, st.c);
                act(LS_ADDDIAG, count);
                potrf();
                trsm(true, 1, LS_X, LS_Q);
                trsm(false, 1, LS_Q, LS_B);
                act(LS_AXPBY, count, 0, 0, 0, st.b / st.c, st.a - st.b / st.c);
            }
        }
        // W = X (square) or Q0 X (tall)
        act(LS_COPY, count, LS_X, LS_W, 1);
        if (any_tall) {
            std::vector<GemmProblem<T>> g;
            for (const auto& b : blk)
                g.push_back(ls_prob<T>(b.m > b.n ? b.m : 0, b.n, b.n, b.Q0, b.m, b.X, b.n, b.W, b.m, 1.0, 0.0, 0, 0, 0));
            gemm(0, 0, g);
        }
        // P = sym(X^H S)   (project_hermitian! of W^H A)
        {
            std::vector<GemmProblem<T>> g;
            for (const auto& b : blk)
                g.push_back(ls_prob<T>(b.n, b.n, b.n, b.X, b.n, b.S, b.lds, b.Z, b.n, 1.0, 0.0, 1, 0, 0));
            gemm(2, 0, g);
        }
        act(LS_SYMM, count);
    }
};

// l0 of a block from its estimate (est = ||G L^-H||_F^2 over LS_NPROBE probes; info != 0: the Cholesky of X0^H X0 broke
// down).  Same rule as the single-matrix driver (polar.cu): l0 = 0.3 / sqrt(tr) when that exceeds 1e-7, else eps.
inline double ls_l0_from_estimate(double est, int info) {
    const double eps_l0 = 2.2e-16;
    const double tr = est / LS_NPROBE;
    if (info != 0 || !(tr > 0.0) || !std::isfinite(tr)) return eps_l0;
    const double l0 = 0.3 / std::sqrt(tr);
    if (!(l0 > 1e-7)) return eps_l0;
    return l0 < 0.9 ? l0 : 0.9;
}
// +-1 entry (row r, column c) of a block's probe matrix: a hash of the position, the same on the device and in the CPU replay
inline
#ifdef __CUDACC__
__host__ __device__
#endif
double ls_probe_entry(int r, int c) {
    unsigned h = ((unsigned)c * LS_NPROBE + (unsigned)r) * 2654435761u + 777u;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return (h & 1u) ? 1.0 : -1.0;
}

// GEMM launches of the plan for the largest block (bound for the descriptor storage: launches * blocks)
template <typename T>
inline int ls_gemm_launch_bound(int nmax, int nb, bool any_tall, const std::vector<QdwhStep>& sched) {
    std::vector<LsBlk<T>> one(1);
    one[0] = LsBlk<T>{};
    one[0].m = any_tall ? nmax + 1 : nmax;
    one[0].n = nmax;
    one[0].lda = one[0].m;
    LsPlan<T> pl;
    LsPlanner<T> p(one, nb, pl);
    p.build_prepare(true);
    p.build_iterate(sched);
    return pl.gemm_launches;
}

}  // namespace mak
