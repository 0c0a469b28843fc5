// Second stage of the two-stage Hermitian tridiagonalisation (round-2 work):
// band (lower bandwidth b) -> tridiagonal by bulge chasing, and the blocked ("diamond") application
// of the chase reflectors Q2 to the eigenvector matrix.
//
// This header holds the task GEOMETRY (which b x b blocks a task touches, the wavefront rule that
// makes tasks independent, the parallelogram layout of a reflector block) and sequential task
// bodies.  Everything is `__host__ __device__`: tests/cpu_harness/sbr_host.cpp drives it with g++
// (sequential order, randomised wavefront order, reference vs diamond back-transform), and the CUDA
// kernels are to call the same geometry functions, so the index logic is validated before any GPU
// time is spent.  Published statements of the algorithm: Schwarz 1968 / Murata-Horikoshi 1975
// (chasing), Haidar-Ltaief-Dongarra 2011 (tile kernels, diamond-shaped blocking of Q2).
//
// Conventions
//   band storage   AB[(i - j) + j*ldab] = B[i, j],  0 <= i - j < ldab,  ldab >= 2b  (room for the bulge)
//   reflector      H = I - tau v v^H, v[0] = 1, generated so that H^H x = beta e1 with beta >= 0
//                  (larfgp_scalars, the convention of every other kernel in this library)
//   chase          B <- H^H B H, tasks (s, k): sweep s = column being reduced, k = block along the chase
//                  k = 0 : rows r0 = s+1 .. : eliminate B[s+2.., s], two-sided update of the diagonal block
//                  k >= 1: r0 = s+1+k b, c0 = r0-b: G = B[r0.., c0..c0+b) <- G H_prev, eliminate G[1.., 0]
//                          with a new H, G <- H^H G, two-sided update of the diagonal block at r0
//   result         T = Q2^H B Q2,  Q2 = product of all H in generation order (s ascending, k ascending)
//   eigenvectors   X = Q2 Z: reflectors applied to Z in REVERSE generation order
//   storage of Q2  V2[r + s*ldv] = component of the sweep-s reflector that acts on row r (v[0] = 1 stored),
//                  tau2[k + s*ldt]
#pragma once
#include "scalar.h"

#ifdef __CUDACC__
#define SBR_HD __host__ __device__ __forceinline__
#else
#define SBR_HD inline
#endif

namespace mak {
namespace sbr {

struct Task {
    int r0;   // first row the task's reflector acts on
    int c0;   // first column of the off-diagonal block (k >= 1), or the reduced column s (k = 0)
    int L;    // reflector length = rows in the block (0: task does not exist)
    int Lp;   // length of the previous reflector of the chain (k >= 1)
};

// geometry of task (s, k) for an n x n matrix of bandwidth b
SBR_HD Task task_geometry(int n, int b, int s, int k) {
    Task t;
    t.r0 = s + 1 + k * b;
    t.c0 = (k == 0) ? s : t.r0 - b;
    int L = n - t.r0;
    t.L = L < 0 ? 0 : (L < b ? L : b);
    t.Lp = (k == 0) ? 0 : b;   // r0 <= n-1 implies the previous block was full
    return t;
}
// number of tasks of sweep s (k = 0 .. ntasks-1); sweeps are s = 0 .. n-2
SBR_HD int sweep_ntasks(int n, int b, int s) {
    const int rows = n - 1 - s;   // rows below the diagonal in column s
    return rows <= 0 ? 0 : (rows + b - 1) / b;
}
// Tasks with equal wavefront index are independent: (s+1, k) needs (s, k+1) and (s+1, k-1).
SBR_HD int wavefront(int s, int k) { return 2 * s + k; }

template <typename T>
struct Band {
    int n, b, ldab;
    T* AB;
    SBR_HD T& at(int i, int j) const { return AB[(size_t)j * ldab + (i - j)]; }   // i >= j
};

template <typename T>
struct Q2Store {
    int ldv, ldt;
    T* V2;     // ldv x n
    T* tau2;   // ldt x n
};

// D (L x L Hermitian, lower part stored at B[r0.., r0..]) <- H^H D H,  H = I - tau v v^H
// work: L scalars
template <typename T>
SBR_HD void herm_two_sided(const Band<T>& B, int r0, int L, const T* v, T tau, T* work) {
    if (is_zero(tau)) return;
    // p = D v
    for (int i = 0; i < L; ++i) {
        T p = zero<T>();
        for (int j = 0; j <= i; ++j) fma_(p, B.at(r0 + i, r0 + j), v[j]);
        for (int j = i + 1; j < L; ++j) fmac_(p, B.at(r0 + j, r0 + i), v[j]);   // conj(D[j,i]) v_j
        work[i] = p;
    }
    T alpha = zero<T>();
    for (int i = 0; i < L; ++i) fmac_(alpha, v[i], work[i]);                    // v^H p (real)
    // w = tau p - (|tau|^2 alpha / 2) v
    const double half = 0.5 * abs2_(tau) * real_(alpha);
    for (int i = 0; i < L; ++i) work[i] = sub_(mul_(tau, work[i]), scale_(v[i], half));
    // D <- D - v w^H - w v^H (lower part), diagonal kept real
    for (int j = 0; j < L; ++j)
        for (int i = j; i < L; ++i) {
            T d = B.at(r0 + i, r0 + j);
            d = sub_(d, mul_(v[i], conj_(work[j])));
            d = sub_(d, mul_(work[i], conj_(v[j])));
            if (i == j) d = mk<T>(real_(d));
            B.at(r0 + i, r0 + j) = d;
        }
}

// generate the reflector that maps x (L entries, stride 1 at px) to beta e1; stores v (v[0] = 1), returns tau
template <typename T>
SBR_HD T make_reflector(T* px, int L, T* v) {
    double sigma = 0.0;
    for (int i = 1; i < L; ++i) sigma += abs2_(px[i]);
    double beta; T tau, scale;
    larfgp_scalars<T>(px[0], sigma, beta, tau, scale);
    v[0] = one<T>();
    for (int i = 1; i < L; ++i) { v[i] = mul_(px[i], scale); px[i] = zero<T>(); }
    px[0] = mk<T>(beta);
    return tau;
}

// One chase task.  work: 2b scalars.
template <typename T>
SBR_HD void chase_task(const Band<T>& B, const Q2Store<T>& Q, int s, int k, T* work) {
    const Task t = task_geometry(B.n, B.b, s, k);
    if (t.L <= 0) return;
    T* v = Q.V2 + (size_t)s * Q.ldv + t.r0;
    T tau;
    if (k == 0) {
        // column s: rows r0 .. r0+L-1 are contiguous in band storage
        tau = make_reflector<T>(&B.at(t.r0, s), t.L, v);
    } else {
        const T* vp = Q.V2 + (size_t)s * Q.ldv + t.c0;     // previous reflector of this sweep (length b)
        const T taup = Q.tau2[(size_t)s * Q.ldt + (k - 1)];
        // G <- G H_prev = G - tau (G v) v^H
        if (!is_zero(taup)) {
            for (int i = 0; i < t.L; ++i) {
                T g = zero<T>();
                for (int j = 0; j < t.Lp; ++j) fma_(g, B.at(t.r0 + i, t.c0 + j), vp[j]);
                g = mul_(taup, g);
                for (int j = 0; j < t.Lp; ++j) {
                    T& e = B.at(t.r0 + i, t.c0 + j);
                    e = sub_(e, mul_(g, conj_(vp[j])));
                }
            }
        }
        // new reflector from column 0 of G (contiguous in band storage), G <- H^H G on the other columns
        tau = make_reflector<T>(&B.at(t.r0, t.c0), t.L, v);
        if (!is_zero(tau)) {
            const T ctau = conj_(tau);
            for (int j = 1; j < t.Lp; ++j) {
                T d = zero<T>();
                for (int i = 0; i < t.L; ++i) fmac_(d, v[i], B.at(t.r0 + i, t.c0 + j));
                d = mul_(ctau, d);
                for (int i = 0; i < t.L; ++i) {
                    T& e = B.at(t.r0 + i, t.c0 + j);
                    e = sub_(e, mul_(v[i], d));
                }
            }
        }
    }
    Q.tau2[(size_t)s * Q.ldt + k] = tau;
    herm_two_sided<T>(B, t.r0, t.L, v, tau, work);
}

// ---------------------------------------------------------------------------------------
// Q2 application
// ---------------------------------------------------------------------------------------
// reference: one reflector at a time in reverse generation order.  Z: n x ncols, ld ldz.
template <typename T>
SBR_HD void apply_reflector(int L, const T* v, T tau, T* Zrows, int ldz, int ncols) {
    if (is_zero(tau)) return;
    for (int c = 0; c < ncols; ++c) {
        T* z = Zrows + (size_t)c * ldz;
        T d = zero<T>();
        for (int i = 0; i < L; ++i) fmac_(d, v[i], z[i]);
        d = mul_(tau, d);
        for (int i = 0; i < L; ++i) z[i] = sub_(z[i], mul_(v[i], d));
    }
}

// Diamond blocking: sweeps are cut into groups of g; block (grp, k) = reflectors (s, k), s in the group.
struct DBlock {
    int s0;     // first sweep of the group
    int ns;     // sweeps of the group that have a reflector at this k (a prefix: lengths shrink with s)
    int base;   // first row of the parallelogram = s0 + 1 + k b
    int rows;   // rows spanned = (last reflector's first row + its length) - base
};
SBR_HD DBlock dblock_geometry(int n, int b, int g, int grp, int k) {
    DBlock d;
    d.s0 = grp * g;
    d.base = d.s0 + 1 + k * b;
    int ns = 0, end = d.base;
    for (int j = 0; j < g; ++j) {
        const int s = d.s0 + j;
        if (s > n - 2) break;
        const Task t = task_geometry(n, b, s, k);
        if (t.L <= 0) break;
        ns = j + 1;
        if (t.r0 + t.L > end) end = t.r0 + t.L;
    }
    d.ns = ns;
    d.rows = end - d.base;
    return d;
}
// blocks with equal diamond wavefront are independent; (grp, k) needs (grp, k-1), (grp+1, k), (grp+1, k-1)
SBR_HD int diamond_wavefront(int ngroups, int grp, int k) { return (ngroups - 1 - grp) + k; }

// explicit parallelogram V (rows x ns, zero outside the reflectors) and the compact-WY T (ns x ns, upper)
// of  H_{s0} H_{s0+1} ... H_{s0+ns-1} = I - V T V^H
template <typename T>
SBR_HD void dblock_build(int n, int b, const Q2Store<T>& Q, const DBlock& d, int k, T* V, int ldvv, T* Tm, int ldtt) {
    for (int j = 0; j < d.ns; ++j) {
        const Task t = task_geometry(n, b, d.s0 + j, k);
        const T* v = Q.V2 + (size_t)(d.s0 + j) * Q.ldv + t.r0;
        T* col = V + (size_t)j * ldvv;
        for (int i = 0; i < d.rows; ++i) col[i] = zero<T>();
        for (int i = 0; i < t.L; ++i) col[(t.r0 - d.base) + i] = v[i];
    }
    for (int j = 0; j < d.ns; ++j) {
        const T tj = Q.tau2[(size_t)(d.s0 + j) * Q.ldt + k];
        for (int i = 0; i < d.ns; ++i) Tm[(size_t)j * ldtt + i] = zero<T>();
        // z = V[:, 0:j]^H v_j ;  T[0:j, j] = -tau_j T[0:j,0:j] z
        for (int i = 0; i < j; ++i) {
            T z = zero<T>();
            for (int r = 0; r < d.rows; ++r) fmac_(z, V[(size_t)i * ldvv + r], V[(size_t)j * ldvv + r]);
            Tm[(size_t)j * ldtt + i] = z;   // stash z in column j
        }
        for (int i = 0; i < j; ++i) {
            T acc = zero<T>();
            for (int p = i; p < j; ++p) fma_(acc, Tm[(size_t)p * ldtt + i], Tm[(size_t)j * ldtt + p]);
            // rows are consumed top-down: row i only needs z_p for p >= i, which are still untouched
            Tm[(size_t)j * ldtt + i] = neg_(mul_(tj, acc));
        }
        Tm[(size_t)j * ldtt + j] = tj;
    }
}

// Zrows (rows x ncols) <- (I - V T V^H) Zrows ; W: ns x ncols scratch (ld ns)
template <typename T>
SBR_HD void dblock_apply(const DBlock& d, const T* V, int ldvv, const T* Tm, int ldtt, T* Zrows, int ldz, int ncols, T* W) {
    for (int c = 0; c < ncols; ++c) {
        T* z = Zrows + (size_t)c * ldz;
        T* w = W + (size_t)c * d.ns;
        for (int j = 0; j < d.ns; ++j) {
            T a = zero<T>();
            for (int r = 0; r < d.rows; ++r) fmac_(a, V[(size_t)j * ldvv + r], z[r]);
            w[j] = a;
        }
        // w <- T w (upper triangular, in place top-down)
        for (int i = 0; i < d.ns; ++i) {
            T a = zero<T>();
            for (int p = i; p < d.ns; ++p) fma_(a, Tm[(size_t)p * ldtt + i], w[p]);
            w[i] = a;
        }
        for (int j = 0; j < d.ns; ++j)
            for (int r = 0; r < d.rows; ++r) z[r] = sub_(z[r], mul_(V[(size_t)j * ldvv + r], w[j]));
    }
}

}  // namespace sbr
}  // namespace mak
