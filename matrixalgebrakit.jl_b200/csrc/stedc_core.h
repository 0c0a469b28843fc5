// Divide-and-conquer symmetric tridiagonal eigensolver (stedc class) — work-item bodies.
//
// Every function here is `__host__ __device__`: the CUDA kernels in stedc.cu call them with
// thread indices, and tests/cpu_harness compiles the same header with g++ to exercise the
// deflation / secular-equation logic on the CPU box (there is no GPU where the code is built).
//
// Algorithm (Cuppen 1981; Gu & Eisenstat 1995; the LAPACK dstedc/dlaed0-4 family is the
// published statement): bottom-up binary tree over index ranges, leaves (<= DC_LEAF) by
// implicit QL, each merge = rank-one update D + rho z z^T: sort, deflate, secular roots,
// Loewner re-derivation of z, eigenvectors of the update, one GEMM pair per merge.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define MAK_HD __host__ __device__ __forceinline__
#else
#define MAK_HD inline
#endif

namespace mak {
namespace dc {

constexpr int DC_LEAF = 32;
constexpr double DC_EPS = 1.1102230246251565e-16;  // 2^-53, LAPACK dlamch('E')

// one merge of two adjacent solved sub-problems [lo,mid) and [mid,hi)
struct Merge {
    int lo, mid, hi;
    int K;           // non-deflated count            (written by deflate)
    int k1, k2, k3;  // non-deflated per column type   (written by deflate)
    int nrot;        // rotations recorded             (written by deflate)
    double rho;      // 2*|e[mid-1]| after z normalisation
    double sgn;      // sign(e[mid-1])
};

// device work arrays, all of length n unless noted, indexed by GLOBAL position lo+local
struct Ctx {
    int n;
    double* D;       // current eigenvalues (each sub-problem ascending)
    double* Dn;      // next level's eigenvalues
    double* z;       // rank-one vector (normalised)
    int* perm;       // perm[lo+r] = local index of the r-th smallest D in [lo,hi)
    double* dl;      // after deflate: [0,K) non-deflated d ascending, [K,N) deflated d ascending
    double* zl;      // non-deflated z (same order as dl[0:K])
    int* src;        // src[lo+j] = local source column of entry j of dl
    int* ctype;      // ctype[lo+j] = type (1 top, 2 both, 3 bottom) of entry j of dl
    int* rowpos;     // rowpos[lo+j], j<K: row of S / column of the packed operand
    int* rot_p;      // rotation list: source columns (local)
    int* rot_q;
    double* rot_c;
    double* rot_s;
    int* rot_tp;     // column types before the rotation
    int* rot_tq;
    double* tau;     // secular roots: lambda_j = dl[orig_j] + tau_j
    int* orig;
    double* zhat;    // Loewner-corrected z
    int* pos;        // pos[lo+j]: final (sorted) local column of entry j
};

MAK_HD double sign1(double x) { return x < 0.0 ? -1.0 : 1.0; }

// ---------------------------------------------------------------------------------------
// leaf solver: implicit QL with Wilkinson shift on (d,e), accumulating into Z (n x n, ld ldz,
// identity on entry).  e has length n (e[n-1] is scratch).  Ascending sort at the end.
// Returns 0, or l+1 if eigenvalue l failed to converge in 60 iterations.
// ---------------------------------------------------------------------------------------
MAK_HD int leaf_ql(int n, double* d, double* e, double* Z, int ldz) {
    for (int l = 0; l < n; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; ++m) {
                double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= DC_EPS * dd) break;
            }
            if (m != l) {
                if (iter++ == 60) return l + 1;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + copysign(r, g));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i], b = c * e[i];
                    r = hypot(f, g);
                    e[i + 1] = r;
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - b;
                    double* zi = Z + (size_t)i * ldz;
                    double* zi1 = Z + (size_t)(i + 1) * ldz;
                    for (int k = 0; k < n; ++k) {
                        double f2 = zi1[k];
                        zi1[k] = s * zi[k] + c * f2;
                        zi[k] = c * zi[k] - s * f2;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    // selection sort ascending, swapping columns
    for (int i = 0; i < n - 1; ++i) {
        int k = i;
        double p = d[i];
        for (int j = i + 1; j < n; ++j)
            if (d[j] < p) { k = j; p = d[j]; }
        if (k != i) {
            d[k] = d[i];
            d[i] = p;
            double* a = Z + (size_t)i * ldz;
            double* b = Z + (size_t)k * ldz;
            for (int r = 0; r < n; ++r) { double t = a[r]; a[r] = b[r]; b[r] = t; }
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// merge step 1: z = [last row of Q1 ; sgn * first row of Q2] / sqrt(2)      (item = local i)
// ---------------------------------------------------------------------------------------
MAK_HD void merge_z_item(const Ctx& c, const Merge& mg, const double* Z, int ldz, int i) {
    const int N1 = mg.mid - mg.lo;
    const double is2 = 0.70710678118654752440;
    double v;
    if (i < N1) v = Z[(size_t)(mg.lo + i) * ldz + (mg.mid - 1)];
    else v = mg.sgn * Z[(size_t)(mg.lo + i) * ldz + mg.mid];
    c.z[mg.lo + i] = v * is2;
}

// merge step 2: rank of local element i in the merged ascending order (stable: child 1 first)
MAK_HD void merge_rank_item(const Ctx& c, const Merge& mg, int i) {
    const int N1 = mg.mid - mg.lo, N2 = mg.hi - mg.mid;
    const double* d1 = c.D + mg.lo;
    const double* d2 = c.D + mg.mid;
    int rank;
    if (i < N1) {
        double v = d1[i];
        int lo = 0, hi = N2;  // count of d2 < v
        while (lo < hi) { int md = (lo + hi) >> 1; if (d2[md] < v) lo = md + 1; else hi = md; }
        rank = i + lo;
    } else {
        double v = d2[i - N1];
        int lo = 0, hi = N1;  // count of d1 <= v
        while (lo < hi) { int md = (lo + hi) >> 1; if (d1[md] <= v) lo = md + 1; else hi = md; }
        rank = (i - N1) + lo;
    }
    c.perm[mg.lo + rank] = i;
}

// ---------------------------------------------------------------------------------------
// merge step 3: deflation scan (serial per merge; dlaed2's logic)
// ---------------------------------------------------------------------------------------
MAK_HD void deflate_scan(const Ctx& c, Merge& mg) {
    const int lo = mg.lo, N = mg.hi - mg.lo, N1 = mg.mid - mg.lo;
    const double* D = c.D + lo;
    double* z = c.z + lo;
    const int* perm = c.perm + lo;
    double* dl = c.dl + lo;
    double* zl = c.zl + lo;
    int* src = c.src + lo;
    int* ctype = c.ctype + lo;
    int* rowpos = c.rowpos + lo;
    double dmax = 0.0, zmax = 0.0;
    for (int i = 0; i < N; ++i) {
        dmax = fmax(dmax, fabs(D[i]));
        zmax = fmax(zmax, fabs(z[i]));
    }
    const double rho = mg.rho;
    const double tol = 8.0 * DC_EPS * fmax(dmax, zmax);
    int K = 0, K2 = N;  // deflated entries fill [K2, N) from the back
    int nrot = 0;
    mg.k1 = mg.k2 = mg.k3 = 0;
    // We keep a scratch copy of d values that rotations modify: use dl's tail region carefully.
    // Working arrays: wd[i] (current d of local column i) and wt[i] (current type) live in
    // c.tau / c.orig storage temporarily (they are produced later in the pipeline).
    double* wd = c.tau + lo;
    int* wt = c.orig + lo;
    for (int i = 0; i < N; ++i) { wd[i] = D[i]; wt[i] = (i < N1) ? 1 : 3; }

    if (rho * zmax <= tol) {
        // everything deflates: eigenvalues are the sorted d's, vectors are Q's columns
        for (int r = 0; r < N; ++r) {
            int i = perm[r];
            dl[r] = wd[i]; src[r] = i; ctype[r] = wt[i];
        }
        mg.K = 0; mg.nrot = 0;
        return;
    }
    // deflated entries are collected in scan order (ascending d, with insertion to keep order)
    int pj = -1;
    for (int r = 0; r < N; ++r) {
        int nj = perm[r];
        if (rho * fabs(z[nj]) <= tol) {
            // deflate: tiny z component.  Insert into the deflated list keeping ascending order.
            --K2;
            // deflated list grows downward in dl[K2..N); we fill it in scan (ascending) order, so
            // store reversed now and fix the order at the end.
            dl[K2] = wd[nj]; src[K2] = nj; ctype[K2] = wt[nj];
            continue;
        }
        if (pj < 0) { pj = nj; continue; }
        // check whether d[pj] and d[nj] are close enough to deflate pj by a rotation
        double s = z[pj], cc = z[nj];
        double tau = hypot(cc, s);
        double t = wd[nj] - wd[pj];
        cc /= tau;
        s = -s / tau;
        if (fabs(t * cc * s) <= tol) {
            z[nj] = tau;
            z[pj] = 0.0;
            c.rot_p[lo + nrot] = pj; c.rot_q[lo + nrot] = nj;
            c.rot_c[lo + nrot] = cc; c.rot_s[lo + nrot] = s;
            c.rot_tp[lo + nrot] = wt[pj]; c.rot_tq[lo + nrot] = wt[nj];
            ++nrot;
            if (wt[pj] != wt[nj]) { wt[pj] = 2; wt[nj] = 2; }
            double tnew = wd[pj] * cc * cc + wd[nj] * s * s;
            wd[nj] = wd[pj] * s * s + wd[nj] * cc * cc;
            wd[pj] = tnew;
            // pj is deflated with value wd[pj]; keep the deflated list sorted (insertion)
            --K2;
            int i = K2;
            // list is stored reversed-in-scan-order: entries at higher index were inserted earlier
            // (smaller d).  Insert so that values DEcrease with increasing index... we normalise
            // ordering after the scan instead; just append here.
            dl[i] = wd[pj]; src[i] = pj; ctype[i] = wt[pj];
            pj = nj;
        } else {
            dl[K] = wd[pj]; zl[K] = z[pj]; src[K] = pj; ctype[K] = wt[pj];
            ++K;
            pj = nj;
        }
    }
    if (pj >= 0) {
        dl[K] = wd[pj]; zl[K] = z[pj]; src[K] = pj; ctype[K] = wt[pj];
        ++K;
    }
    // deflated block [K2, N) was filled back-to-front in (nearly) ascending order: reverse it,
    // then insertion-sort (rotations can perturb the order slightly)
    for (int a = K2, b = N - 1; a < b; ++a, --b) {
        double td = dl[a]; dl[a] = dl[b]; dl[b] = td;
        int ti = src[a]; src[a] = src[b]; src[b] = ti;
        ti = ctype[a]; ctype[a] = ctype[b]; ctype[b] = ti;
    }
    for (int a = K2 + 1; a < N; ++a) {
        double td = dl[a]; int ts = src[a], tt = ctype[a];
        int b = a - 1;
        while (b >= K2 && dl[b] > td) { dl[b + 1] = dl[b]; src[b + 1] = src[b]; ctype[b + 1] = ctype[b]; --b; }
        dl[b + 1] = td; src[b + 1] = ts; ctype[b + 1] = tt;
    }
    // non-deflated d's must be ascending for the secular solver (a rotation can make the new
    // d[nj] overtake nothing that follows, but keep a guard)
    for (int a = 1; a < K; ++a) {
        double td = dl[a], tz = zl[a]; int ts = src[a], tt = ctype[a];
        int b = a - 1;
        while (b >= 0 && dl[b] > td) {
            dl[b + 1] = dl[b]; zl[b + 1] = zl[b]; src[b + 1] = src[b]; ctype[b + 1] = ctype[b]; --b;
        }
        dl[b + 1] = td; zl[b + 1] = tz; src[b + 1] = ts; ctype[b + 1] = tt;
    }
    // type-grouped row positions of the non-deflated entries: [type1 | type2 | type3]
    int k1 = 0, k2 = 0, k3 = 0;
    for (int j = 0; j < K; ++j) {
        if (ctype[j] == 1) ++k1; else if (ctype[j] == 2) ++k2; else ++k3;
    }
    int p1 = 0, p2 = k1, p3 = k1 + k2;
    for (int j = 0; j < K; ++j) {
        if (ctype[j] == 1) rowpos[j] = p1++; else if (ctype[j] == 2) rowpos[j] = p2++; else rowpos[j] = p3++;
    }
    mg.K = K; mg.k1 = k1; mg.k2 = k2; mg.k3 = k3; mg.nrot = nrot;
}

// ---------------------------------------------------------------------------------------
// merge step 4: apply the recorded Givens rotations to row r of Q (serial over the list).
// Columns are read through their support mask (type 1: rows < N1, type 3: rows >= N1).
// ---------------------------------------------------------------------------------------
MAK_HD void rotate_row_item(const Ctx& c, const Merge& mg, double* Z, int ldz, int r) {
    const int lo = mg.lo, N1 = mg.mid - mg.lo;
    const bool top = r < N1;
    for (int t = 0; t < mg.nrot; ++t) {
        int p = c.rot_p[lo + t], q = c.rot_q[lo + t];
        double cc = c.rot_c[lo + t], s = c.rot_s[lo + t];
        int tp = c.rot_tp[lo + t], tq = c.rot_tq[lo + t];
        double* zp = Z + (size_t)(lo + p) * ldz + lo + r;
        double* zq = Z + (size_t)(lo + q) * ldz + lo + r;
        double a = ((tp == 1 && !top) || (tp == 3 && top)) ? 0.0 : *zp;
        double b = ((tq == 1 && !top) || (tq == 3 && top)) ? 0.0 : *zq;
        if (tp == tq && ((tp == 1 && !top) || (tp == 3 && top))) continue;  // both outside support
        // drot(x=Q[:,pj], y=Q[:,nj], c, s): x' = c x + s y ; y' = c y - s x
        *zp = cc * a + s * b;
        *zq = cc * b - s * a;
    }
}

// ---------------------------------------------------------------------------------------
// merge step 5: secular equation root j of  1 + rho * sum_i zl_i^2 / (dl_i - lambda) = 0
//   roots interlace: lambda_j in (dl_j, dl_{j+1}), last root in (dl_{K-1}, dl_{K-1}+rho*|z|^2).
//   Output: orig (index of the nearer pole) and tau with lambda = dl[orig] + tau; all later
//   differences dl_i - lambda are formed as (dl_i - dl[orig]) - tau (relative accuracy).
// ---------------------------------------------------------------------------------------
struct SecEval { double f, psi, dpsi, phi, dphi, erretm; };

MAK_HD SecEval sec_eval(int K, const double* d, const double* z, double rho, int io, double tau, int ks) {
    // psi: poles 0..ks, phi: poles ks+1..K-1
    SecEval e;
    double psi = 0.0, dpsi = 0.0, phi = 0.0, dphi = 0.0, err = 0.0;
    const double dio = d[io];
    for (int i = 0; i <= ks; ++i) {
        double del = (d[i] - dio) - tau;
        double t = z[i] / del;
        double zt = z[i] * t;
        psi += zt; dpsi += t * t; err += fabs(zt);
    }
    for (int i = K - 1; i > ks; --i) {
        double del = (d[i] - dio) - tau;
        double t = z[i] / del;
        double zt = z[i] * t;
        phi += zt; dphi += t * t; err += fabs(zt);
    }
    e.psi = rho * psi; e.dpsi = rho * dpsi; e.phi = rho * phi; e.dphi = rho * dphi;
    e.f = 1.0 + e.psi + e.phi;
    e.erretm = 8.0 * (rho * err + 1.0) + fabs(tau) * (e.dpsi + e.dphi);
    return e;
}

MAK_HD void secular_root(int K, int j, const double* d, const double* z, double rho, double* tau_out,
                         int* orig_out) {
    if (K == 1) {
        *orig_out = 0;
        *tau_out = rho * z[0] * z[0];
        return;
    }
    const bool last = (j == K - 1);
    int io;          // origin pole
    double lo, hi;   // bracket on tau (relative to d[io])
    int ks;          // split index for psi/phi
    if (!last) {
        ks = j;
        const double del = d[j + 1] - d[j];
        // f at the midpoint decides which pole is nearer to the root
        SecEval em = sec_eval(K, d, z, rho, j, 0.5 * del, ks);
        if (em.f >= 0.0) { io = j; lo = 0.0; hi = 0.5 * del; }
        else { io = j + 1; lo = -0.5 * del; hi = 0.0; }
    } else {
        ks = K - 2;
        double zz = 0.0;
        for (int i = 0; i < K; ++i) zz += z[i] * z[i];
        io = K - 1; lo = 0.0; hi = rho * zz;
    }
    // initial guess: solve the two-pole model with the remaining poles frozen at the midpoint
    double tau = 0.5 * (lo + hi);
    if (hi <= lo) { *orig_out = io; *tau_out = tau; return; }
    for (int it = 0; it < 80; ++it) {
        SecEval e = sec_eval(K, d, z, rho, io, tau, ks);
        if (fabs(e.f) <= DC_EPS * e.erretm) break;
        if (e.f < 0.0) lo = fmax(lo, tau); else hi = fmin(hi, tau);
        if (!(hi - lo > 2.0 * DC_EPS * fmax(fabs(lo), fabs(hi)))) { tau = 0.5 * (lo + hi); break; }
        // two-pole rational model ("middle way"): psi ~ s + a/(d_p - x), phi ~ r + b/(d_q - x)
        const int ip = ks, iq = ks + 1;
        const double dp = (d[ip] - d[io]) - tau;  // d_p - lambda
        const double dq = (d[iq] - d[io]) - tau;  // d_q - lambda
        double eta;
        {
            const double a = e.dpsi * dp * dp, b = e.dphi * dq * dq;
            const double cc = e.f - e.dpsi * dp - e.dphi * dq;  // 1 + s + r
            const double B = cc * (dp + dq) + a + b;
            const double Cc = dp * dq * e.f;
            // cc*eta^2 - B*eta + Cc = 0
            double disc = B * B - 4.0 * cc * Cc;
            if (disc < 0.0) disc = 0.0;
            const double sq = sqrt(disc);
            double e1, e2;
            if (cc == 0.0) {
                e1 = e2 = (B != 0.0) ? Cc / B : 0.0;
            } else {
                // numerically stable pair of roots
                double qd = (B >= 0.0) ? 0.5 * (B + sq) : 0.5 * (B - sq);
                e1 = (qd != 0.0) ? Cc / qd : 0.0;
                e2 = qd / cc;
            }
            // pick the root that lands strictly inside the bracket; prefer the smaller step
            const double t1 = tau + e1, t2 = tau + e2;
            const bool ok1 = (t1 > lo && t1 < hi), ok2 = (t2 > lo && t2 < hi);
            if (ok1 && ok2) eta = (fabs(e1) <= fabs(e2)) ? e1 : e2;
            else if (ok1) eta = e1;
            else if (ok2) eta = e2;
            else {
                // Newton step, else bisection
                double dw = e.dpsi + e.dphi;
                double en = (dw > 0.0) ? -e.f / dw : 0.0;
                double tn = tau + en;
                eta = (tn > lo && tn < hi) ? en : (0.5 * (lo + hi) - tau);
            }
        }
        tau += eta;
    }
    *orig_out = io;
    *tau_out = tau;
}

// d_i - lambda_j with full relative accuracy
MAK_HD double sec_delta(const double* d, const double* tau, const int* orig, int i, int j) {
    return (d[i] - d[orig[j]]) - tau[j];
}

// ---------------------------------------------------------------------------------------
// merge step 6: Loewner formula (Gu & Eisenstat): zhat_i^2 = prod_j (lambda_j - d_i) /
//   (rho * prod_{j != i} (d_j - d_i)), sign taken from z_i
// ---------------------------------------------------------------------------------------
MAK_HD void zhat_item(int K, const double* d, const double* z, double rho, const double* tau, const int* orig,
                      double* zhat, int i) {
    // lambda_i - d_i first, then ratios (lambda_j - d_i)/(d_j - d_i), each close to 1 in magnitude
    double prod = -sec_delta(d, tau, orig, i, i) / rho;
    for (int j = 0; j < K; ++j) {
        if (j == i) continue;
        prod *= (-sec_delta(d, tau, orig, i, j)) / (d[j] - d[i]);
    }
    zhat[i] = copysign(sqrt(fabs(prod)), z[i]);
}

// merge step 7: final (ascending) position of every entry: roots [0,K) vs deflated [K,N)
MAK_HD void final_pos_item(const Ctx& c, const Merge& mg, int j) {
    const int lo = mg.lo, N = mg.hi - mg.lo, K = mg.K;
    const double* dl = c.dl + lo;
    const double* tau = c.tau + lo;
    const int* orig = c.orig + lo;
    int rank;
    double v;
    if (j < K) {
        v = dl[orig[j]] + tau[j];
        int a = K, b = N;  // count of deflated < v
        while (a < b) { int md = (a + b) >> 1; if (dl[md] < v) a = md + 1; else b = md; }
        rank = j + (a - K);
    } else {
        v = dl[j];
        int a = 0, b = K;  // count of roots <= v
        while (a < b) {
            int md = (a + b) >> 1;
            double lm = dl[orig[md]] + tau[md];
            if (lm <= v) a = md + 1; else b = md;
        }
        rank = (j - K) + a;
    }
    c.pos[lo + j] = rank;
    c.Dn[lo + rank] = v;
}

}  // namespace dc
}  // namespace mak
