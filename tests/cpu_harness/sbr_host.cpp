// CPU harness for csrc/sbr_core.h (band -> tridiagonal bulge chasing and the diamond-blocked
// application of Q2): the product header compiled with g++ and driven by plain loops.
//   order = 0: tasks in sequential (generation) order
//   order = 1: wavefront order (sbr::wavefront), tasks of one wavefront in a seeded random order --
//              must give bit-identical results if the independence rule is right
//   q2mode = 0: one reflector at a time in reverse generation order (reference)
//   q2mode = 1: diamond blocks (group size g) in diamond-wavefront order, random inside a wavefront
#include <algorithm>
#include <random>
#include <vector>

#include "../../matrixalgebrakit.jl_b200/csrc/sbr_core.h"

using namespace mak;

template <typename T>
static int chase_t(int n, int b, const T* A, int lda, double* d, double* e, T* V2, T* tau2, int ldt, int order,
                   unsigned seed) {
    const int ldab = 2 * b;
    std::vector<T> AB((size_t)ldab * n, zero<T>());
    sbr::Band<T> B{n, b, ldab, AB.data()};
    for (int j = 0; j < n; ++j)
        for (int i = j; i < n && i - j <= b; ++i) B.at(i, j) = A[(size_t)j * lda + i];
    for (int j = 0; j < n; ++j) B.at(j, j) = mk<T>(real_(B.at(j, j)));
    sbr::Q2Store<T> Q{n, ldt, V2, tau2};
    std::vector<T> work(2 * b + 2);
    if (order == 0) {
        for (int s = 0; s <= n - 2; ++s)
            for (int k = 0; k < sbr::sweep_ntasks(n, b, s); ++k) sbr::chase_task<T>(B, Q, s, k, work.data());
    } else {
        int maxw = 0;
        for (int s = 0; s <= n - 2; ++s) maxw = std::max(maxw, sbr::wavefront(s, sbr::sweep_ntasks(n, b, s) - 1));
        std::vector<std::vector<std::pair<int, int>>> waves(maxw + 1);
        for (int s = 0; s <= n - 2; ++s)
            for (int k = 0; k < sbr::sweep_ntasks(n, b, s); ++k) waves[sbr::wavefront(s, k)].push_back({s, k});
        std::mt19937 rng(seed);
        for (auto& w : waves) {
            std::shuffle(w.begin(), w.end(), rng);
            for (auto& t : w) sbr::chase_task<T>(B, Q, t.first, t.second, work.data());
        }
    }
    // everything below the first subdiagonal must be exactly zero now
    int bad = 0;
    for (int j = 0; j < n; ++j)
        for (int dd = 2; dd < ldab && j + dd < n; ++dd)
            if (!is_zero(AB[(size_t)j * ldab + dd])) ++bad;
    for (int j = 0; j < n; ++j) {
        d[j] = real_(B.at(j, j));
        if (j + 1 < n) {
            e[j] = real_(B.at(j + 1, j));
            if (imag_(B.at(j + 1, j)) != 0.0) ++bad;
        }
    }
    return bad;
}

template <typename T>
static int apply_q2_t(int n, int b, const T* V2, const T* tau2, int ldt, T* Z, int ldz, int ncols, int q2mode, int g,
                      unsigned seed) {
    sbr::Q2Store<T> Q{n, ldt, const_cast<T*>(V2), const_cast<T*>(tau2)};
    if (q2mode == 0) {
        for (int s = n - 2; s >= 0; --s)
            for (int k = sbr::sweep_ntasks(n, b, s) - 1; k >= 0; --k) {
                const sbr::Task t = sbr::task_geometry(n, b, s, k);
                sbr::apply_reflector<T>(t.L, V2 + (size_t)s * n + t.r0, tau2[(size_t)s * ldt + k], Z + t.r0, ldz, ncols);
            }
        return 0;
    }
    const int nsweeps = n - 1;
    const int ngroups = (nsweeps + g - 1) / g;
    const int kmax = (n + b - 1) / b + 1;
    std::vector<std::vector<std::pair<int, int>>> waves(ngroups + kmax + 1);
    for (int grp = 0; grp < ngroups; ++grp)
        for (int k = 0; k < kmax; ++k) {
            const sbr::DBlock d = sbr::dblock_geometry(n, b, g, grp, k);
            if (d.ns > 0) waves[sbr::diamond_wavefront(ngroups, grp, k)].push_back({grp, k});
        }
    std::mt19937 rng(seed);
    std::vector<T> V((size_t)(b + g) * g), Tm((size_t)g * g), W((size_t)g * ncols);
    for (auto& w : waves) {
        std::shuffle(w.begin(), w.end(), rng);
        for (auto& t : w) {
            const sbr::DBlock d = sbr::dblock_geometry(n, b, g, t.first, t.second);
            if (d.rows > b + g) return -1;
            sbr::dblock_build<T>(n, b, Q, d, t.second, V.data(), b + g, Tm.data(), g);
            sbr::dblock_apply<T>(d, V.data(), b + g, Tm.data(), g, Z + d.base, ldz, ncols, W.data());
        }
    }
    return 0;
}

extern "C" {
// dtype 0: double, 1: interleaved complex.  A: n x n (lda), only the lower band of width b is read.
// V2: n x n, tau2: ldt x n (ldt >= ceil(n/b) + 1), both zero-initialised by the caller.
int sbr_host_chase(int dtype, int n, int b, const void* A, int lda, double* d, double* e, void* V2, void* tau2, int ldt,
                   int order, unsigned seed) {
    if (dtype == 0) return chase_t<double>(n, b, (const double*)A, lda, d, e, (double*)V2, (double*)tau2, ldt, order, seed);
    return chase_t<cplx>(n, b, (const cplx*)A, lda, d, e, (cplx*)V2, (cplx*)tau2, ldt, order, seed);
}
int sbr_host_apply_q2(int dtype, int n, int b, const void* V2, const void* tau2, int ldt, void* Z, int ldz, int ncols,
                      int q2mode, int g, unsigned seed) {
    if (dtype == 0)
        return apply_q2_t<double>(n, b, (const double*)V2, (const double*)tau2, ldt, (double*)Z, ldz, ncols, q2mode, g, seed);
    return apply_q2_t<cplx>(n, b, (const cplx*)V2, (const cplx*)tau2, ldt, (cplx*)Z, ldz, ncols, q2mode, g, seed);
}
}
