// QDWH step schedule (host only, plain C++: shared by the single-matrix driver in polar.cu, the lock-step batched
// planner in polar_lockstep_plan.h and the CPU logic tests).
#pragma once
#include <vector>
#include <cmath>
namespace mak {
struct QdwhStep { double a, b, c; bool qr; };
constexpr double QDWH_CHOLQR_MAX_C = 1e12;

// QR-type step when c > cmax.  The Cholesky-type step forms X Z^-1 with kappa(Z) <= 1 + c, i.e. an error of
// ~ c*eps in X; the classical threshold is c > 100.  The parity bound on this path is 10*n*eps, so for
// large n the threshold is raised to n/8 (error <= n*eps/8): on Gaussian 8192^2 input the second
// iteration (c ~ 2e2) becomes a Cholesky step, 90 ms instead of 227 ms.
inline std::vector<QdwhStep> qdwh_schedule(double l, int maxiter, double cmax = 100.0) {
    std::vector<QdwhStep> v;
    for (int it = 0; it < maxiter; ++it) {
        if (fabs(1.0 - l) <= 1e-15) {
            // one Halley step past the nominal convergence point costs little and polishes X^H X = I
            if (!v.empty() && v.back().c > 3.0 + 1e-6) v.push_back(QdwhStep{3.0, 1.0, 3.0, false});
            break;
        }
        double l2 = l * l;
        double dd = cbrt(4.0 * (1.0 - l2) / (l2 * l2));
        double a = sqrt(1.0 + dd) + 0.5 * sqrt(8.0 - 4.0 * dd + 8.0 * (2.0 - l2) / (l2 * sqrt(1.0 + dd)));
        double b = (a - 1.0) * (a - 1.0) / 4.0, c = a + b - 1.0;
        v.push_back(QdwhStep{a, b, c, c > cmax});
        l = l * (a + b * l2) / (1.0 + c * l2);
        if (l > 1.0) l = 1.0;
    }
    return v;
}

}  // namespace mak
