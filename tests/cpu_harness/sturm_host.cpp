// CPU harness for the values-only tridiagonal solver (matrixalgebrakit.jl_b200/csrc/sturm_core.h):
// the loop below is what sturm_eigvals_kernel does with one thread per k.  Test infrastructure only.
#include "../../matrixalgebrakit.jl_b200/csrc/sturm_core.h"

using namespace mak::sturm;

extern "C" int sturm_host(int n, const double* d, const double* e, double* w) {
    if (n <= 0) return 0;
    const Bounds b = bounds(n, d, e);
    for (int k = 0; k < n; ++k) w[k] = kth_eigenvalue(n, d, e, b, k);
    return 0;
}

// number of eigenvalues of T below x (unscaled x), for the count's own test
extern "C" int sturm_count_host(int n, const double* d, const double* e, double x) {
    const Bounds b = bounds(n, d, e);
    double xs[1] = {x * b.inv};
    int c[1];
    count_below<1>(n, d, e, b.inv, xs, c);
    return c[0];
}
