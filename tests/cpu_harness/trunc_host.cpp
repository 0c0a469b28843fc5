// CPU harness for the batched truncation search (matrixalgebrakit.jl_b200/csrc/trunc_core.h): the loop
// below is trunc_select_kernel with one thread per block.  Test infrastructure only.
#include "../../matrixalgebrakit.jl_b200/csrc/trunc_core.h"

extern "C" int trunc_select_host(int batch, const int* k, const double* const* S, const makb200_trunc_spec* spec,
                                 int* rank, double* eps) {
    for (int i = 0; i < batch; ++i) mak::trunc::select(k[i], S[i], *spec, &rank[i], &eps[i]);
    return 0;
}
