// CPU harness for the batched truncation search (matrixalgebrakit.jl_b200/csrc/trunc_core.h): the loop
// below is trunc_select_kernel with one thread per block.  Test infrastructure only.
#include "../../matrixalgebrakit.jl_b200/csrc/trunc_core.h"

extern "C" int trunc_select_host(int batch, const int* k, const double* const* S, const makb200_trunc_spec* spec,
                                 const int* maxrank_blk, int* rank, double* eps) {
    for (int i = 0; i < batch; ++i) {
        makb200_trunc_spec sp = *spec;   // the kernel's per-thread copy
        const int cap = maxrank_blk ? maxrank_blk[i] : -1;
        if (cap >= 0) sp.maxrank = (sp.maxrank >= 0 && sp.maxrank < cap) ? sp.maxrank : cap;
        mak::trunc::select(k[i], S[i], sp, &rank[i], &eps[i]);
    }
    return 0;
}
