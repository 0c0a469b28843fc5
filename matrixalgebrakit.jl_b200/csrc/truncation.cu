// Batched truncation search for svd_trunc! over many blocks (BASELINE config 3): rank and truncation
// error of every block from ONE launch and ONE device->host read, instead of a host copy of every
// block's values (the reference's GPU path, MatrixAlgebraKitCUDAExt.jl:64-66, searches on the host).
#include "common.cuh"
#include "trunc_core.h"
#include <vector>

namespace mak {

struct TruncDesc {
    const double* S;
    int k;
    int maxrank;   // per-block cap, < 0: none
};

__global__ void trunc_select_kernel(int batch, const TruncDesc* __restrict__ descs, makb200_trunc_spec sp,
                                    int* __restrict__ rank, double* __restrict__ eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    int r;
    double e;
    const int cap = descs[i].maxrank;
    if (cap >= 0) sp.maxrank = (sp.maxrank >= 0 && sp.maxrank < cap) ? sp.maxrank : cap;   // sp is this thread's copy
    trunc::select(descs[i].k, descs[i].S, sp, &r, &e);
    rank[i] = r;
    eps[i] = e;
}

}  // namespace mak

extern "C" {

size_t makb200_trunc_select_batched_worksize(makb200_handle_t* h, int batch) {
    if (!h || batch < 0) return 0;
    return mak::align_up(sizeof(mak::TruncDesc) * (size_t)(batch > 0 ? batch : 1), 256) + 256;
}

int makb200_trunc_select_batched(makb200_handle_t* h, int batch, const int* k, double* const* S,
                                 const makb200_trunc_spec* spec, const int* maxrank_blk, int* rank_dev,
                                 double* eps_dev, void* work, size_t lwork) {
    if (!h) return -1;
    if (batch < 0) return -2;
    if (batch == 0) return 0;
    if (!k) return -3;
    if (!S) return -4;
    if (!spec) return -5;
    if (!rank_dev) return -7;
    if (!eps_dev) return -8;
    if (spec->by_value && !(spec->vp > 0.0 && spec->vp < 1e300)) return -5;   // finite p only
    if (spec->by_error && !(spec->ep > 0.0 && spec->ep < 1e300)) return -5;
    std::vector<mak::TruncDesc> d((size_t)batch);
    for (int i = 0; i < batch; ++i) {
        if (k[i] < 0) return -3;
        if (k[i] > 0 && !S[i]) return -4;
        d[i].S = S[i];
        d[i].k = k[i];
        d[i].maxrank = maxrank_blk ? maxrank_blk[i] : -1;
    }
    mak::Arena ar(work, lwork);
    mak::TruncDesc* ddev = ar.get<mak::TruncDesc>((size_t)batch);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    {
        mak::Stager st(h, d.size() * sizeof(mak::TruncDesc) + 1024);
        MAK_CUDA(h, st.put(ddev, d.data(), d.size() * sizeof(mak::TruncDesc), h->stream));
    }
    mak::trunc_select_kernel<<<(batch + 127) / 128, 128, 0, h->stream>>>(batch, ddev, *spec, rank_dev, eps_dev);
    mak::count_launch();
    MAK_LAUNCH_CHECK(h, "trunc_select_kernel");
    return 0;
}

}  // extern "C"
