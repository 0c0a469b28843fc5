// minimal TMA 2D load test: box (BR x BC) doubles from an n x n column-major matrix
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int BR, int BC, int FENCE>
__global__ void k(const __grid_constant__ CUtensorMap tmap, double* out, int c0, int c1) {
    extern __shared__ __align__(128) unsigned char raw[];
    unsigned char* sm = raw + ((128u - (smem_u32(raw) & 127u)) & 127u);
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1));
        if (FENCE == 0) asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        else asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(BR * BC * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
                     "r"(smem_u32(sm)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(c0), "r"(c1) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    const double* t = (const double*)sm;
    for (int i = threadIdx.x; i < BR * BC; i += blockDim.x) out[i] = t[i];
}
template <int BR, int BC, int FENCE, int L2P>
int run(PFN_cuTensorMapEncodeTiled_v12000 enc, double* dA, int n, double* dout, const std::vector<double>& hA) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)n};
    cuuint64_t gstr[1] = {(cuuint64_t)n * 8};
    cuuint32_t box[2] = {BR, BC};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, dA, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, L2P ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %dx%d fence=%d l2p=%d encode rc=%d ", BR, BC, FENCE, L2P, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
    size_t smem = (size_t)BR * BC * 8 + 128;
    cudaFuncSetAttribute(k<BR, BC, FENCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int c0 = (BC == 16 && FENCE == 1 && L2P == 0) ? 4 : 5, c1 = 3;
    k<BR, BC, FENCE><<<1, 128, smem>>>(tm, dout, c0, c1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s ", cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 2; }
    std::vector<double> h((size_t)BR * BC);
    cudaMemcpy(h.data(), dout, h.size() * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < BC; ++c)
        for (int rr = 0; rr < BR; ++rr) {
            double want = (c0 + rr < n && c1 + c < n) ? hA[(size_t)(c1 + c) * n + c0 + rr] : 0.0;
            if (h[(size_t)c * BR + rr] != want) ++bad;
        }
    printf("mismatches=%d\n", bad);
    return 0;
}
int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    int n = 1024;
    std::vector<double> hA((size_t)n * n);
    for (size_t i = 0; i < hA.size(); ++i) hA[i] = (double)i;
    double *dA, *dout;
    cudaMalloc(&dA, hA.size() * 8);
    cudaMalloc(&dout, 256 * 64 * 8);
    cudaMemcpy(dA, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice);
    run<256, 16, 1, 0>(enc, dA, n, dout, hA);
    run<256, 16, 1, 1>(enc, dA, n, dout, hA);
    run<256, 16, 0, 0>(enc, dA, n, dout, hA);
    return 0;
}
