// CUDA programming guide style TMA test using libcu++ wrappers
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
template <typename E, int BR, int BC>
__global__ void k(const __grid_constant__ CUtensorMap tensor_map, E* out, int x, int y) {
    __shared__ alignas(128) E smem_buffer[BC][BR];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BR * BC; i += blockDim.x) out[i] = (&smem_buffer[0][0])[i];
}
template <typename E, int BR, int BC>
void run(PFN_cuTensorMapEncodeTiled_v12000 enc, CUtensorMapDataType dt, E* dA, int n, E* dout, const std::vector<E>& hA) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)n};
    cuuint64_t gstr[1] = {(cuuint64_t)n * sizeof(E)};
    cuuint32_t box[2] = {BR, BC};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, dt, 2, dA, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("elem %zu box %dx%d encode rc=%d ", sizeof(E), BR, BC, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return; }
    k<E, BR, BC><<<1, 128>>>(tm, dout, 4, 3);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s ", cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return; }
    std::vector<E> h((size_t)BR * BC);
    cudaMemcpy(h.data(), dout, h.size() * sizeof(E), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < BC; ++c)
        for (int rr = 0; rr < BR; ++rr)
            if (h[(size_t)c * BR + rr] != hA[(size_t)(3 + c) * n + 4 + rr]) ++bad;
    printf("mismatches=%d\n", bad);
}
int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    int n = 1024;
    {
        std::vector<int> hA((size_t)n * n);
        for (size_t i = 0; i < hA.size(); ++i) hA[i] = (int)i;
        int *dA, *dout;
        cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dout, 64 * 64 * 4);
        cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
        run<int, 64, 16>(enc, CU_TENSOR_MAP_DATA_TYPE_INT32, dA, n, dout, hA);
    }
    {
        std::vector<double> hA((size_t)n * n);
        for (size_t i = 0; i < hA.size(); ++i) hA[i] = (double)i;
        double *dA, *dout;
        cudaMalloc(&dA, hA.size() * 8); cudaMalloc(&dout, 256 * 16 * 8);
        cudaMemcpy(dA, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice);
        run<double, 16, 16>(enc, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, dA, n, dout, hA);
        run<double, 64, 16>(enc, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, dA, n, dout, hA);
        run<double, 256, 16>(enc, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, dA, n, dout, hA);
    }
    return 0;
}
