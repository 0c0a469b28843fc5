// One-CTA-per-block batched QR for the many small blocks tensor-network codes emit.
// The whole block lives in shared memory: A is read from HBM once, Q and R are written once
// (algorithmic bytes = sz*(mn + mk + kn)), so small blocks run at HBM speed.
// Same reflector convention as the large path (beta >= 0  =>  gauge-fixed Q, R).
#include "batched.cuh"

namespace mak {

constexpr int BQ_THREADS = 256;
constexpr size_t BQ_SMEM_BYTES = 200 * 1024;

// shared-memory elements a block needs: padded tile + tau
size_t batched_qr_smem_elems(int m, int n) { return (size_t)(m | 1) * n + (m < n ? m : n) + 1; }
template <typename T> size_t batched_qr_max_smem_elems() { return BQ_SMEM_BYTES / sizeof(T) - 64; }
template size_t batched_qr_max_smem_elems<double>();
template size_t batched_qr_max_smem_elems<cplx>();

template <typename T>
__global__ void __launch_bounds__(BQ_THREADS)
batched_qr_kernel(const QrBlockDesc<T>* __restrict__ descs, int* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const QrBlockDesc<T> d = descs[blockIdx.x];
    const int m = d.m, n = d.n, k = m < n ? m : n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = BQ_THREADS / 32;
    if (m <= 0 || n < 0) { if (tid == 0 && info) info[blockIdx.x] = 0; return; }
    const int lds = m | 1;  // odd leading dimension
    T* S = reinterpret_cast<T*>(smem_raw);        // lds x n
    T* tau = S + (size_t)lds * n;                 // k
    T* red = tau + (k > 0 ? k : 1);               // 32
    T* scal = red + 32;                           // 4

    for (int idx = tid; idx < m * n; idx += BQ_THREADS) {
        int c = idx / m, r = idx - c * m;
        S[(size_t)c * lds + r] = d.A[(size_t)c * d.lda + r];
    }
    __syncthreads();

    // ---- factorization: S -> V \ R ----
    for (int j = 0; j < k; ++j) {
        T* cj = S + (size_t)j * lds;
        double part = 0.0;
        for (int r = j + 1 + tid; r < m; r += BQ_THREADS) part += abs2_(cj[r]);
        T tot = block_sum<T>(mk<T>(part), red);
        double beta; T tj, scale;
        larfgp_scalars<T>(cj[j], real_(tot), beta, tj, scale);
        __syncthreads();  // everyone has read cj[j]
        for (int r = j + 1 + tid; r < m; r += BQ_THREADS) cj[r] = mul_(cj[r], scale);
        if (tid == 0) { cj[j] = mk<T>(beta); tau[j] = tj; }
        __syncthreads();
        // apply H_j^H to columns l > j: one warp per column
        const T ctau = conj_(tj);
        for (int l = j + 1 + warp; l < n; l += NW) {
            T* cl = S + (size_t)l * lds;
            T s = zero<T>();
            for (int r = j + 1 + lane; r < m; r += 32) fmac_(s, cj[r], cl[r]);
            s = warp_sum(s);
            T f = mul_(ctau, add_(cl[j], s));
            __syncwarp();
            for (int r = j + 1 + lane; r < m; r += 32) cl[r] = sub_(cl[r], mul_(f, cj[r]));
            if (lane == 0) cl[j] = sub_(cl[j], f);
        }
        __syncthreads();
    }

    // ---- R out (upper triangle, zeros below) ----
    if (d.R) {
        for (int idx = tid; idx < k * n; idx += BQ_THREADS) {
            int c = idx / k, r = idx - c * k;
            d.R[(size_t)c * d.ldr + r] = (r <= c) ? S[(size_t)c * lds + r] : zero<T>();
        }
    }
    __syncthreads();

    // ---- form Q (m x k) in place over V, backward accumulation (org2r) ----
    for (int j = k - 1; j >= 0; --j) {
        T* cj = S + (size_t)j * lds;
        const T tj = tau[j];
        // apply H_j to columns l in (j, k): Q[j:, l] -= tau * w (w^H Q[j:, l]); rows < j of those
        // columns are already final
        for (int l = j + 1 + warp; l < k; l += NW) {
            T* cl = S + (size_t)l * lds;
            T s = zero<T>();
            for (int r = j + 1 + lane; r < m; r += 32) fmac_(s, cj[r], cl[r]);
            s = warp_sum(s);
            T f = mul_(tj, add_(cl[j], s));
            __syncwarp();
            for (int r = j + 1 + lane; r < m; r += 32) cl[r] = sub_(cl[r], mul_(f, cj[r]));
            if (lane == 0) cl[j] = sub_(cl[j], f);
        }
        __syncthreads();
        // column j itself: H_j e_j = e_j - tau * w
        for (int r = j + 1 + tid; r < m; r += BQ_THREADS) cj[r] = neg_(mul_(tj, cj[r]));
        for (int r = tid; r < j; r += BQ_THREADS) cj[r] = zero<T>();
        if (tid == 0) cj[j] = sub_(one<T>(), tj);
        __syncthreads();
    }
    for (int idx = tid; idx < m * k; idx += BQ_THREADS) {
        int c = idx / m, r = idx - c * m;
        d.Q[(size_t)c * d.ldq + r] = S[(size_t)c * lds + r];
    }
    if (tid == 0 && info) info[blockIdx.x] = 0;
}

int batched_init(makb200_handle* h) {
    MAK_CUDA(h, cudaFuncSetAttribute(batched_qr_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    MAK_CUDA(h, cudaFuncSetAttribute(batched_qr_kernel<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)BQ_SMEM_BYTES));
    return 0;
}

template <typename T>
int batched_qr_smem(makb200_handle* h, int batch, size_t max_smem_elems, const QrBlockDesc<T>* descs, int* info) {
    if (batch <= 0) return 0;
    size_t smem = (max_smem_elems + 64) * sizeof(T);
    if (smem > BQ_SMEM_BYTES) return MAKB200_ERR_WORKSPACE;
    batched_qr_kernel<T><<<batch, BQ_THREADS, smem, h->stream>>>(descs, info);
    count_launch();
    MAK_LAUNCH_CHECK(h, "batched_qr_kernel");
    return 0;
}
template int batched_qr_smem<double>(makb200_handle*, int, size_t, const QrBlockDesc<double>*, int*);
template int batched_qr_smem<cplx>(makb200_handle*, int, size_t, const QrBlockDesc<cplx>*, int*);

}  // namespace mak
