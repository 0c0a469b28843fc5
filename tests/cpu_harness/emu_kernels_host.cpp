// CPU logic tests of the product's emulatable kernel headers: the SAME device code nvcc compiles into
// libmakb200, compiled here with g++ on top of cuda_emu.h (fibers) and driven through ctypes by
// tests/test_emu_kernels_cpu.py.  Test infrastructure only.
#include "cuda_emu.h"
#include "../../matrixalgebrakit.jl_b200/csrc/batched_qr_warp.cuh"
#include "../../matrixalgebrakit.jl_b200/csrc/sbr_chase_persistent.cuh"

using mak::cplx;

template <typename T>
static int run_bqr_warp(int variant, int batch, const int* m, const int* n, void** A, const int* lda, void** Q,
                        const int* ldq, void** R, const int* ldr) {
    std::vector<mak::QrBlockDesc<T>> d(batch);
    int cap = 0, rmax = 0;
    for (int i = 0; i < batch; ++i) {
        d[i] = mak::QrBlockDesc<T>{m[i], n[i], (T*)A[i], lda[i], (T*)Q[i], ldq[i], (T*)R[i], ldr[i]};
        const int e = (m[i] | 1) * n[i];
        if (e > cap) cap = e;
        if (m[i] > rmax) rmax = m[i];
        if (n[i] > rmax) rmax = n[i];
    }
    const int grid = (batch + 3) / 4;
    if (variant == 0) {
        emu::launch(mak::batched_qr_warp_kernel<T>, dim3(grid), dim3(128), 4 * (size_t)cap * sizeof(T),
                    (const mak::QrBlockDesc<T>*)d.data(), batch, cap);
    } else if (variant == 1) {
        if (rmax <= 16) emu::launch(mak::batched_qr_warp_reg_kernel<T, 16>, dim3(grid), dim3(128), 0, (const mak::QrBlockDesc<T>*)d.data(), batch);
        else if (rmax <= 24) emu::launch(mak::batched_qr_warp_reg_kernel<T, 24>, dim3(grid), dim3(128), 0, (const mak::QrBlockDesc<T>*)d.data(), batch);
        else emu::launch(mak::batched_qr_warp_reg_kernel<T, 32>, dim3(grid), dim3(128), 0, (const mak::QrBlockDesc<T>*)d.data(), batch);
    } else {
        return -1;
    }
    return 0;
}

extern "C" int emu_batched_qr_warp(int dt, int variant, int batch, const int* m, const int* n, void** A, const int* lda,
                                   void** Q, const int* ldq, void** R, const int* ldr, int order, uint64_t seed) {
    emu::set_order(order, seed);
    return dt == 0 ? run_bqr_warp<double>(variant, batch, m, n, A, lda, Q, ldq, R, ldr)
                   : run_bqr_warp<cplx>(variant, batch, m, n, A, lda, Q, ldq, R, ldr);
}

// self-test of the emulator's collectives: warp_sum / shuffles / ballot / block_sum / dmma against closed forms
__global__ void emu_selftest_kernel(double* out, const double* Am, const double* Bm, double* Cm) {
    __shared__ double scratch[32];
    const int tid = threadIdx.x, lane = tid & 31;
    const double ws = mak::warp_sum((double)(tid + 1));
    const double bs = mak::block_sum<double>((double)(tid + 1), scratch);
    const unsigned bal = __ballot_sync(0xffffffffu, lane % 3 == 0);
    const double dn = __shfl_down_sync(0xffffffffu, (double)tid, 1);
    const double x16 = __shfl_xor_sync(0xffffffffu, (double)lane, 5, 16);
    out[tid * 5 + 0] = ws;
    out[tid * 5 + 1] = bs;
    out[tid * 5 + 2] = (double)bal;
    out[tid * 5 + 3] = dn;
    out[tid * 5 + 4] = x16;
    if (tid < 32) {
        // C(8x8) += A(8x4) B(4x8), all row-major in memory
        double d0 = Cm[(lane >> 2) * 8 + (lane & 3) * 2], d1 = Cm[(lane >> 2) * 8 + (lane & 3) * 2 + 1];
        emu::dmma(d0, d1, Am[(lane >> 2) * 4 + (lane & 3)], Bm[(lane & 3) * 8 + (lane >> 2)]);
        Cm[(lane >> 2) * 8 + (lane & 3) * 2] = d0;
        Cm[(lane >> 2) * 8 + (lane & 3) * 2 + 1] = d1;
    }
}
extern "C" int emu_selftest(int nthreads, double* out, const double* Am, const double* Bm, double* Cm, int order, uint64_t seed) {
    emu::set_order(order, seed);
    emu::launch(emu_selftest_kernel, dim3(1), dim3(nthreads), 0, out, Am, Bm, Cm);
    return 0;
}

// persistent bulge chasing: `grid` co-resident CTAs of 256 fibers; AB is the 2b x n band storage (in/out),
// V2 (ldv x n), tau2 (ldt x n) zero-initialised by the caller.  Returns prog[n] (non-zero: a consumer gave up).
template <typename T>
static int run_chase_persistent(int n, int b, T* AB, int ldab, T* V2, int ldv, T* tau2, int ldt, int grid) {
    std::vector<int> prog(n + 1, 0);
    emu::launch_coresident(mak::chase_persistent_kernel<T>, dim3(grid), dim3(mak::SBRP_THREADS),
                           mak::chase_persistent_smem_elems(b) * sizeof(T), n, b, AB, ldab, V2, ldv, tau2, ldt, prog.data());
    for (int s = 0; s <= n - 2; ++s)
        if (prog[s] != mak::sbr::sweep_ntasks(n, b, s)) return -(s + 1);
    return prog[n];
}
extern "C" int emu_chase_persistent(int dt, int n, int b, void* AB, int ldab, void* V2, int ldv, void* tau2, int ldt,
                                    int grid, int order, uint64_t seed) {
    emu::set_order(order, seed);
    return dt == 0 ? run_chase_persistent<double>(n, b, (double*)AB, ldab, (double*)V2, ldv, (double*)tau2, ldt, grid)
                   : run_chase_persistent<cplx>(n, b, (cplx*)AB, ldab, (cplx*)V2, ldv, (cplx*)tau2, ldt, grid);
}
