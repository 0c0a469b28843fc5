// CPU logic tests of the product's emulatable kernel headers: the SAME device code nvcc compiles into
// libmakb200, compiled here with g++ on top of cuda_emu.h (fibers) and driven through ctypes by
// tests/test_emu_kernels_cpu.py.  Test infrastructure only.
#include "cuda_emu.h"
#include "../../matrixalgebrakit.jl_b200/csrc/batched_qr_warp.cuh"
#include "../../matrixalgebrakit.jl_b200/csrc/sbr_chase_persistent.cuh"
#include "../../matrixalgebrakit.jl_b200/csrc/sbr_q2_slab.cuh"
#include "../../matrixalgebrakit.jl_b200/csrc/projections.cuh"
#include "../../matrixalgebrakit.jl_b200/csrc/bhetrd.cuh"

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
    } else if (variant == 2) {
        emu::launch(mak::batched_qr_warp_blk_kernel<T>, dim3(grid), dim3(128), 4 * (size_t)(cap + mak::BQW_TF_ELEMS) * sizeof(T),
                    (const mak::QrBlockDesc<T>*)d.data(), batch, cap);
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

// fused Q2 application (csrc/sbr_q2_slab.cuh): the V/T pools are built with sbr_core.h's dblock_build in the
// layout of q2_build_kernel (V: ld b+g, zero filled; T: ld g), blocks listed in (grp, k) order.
template <typename T>
static int run_q2_slab(int n, int b, int g, int cw, const T* V2, int ldv, const T* tau2, int ldt, T* Z, int ldz, int ncols) {
    if (n < 2) return 0;
    const int ngroups = (n - 1 + g - 1) / g, kmax = (n - 1 + b - 1) / b, ldvb = b + g;
    mak::sbr::Q2Store<T> Q{ldv, ldt, const_cast<T*>(V2), const_cast<T*>(tau2)};
    std::vector<mak::Q2BlockDesc> descs;
    std::vector<int> map((size_t)ngroups * kmax, -1);
    for (int grp = 0; grp < ngroups; ++grp)
        for (int k = 0; k < kmax; ++k) {
            const mak::sbr::DBlock d = mak::sbr::dblock_geometry(n, b, g, grp, k);
            if (d.ns <= 0) continue;
            mak::Q2BlockDesc q{d.s0, d.ns, d.base, d.rows, k, 0, descs.size() * (size_t)ldvb * g, descs.size() * (size_t)g * g};
            map[(size_t)grp * kmax + k] = (int)descs.size();
            descs.push_back(q);
        }
    std::vector<T> Vpool(descs.size() * (size_t)ldvb * g, mak::zero<T>()), Tpool(descs.size() * (size_t)g * g, mak::zero<T>());
    for (auto& q : descs) {
        const mak::sbr::DBlock d{q.s0, q.ns, q.base, q.rows};
        mak::sbr::dblock_build<T>(n, b, Q, d, q.k, Vpool.data() + q.voff, ldvb, Tpool.data() + q.toff, g);
        // dblock_build zeroes only d.rows rows of the ns columns: the pool is zero-initialised, as q2_build_kernel leaves it
    }
    const mak::Q2SlabSmem sm = mak::q2_slab_smem(b, g, cw);
    emu::launch(mak::q2_slab_kernel<T>, dim3((ncols + cw - 1) / cw), dim3(mak::Q2S_THREADS), sm.total * sizeof(T), n, b, g, cw,
                ngroups, kmax, (const int*)map.data(), (const mak::Q2BlockDesc*)descs.data(), (const T*)Vpool.data(),
                (const T*)Tpool.data(), Z, ldz, ncols);
    return 0;
}
extern "C" int emu_q2_slab(int dt, int n, int b, int g, int cw, const void* V2, int ldv, const void* tau2, int ldt, void* Z,
                           int ldz, int ncols, int order, uint64_t seed) {
    emu::set_order(order, seed);
    return dt == 0 ? run_q2_slab<double>(n, b, g, cw, (const double*)V2, ldv, (const double*)tau2, ldt, (double*)Z, ldz, ncols)
                   : run_q2_slab<cplx>(n, b, g, cw, (const cplx*)V2, ldv, (const cplx*)tau2, ldt, (cplx*)Z, ldz, ncols);
}

// projections.cuh: the launch shapes are those of project_herm_t / herm_props_t / gram_defect_t (eigh.cu)
template <typename T>
static void run_project(int anti, int n, const T* A, int lda, T* B, int ldb) {
    const int nb = (n + 31) / 32;
    if (anti) emu::launch(mak::project_herm_kernel<T, true>, dim3(nb, nb), dim3(32, 8), 0, n, A, lda, B, ldb);
    else emu::launch(mak::project_herm_kernel<T, false>, dim3(nb, nb), dim3(32, 8), 0, n, A, lda, B, ldb);
}
extern "C" int emu_project_herm(int dt, int anti, int n, const void* A, int lda, void* B, int ldb, int order, uint64_t seed) {
    emu::set_order(order, seed);
    if (dt == 0) run_project<double>(anti, n, (const double*)A, lda, (double*)B, ldb);
    else run_project<cplx>(anti, n, (const cplx*)A, lda, (cplx*)B, ldb);
    return 0;
}
template <typename T>
static void run_props(int anti, int n, const T* A, int lda, double* out4) {
    const int nb = (n + 31) / 32;
    for (int i = 0; i < 4; ++i) out4[i] = 0.0;
    if (anti) emu::launch(mak::herm_props_kernel<T, true>, dim3(nb, nb), dim3(32, 8), 0, n, A, lda, out4);
    else emu::launch(mak::herm_props_kernel<T, false>, dim3(nb, nb), dim3(32, 8), 0, n, A, lda, out4);
}
extern "C" int emu_herm_props(int dt, int anti, int n, const void* A, int lda, double* out4, int order, uint64_t seed) {
    emu::set_order(order, seed);
    if (dt == 0) run_props<double>(anti, n, (const double*)A, lda, out4);
    else run_props<cplx>(anti, n, (const cplx*)A, lda, out4);
    return 0;
}
extern "C" int emu_gram_defect(int dt, int n, const void* P, int ldp, double* out2, int grid, int order, uint64_t seed) {
    emu::set_order(order, seed);
    out2[0] = out2[1] = 0.0;
    if (dt == 0) emu::launch(mak::gram_defect_kernel<double>, dim3(grid), dim3(256), 0, n, (const double*)P, ldp, out2);
    else emu::launch(mak::gram_defect_kernel<cplx>, dim3(grid), dim3(256), 0, n, (const cplx*)P, ldp, out2);
    return 0;
}

// bhetrd.cuh: one CTA per block, launch shape of bhetrd_batched_t (eigh.cu)
template <typename T>
static int run_bhetrd(int batch, const int* n, void** A, const int* lda, void** d, void** e, void** tau, int mirror) {
    std::vector<mak::BhetrdDesc<T>> descs(batch);
    int nmax = 1;
    for (int i = 0; i < batch; ++i) {
        descs[i] = mak::BhetrdDesc<T>{n[i], (T*)A[i], lda[i], (double*)d[i], (double*)e[i], (T*)tau[i]};
        if (n[i] > nmax) nmax = n[i];
    }
    emu::launch(mak::bhetrd_kernel<T>, dim3(batch), dim3(mak::BHETRD_THREADS), mak::bhetrd_smem_elems(nmax) * sizeof(T),
                (const mak::BhetrdDesc<T>*)descs.data(), nmax, mirror);
    return 0;
}
extern "C" int emu_bhetrd(int dt, int batch, const int* n, void** A, const int* lda, void** d, void** e, void** tau,
                          int mirror, int order, uint64_t seed) {
    emu::set_order(order, seed);
    return dt == 0 ? run_bhetrd<double>(batch, n, A, lda, d, e, tau, mirror) : run_bhetrd<cplx>(batch, n, A, lda, d, e, tau, mirror);
}

extern "C" int emu_tri_init(int dt, int mode, int m, int n, void* A, int lda, int grid, int order, uint64_t seed) {
    emu::set_order(order, seed);
    if (dt == 0) emu::launch(mak::tri_init_kernel<double>, dim3(grid), dim3(256), 0, mode, m, n, (double*)A, lda);
    else emu::launch(mak::tri_init_kernel<cplx>, dim3(grid), dim3(256), 0, mode, m, n, (cplx*)A, lda);
    return 0;
}

extern "C" int emu_fro2(int dt, int m, int n, const void* A, int lda, double* out1, int grid, int order, uint64_t seed) {
    emu::set_order(order, seed);
    out1[0] = 0.0;
    if (dt == 0) emu::launch(mak::fro2_atomic_kernel<double>, dim3(grid), dim3(256), 0, m, n, (const double*)A, lda, out1);
    else emu::launch(mak::fro2_atomic_kernel<cplx>, dim3(grid), dim3(256), 0, m, n, (const cplx*)A, lda, out1);
    return 0;
}

extern "C" int emu_mirror_lower(int dt, int n, void* A, int lda, int order, uint64_t seed) {
    emu::set_order(order, seed);
    const int nb = (n + 31) / 32;
    if (dt == 0) emu::launch(mak::mirror_lower_kernel<double>, dim3(nb, nb), dim3(32, 8), 0, n, (double*)A, lda);
    else emu::launch(mak::mirror_lower_kernel<cplx>, dim3(nb, nb), dim3(32, 8), 0, n, (cplx*)A, lda);
    return 0;
}

extern "C" int emu_col_norm_defect(int dt, int m, int ncols, const void* U, int ldu, double* out, int order, uint64_t seed) {
    emu::set_order(order, seed);
    out[0] = 0.0;
    if (dt == 0) emu::launch(mak::col_norm_defect_kernel<double>, dim3((ncols + 7) / 8), dim3(256), 0, m, ncols, (const double*)U, ldu, out);
    else emu::launch(mak::col_norm_defect_kernel<cplx>, dim3((ncols + 7) / 8), dim3(256), 0, m, ncols, (const cplx*)U, ldu, out);
    return 0;
}
