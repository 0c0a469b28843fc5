// C ABI (include/makb200.h): handle management and argument checking; the numerical work
// lives in gemm.cu / qr.cu / ...
#include "common.cuh"
#include "gemm.cuh"
#include "qr.cuh"
#include "batched.cuh"
#include "eigh.cuh"
#include "stedc.cuh"
#include "polar.cuh"
#include "sbr.cuh"
#include "bhetrd.cuh"
#include "nccl_dl.h"
#include <dlfcn.h>
#include <vector>
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <map>
#include <mutex>
#include <tuple>

using mak::cplx;

extern "C" {

int makb200_version(void) { return 100; }

int makb200_create(makb200_handle_t** out, int device) {
    if (!out) return -1;
    *out = nullptr;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return MAKB200_ERR_CUDA;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return MAKB200_ERR_CUDA;
    if (prop.major != 10) return MAKB200_ERR_CUDA;  // sm_100a SASS only: fail loudly elsewhere
    makb200_handle* h = new makb200_handle();
    h->device = device;
    h->stream = 0;
    h->num_sms = prop.multiProcessorCount;
    h->max_cluster = 8;
    h->err[0] = 0;
    {
        // high-priority auxiliary stream: latency-bound panel kernels must be able to claim SMs while
        // bulk GEMMs of the caller's (default-priority) stream are still draining
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        if (cudaStreamCreateWithPriority(&h->aux_stream, cudaStreamNonBlocking, greatest) != cudaSuccess) {
            delete h;
            return MAKB200_ERR_CUDA;
        }
    }
    h->no_lookahead = false;
    h->graph_cache = nullptr;
    h->defect_dev = nullptr;
    h->stage = nullptr;
    h->stage_bytes = 0;
    if (cudaEventCreateWithFlags(&h->stage_ev, cudaEventDisableTiming) != cudaSuccess) { delete h; return MAKB200_ERR_CUDA; }
    for (int i = 0; i < 8; ++i)
        if (cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming) != cudaSuccess) { delete h; return MAKB200_ERR_CUDA; }
    for (int i = 0; i < MAK_NPOOL; ++i) {
        if (cudaEventCreateWithFlags(&h->pool_ev[i], cudaEventDisableTiming) != cudaSuccess) { delete h; return MAKB200_ERR_CUDA; }
        if (cudaStreamCreateWithFlags(&h->pool[i], cudaStreamNonBlocking) != cudaSuccess) { delete h; return MAKB200_ERR_CUDA; }
    }
    int rc = mak::qr_init(h);
    if (rc == 0) rc = mak::batched_init(h);
    if (rc == 0) rc = mak::batched_blocked_init(h);
    if (rc == 0) rc = mak::polar_init(h);
    if (rc != 0) { delete h; return rc; }
    *out = h;
    return 0;
}

static void graph_cache_destroy(makb200_handle_t* h);

int makb200_destroy(makb200_handle_t* h) {
    if (!h) return -1;
    graph_cache_destroy(h);
    cudaStreamDestroy(h->aux_stream);
    cudaEventDestroy(h->stage_ev);
    if (h->stage) cudaFreeHost(h->stage);
    for (int i = 0; i < 8; ++i) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < MAK_NPOOL; ++i) { cudaEventDestroy(h->pool_ev[i]); cudaStreamDestroy(h->pool[i]); }
    delete h;
    return 0;
}

int makb200_set_stream(makb200_handle_t* h, void* s) {
    if (!h) return -1;
    h->stream = (cudaStream_t)s;
    return 0;
}

const char* makb200_last_error(makb200_handle_t* h) { return h ? h->err : "null handle"; }

unsigned long long makb200_launch_count(void) { return mak::g_launches; }

int makb200_kernel_timing(int enable) {
    mak::g_clock_dots.on = enable != 0;
    mak::g_clock_gemm.on = enable != 0;
    mak::g_clock_w.on = enable != 0;
    mak::g_clock_w.n = 0;
    mak::g_clock_dots.n = 0;
    mak::g_clock_gemm.n = 0;
    mak::g_gemm_flops = 0.0;
    return 0;
}

double makb200_gemm_flops(void) { return mak::g_gemm_flops; }

int makb200_kernel_time(int which, double* ms, int* launches) {
    if (!ms || !launches) return -2;
    if (which == 0) mak::g_clock_dots.collect(ms, launches);
    else if (which == 1) mak::g_clock_gemm.collect(ms, launches);
    else if (which == 2) mak::g_clock_w.collect(ms, launches);
    else return -1;
    return 0;
}

static bool op_ok(int op) { return op == MAKB200_OP_N || op == MAKB200_OP_T || op == MAKB200_OP_C; }

int makb200_gemm(makb200_handle_t* h, int dtype, int opa, int opb, int m, int n, int k, const void* alpha,
                 const void* A, int lda, const void* B, int ldb, const void* beta, void* C, int ldc) {
    if (!h) return -1;
    if (dtype != MAKB200_F64 && dtype != MAKB200_C128) return -2;
    if (!op_ok(opa)) return -3;
    if (!op_ok(opb)) return -4;
    if (m < 0) return -5;
    if (n < 0) return -6;
    if (k < 0) return -7;
    if (!alpha) return -8;
    int arows = (opa == MAKB200_OP_N) ? m : k, brows = (opb == MAKB200_OP_N) ? k : n;
    if (lda < (arows > 1 ? arows : 1)) return -10;
    if (ldb < (brows > 1 ? brows : 1)) return -12;
    if (!beta) return -13;
    if (ldc < (m > 1 ? m : 1)) return -15;
    if (m == 0 || n == 0) return 0;
    if ((k > 0 && (!A || !B)) || !C) return -9;
    cudaError_t e;
    if (dtype == MAKB200_F64)
        e = mak::gemm<double>(h->stream, h->num_sms, opa, opb, m, n, k, *(const double*)alpha, (const double*)A, lda,
                              (const double*)B, ldb, *(const double*)beta, (double*)C, ldc);
    else
        e = mak::gemm<cplx>(h->stream, h->num_sms, opa, opb, m, n, k, *(const cplx*)alpha, (const cplx*)A, lda,
                            (const cplx*)B, ldb, *(const cplx*)beta, (cplx*)C, ldc);
    if (e != cudaSuccess) return mak::cuda_fail(h, e, "makb200_gemm");
    return 0;
}

static bool dtype_ok(int d) { return d == MAKB200_F64 || d == MAKB200_C128; }
static int maxi(int a, int b) { return a > b ? a : b; }

size_t makb200_qr_worksize(makb200_handle_t* h, int dtype, int mode, int m, int n) {
    if (!h || !dtype_ok(dtype) || m < 0 || n < 0) return 0;
    int k = m < n ? m : n, ncq = (mode == MAKB200_QR_FULL) ? m : k;
    return dtype == MAKB200_F64 ? mak::qr_worksize_t<double>(h, m, n, ncq) : mak::qr_worksize_t<cplx>(h, m, n, ncq);
}

int makb200_qr(makb200_handle_t* h, int dtype, int mode, int positive, int m, int n, void* A, int lda, void* Q,
               int ldq, void* R, int ldr, void* work, size_t lwork) {
    (void)positive;
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (mode != MAKB200_QR_COMPACT && mode != MAKB200_QR_FULL) return -3;
    if (m < 0) return -5;
    if (n < 0) return -6;
    int k = m < n ? m : n, ncq = (mode == MAKB200_QR_FULL) ? m : k;
    if (lda < maxi(1, m)) return -8;
    if (ldq < maxi(1, m)) return -10;
    if (R && ldr > 0 && ldr < maxi(1, ncq)) return -12;
    if (m == 0) return 0;
    if ((!A && n > 0) || !Q) return -7;
    if (Q == A) return -9;  // in-place Q is not provided (qr.jl:150-153 rejects it for R/positive anyway)
    if (dtype == MAKB200_F64)
        return mak::qr_fused_t<double>(h, mode, m, n, (double*)A, lda, (double*)Q, ldq, (double*)R, ldr, work, lwork);
    return mak::qr_fused_t<cplx>(h, mode, m, n, (cplx*)A, lda, (cplx*)Q, ldq, (cplx*)R, ldr, work, lwork);
}

size_t makb200_geqrf_worksize(makb200_handle_t* h, int dtype, int m, int n) {
    if (!h || !dtype_ok(dtype) || m < 0 || n < 0) return 0;
    int k = m < n ? m : n;
    return dtype == MAKB200_F64 ? mak::qr_worksize_t<double>(h, m, n, k) : mak::qr_worksize_t<cplx>(h, m, n, k);
}

int makb200_geqrf(makb200_handle_t* h, int dtype, int m, int n, void* A, int lda, void* tau, void* work,
                  size_t lwork) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (m < 0) return -3;
    if (n < 0) return -4;
    if (lda < maxi(1, m)) return -6;
    if (m == 0 || n == 0) return 0;  // yalapack.jl:177
    if (!A) return -5;
    if (!tau) return -7;
    if (dtype == MAKB200_F64) return mak::geqrf_t<double>(h, m, n, (double*)A, lda, (double*)tau, work, lwork);
    return mak::geqrf_t<cplx>(h, m, n, (cplx*)A, lda, (cplx*)tau, work, lwork);
}

size_t makb200_orgqr_worksize(makb200_handle_t* h, int dtype, int m, int ncols, int k) {
    if (!h || !dtype_ok(dtype) || m < 0 || ncols < 0 || k < 0) return 0;
    return dtype == MAKB200_F64 ? mak::qr_worksize_t<double>(h, m, k, ncols) : mak::qr_worksize_t<cplx>(h, m, k, ncols);
}

int makb200_orgqr(makb200_handle_t* h, int dtype, int m, int ncols, int k, const void* A, int lda, const void* tau,
                  void* Q, int ldq, void* work, size_t lwork) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (m < 0) return -3;
    if (ncols < 0 || ncols > m) return -4;
    if (k < 0 || k > m) return -5;
    if (lda < maxi(1, m)) return -7;
    if (ldq < maxi(1, m)) return -10;
    if (m == 0 || ncols == 0) return 0;
    if (k > 0 && (!A || !tau)) return -6;
    if (!Q || Q == A) return -9;
    if (dtype == MAKB200_F64)
        return mak::orgqr_t<double>(h, m, ncols, k, (const double*)A, lda, (const double*)tau, (double*)Q, ldq, work,
                                    lwork);
    return mak::orgqr_t<cplx>(h, m, ncols, k, (const cplx*)A, lda, (const cplx*)tau, (cplx*)Q, ldq, work, lwork);
}

}  // extern "C"

// ---- batched -----------------------------------------------------------------------
// Run `fn(slot_work, slot_lwork, i)` for every index in `big` round-robin over the handle's stream
// pool (the per-block paths are launch-latency bound, so independent blocks overlap almost freely).
constexpr int NPOOL = MAK_NPOOL;
static int env_int(const char* name, int dflt, int lo, int hi) {
    const char* e = getenv(name);
    int v = e ? atoi(e) : dflt;
    return v < lo ? lo : (v > hi ? hi : v);
}
// measured on B200 (round 1, 257-512 c128 blocks): 8 streams / 8 threads 229 svd/s, 561 eigh/s;
// 32 streams / 16 threads 176 / 351 (driver lock contention) -> defaults 8 / 8
static int pool_streams() { static int v = env_int("MAKB200_POOL_STREAMS", 8, 1, 32); return v; }
static int pool_threads() {   // host threads feeding the stream pool (1 = single-threaded round robin)
    static int v = -1;
    if (v < 0) {
        v = env_int("MAKB200_POOL_THREADS", 8, 1, NPOOL);
        unsigned hc = std::thread::hardware_concurrency();
        const int lws = env_int("LOCAL_WORLD_SIZE", 1, 1, 1024);   // ranks sharing this host (torchrun)
        if (hc > 0) {
            int cap = (int)hc / lws;
            if (cap < 1) cap = 1;
            if (v > cap) v = cap;
        }
    }
    return v;
}
// Workspace of a pooled call: one slice per stream.
static int graph_slots();
static size_t pooled_worksize(size_t per_block, size_t nbig, size_t staging = 0) {
    size_t np = (size_t)pool_streams();
    if (staging > 0 && (size_t)graph_slots() > np) np = (size_t)graph_slots();   // graph-replayed path: more, cheaper slots
    if (nbig < np) np = nbig;
    return (per_block + staging + 1024) * np;
}

template <typename F>
static int run_pooled(makb200_handle_t* h, const std::vector<int>& big, char* work, size_t lwork, F fn) {
    if (big.empty()) return 0;
    cudaStream_t main = h->stream;
    const int np = (int)big.size() < pool_streams() ? (int)big.size() : pool_streams();
    const size_t slice = (lwork / np) & ~(size_t)255;
    MAK_CUDA(h, cudaEventRecord(h->pool_ev[0], main));
    for (int s = 0; s < np; ++s) MAK_CUDA(h, cudaStreamWaitEvent(h->pool[s], h->pool_ev[0], 0));
    h->no_lookahead = true;
    int rc = 0;
    // The per-block paths are serial chains of thousands of tiny launches: one stream keeps only a
    // sliver of the GPU busy and ONE host thread cannot feed many streams (~2 us per launch).  So the
    // pool has up to 32 streams fed by up to 16 host threads; thread t owns streams t, t+nt, ... and
    // alternates between them block by block, each stream with its own workspace slice and each thread
    // with a private copy of the handle.  Blocks are handed out dynamically.  The first block runs on
    // the calling thread so that every lazily configured kernel attribute is set before the fan-out.
    const bool timing = mak::g_clock_gemm.on || mak::g_clock_dots.on || mak::g_clock_w.on ||
                        (getenv("MAKB200_PROFILE") && getenv("MAKB200_PROFILE")[0] == '1');
    const int nthreads = timing ? 1 : (pool_threads() < np ? pool_threads() : np);
    size_t first = 0;
    if (nthreads > 1) {
        h->stream = h->pool[0];
        rc = fn(h, work, slice, big[0]);
        first = 1;
    }
    if (nthreads > 1 && rc == 0 && big.size() > 1) {
        std::atomic<size_t> next(first);
        std::vector<int> rcs(nthreads, 0);
        std::vector<makb200_handle> hs(nthreads, *h);
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) {
            hs[t].no_lookahead = true;
            hs[t].err[0] = 0;
            th.emplace_back([&, t]() {
                if (cudaSetDevice(h->device) != cudaSuccess) { rcs[t] = MAKB200_ERR_CUDA; return; }
                int s = t;
                for (;;) {
                    const size_t idx = next.fetch_add(1);
                    if (idx >= big.size()) break;
                    hs[t].stream = h->pool[s];
                    const int r = fn(&hs[t], work + s * slice, slice, big[idx]);
                    if (r) { rcs[t] = r; break; }
                    s += nthreads;
                    if (s >= np) s = t;
                }
            });
        }
        for (auto& x : th) x.join();
        for (int t = 0; t < nthreads && rc == 0; ++t)
            if (rcs[t]) { rc = rcs[t]; memcpy(h->err, hs[t].err, sizeof(h->err)); }
    } else if (rc == 0) {
        for (size_t idx = first; idx < big.size() && rc == 0; ++idx) {
            const int s = (int)(idx % np);
            h->stream = h->pool[s];
            rc = fn(h, work + s * slice, slice, big[idx]);
        }
    }
    h->stream = main;
    h->no_lookahead = false;
    // join on the error path too: blocks of other streams may still be running, and the caller is about to reuse (or
    // free) the workspace and the outputs
    for (int s = 0; s < np; ++s) {
        const cudaError_t e1 = cudaEventRecord(h->pool_ev[s], h->pool[s]);
        const cudaError_t e2 = e1 == cudaSuccess ? cudaStreamWaitEvent(main, h->pool_ev[s], 0) : e1;
        if (e2 != cudaSuccess && rc == 0) rc = mak::cuda_fail(h, e2, "run_pooled: join");
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// Graph replay of the per-block paths.  A mid-size block (65..512) runs the single-matrix svd_t / eigh_t: a serial
// chain of ~1000 tiny launches (n = 384 c128: 1326 launches, 22 ms), so the pooled path is bound by what host threads
// can launch.  Blocks of one shape run the SAME launch sequence; here it is captured ONCE per (op, shape, stream slot)
// against fixed staging buffers in that slot's workspace slice and replayed per block between a copy-in and a copy-out:
// per block the host issues one cudaGraphLaunch and a few small copies instead of a thousand launches, the tensor maps
// of the GEMMs are encoded once at capture, and the GPU runs the chain back to back.  Shapes are dealt to the slots
// (longest processing time first) so every graph is reused for all blocks of its shape.
// The rank-defect check of svd_t is a host read: captured sequences store the indicator per block instead, ONE read
// after the batch finds deficient blocks, and those are redone through the uncaptured path (the staged copy-in leaves
// the caller's A intact).
// ---------------------------------------------------------------------------------------------------------------
struct GraphEntry { cudaGraphExec_t exec; void* stage; };
struct GraphCache {
    std::mutex mu;
    std::map<std::tuple<int, int, int, int, int, int>, GraphEntry> map;   // (kind, dtype, m, n, flags, slot)
};
static GraphCache* graph_cache(makb200_handle_t* h) {
    if (!h->graph_cache) h->graph_cache = new GraphCache();
    return (GraphCache*)h->graph_cache;
}
static void graph_cache_destroy(makb200_handle_t* h) {
    if (!h->graph_cache) return;
    GraphCache* c = (GraphCache*)h->graph_cache;
    for (auto& kv : c->map) cudaGraphExecDestroy(kv.second.exec);
    delete c;
    h->graph_cache = nullptr;
}
// MEASURED (round 2, profiles/r2_batched_graphs.log): replay works and is parity-green, but the throughput does not move
// (257-512 c128 svd: 263 vs 231 blocks/s): with 32 chains in flight the device itself dispatches only ~350 k kernels/s
// (263 blocks/s x 1326 kernels), so the bound is the NUMBER of kernels per block, not who launches them.  Replay stays
// opt-in (MAKB200_BATCH_GRAPHS=1); the default remains the launch-per-kernel pool, which has no first-call capture cost.
static bool graphs_enabled() {
    const char* e = getenv("MAKB200_BATCH_GRAPHS");   // read per call: tests run both paths
    return e && e[0] == '1';
}
static int graph_slots() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MAKB200_GRAPH_SLOTS"); v = e ? atoi(e) : 32; if (v < 1) v = 1; if (v > MAK_NPOOL) v = MAK_NPOOL; }
    return v;
}

// One shape group of a graph-replayed batch: blocks `idx` (indices into the caller's arrays) of shape m x n.
struct ShapeGroup { int m, n; std::vector<int> idx; double cost; };

// capture(hh, m, n, stage, stage_bytes) must issue the whole per-block sequence on hh->stream against the staging area;
// copy_in(hh, i, m, n, stage) / copy_out(hh, i, m, n, stage, ordinal) move block i in and out (ordinal = position in `big`).
// Returns 0, an error code, or -100000 when capture is not possible (caller falls back to run_pooled).
template <typename FC, typename FI, typename FO>
static int run_graphed(makb200_handle_t* h, int kind, int dtype, int flags, std::vector<ShapeGroup>& groups, char* work,
                       size_t lwork, size_t slot_bytes, FC capture, FI copy_in, FO copy_out) {
    if (groups.empty()) return 0;
    int np = graph_slots();
    if ((size_t)np > groups.size()) np = (int)groups.size();
    if (slot_bytes == 0 || lwork / slot_bytes < 1) return -100000;
    if ((size_t)np > lwork / slot_bytes) np = (int)(lwork / slot_bytes);
    const size_t slice = slot_bytes & ~(size_t)255;
    // shapes -> slots, longest processing time first
    std::sort(groups.begin(), groups.end(), [](const ShapeGroup& a, const ShapeGroup& b) { return a.cost > b.cost; });
    std::vector<std::vector<int>> slot_groups(np);
    {
        std::vector<double> load(np, 0.0);
        for (size_t g = 0; g < groups.size(); ++g) {
            int best = 0;
            for (int s = 1; s < np; ++s) if (load[s] < load[best]) best = s;
            slot_groups[best].push_back((int)g);
            load[best] += groups[g].cost;
        }
    }
    cudaStream_t main = h->stream;
    MAK_CUDA(h, cudaEventRecord(h->pool_ev[0], main));
    for (int s = 0; s < np; ++s) MAK_CUDA(h, cudaStreamWaitEvent(h->pool[s], h->pool_ev[0], 0));
    GraphCache* cache = graph_cache(h);
    const int nthreads = pool_threads() < np ? pool_threads() : np;
    std::vector<int> rcs(nthreads, 0);
    std::vector<makb200_handle> hs(nthreads, *h);
    auto worker = [&](int t) {
        if (cudaSetDevice(h->device) != cudaSuccess) { rcs[t] = MAKB200_ERR_CUDA; return; }
        makb200_handle* hh = &hs[t];
        hh->no_lookahead = true;
        hh->err[0] = 0;
        for (int s = t; s < np && rcs[t] == 0; s += nthreads) {
            hh->stream = h->pool[s];
            char* stage = work + (size_t)s * slice;
            for (int g : slot_groups[s]) {
                const ShapeGroup& sg = groups[g];
                const auto key = std::make_tuple(kind, dtype, sg.m, sg.n, flags, s);
                cudaGraphExec_t exec = nullptr;
                {
                    std::lock_guard<std::mutex> lk(cache->mu);
                    auto it = cache->map.find(key);
                    if (it != cache->map.end()) {
                        if (it->second.stage == (void*)stage) exec = it->second.exec;
                        else { cudaGraphExecDestroy(it->second.exec); cache->map.erase(it); }   // the caller's workspace moved
                    }
                }
                if (!exec) {
                    cudaGraph_t graph = nullptr;
                    if (cudaStreamBeginCapture(hh->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { rcs[t] = -100000; break; }
                    const int rc = capture(hh, sg.m, sg.n, stage, slice);
                    const cudaError_t ce = cudaStreamEndCapture(hh->stream, &graph);
                    if (rc != 0 || ce != cudaSuccess || !graph) {
                        if (graph) cudaGraphDestroy(graph);
                        cudaGetLastError();
                        rcs[t] = rc > 0 ? rc : -100000;
                        break;
                    }
                    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
                    cudaGraphDestroy(graph);
                    if (ie != cudaSuccess) { cudaGetLastError(); rcs[t] = -100000; break; }
                    std::lock_guard<std::mutex> lk(cache->mu);
                    cache->map[key] = GraphEntry{exec, (void*)stage};
                }
                for (size_t q = 0; q < sg.idx.size() && rcs[t] == 0; ++q) {
                    int rc = copy_in(hh, sg.idx[q], sg.m, sg.n, stage);
                    if (rc == 0 && cudaGraphLaunch(exec, hh->stream) != cudaSuccess) rc = mak::cuda_fail(hh, cudaGetLastError(), "cudaGraphLaunch");
                    if (rc == 0) rc = copy_out(hh, sg.idx[q], sg.m, sg.n, stage);
                    if (rc) rcs[t] = rc;
                    else mak::count_launch();
                }
                if (rcs[t]) break;
            }
        }
    };
    if (nthreads > 1) {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
        for (auto& x : th) x.join();
    } else {
        worker(0);
    }
    int rc = 0;
    for (int t = 0; t < nthreads && rc == 0; ++t)
        if (rcs[t]) { rc = rcs[t]; memcpy(h->err, hs[t].err, sizeof(h->err)); }
    // join the pool streams on the error path too: kernels of other blocks may still be running on them
    for (int s = 0; s < np; ++s) {
        if (cudaEventRecord(h->pool_ev[s], h->pool[s]) == cudaSuccess) cudaStreamWaitEvent(main, h->pool_ev[s], 0);
    }
    return rc;
}

// size classes of a batched QR: three warp-kernel classes (m,n <= 32 by row capacity), the one-CTA
// shared-memory kernel, the lock-step blocked path (panel fits one CTA), and per-block "big"
template <typename T>
struct QrClasses {
    std::vector<int> warp[3], smem[3], blocked, big;   // smem[]: by shared-memory footprint (CTA size class)
    size_t max_se[3] = {0, 0, 0};
    int warp_cap[3] = {0, 0, 0};                       // per-warp shared-memory elements of each warp class
    mak::BqrSchedule sched;            // two-level column schedule of the blocked class (sorted by k descending)
};
static int bqr_warp_max() {   // largest dimension served by the warp-per-block register kernel
    static int v = -1;
    if (v < 0) { const char* e = getenv("MAKB200_BQR_WARP_MAX"); v = e ? atoi(e) : 32; if (v > 32) v = 32; }
    return v;
}
static int bqr_min_dim() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MAKB200_BQR_MIN_DIM"); v = e ? atoi(e) : 65; if (v < 33) v = 33; }   // 65 measured: 65-128 bucket 16.1 -> 12.8 ms vs 96
    return v;
}
template <typename T>
static int classify_qr(int batch, const int* m, const int* n, QrClasses<T>& c) {
    for (int i = 0; i < batch; ++i) {
        if (m[i] < 0 || n[i] < 0) return -4;
        if (m[i] == 0) continue;
        const int k = m[i] < n[i] ? m[i] : n[i];
        size_t se = mak::batched_qr_smem_elems(m[i], n[i]);
        const bool fits_smem = se <= mak::batched_qr_max_smem_elems<T>();
        if (m[i] <= bqr_warp_max() && n[i] <= bqr_warp_max() && n[i] > 0) {
            const int big_ = m[i] > n[i] ? m[i] : n[i];
            const int wc = big_ <= 16 ? 0 : (big_ <= 24 ? 1 : 2);
            c.warp[wc].push_back(i);
            const int ce = (m[i] | 1) * n[i];
            if (ce > c.warp_cap[wc]) c.warp_cap[wc] = ce;
        } else if (fits_smem && (k < bqr_min_dim() || !mak::bqr_fits<T>(m[i], n[i]))) {
            const size_t bytes = (se + 64) * sizeof(T);
            const int cl = bytes <= 24 * 1024 ? 0 : (bytes <= 72 * 1024 ? 1 : 2);
            c.smem[cl].push_back(i);
            if (se > c.max_se[cl]) c.max_se[cl] = se;
        } else if (n[i] > 0 && mak::bqr_fits<T>(m[i], n[i])) {
            c.blocked.push_back(i);
        } else {
            c.big.push_back(i);
        }
    }
    std::stable_sort(c.blocked.begin(), c.blocked.end(), [&](int a, int b) {
        return (m[a] < n[a] ? m[a] : n[a]) > (m[b] < n[b] ? m[b] : n[b]);
    });
    std::vector<int> ms, ns, ks;
    for (int i : c.blocked) { ms.push_back(m[i]); ns.push_back(n[i]); ks.push_back(m[i] < n[i] ? m[i] : n[i]); }
    c.sched = mak::bqr_schedule<T>(ms, ns, ks);
    return 0;
}
template <typename T>
static size_t qr_batched_worksize_t(makb200_handle_t* h, int batch, const int* m, const int* n) {
    QrClasses<T> c;
    if (classify_qr<T>(batch, m, n, c)) return 0;
    const size_t nb_ = (size_t)(batch > 0 ? batch : 1);
    size_t bytes = mak::align_up(sizeof(mak::QrBlockDesc<T>) * nb_, 256) + mak::align_up(sizeof(mak::BqrBlock<T>) * nb_, 256) +
                   mak::align_up(sizeof(mak::GemmProblem<T>) * 4 * nb_, 256);
    size_t welems = 0;
    for (int i : c.blocked) {
        const int k = m[i] < n[i] ? m[i] : n[i];
        welems += mak::bqr_block_work_elems<T>(c.sched, m[i], n[i], k);
    }
    bytes += mak::align_up(welems * sizeof(T), 256);
    size_t big = 0;
    for (int i : c.big) {
        size_t w = mak::qr_worksize_t<T>(h, m[i], n[i], m[i] < n[i] ? m[i] : n[i]);
        if (w > big) big = w;
    }
    return bytes + (c.big.empty() ? 0 : pooled_worksize(big, c.big.size())) + 1024;
}

// A batched-QR plan: classification, workspace carving and the descriptor upload done once;
// run() only launches.  (Block-sparse tensors keep their block structure across many calls.)
struct makb200_qr_batched_plan {
    int dtype;
    virtual ~makb200_qr_batched_plan() {}
    virtual int run(makb200_handle_t* h, int* info) = 0;
};
template <typename T>
struct QrPlan : makb200_qr_batched_plan {
    QrClasses<T> c;
    int batch = 0;
    mak::QrBlockDesc<T>* ddev = nullptr;
    mak::BqrBlock<T>* bdev = nullptr;
    mak::GemmProblem<T>* pdev = nullptr;
    size_t nblocked = 0;
    char* wbig = nullptr;
    size_t lbig = 0;
    std::vector<int> m, n, lda, ldq, ldr;   // host copies for the per-block "big" path
    std::vector<void*> A, Q, R;

    int run(makb200_handle_t* h, int* info) override {
        if (info) MAK_CUDA(h, cudaMemsetAsync(info, 0, sizeof(int) * batch, h->stream));
        size_t off = 0;
        for (int cl = 0; cl < 3; ++cl) {
            if (c.warp[cl].empty()) continue;
            int rc = mak::batched_qr_warp<T>(h, (int)c.warp[cl].size(), c.warp_cap[cl], ddev + off, cl == 0 ? 16 : (cl == 1 ? 24 : 32));
            if (rc) return rc;
            off += c.warp[cl].size();
        }
        for (int cl = 0; cl < 3; ++cl) {
            if (c.smem[cl].empty()) continue;
            int rc = mak::batched_qr_smem<T>(h, (int)c.smem[cl].size(), c.max_se[cl], ddev + off, nullptr);
            if (rc) return rc;
            off += c.smem[cl].size();
        }
        if (nblocked) {
            int rc = mak::batched_qr_blocked<T>(h, (int)nblocked, bdev, c.sched, pdev);
            if (rc) return rc;
        }
        // blocks whose panel does not fit one CTA take the single-matrix blocked DMMA path
        return run_pooled(h, c.big, wbig, lbig, [&](makb200_handle_t* hh, char* w, size_t lw, int i) {
            return mak::qr_fused_t<T>(hh, MAKB200_QR_COMPACT, m[i], n[i], (T*)A[i], lda[i], (T*)Q[i], ldq[i],
                                      R[i] ? (T*)R[i] : nullptr, ldr[i], w, lw);
        });
    }
};

template <typename T>
static int qr_plan_build(makb200_handle_t* h, int batch, const int* m, const int* n, void* const* A, const int* lda,
                         void* const* Q, const int* ldq, void* const* R, const int* ldr, void* work, size_t lwork,
                         QrPlan<T>* pl) {
    QrClasses<T>& c = pl->c;
    int rcc = classify_qr<T>(batch, m, n, c);
    if (rcc) return rcc;
    pl->batch = batch;
    const size_t nb_ = (size_t)(batch > 0 ? batch : 1);
    mak::Arena ar(work, lwork);
    pl->ddev = ar.get<mak::QrBlockDesc<T>>(nb_);
    pl->bdev = ar.get<mak::BqrBlock<T>>(nb_);
    pl->pdev = ar.get<mak::GemmProblem<T>>(4 * nb_);
    size_t welems = 0;
    for (int i : c.blocked) {
        const int k = m[i] < n[i] ? m[i] : n[i];
        welems += mak::bqr_block_work_elems<T>(c.sched, m[i], n[i], k);
    }
    T* wblk = ar.get<T>(welems);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    pl->wbig = (char*)work + ar.off;
    pl->lbig = lwork > ar.off ? lwork - ar.off : 0;
    if (!c.big.empty()) {
        pl->m.assign(m, m + batch); pl->n.assign(n, n + batch);
        pl->lda.assign(lda, lda + batch); pl->ldq.assign(ldq, ldq + batch);
        pl->ldr.assign(batch, 0);
        pl->A.assign(A, A + batch); pl->Q.assign(Q, Q + batch); pl->R.assign(batch, nullptr);
        for (int i = 0; i < batch; ++i) {
            if (R && R[i] && ldr) { pl->R[i] = R[i]; pl->ldr[i] = ldr[i]; }
        }
    }
    // descriptors of every class, one pinned staging upload
    std::vector<mak::QrBlockDesc<T>> descs;
    descs.reserve(batch);
    auto push_desc = [&](int i) {
        mak::QrBlockDesc<T> d;
        d.m = m[i]; d.n = n[i];
        d.A = (T*)A[i]; d.lda = lda[i];
        d.Q = (T*)Q[i]; d.ldq = ldq[i];
        d.R = (R && R[i]) ? (T*)R[i] : nullptr; d.ldr = ldr ? ldr[i] : 0;
        descs.push_back(d);
    };
    for (int cl = 0; cl < 3; ++cl) for (int i : c.warp[cl]) push_desc(i);
    for (int cl = 0; cl < 3; ++cl) for (int i : c.smem[cl]) push_desc(i);
    std::vector<mak::BqrBlock<T>> bl;
    bl.reserve(c.blocked.size());
    {
        T* p = wblk;
        for (int i : c.blocked) {
            const int k = m[i] < n[i] ? m[i] : n[i];
            mak::BqrBlock<T> b;
            b.m = m[i]; b.n = n[i]; b.k = k;
            b.A = (T*)A[i]; b.lda = lda[i];
            b.Q = (T*)Q[i]; b.ldq = ldq[i];
            b.R = (R && R[i]) ? (T*)R[i] : nullptr; b.ldr = ldr ? ldr[i] : 0;
            mak::bqr_carve_block<T>(c.sched, b, p);
            bl.push_back(b);
        }
    }
    pl->nblocked = bl.size();
    {
        mak::Stager st(h, descs.size() * sizeof(mak::QrBlockDesc<T>) + bl.size() * sizeof(mak::BqrBlock<T>) + 1024);
        MAK_CUDA(h, st.put(pl->ddev, descs.data(), descs.size() * sizeof(mak::QrBlockDesc<T>), h->stream));
        MAK_CUDA(h, st.put(pl->bdev, bl.data(), bl.size() * sizeof(mak::BqrBlock<T>), h->stream));
    }
    return 0;
}

template <typename T>
static int qr_batched_t(makb200_handle_t* h, int batch, const int* m, const int* n, void* const* A, const int* lda,
                        void* const* Q, const int* ldq, void* const* R, const int* ldr, int* info, void* work,
                        size_t lwork) {
    QrPlan<T> pl;
    int rc = qr_plan_build<T>(h, batch, m, n, A, lda, Q, ldq, R, ldr, work, lwork, &pl);
    if (rc) return rc;
    return pl.run(h, info);
}

extern "C" {

size_t makb200_qr_batched_worksize(makb200_handle_t* h, int dtype, int batch, const int* m, const int* n) {
    if (!h || !dtype_ok(dtype) || batch < 0 || (batch > 0 && (!m || !n))) return 0;
    return dtype == MAKB200_F64 ? qr_batched_worksize_t<double>(h, batch, m, n)
                                : qr_batched_worksize_t<cplx>(h, batch, m, n);
}

int makb200_qr_batched(makb200_handle_t* h, int dtype, int batch, const int* m, const int* n, void* const* A,
                       const int* lda, void* const* Q, const int* ldq, void* const* R, const int* ldr, int* info,
                       void* work, size_t lwork) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (batch < 0) return -3;
    if (batch == 0) return 0;
    if (!m) return -4;
    if (!n) return -5;
    if (!A) return -6;
    if (!lda) return -7;
    if (!Q) return -8;
    if (!ldq) return -9;
    if (dtype == MAKB200_F64) return qr_batched_t<double>(h, batch, m, n, A, lda, Q, ldq, R, ldr, info, work, lwork);
    return qr_batched_t<cplx>(h, batch, m, n, A, lda, Q, ldq, R, ldr, info, work, lwork);
}


int makb200_qr_batched_plan_create(makb200_handle_t* h, int dtype, int batch, const int* m, const int* n,
                                   void* const* A, const int* lda, void* const* Q, const int* ldq, void* const* R,
                                   const int* ldr, void* work, size_t lwork, makb200_qr_batched_plan_t** plan) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (batch <= 0) return -3;
    if (!m) return -4;
    if (!n) return -5;
    if (!A) return -6;
    if (!lda) return -7;
    if (!Q) return -8;
    if (!ldq) return -9;
    if (!plan) return -14;
    *plan = nullptr;
    int rc;
    if (dtype == MAKB200_F64) {
        auto* pl = new QrPlan<double>();
        pl->dtype = dtype;
        rc = qr_plan_build<double>(h, batch, m, n, A, lda, Q, ldq, R, ldr, work, lwork, pl);
        if (rc) { delete pl; return rc; }
        *plan = pl;
    } else {
        auto* pl = new QrPlan<cplx>();
        pl->dtype = dtype;
        rc = qr_plan_build<cplx>(h, batch, m, n, A, lda, Q, ldq, R, ldr, work, lwork, pl);
        if (rc) { delete pl; return rc; }
        *plan = pl;
    }
    return 0;
}

int makb200_qr_batched_plan_run(makb200_handle_t* h, makb200_qr_batched_plan_t* plan, int* info) {
    if (!h) return -1;
    if (!plan) return -2;
    return plan->run(h, info);
}

int makb200_qr_batched_plan_destroy(makb200_qr_batched_plan_t* plan) {
    delete plan;
    return 0;
}


// ---- eigh ---------------------------------------------------------------------------
int makb200_hermitian_defect(makb200_handle_t* h, int dtype, int n, const void* A, int lda, double* out2_dev) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (n < 0) return -3;
    if (lda < maxi(1, n)) return -5;
    if (!out2_dev) return -6;
    if (n > 0 && !A) return -4;
    if (dtype == MAKB200_F64) return mak::herm_defect_t<double>(h, n, (const double*)A, lda, out2_dev);
    return mak::herm_defect_t<cplx>(h, n, (const cplx*)A, lda, out2_dev);
}

int makb200_project_hermitian(makb200_handle_t* h, int dtype, int anti, int n, const void* A, int lda, void* B,
                              int ldb) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (n < 0) return -4;
    if (lda < maxi(1, n)) return -6;
    if (ldb < maxi(1, n)) return -8;
    if (n == 0) return 0;
    if (!A) return -5;
    if (!B) return -7;
    if (B == A && ldb != lda) return -8;   // in place means the same matrix, not an overlapping one
    if (dtype == MAKB200_F64) return mak::project_herm_t<double>(h, anti != 0, n, (const double*)A, lda, (double*)B, ldb);
    return mak::project_herm_t<cplx>(h, anti != 0, n, (const cplx*)A, lda, (cplx*)B, ldb);
}

int makb200_hermitian_props(makb200_handle_t* h, int dtype, int anti, int n, const void* A, int lda,
                            double* out4_dev) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (n < 0) return -4;
    if (lda < maxi(1, n)) return -6;
    if (!out4_dev) return -7;
    if (n > 0 && !A) return -5;
    if (dtype == MAKB200_F64) return mak::herm_props_t<double>(h, anti != 0, n, (const double*)A, lda, out4_dev);
    return mak::herm_props_t<cplx>(h, anti != 0, n, (const cplx*)A, lda, out4_dev);
}

int makb200_tri_init(makb200_handle_t* h, int dtype, int mode, int m, int n, void* A, int lda) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (mode < 0 || mode > 2) return -3;
    if (m < 0) return -4;
    if (n < 0) return -5;
    if (lda < maxi(1, m)) return -7;
    if (m == 0 || n == 0) return 0;
    if (!A) return -6;
    if (dtype == MAKB200_F64) return mak::tri_init_t<double>(h, mode, m, n, (double*)A, lda);
    return mak::tri_init_t<cplx>(h, mode, m, n, (cplx*)A, lda);
}

int makb200_fro2(makb200_handle_t* h, int dtype, int m, int n, const void* A, int lda, double* out1_dev) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (m < 0) return -3;
    if (n < 0) return -4;
    if (lda < maxi(1, m)) return -6;
    if (!out1_dev) return -7;
    if (m > 0 && n > 0 && !A) return -5;
    if (dtype == MAKB200_F64) return mak::fro2_t<double>(h, m, n, (const double*)A, lda, out1_dev);
    return mak::fro2_t<cplx>(h, m, n, (const cplx*)A, lda, out1_dev);
}

int makb200_gram_defect(makb200_handle_t* h, int dtype, int n, const void* P, int ldp, double* out2_dev) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (n < 0) return -3;
    if (ldp < maxi(1, n)) return -5;
    if (!out2_dev) return -6;
    if (n > 0 && !P) return -4;
    if (dtype == MAKB200_F64) return mak::gram_defect_t<double>(h, n, (const double*)P, ldp, out2_dev);
    return mak::gram_defect_t<cplx>(h, n, (const cplx*)P, ldp, out2_dev);
}

size_t makb200_eigh_worksize(makb200_handle_t* h, int dtype, int n) {
    if (!h || !dtype_ok(dtype) || n < 0) return 0;
    return dtype == MAKB200_F64 ? mak::eigh_worksize_t<double>(h, n) : mak::eigh_worksize_t<cplx>(h, n);
}

int makb200_eigh(makb200_handle_t* h, int dtype, int fixgauge, int n, void* A, int lda, double* W, void* V, int ldv,
                 void* work, size_t lwork, int* info_dev) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (n < 0) return -4;
    if (lda < maxi(1, n)) return -6;
    if (V && ldv < maxi(1, n)) return -9;
    if (n == 0) return 0;
    if (!A) return -5;
    if (!W) return -7;
    if (V == A) return -8;   // V == NULL: values only (job 'N')
    if (dtype == MAKB200_F64)
        return mak::eigh_t<double>(h, n, (double*)A, lda, W, (double*)V, ldv, fixgauge, work, lwork, info_dev);
    return mak::eigh_t<cplx>(h, n, (cplx*)A, lda, W, (cplx*)V, ldv, fixgauge, work, lwork, info_dev);
}

size_t makb200_stedc_worksize(makb200_handle_t* h, int n) {
    if (!h || n < 0) return 0;
    return mak::stedc_worksize(n);
}

int makb200_stedc(makb200_handle_t* h, int n, const double* d, const double* e, double* W, double* Z, int ldz,
                  void* work, size_t lwork, int* info_dev) {
    if (!h) return -1;
    if (n < 0) return -2;
    if (ldz < maxi(1, n)) return -7;
    if (n == 0) return 0;
    if (!d) return -3;
    if (n > 1 && !e) return -4;
    if (!W) return -5;
    if (!Z) return -6;
    return mak::stedc(h, n, d, e, W, Z, ldz, work, lwork, info_dev);
}


// ---- polar / svd ---------------------------------------------------------------------
static double qdwh_l0(double l0) { return (l0 > 0.0 && l0 < 1.0) ? l0 : 2.2e-16; }

size_t makb200_polar_worksize(makb200_handle_t* h, int dtype, int m, int n) {
    if (!h || !dtype_ok(dtype) || m < 0 || n < 0 || m < n) return 0;
    return dtype == MAKB200_F64 ? mak::polar_worksize_t<double>(h, m, n) : mak::polar_worksize_t<cplx>(h, m, n);
}

int makb200_polar_qdwh(makb200_handle_t* h, int dtype, int m, int n, void* A, int lda, void* W, int ldw, void* P,
                       int ldp, double l0, int maxiter, void* work, size_t lwork, int* iters_host, int* info_dev) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (m < 0) return -3;
    if (n < 0 || n > m) return -4;  // `left_polar!` requires m >= n (implementations/polar.jl:9-10)
    if (lda < maxi(1, m)) return -6;
    if (ldw < maxi(1, m)) return -8;
    if (P && ldp > 0 && ldp < maxi(1, n)) return -10;
    if (m == 0 || n == 0) return 0;
    if (!A) return -5;
    if (!W || W == A) return -7;
    if (maxiter <= 0) maxiter = 12;
    if (dtype == MAKB200_F64)
        return mak::polar_qdwh_t<double>(h, m, n, (double*)A, lda, (double*)W, ldw, (double*)P, ldp, qdwh_l0(l0),
                                         maxiter, work, lwork, iters_host, info_dev);
    return mak::polar_qdwh_t<cplx>(h, m, n, (cplx*)A, lda, (cplx*)W, ldw, (cplx*)P, ldp, qdwh_l0(l0), maxiter, work,
                                   lwork, iters_host, info_dev);
}

size_t makb200_svd_worksize(makb200_handle_t* h, int dtype, int m, int n) {
    if (!h || !dtype_ok(dtype) || m < 0 || n < 0) return 0;
    return dtype == MAKB200_F64 ? mak::svd_worksize_t<double>(h, m, n) : mak::svd_worksize_t<cplx>(h, m, n);
}

int makb200_svd(makb200_handle_t* h, int dtype, int fixgauge, int m, int n, void* A, int lda, double* S, void* U,
                int ldu, void* Vh, int ldvh, double l0, void* work, size_t lwork, int* info_dev) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (m < 0) return -4;
    if (n < 0) return -5;
    int k = m < n ? m : n;
    if (lda < maxi(1, m)) return -7;
    if (U && ldu < maxi(1, m)) return -10;
    if (Vh && ldvh < maxi(1, k)) return -12;
    if ((U == nullptr) != (Vh == nullptr)) return -11;  // both or neither (job 'S' or 'N', yalapack.jl:2105-2127)
    if (m == 0 || n == 0) return 0;                       // empty: the host layer fills identities (svd.jl:197)
    if (!A) return -6;
    if (!S) return -8;
    if (U == A || Vh == A) return -9;
    if (dtype == MAKB200_F64)
        return mak::svd_t<double>(h, m, n, (double*)A, lda, S, (double*)U, ldu, (double*)Vh, ldvh, fixgauge,
                                  qdwh_l0(l0), work, lwork, info_dev);
    return mak::svd_t<cplx>(h, m, n, (cplx*)A, lda, S, (cplx*)U, ldu, (cplx*)Vh, ldvh, fixgauge, qdwh_l0(l0), work,
                            lwork, info_dev);
}

int makb200_svd_leading(makb200_handle_t* h, int dtype, int fixgauge, int m, int n, int r, void* A, int lda, double* S,
                        void* U, int ldu, void* Vh, int ldvh, double l0, void* work, size_t lwork, int* info_dev) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (m < 0) return -4;
    if (n < 0) return -5;
    const int k = m < n ? m : n;
    if (r < 0 || r > k) return -6;
    if (lda < maxi(1, m)) return -8;
    if (ldu < maxi(1, m)) return -11;
    if (ldvh < maxi(1, r)) return -13;
    if (m == 0 || n == 0 || r == 0) return r == 0 && m > 0 && n > 0 ? -6 : 0;   // r = 0: use makb200_svd with U = Vh = NULL
    if (!A) return -7;
    if (!S) return -9;
    if (!U || U == A) return -10;
    if (!Vh || Vh == A) return -12;
    if (dtype == MAKB200_F64)
        return mak::svd_t<double>(h, m, n, (double*)A, lda, S, (double*)U, ldu, (double*)Vh, ldvh, fixgauge,
                                  qdwh_l0(l0), work, lwork, info_dev, r);
    return mak::svd_t<cplx>(h, m, n, (cplx*)A, lda, S, (cplx*)U, ldu, (cplx*)Vh, ldvh, fixgauge, qdwh_l0(l0), work,
                            lwork, info_dev, r);
}


// ---- tall-skinny local QR for TSQR ------------------------------------------------------
size_t makb200_tsqr_local_worksize(makb200_handle_t* h, int dtype, int m, int n) {
    if (!h || !dtype_ok(dtype) || m < 0 || n < 0) return 0;
    return dtype == MAKB200_F64 ? mak::cholqr2_worksize_t<double>(h, m, n) : mak::cholqr2_worksize_t<cplx>(h, m, n);
}

int makb200_tsqr_local_ex(makb200_handle_t* h, int dtype, int m, int n, void* A, int lda, void* Q, int ldq, void* R,
                          int ldr, int nshift, void* work, size_t lwork, int* info_dev) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (m < 0) return -3;
    if (n < 0 || n > m) return -4;
    if (lda < maxi(1, m)) return -6;
    if (ldq < maxi(1, m)) return -8;
    if (R && ldr > 0 && ldr < maxi(1, n)) return -10;
    if (nshift < 0 || nshift > 3) return -11;
    if (m == 0 || n == 0) return 0;
    if (!A) return -5;
    if (!Q || Q == A) return -7;
    if (dtype == MAKB200_F64)
        return mak::cholqr2_t<double>(h, m, n, (double*)A, lda, (double*)Q, ldq, (double*)R, ldr, work, lwork, info_dev,
                                      nshift);
    return mak::cholqr2_t<cplx>(h, m, n, (cplx*)A, lda, (cplx*)Q, ldq, (cplx*)R, ldr, work, lwork, info_dev, nshift);
}

int makb200_tsqr_local(makb200_handle_t* h, int dtype, int m, int n, void* A, int lda, void* Q, int ldq, void* R,
                       int ldr, void* work, size_t lwork, int* info_dev) {
    int rc = makb200_tsqr_local_ex(h, dtype, m, n, A, lda, Q, ldq, R, ldr, 0, work, lwork, info_dev);
    return rc == -11 ? -1 : (rc <= -12 ? rc + 1 : rc);
}

int makb200_gauge_columns(makb200_handle_t* h, int dtype, int m, int ncols, void* V, int ldv) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (m < 0) return -3;
    if (ncols < 0) return -4;
    if (ldv < maxi(1, m)) return -6;
    if (m == 0 || ncols == 0) return 0;
    if (!V) return -5;
    mak::count_launch();
    if (dtype == MAKB200_F64) return mak::gauge_columns<double>(h, m, ncols, (double*)V, ldv, (double*)nullptr, 0, 0);
    return mak::gauge_columns<cplx>(h, m, ncols, (cplx*)V, ldv, (cplx*)nullptr, 0, 0);
}

// ---- L1 shim: ormqr / unmqr (left side) -----------------------------------------------------
size_t makb200_ormqr_worksize(makb200_handle_t* h, int dtype, int m, int n, int k) {
    if (!h || !dtype_ok(dtype) || m < 0 || n < 0 || k < 0) return 0;
    return dtype == MAKB200_F64 ? mak::ormqr_worksize_t<double>(h, m, k, n) : mak::ormqr_worksize_t<cplx>(h, m, k, n);
}

int makb200_ormqr(makb200_handle_t* h, int dtype, int side, int trans, int m, int n, int k, const void* A, int lda,
                  const void* tau, void* C, int ldc, void* work, size_t lwork) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (side != 0) return -3;                               // left side only
    if (!op_ok(trans)) return -4;
    if (trans == MAKB200_OP_T && dtype == MAKB200_C128) return -4;   // LAPACK zunmqr: 'N' or 'C'
    if (m < 0) return -5;
    if (n < 0) return -6;
    if (k < 0 || k > m) return -7;
    if (lda < maxi(1, m)) return -9;
    if (ldc < maxi(1, m)) return -12;
    if (m == 0 || n == 0 || k == 0) return 0;
    if (!A) return -8;
    if (!tau) return -10;
    if (!C || C == A) return -11;
    const bool adj = trans != MAKB200_OP_N;
    if (dtype == MAKB200_F64)
        return mak::ormqr_left_t<double>(h, m, k, (const double*)A, lda, (const double*)tau, (double*)C, ldc, n, work, lwork, adj);
    return mak::ormqr_left_t<cplx>(h, m, k, (const cplx*)A, lda, (const cplx*)tau, (cplx*)C, ldc, n, work, lwork, adj);
}

// ---- multi-GPU TSQR (NCCL resolved at run time) --------------------------------------------
size_t makb200_tsqr_worksize(makb200_handle_t* h, int dtype, int m, int n, int nranks) {
    if (!h || !dtype_ok(dtype) || m < 0 || n < 0 || nranks < 1) return 0;
    return dtype == MAKB200_F64 ? mak::tsqr_worksize_t<double>(h, m, n, nranks) : mak::tsqr_worksize_t<cplx>(h, m, n, nranks);
}

int makb200_nccl_unique_id(void* id128) {
    if (!id128) return -1;
    const mak::NcclApi* api = mak::nccl_api();
    if (!api) return MAKB200_ERR_NCCL;
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return MAKB200_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return 0;
}

int makb200_comm_create(void** comm, int nranks, int rank, const void* id128) {
    if (!comm) return -1;
    if (nranks < 1) return -2;
    if (rank < 0 || rank >= nranks) return -3;
    if (!id128) return -4;
    const mak::NcclApi* api = mak::nccl_api();
    if (!api) return MAKB200_ERR_NCCL;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t c = nullptr;
    if (api->CommInitRank(&c, nranks, id, rank) != ncclSuccess) return MAKB200_ERR_NCCL;
    *comm = (void*)c;
    return 0;
}

int makb200_comm_destroy(void* comm) {
    if (!comm) return -1;
    const mak::NcclApi* api = mak::nccl_api();
    if (!api) return MAKB200_ERR_NCCL;
    return api->CommDestroy((ncclComm_t)comm) == ncclSuccess ? 0 : MAKB200_ERR_NCCL;
}

int makb200_tsqr(makb200_handle_t* h, void* comm, int dtype, int m, int n, void* A, int lda, void* Q, int ldq, void* R,
                 int ldr, void* work, size_t lwork, int* info_dev) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -3;
    if (m < 0) return -4;
    if (n < 0) return -5;
    if (lda < maxi(1, m)) return -7;
    if (ldq < maxi(1, m)) return -9;
    if (R && ldr > 0 && ldr < maxi(1, n)) return -11;
    if (m > 0 && n > 0 && !A) return -6;
    if (m > 0 && n > 0 && (!Q || Q == A)) return -8;
    if (!comm && n > m) return -5;
    const mak::NcclApi* api = nullptr;
    if (comm) {
        const char* why = nullptr;
        api = mak::nccl_api(&why);
        if (!api) {
            snprintf(h->err, sizeof(h->err), "libnccl.so.2 not loadable: %s", why ? why : "?");
            return MAKB200_ERR_NCCL;
        }
    }
    if (dtype == MAKB200_F64)
        return mak::tsqr_t<double>(h, api, (ncclComm_t)comm, m, n, (double*)A, lda, (double*)Q, ldq, (double*)R, ldr, work,
                                   lwork, info_dev);
    return mak::tsqr_t<cplx>(h, api, (ncclComm_t)comm, m, n, (cplx*)A, lda, (cplx*)Q, ldq, (cplx*)R, ldr, work, lwork,
                             info_dev);
}

}  // extern "C"

#include "polar_lockstep.cuh"
#include "eigh_lockstep.cuh"
// ---- lock-step QDWH of a chunk of mid-size blocks (phase 1 of the phased batched SVD) -------------------------
static bool lockstep_enabled() {
    const char* e = getenv("MAKB200_SVD_LOCKSTEP");   // read per call: tests run both paths
    return !(e && e[0] == '0');
}
// workspace of the lock-step phase for `chunk` blocks no larger than mmax x nmax: tables, per-block buffers and the
// batched QR of the stacked [sqrt(c) X; I] (and of the tall blocks)
template <typename T>
static size_t svd_lockstep_bytes(makb200_handle_t* h, size_t chunk, int mmax, int nmax, bool any_tall, size_t* blocks_bytes,
                                 size_t* tables_bytes, size_t* qr_bytes) {
    constexpr int nb = mak::CholNB<T>::value;
    const size_t tb = mak::ls_tables_bytes<T>((int)chunk, nmax, any_tall);
    const int mrow = any_tall ? std::max(mmax, nmax + 1) : nmax;
    const size_t bb = chunk * mak::align_up(mak::ls_block_elems<T>(mrow, nmax, nb) * sizeof(T), 256);
    std::vector<int> m2(chunk, 2 * nmax), nn(chunk, nmax);
    size_t qb = qr_batched_worksize_t<T>(h, (int)chunk, m2.data(), nn.data());
    if (any_tall) {
        std::vector<int> mt(chunk, mrow);
        qb = std::max(qb, qr_batched_worksize_t<T>(h, (int)chunk, mt.data(), nn.data()));
    }
    qb = mak::align_up(qb, 256) + 4096;
    if (blocks_bytes) *blocks_bytes = bb;
    if (tables_bytes) *tables_bytes = tb;
    if (qr_bytes) *qr_bytes = qb;
    // the same region then serves the lock-step eigensolve and SVD tail of the chunk
    const size_t phase3 = mak::eigh_lockstep_bytes<T>((int)chunk, nmax) + mak::svd_tail_lockstep_bytes<T>((int)chunk) + 1024;
    return std::max(tb + bb + qb, phase3) + 1024;
}

// ---- batched svd --------------------------------------------------------------------------
constexpr int SVD_PHASE_CHUNK = 192;     // blocks per one-launch tridiagonalisation (>= one CTA per SM)
static bool phased_enabled() {
    const char* e = getenv("MAKB200_SVD_PHASED");   // read per call: tests run both paths
    return !(e && e[0] == '0');
}
// per-block storage that lives across the phases: W (m x n), P, V (n x n), tau, wv, d, e, flag
template <typename T>
static size_t svd_phase_persist_bytes(int m, int n) {
    const size_t mm = (size_t)m, nn = (size_t)n;
    return mak::align_up(mm * nn * sizeof(T), 256) + 2 * mak::align_up(nn * nn * sizeof(T), 256) + mak::align_up(nn * sizeof(T), 256) +
           3 * mak::align_up(nn * 8, 256) + 256;
}

// ---- chunks of the phased / lock-step batched SVD ---------------------------------------------------------------
// `ph` (phased blocks: m >= n, 3 <= n <= BHETRD_MAX_N) is sorted by n descending.  A chunk takes blocks while its
// per-block buffers stay under SVD_LS_BUDGET (at least SVD_PHASE_CHUNK, at most SVD_LS_MAX_CHUNK blocks): 192 blocks of
// 512 x 512 but ~2000 of 100 x 100, because a chunk costs ~700 launches whatever its size and the small-block buckets
// were bound by exactly that.  Used by the workspace query and the run, so both see the same chunks.
constexpr size_t SVD_LS_BUDGET = (size_t)8 << 30;
constexpr size_t SVD_LS_MAX_CHUNK = 2048;
constexpr int SVD_LS_MIN_CHUNK_DEFAULT = 296;   // 2 x 148 SMs: 2412 -> 2300 ms on the 257-512 bucket of config 3 (192 blocks = 1.3 waves of the one-launch tridiagonalisation)
// least blocks per lock-step chunk: two CTAs of the one-launch tridiagonalisation per SM (MAKB200_SVD_CHUNK_MIN overrides)
static size_t svd_ls_min_chunk() {
    const char* e = getenv("MAKB200_SVD_CHUNK_MIN");   // read per call (workspace query and run see the same value)
    const int v = e ? atoi(e) : SVD_LS_MIN_CHUNK_DEFAULT;
    return (size_t)(v < 8 ? 8 : v);
}
struct SvdChunk { size_t c0, nc; int mmax, nmax; bool tall; size_t persist, ls_total, ls_bb, ls_tb, ls_qb; };
template <typename T>
static std::vector<SvdChunk> svd_chunk_plan(makb200_handle_t* h, const std::vector<int>& ph, const int* m, const int* n,
                                            size_t* region) {
    constexpr int lnb = mak::CholNB<T>::value;
    std::vector<SvdChunk> plan;
    size_t reg = 0;
    for (size_t c0 = 0; c0 < ph.size();) {
        SvdChunk c{};
        c.c0 = c0;
        size_t bytes = 0;
        while (c0 + c.nc < ph.size() && c.nc < SVD_LS_MAX_CHUNK) {
            const int i = ph[c0 + c.nc];
            const size_t bi = mak::ls_block_elems<T>(m[i], n[i], lnb) * sizeof(T) + svd_phase_persist_bytes<T>(m[i], n[i]);
            if (c.nc >= svd_ls_min_chunk() && bytes + bi > SVD_LS_BUDGET) break;
            bytes += bi;
            c.mmax = std::max(c.mmax, m[i]); c.nmax = std::max(c.nmax, n[i]);
            c.tall = c.tall || m[i] > n[i];
            c.persist = std::max(c.persist, svd_phase_persist_bytes<T>(m[i], n[i]));
            ++c.nc;
        }
        c.ls_total = svd_lockstep_bytes<T>(h, c.nc, c.mmax, c.nmax, c.tall, &c.ls_bb, &c.ls_tb, &c.ls_qb);
        reg = std::max(reg, c.nc * (c.persist + sizeof(mak::BhetrdDesc<T>) + 64) + 512 + c.ls_total);
        plan.push_back(c);
        c0 += c.nc;
    }
    if (region) *region = reg;
    return plan;
}
// the phased blocks of a batch (indices), n descending
template <typename T>
static std::vector<int> svd_phased_blocks(int batch, const int* m, const int* n) {
    std::vector<int> ph;
    for (int i = 0; i < batch; ++i) {
        if (m[i] <= 0 || n[i] <= 0) continue;
        if (mak::batched_svd_smem_bytes(m[i], n[i], sizeof(T)) <= mak::batched_svd_max_smem_bytes()) continue;
        if (m[i] >= n[i] && n[i] >= 3 && n[i] <= mak::BHETRD_MAX_N) ph.push_back(i);
    }
    std::stable_sort(ph.begin(), ph.end(), [&](int a, int b) { return n[a] > n[b]; });
    return ph;
}

template <typename T>
static int svd_batched_t(makb200_handle_t* h, int fixgauge, int batch, const int* m, const int* n, void* const* A,
                         const int* lda, void* const* S, void* const* U, const int* ldu, void* const* Vh,
                         const int* ldvh, int* info, void* work, size_t lwork) {
    std::vector<mak::SvdBlockDesc<T>> small, small2;   // small: <= 40 KB of shared memory (128-thread CTAs)
    std::vector<int> big;
    size_t max_b = 0, max_b2 = 0;
    for (int i = 0; i < batch; ++i) {
        if (m[i] < 0 || n[i] < 0) return -5;
        if (m[i] == 0 || n[i] == 0) continue;
        size_t sb = mak::batched_svd_smem_bytes(m[i], n[i], sizeof(T));
        if (sb > mak::batched_svd_max_smem_bytes()) { big.push_back(i); continue; }
        mak::SvdBlockDesc<T> d;
        d.m = m[i]; d.n = n[i]; d.fixgauge = fixgauge;
        d.A = (const T*)A[i]; d.lda = lda[i];
        d.S = (double*)S[i];
        d.U = U ? (T*)U[i] : nullptr; d.ldu = ldu ? ldu[i] : 0;
        d.Vh = Vh ? (T*)Vh[i] : nullptr; d.ldvh = ldvh ? ldvh[i] : 0;
        if (!d.U || !d.Vh) { d.U = nullptr; d.Vh = nullptr; }
        if (sb <= 40 * 1024) { small.push_back(d); if (sb > max_b) max_b = sb; }
        else { small2.push_back(d); if (sb > max_b2) max_b2 = sb; }
    }
    mak::Arena ar(work, lwork);
    mak::SvdBlockDesc<T>* ddev = ar.get<mak::SvdBlockDesc<T>>(batch > 0 ? batch : 1);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    if (info) MAK_CUDA(h, cudaMemsetAsync(info, 0, sizeof(int) * batch, h->stream));
    if (!small.empty() || !small2.empty()) {
        {
            mak::Stager st(h, (small.size() + small2.size()) * sizeof(mak::SvdBlockDesc<T>) + 1024);
            MAK_CUDA(h, st.put(ddev, small.data(), small.size() * sizeof(mak::SvdBlockDesc<T>), h->stream));
            MAK_CUDA(h, st.put(ddev + small.size(), small2.data(), small2.size() * sizeof(mak::SvdBlockDesc<T>), h->stream));
        }
        int rc = mak::batched_svd_smem<T>(h, (int)small.size(), max_b, ddev, nullptr);
        if (rc) return rc;
        rc = mak::batched_svd_smem<T>(h, (int)small2.size(), max_b2, ddev + small.size(), nullptr);
        if (rc) return rc;
    }
    // blocks too large for one CTA's shared memory: QDWH + D&C path, one block at a time
    char* wbig = (char*)work + ar.off;
    size_t lbig = lwork > ar.off ? lwork - ar.off : 0;
    auto per_block = [&](makb200_handle_t* hh, char* w, size_t lw, int i) {
        return mak::svd_t<T>(hh, m[i], n[i], (T*)A[i], lda[i], (double*)S[i], U ? (T*)U[i] : nullptr, ldu ? ldu[i] : 0,
                             Vh ? (T*)Vh[i] : nullptr, ldvh ? ldvh[i] : 0, fixgauge, 2.2e-16, w, lw, nullptr);
    };
    if (big.empty()) return 0;
    const bool vectors = (U != nullptr && Vh != nullptr);
    // ---- phased path (default): the device dispatches only ~350 k kernels/s however many streams feed it
    // (profiles/r2_batched_graphs.log), so what counts is kernels per block, and 2n of a block's ~3.5n launches are the
    // per-column kernels of the tridiagonalisation of P.  Chunks of blocks go through
    //   1. QDWH per block on the stream pool            -> W, P in per-block storage
    //   2. ONE launch that tridiagonalises every P of the chunk (one CTA per block, csrc/bhetrd.cuh)
    //   3. tridiagonal D&C, back-transformation, U = W V, gauge per block on the pool
    if (vectors && phased_enabled() && !graphs_enabled()) {
        std::vector<int> ph, rest;
        for (int i : big) ((m[i] >= n[i] && n[i] >= 3 && n[i] <= mak::BHETRD_MAX_N) ? ph : rest).push_back(i);
        // n descending: a chunk holds blocks of similar size (the lock-step launches are sized by the largest)
        std::stable_sort(ph.begin(), ph.end(), [&](int a, int b) { return n[a] > n[b]; });
        size_t scratch = 0, persist = 0;
        for (int i : ph) {
            scratch = std::max(scratch, mak::svd_phase_scratch_t<T>(h, m[i], n[i]));
            persist = std::max(persist, svd_phase_persist_bytes<T>(m[i], n[i]));
        }
        const size_t np_pool = (size_t)pool_streams();
        const size_t pool_region = (scratch + 1024) * np_pool;
        // lock-step phases (default): chunks sized by their buffers, the lock-step region at the end of the workspace;
        // without room for it every chunk is SVD_PHASE_CHUNK blocks and runs block by block on the pool as before
        std::vector<SvdChunk> plan;
        bool lockstep = lockstep_enabled() && ph.size() >= 8;
        if (lockstep) {
            size_t region = 0;
            plan = svd_chunk_plan<T>(h, ph, m, n, &region);
            if (lbig < pool_region + 4096 + region) { lockstep = false; plan.clear(); }
        }
        if (!lockstep && !ph.empty() && lbig > pool_region + 4096) {
            const size_t chunk = std::min<size_t>({(size_t)SVD_PHASE_CHUNK, ph.size(), (lbig - pool_region - 4096) / (persist + sizeof(mak::BhetrdDesc<T>) + 64)});
            if (chunk >= 8)
                for (size_t c0 = 0; c0 < ph.size(); c0 += chunk) {
                    SvdChunk c{};
                    c.c0 = c0; c.nc = std::min(chunk, ph.size() - c0); c.persist = persist;
                    plan.push_back(c);
                }
        }
        if (ph.size() >= 8 && !plan.empty()) {
            size_t chunk_max = 0;
            for (const SvdChunk& c : plan) chunk_max = std::max(chunk_max, c.nc);
            char* pbase = wbig + pool_region;
            mak::BhetrdDesc<T>* ddev = (mak::BhetrdDesc<T>*)pbase;
            struct Blk { T *W, *P, *V, *tau; double *wv, *flag, *d, *e; };
            std::vector<Blk> blk(chunk_max);
            std::vector<mak::TrdPre<T>> pre(chunk_max);
            std::vector<int> slot_of(batch, -1);
            for (const SvdChunk& ck : plan) {
                const size_t c0 = ck.c0, nc = ck.nc, persist = ck.persist;
                const size_t ls_total = ck.ls_total, ls_bb = ck.ls_bb, ls_tb = ck.ls_tb, ls_qb = ck.ls_qb;
                char* store = pbase + mak::align_up(sizeof(mak::BhetrdDesc<T>) * nc, 256);
                char* ls_base = wbig + lbig - ls_total;   // [tables | per-block buffers | batched-QR workspace]
                std::vector<int> ids(ph.begin() + c0, ph.begin() + c0 + nc);
                std::vector<mak::BhetrdDesc<T>> bd(nc);
                int nmax = 0;
                for (size_t q = 0; q < nc; ++q) {
                    const int i = ids[q];
                    const size_t mm = (size_t)m[i], nn = (size_t)n[i];
                    char* pq = store + q * persist;
                    Blk& b = blk[q];
                    size_t off = 0;
                    auto take = [&](size_t bytes) { char* r = pq + off; off += mak::align_up(bytes, 256); return r; };
                    b.W = (T*)take(mm * nn * sizeof(T));
                    b.P = (T*)take(nn * nn * sizeof(T));
                    b.V = (T*)take(nn * nn * sizeof(T));
                    b.tau = (T*)take(nn * sizeof(T));
                    b.wv = (double*)take(nn * 8);
                    b.d = (double*)take(nn * 8);
                    b.e = (double*)take(nn * 8);
                    b.flag = (double*)take(16);
                    pre[q] = mak::TrdPre<T>{b.d, b.e, b.tau};
                    bd[q] = mak::BhetrdDesc<T>{n[i], b.P, n[i], b.d, b.e, b.tau};
                    slot_of[i] = (int)q;
                    if (n[i] > nmax) nmax = n[i];
                }
                int rc = 0;
                bool ls_done = false;
                if (lockstep) {
                    // phase 1 in lock-step: every block of the chunk advances through the same QDWH schedule, one grouped
                    // GEMM / batched kernel per step (csrc/polar_lockstep_plan.h)
                    constexpr int lnb = mak::CholNB<T>::value;
                    std::vector<mak::LsBlk<T>> lb(nc);
                    char* ls_tables = (char*)mak::align_up((size_t)(uintptr_t)ls_base, 256);
                    T* bp = (T*)(ls_tables + ls_tb);
                    char* ls_qr = (char*)bp + ls_bb;
                    std::vector<int> m2(nc), nn(nc), mt, nt;
                    for (size_t q = 0; q < nc; ++q) {
                        const int i = ids[q];
                        mak::LsBlk<T>& l = lb[q];
                        l = mak::LsBlk<T>{};
                        l.m = m[i]; l.n = n[i];
                        l.A = (const T*)A[i]; l.lda = lda[i];
                        l.W = blk[q].W; l.P = blk[q].P;
                        mak::ls_carve_block<T>(l, bp, lnb);
                        bp = (T*)mak::align_up((size_t)(uintptr_t)bp, 256);
                        m2[q] = 2 * n[i]; nn[q] = n[i];
                        if (m[i] > n[i]) { mt.push_back(m[i]); nt.push_back(n[i]); }
                    }
                    size_t qneed = qr_batched_worksize_t<T>(h, (int)nc, m2.data(), nn.data());
                    if (!mt.empty()) qneed = std::max(qneed, qr_batched_worksize_t<T>(h, (int)mt.size(), mt.data(), nt.data()));
                    if ((char*)bp <= ls_qr && qneed <= ls_qb) {
                        rc = mak::polar_lockstep_run<T>(h, lb, ls_tables, ls_tb,
                            [&](int nbat, const int* qm, const int* qn, void* const* qA, const int* qlda, void* const* qQ,
                                const int* qldq, void* const* qR, const int* qldr) {
                                return qr_batched_t<T>(h, nbat, qm, qn, qA, qlda, qQ, qldq, qR, qldr, nullptr, ls_qr, ls_qb);
                            });
                        if (rc) return rc;
                        ls_done = true;
                    }
                }
                if (getenv("MAKB200_LOCKSTEP_VERBOSE"))
                    fprintf(stderr, "[makb200] batched svd chunk of %zu blocks (n <= %d): QDWH %s\n", nc, n[ids[0]],
                            ls_done ? "in lock-step" : "per block on the stream pool");
                if (!ls_done) {
                    rc = run_pooled(h, ids, wbig, pool_region, [&](makb200_handle_t* hh, char* w, size_t lw, int i) {
                        const Blk& b = blk[slot_of[i]];
                        return mak::svd_phase1_t<T>(hh, m[i], n[i], (T*)A[i], lda[i], b.W, b.P, 2.2e-16, w, lw);
                    });
                    if (rc) return rc;
                }
                {
                    mak::Stager st(h, nc * sizeof(mak::BhetrdDesc<T>) + 1024);
                    MAK_CUDA(h, st.put(ddev, bd.data(), nc * sizeof(mak::BhetrdDesc<T>), h->stream));
                }
                rc = mak::bhetrd_batched_t<T>(h, (int)nc, ddev, nmax);
                if (rc) return rc;
                bool tail_done = false;
                if (lockstep) {
                    // phases 2-3 in lock-step too: batched D&C, batched back-transformation, then reorder / U = W V /
                    // defect check / gauge for the whole chunk (csrc/eigh_lockstep.cuh); the region of phase 1 is reused
                    std::vector<int> nn(nc);
                    for (size_t q = 0; q < nc; ++q) nn[q] = n[ids[q]];
                    const size_t e_bytes = mak::eigh_lockstep_layout<T>((int)nc, nn.data()).total;
                    const size_t t_bytes = mak::svd_tail_lockstep_bytes<T>((int)nc);
                    if (e_bytes + t_bytes + 512 <= ls_total) {
                        std::vector<mak::EighLsBlk<T>> eb(nc);
                        std::vector<mak::SvdTailBlk<T>> tb(nc);
                        for (size_t q = 0; q < nc; ++q) {
                            const int i = ids[q];
                            const Blk& b = blk[q];
                            eb[q] = mak::EighLsBlk<T>{n[i], b.P, n[i], b.d, b.e, b.tau, b.wv, b.V, n[i]};
                            tb[q] = mak::SvdTailBlk<T>{m[i], n[i], b.W, b.wv, b.V, n[i], (double*)S[i], (T*)U[i], ldu[i], (T*)Vh[i], ldvh[i], nullptr};
                        }
                        char* lsb = (char*)mak::align_up((size_t)(uintptr_t)ls_base, 256);
                        rc = mak::eigh_lockstep_run<T>(h, (int)nc, eb.data(), 0, lsb, e_bytes);
                        if (rc) return rc;
                        rc = mak::svd_tail_lockstep_run<T>(h, (int)nc, tb.data(), fixgauge, lsb + mak::align_up(e_bytes, 256), t_bytes,
                            [&](int q) {
                                // rank-deficient block: U <- Q of its positive-diagonal Householder QR (polar.cu: svd_tail)
                                const int i = ids[q];
                                const Blk& b = blk[q];
                                int r2 = mak::qr_fused_t<T>(h, MAKB200_QR_COMPACT, m[i], n[i], (T*)U[i], ldu[i], b.W, m[i], (T*)nullptr, 0,
                                                            wbig, pool_region);
                                if (r2) return r2;
                                MAK_CUDA(h, cudaMemcpy2DAsync(U[i], (size_t)ldu[i] * sizeof(T), b.W, (size_t)m[i] * sizeof(T),
                                                              (size_t)m[i] * sizeof(T), n[i], cudaMemcpyDeviceToDevice, h->stream));
                                return 0;
                            });
                        if (rc) return rc;
                        tail_done = true;
                    }
                }
                if (!tail_done) {
                    rc = run_pooled(h, ids, wbig, pool_region, [&](makb200_handle_t* hh, char* w, size_t lw, int i) {
                        const int q = slot_of[i];
                        const Blk& b = blk[q];
                        return mak::svd_phase2_t<T>(hh, m[i], n[i], b.W, b.P, b.V, b.wv, b.flag, &pre[q], (double*)S[i], (T*)U[i], ldu[i],
                                                    (T*)Vh[i], ldvh[i], fixgauge, w, lw);
                    });
                    if (rc) return rc;
                }
            }
            if (rest.empty()) return 0;
            return run_pooled(h, rest, wbig, lbig, per_block);
        }
    }
    if (graphs_enabled() && big.size() >= 4) {
        // graph replay: one captured svd_t per (shape, slot) against staging buffers
        size_t st_a = 0, st_u = 0, st_v = 0, st_s = 0, wmax = 0;
        std::map<std::pair<int, int>, int> gid;
        std::vector<ShapeGroup> groups;
        std::vector<int> ordinal(batch, -1);
        for (size_t q = 0; q < big.size(); ++q) {
            const int i = big[q], k = m[i] < n[i] ? m[i] : n[i];
            ordinal[i] = (int)q;
            st_a = std::max(st_a, (size_t)m[i] * n[i]);
            st_u = std::max(st_u, (size_t)m[i] * k);
            st_v = std::max(st_v, (size_t)k * n[i]);
            st_s = std::max(st_s, (size_t)k);
            wmax = std::max(wmax, mak::svd_worksize_t<T>(h, m[i], n[i]));
            auto key = std::make_pair(m[i], n[i]);
            auto it = gid.find(key);
            if (it == gid.end()) {
                gid[key] = (int)groups.size();
                groups.push_back(ShapeGroup{m[i], n[i], {}, 0.0});
                it = gid.find(key);
            }
            groups[it->second].idx.push_back(i);
            groups[it->second].cost += (double)m[i] * n[i] * k + 2.0e5 * k;   // flops-ish + launch-chain term
        }
        const size_t oA = 0, oU = mak::align_up(st_a * sizeof(T), 256), oV = oU + mak::align_up(st_u * sizeof(T), 256),
                     oS = oV + mak::align_up(st_v * sizeof(T), 256), oF = oS + mak::align_up(st_s * sizeof(double), 256),
                     oW = oF + 256;
        const size_t slot_bytes = oW + mak::align_up(wmax, 256) + 256;
        // per-block rank-defect indicators, collected with ONE read after the batch
        double* defects = nullptr;
        size_t tail = mak::align_up(sizeof(double) * big.size(), 256);
        if (lbig > tail) { defects = (double*)(wbig + lbig - tail); lbig -= tail; }
        if (defects) {
            MAK_CUDA(h, cudaMemsetAsync(defects, 0, sizeof(double) * big.size(), h->stream));
            const size_t esz = sizeof(T);
            int rc = run_graphed(h, /*kind*/ 0, std::is_same<T, double>::value ? 0 : 1, (fixgauge ? 1 : 0) | (vectors ? 2 : 0), groups,
                wbig, lbig, slot_bytes,
                [&](makb200_handle_t* hh, int mm, int nn, char* st, size_t) {
                    const int k = mm < nn ? mm : nn;
                    hh->defect_dev = (double*)(st + oF);
                    const int r = mak::svd_t<T>(hh, mm, nn, (T*)(st + oA), mm, (double*)(st + oS), vectors ? (T*)(st + oU) : nullptr, mm,
                                                vectors ? (T*)(st + oV) : nullptr, k, fixgauge, 2.2e-16, st + oW, slot_bytes - oW, nullptr);
                    hh->defect_dev = nullptr;
                    return r;
                },
                [&](makb200_handle_t* hh, int i, int mm, int nn, char* st) {
                    MAK_CUDA(hh, cudaMemsetAsync(st + oF, 0, sizeof(double), hh->stream));
                    MAK_CUDA(hh, cudaMemcpy2DAsync(st + oA, (size_t)mm * esz, A[i], (size_t)lda[i] * esz, (size_t)mm * esz, nn,
                                                   cudaMemcpyDeviceToDevice, hh->stream));
                    return 0;
                },
                [&](makb200_handle_t* hh, int i, int mm, int nn, char* st) {
                    const int k = mm < nn ? mm : nn;
                    MAK_CUDA(hh, cudaMemcpyAsync(S[i], st + oS, sizeof(double) * k, cudaMemcpyDeviceToDevice, hh->stream));
                    if (vectors) {
                        MAK_CUDA(hh, cudaMemcpy2DAsync(U[i], (size_t)ldu[i] * esz, st + oU, (size_t)mm * esz, (size_t)mm * esz, k,
                                                       cudaMemcpyDeviceToDevice, hh->stream));
                        MAK_CUDA(hh, cudaMemcpy2DAsync(Vh[i], (size_t)ldvh[i] * esz, st + oV, (size_t)k * esz, (size_t)k * esz, nn,
                                                       cudaMemcpyDeviceToDevice, hh->stream));
                        MAK_CUDA(hh, cudaMemcpyAsync(defects + ordinal[i], st + oF, sizeof(double), cudaMemcpyDeviceToDevice, hh->stream));
                    }
                    return 0;
                });
            if (rc == 0) {
                if (!vectors) return 0;
                // deficient blocks (U not isometric): redo through the uncaptured path, which repairs U
                std::vector<double> hd(big.size());
                MAK_CUDA(h, cudaMemcpyAsync(hd.data(), defects, sizeof(double) * big.size(), cudaMemcpyDeviceToHost, h->stream));
                MAK_CUDA(h, cudaStreamSynchronize(h->stream));
                std::vector<int> redo;
                for (size_t q = 0; q < big.size(); ++q)
                    if (hd[q] > 1e-6) redo.push_back(big[q]);
                return run_pooled(h, redo, wbig, lbig, per_block);
            }
            if (rc != -100000) return rc;
            // capture not possible here: fall through to the launch-per-kernel pool
        }
    }
    return run_pooled(h, big, wbig, lbig, per_block);
}

extern "C" {

size_t makb200_svd_batched_worksize(makb200_handle_t* h, int dtype, int batch, const int* m, const int* n) {
    if (!h || !dtype_ok(dtype) || batch < 0 || (batch > 0 && (!m || !n))) return 0;
    size_t esz = dtype == MAKB200_F64 ? sizeof(double) : sizeof(cplx);
    size_t bytes = mak::align_up(sizeof(mak::SvdBlockDesc<cplx>) * (size_t)(batch > 0 ? batch : 1), 256), big = 0, nbig = 0;
    size_t st_a = 0, st_u = 0, st_v = 0, st_s = 0;
    for (int i = 0; i < batch; ++i) {
        if (m[i] <= 0 || n[i] <= 0) continue;
        if (mak::batched_svd_smem_bytes(m[i], n[i], esz) > mak::batched_svd_max_smem_bytes()) {
            size_t w = dtype == MAKB200_F64 ? mak::svd_worksize_t<double>(h, m[i], n[i])
                                            : mak::svd_worksize_t<cplx>(h, m[i], n[i]);
            if (w > big) big = w;
            ++nbig;
            const size_t k = (size_t)(m[i] < n[i] ? m[i] : n[i]);
            st_a = std::max(st_a, (size_t)m[i] * n[i]);
            st_u = std::max(st_u, (size_t)m[i] * k);
            st_v = std::max(st_v, k * (size_t)n[i]);
            st_s = std::max(st_s, k);
        }
    }
    // graph-replayed path: staging copies of A, U, Vh, S per slot + one defect indicator per big block
    const size_t staging = mak::align_up(st_a * esz, 256) + mak::align_up(st_u * esz, 256) + mak::align_up(st_v * esz, 256) +
                           mak::align_up(st_s * 8, 256) + 1024;
    // phased path: per-block W, P, V, ... of one chunk + its descriptors (fixed chunks: the fallback), or the chunk plan
    // of the lock-step path with its region (svd_chunk_plan: the run walks the same chunks)
    size_t persist = 0;
    for (int i = 0; i < batch; ++i) {
        if (m[i] <= 0 || n[i] <= 0 || m[i] < n[i] || n[i] > mak::BHETRD_MAX_N) continue;
        if (mak::batched_svd_smem_bytes(m[i], n[i], esz) <= mak::batched_svd_max_smem_bytes()) continue;
        const size_t pb = dtype == MAKB200_F64 ? svd_phase_persist_bytes<double>(m[i], n[i]) : svd_phase_persist_bytes<cplx>(m[i], n[i]);
        if (pb > persist) persist = pb;
    }
    const size_t chunk = nbig < (size_t)SVD_PHASE_CHUNK ? nbig : (size_t)SVD_PHASE_CHUNK;
    size_t phased = chunk * (persist + sizeof(mak::BhetrdDesc<cplx>) + 64) + 8192;
    {
        size_t region = 0;
        if (dtype == MAKB200_F64) {
            const std::vector<int> ph = svd_phased_blocks<double>(batch, m, n);
            if (ph.size() >= 8) svd_chunk_plan<double>(h, ph, m, n, &region);
        } else {
            const std::vector<int> ph = svd_phased_blocks<cplx>(batch, m, n);
            if (ph.size() >= 8) svd_chunk_plan<cplx>(h, ph, m, n, &region);
        }
        phased = std::max(phased, region + 8192);
    }
    return bytes + pooled_worksize(big, nbig, staging) + phased + mak::align_up(sizeof(double) * nbig, 256) + 512;
}

int makb200_svd_batched(makb200_handle_t* h, int dtype, int fixgauge, int batch, const int* m, const int* n,
                        void* const* A, const int* lda, void* const* S, void* const* U, const int* ldu,
                        void* const* Vh, const int* ldvh, int* info, void* work, size_t lwork) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (batch < 0) return -4;
    if (batch == 0) return 0;
    if (!m) return -5;
    if (!n) return -6;
    if (!A) return -7;
    if (!lda) return -8;
    if (!S) return -9;
    if ((U == nullptr) != (Vh == nullptr)) return -10;
    if (dtype == MAKB200_F64)
        return svd_batched_t<double>(h, fixgauge, batch, m, n, A, lda, S, U, ldu, Vh, ldvh, info, work, lwork);
    return svd_batched_t<cplx>(h, fixgauge, batch, m, n, A, lda, S, U, ldu, Vh, ldvh, info, work, lwork);
}


}  // extern "C"

// ---- batched eigh -------------------------------------------------------------------------
// One-launch tridiagonalisation of every mid-size block of a batch (one CTA per block, csrc/bhetrd.cuh).  Measured at
// scale in round 2 (profiles/r2_bhetrd_scale.log, 600-2000 c128 blocks per bucket): eigh 65-128: 3327 -> 6377 blocks/s,
// 129-256: 1412 -> 2826, 257-512: 708 -> 1365 - default on (MAKB200_BHETRD=0 disables).
static bool bhetrd_enabled() {
    const char* e = getenv("MAKB200_BHETRD");   // read per call: the tests toggle it
    return !(e && e[0] == '0');
}
constexpr int EIGH_LS_CHUNK = 192;       // least blocks per lock-step eigensolve (when that many are left)
constexpr size_t EIGH_LS_BUDGET = (size_t)4 << 30, EIGH_LS_MAX_CHUNK = 4096;
// chunks of the lock-step eigensolve over block orders sorted descending: as many blocks as fit EIGH_LS_BUDGET bytes of
// workspace (bounded below / above); returns (first, count) pairs and the largest workspace any chunk needs
template <typename T>
static std::vector<std::pair<size_t, size_t>> eigh_ls_chunks(const std::vector<int>& ns, size_t* region) {
    std::vector<std::pair<size_t, size_t>> out;
    size_t reg = 0;
    for (size_t c0 = 0; c0 < ns.size();) {
        const int nmax = ns[c0];
        const size_t per = mak::eigh_lockstep_bytes<T>(64, nmax) / 64 + 1;
        size_t nc = std::min<size_t>(std::max<size_t>(EIGH_LS_BUDGET / per, (size_t)EIGH_LS_CHUNK), EIGH_LS_MAX_CHUNK);
        nc = std::min(nc, ns.size() - c0);
        reg = std::max(reg, mak::eigh_lockstep_bytes<T>((int)nc, nmax));
        out.emplace_back(c0, nc);
        c0 += nc;
    }
    if (region) *region = reg;
    return out;
}
template <typename T>
static int eigh_batched_t(makb200_handle_t* h, int fixgauge, int batch, const int* n, void* const* A, const int* lda,
                          void* const* W, void* const* V, const int* ldv, int* info, void* work, size_t lwork) {
    std::vector<mak::EighBlockDesc<T>> small, small2;
    std::vector<int> big;
    size_t max_b = 0, max_b2 = 0;
    for (int i = 0; i < batch; ++i) {
        if (n[i] < 0) return -5;
        if (n[i] == 0) continue;
        size_t sb = mak::batched_eigh_smem_bytes(n[i], sizeof(T));
        if (sb > mak::batched_eigh_max_smem_bytes()) { big.push_back(i); continue; }
        mak::EighBlockDesc<T> d;
        d.n = n[i]; d.fixgauge = fixgauge;
        d.A = (const T*)A[i]; d.lda = lda[i];
        d.W = (double*)W[i];
        d.V = V ? (T*)V[i] : nullptr; d.ldv = ldv ? ldv[i] : 0;
        if (sb <= 40 * 1024) { small.push_back(d); if (sb > max_b) max_b = sb; }
        else { small2.push_back(d); if (sb > max_b2) max_b2 = sb; }
    }
    mak::Arena ar(work, lwork);
    mak::EighBlockDesc<T>* ddev = ar.get<mak::EighBlockDesc<T>>(batch > 0 ? batch : 1);
    if (!ar.ok) return MAKB200_ERR_WORKSPACE;
    if (info) MAK_CUDA(h, cudaMemsetAsync(info, 0, sizeof(int) * batch, h->stream));
    if (!small.empty() || !small2.empty()) {
        {
            mak::Stager st(h, (small.size() + small2.size()) * sizeof(mak::EighBlockDesc<T>) + 1024);
            MAK_CUDA(h, st.put(ddev, small.data(), small.size() * sizeof(mak::EighBlockDesc<T>), h->stream));
            MAK_CUDA(h, st.put(ddev + small.size(), small2.data(), small2.size() * sizeof(mak::EighBlockDesc<T>), h->stream));
        }
        int rc = mak::batched_eigh_smem<T>(h, (int)small.size(), max_b, ddev, nullptr);
        if (rc) return rc;
        rc = mak::batched_eigh_smem<T>(h, (int)small2.size(), max_b2, ddev + small.size(), nullptr);
        if (rc) return rc;
    }
    // EXPERIMENTAL (MAKB200_BHETRD=1, round-2 bring-up): tridiagonalise every big block with n <= BHETRD_MAX_N in
    // ONE launch (one CTA per block, csrc/bhetrd.cuh) instead of two launches per column per block; the pooled
    // per-block path then starts at the tridiagonal solver.
    std::vector<mak::TrdPre<T>> pre(batch, mak::TrdPre<T>{nullptr, nullptr, nullptr});
    if (bhetrd_enabled() && !big.empty()) {
        std::vector<mak::BhetrdDesc<T>> bd;
        int nmax = 0;
        for (int i : big) {
            if (n[i] > mak::BHETRD_MAX_N) continue;
            const size_t nn = (size_t)n[i];
            pre[i].d = ar.get<double>(nn);
            pre[i].e = ar.get<double>(nn);
            pre[i].tau = ar.get<T>(nn);
            bd.push_back(mak::BhetrdDesc<T>{n[i], (T*)A[i], lda[i], pre[i].d, pre[i].e, pre[i].tau});
            if (n[i] > nmax) nmax = n[i];
        }
        mak::BhetrdDesc<T>* bdev = ar.get<mak::BhetrdDesc<T>>(bd.size() > 0 ? bd.size() : 1);
        if (!ar.ok) return MAKB200_ERR_WORKSPACE;
        if (!bd.empty()) {
            {
                mak::Stager st(h, bd.size() * sizeof(mak::BhetrdDesc<T>) + 1024);
                MAK_CUDA(h, st.put(bdev, bd.data(), bd.size() * sizeof(mak::BhetrdDesc<T>), h->stream));
            }
            int rc = mak::bhetrd_batched_t<T>(h, (int)bd.size(), bdev, nmax);
            if (rc) return rc;
        }
    }
    char* wbig = (char*)work + ar.off;
    size_t lbig = lwork > ar.off ? lwork - ar.off : 0;
    // lock-step eigensolve (default): every tridiagonalised block goes through ONE batched D&C and ONE batched
    // back-transformation per chunk instead of ~100 launches per block on the pool (csrc/eigh_lockstep.cuh)
    if (V && lockstep_enabled() && bhetrd_enabled() && !graphs_enabled()) {
        std::vector<int> ls;
        for (int i : big) if (pre[i].d && n[i] >= 3) ls.push_back(i);
        if (ls.size() >= 8) {
            std::stable_sort(ls.begin(), ls.end(), [&](int a, int b) { return n[a] > n[b]; });
            std::vector<int> lsn(ls.size());
            for (size_t q = 0; q < ls.size(); ++q) lsn[q] = n[ls[q]];
            size_t need = 0;
            const std::vector<std::pair<size_t, size_t>> chunks = eigh_ls_chunks<T>(lsn, &need);
            if (need + 512 <= lbig) {
                char* lsb = (char*)mak::align_up((size_t)(uintptr_t)wbig, 256);
                for (const auto& ck : chunks) {
                    const size_t c0 = ck.first, nc = ck.second;
                    std::vector<mak::EighLsBlk<T>> eb(nc);
                    for (size_t q = 0; q < nc; ++q) {
                        const int i = ls[c0 + q];
                        eb[q] = mak::EighLsBlk<T>{n[i], (T*)A[i], lda[i], pre[i].d, pre[i].e, pre[i].tau, (double*)W[i], (T*)V[i], ldv[i]};
                    }
                    int rc = mak::eigh_lockstep_run<T>(h, (int)nc, eb.data(), fixgauge, lsb, need);
                    if (rc) return rc;
                }
                if (getenv("MAKB200_LOCKSTEP_VERBOSE"))
                    fprintf(stderr, "[makb200] batched eigh: %zu blocks (n <= %d) solved in lock-step\n", ls.size(), n[ls[0]]);
                std::vector<int> rest;
                for (int i : big) if (!(pre[i].d && n[i] >= 3)) rest.push_back(i);
                big.swap(rest);
                if (big.empty()) return 0;
            }
        }
    }
    auto per_block = [&](makb200_handle_t* hh, char* w, size_t lw, int i) {
        return mak::eigh_t<T>(hh, n[i], (T*)A[i], lda[i], (double*)W[i], V ? (T*)V[i] : (T*)nullptr, V ? ldv[i] : n[i],
                              fixgauge, w, lw, nullptr, 0,
                              pre[i].d ? &pre[i] : nullptr);   // V == NULL: values only (Sturm K-section)
    };
    if (graphs_enabled() && !bhetrd_enabled() && big.size() >= 4) {
        // graph replay: one captured eigh_t per (n, slot) against staging buffers (see run_graphed)
        size_t nmax = 0, wmax = 0;
        std::map<int, int> gid;
        std::vector<ShapeGroup> groups;
        for (int i : big) {
            nmax = std::max(nmax, (size_t)n[i]);
            wmax = std::max(wmax, mak::eigh_worksize_t<T>(h, n[i]));
            auto it = gid.find(n[i]);
            if (it == gid.end()) {
                gid[n[i]] = (int)groups.size();
                groups.push_back(ShapeGroup{n[i], n[i], {}, 0.0});
                it = gid.find(n[i]);
            }
            groups[it->second].idx.push_back(i);
            groups[it->second].cost += (double)n[i] * n[i] * n[i] + 2.0e5 * n[i];
        }
        const size_t esz = sizeof(T);
        const size_t oA = 0, oV = mak::align_up(nmax * nmax * esz, 256), oD = oV + mak::align_up(nmax * nmax * esz, 256),
                     oW = oD + mak::align_up(nmax * 8, 256);
        const size_t slot_bytes = oW + mak::align_up(wmax, 256) + 256;
        const bool vectors = V != nullptr;
        int rc = run_graphed(h, /*kind*/ 1, std::is_same<T, double>::value ? 0 : 1, (fixgauge ? 1 : 0) | (vectors ? 2 : 0), groups, wbig,
            lbig, slot_bytes,
            [&](makb200_handle_t* hh, int nn, int, char* st, size_t) {
                return mak::eigh_t<T>(hh, nn, (T*)(st + oA), nn, (double*)(st + oD), vectors ? (T*)(st + oV) : (T*)nullptr, nn, fixgauge,
                                      st + oW, slot_bytes - oW, nullptr, 0, nullptr);
            },
            [&](makb200_handle_t* hh, int i, int nn, int, char* st) {
                MAK_CUDA(hh, cudaMemcpy2DAsync(st + oA, (size_t)nn * esz, A[i], (size_t)lda[i] * esz, (size_t)nn * esz, nn,
                                               cudaMemcpyDeviceToDevice, hh->stream));
                return 0;
            },
            [&](makb200_handle_t* hh, int i, int nn, int, char* st) {
                MAK_CUDA(hh, cudaMemcpyAsync(W[i], st + oD, sizeof(double) * nn, cudaMemcpyDeviceToDevice, hh->stream));
                if (vectors)
                    MAK_CUDA(hh, cudaMemcpy2DAsync(V[i], (size_t)ldv[i] * esz, st + oV, (size_t)nn * esz, (size_t)nn * esz, nn,
                                                   cudaMemcpyDeviceToDevice, hh->stream));
                return 0;
            });
        if (rc != -100000) return rc;
    }
    return run_pooled(h, big, wbig, lbig, per_block);
}

extern "C" {

size_t makb200_eigh_batched_worksize(makb200_handle_t* h, int dtype, int batch, const int* n) {
    if (!h || !dtype_ok(dtype) || batch < 0 || (batch > 0 && !n)) return 0;
    size_t esz = dtype == MAKB200_F64 ? sizeof(double) : sizeof(cplx);
    size_t bytes = mak::align_up(sizeof(mak::EighBlockDesc<cplx>) * (size_t)(batch > 0 ? batch : 1), 256), big = 0, nbig = 0;
    size_t nmax_big = 0, nmax_ls = 0;
    bool any = false;
    for (int i = 0; i < batch; ++i) {
        if (n[i] <= 0) continue;
        if (mak::batched_eigh_smem_bytes(n[i], esz) > mak::batched_eigh_max_smem_bytes()) {
            size_t w = dtype == MAKB200_F64 ? mak::eigh_worksize_t<double>(h, n[i]) : mak::eigh_worksize_t<cplx>(h, n[i]);
            if (w > big) big = w;
            any = true;
            ++nbig;
            if ((size_t)n[i] > nmax_big) nmax_big = (size_t)n[i];
            if (n[i] <= mak::BHETRD_MAX_N && (size_t)n[i] > nmax_ls) nmax_ls = (size_t)n[i];
            // d, e, tau of the one-launch tridiagonalisation (MAKB200_BHETRD) + its descriptor
            bytes += 2 * mak::align_up(sizeof(double) * (size_t)n[i], 256) + mak::align_up(sizeof(cplx) * (size_t)n[i], 256) +
                     sizeof(mak::BhetrdDesc<cplx>);
        }
    }
    // graph-replayed path: staging copies of A, V, W per slot
    const size_t staging = 2 * mak::align_up(nmax_big * nmax_big * esz, 256) + mak::align_up(nmax_big * 8, 256) + 1024;
    size_t pooled = any ? pooled_worksize(big, nbig, staging) : 0;
    // lock-step eigensolve of the tridiagonalised blocks (shares the region of the pooled slices)
    if (nbig >= 8 && nmax_ls > 0) {
        std::vector<int> lsn;
        for (int i = 0; i < batch; ++i)
            if (n[i] >= 3 && n[i] <= mak::BHETRD_MAX_N && mak::batched_eigh_smem_bytes(n[i], esz) > mak::batched_eigh_max_smem_bytes())
                lsn.push_back(n[i]);
        std::sort(lsn.begin(), lsn.end(), [](int a, int b) { return a > b; });
        size_t ls = 0;
        if (dtype == MAKB200_F64) eigh_ls_chunks<double>(lsn, &ls);
        else eigh_ls_chunks<cplx>(lsn, &ls);
        if (ls + 1024 > pooled) pooled = ls + 1024;
    }
    return bytes + pooled + 512;
}

int makb200_eigh_batched(makb200_handle_t* h, int dtype, int fixgauge, int batch, const int* n, void* const* A,
                         const int* lda, void* const* W, void* const* V, const int* ldv, int* info, void* work,
                         size_t lwork) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (batch < 0) return -4;
    if (batch == 0) return 0;
    if (!n) return -5;
    if (!A) return -6;
    if (!lda) return -7;
    if (!W) return -8;
    if (V && !ldv) return -10;
    if (dtype == MAKB200_F64) return eigh_batched_t<double>(h, fixgauge, batch, n, A, lda, W, V, ldv, info, work, lwork);
    return eigh_batched_t<cplx>(h, fixgauge, batch, n, A, lda, W, V, ldv, info, work, lwork);
}

}  // extern "C"

extern "C" {

// ---- adjoint (lq_via_qr!, svd_via_adjoint!) -----------------------------------------------
int makb200_adjoint(makb200_handle_t* h, int dtype, int m, int n, const void* A, int lda, void* B, int ldb) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (m < 0) return -3;
    if (n < 0) return -4;
    if (lda < maxi(1, m)) return -6;
    if (ldb < maxi(1, n)) return -8;
    if (m == 0 || n == 0) return 0;
    if (!A) return -5;
    if (!B || B == A) return -7;
    if (dtype == MAKB200_F64) return mak::adjoint_t<double>(h, m, n, (const double*)A, lda, (double*)B, ldb);
    return mak::adjoint_t<cplx>(h, m, n, (const cplx*)A, lda, (cplx*)B, ldb);
}

}  // extern "C"

extern "C" {

// ---- experimental: second stage of the two-stage tridiagonalisation (band -> tridiagonal) ----------
size_t makb200_sbr_chase_worksize(makb200_handle_t* h, int dtype, int n, int b) {
    if (!h || !dtype_ok(dtype) || n < 0 || b < 1) return 0;
    return dtype == MAKB200_F64 ? mak::sbr_chase_worksize_t<double>(n, b) : mak::sbr_chase_worksize_t<cplx>(n, b);
}

int makb200_sbr_chase(makb200_handle_t* h, int dtype, int n, int b, const void* A, int lda, double* d, double* e,
                      void* V2, int ldv, void* tau2, int ldt, void* work, size_t lwork) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (n < 0) return -3;
    if (b < 1 || b > 64) return -4;
    if (lda < maxi(1, n)) return -6;
    if (ldv < maxi(1, n)) return -10;
    if (ldt < (n + b - 1) / b + 1) return -12;
    if (n == 0) return 0;
    if (!A) return -5;
    if (!d) return -7;
    if (n > 1 && !e) return -8;
    if (!V2) return -9;
    if (!tau2) return -11;
    if (dtype == MAKB200_F64)
        return mak::sbr_chase_t<double>(h, n, b, (const double*)A, lda, d, e, (double*)V2, ldv, (double*)tau2, ldt, work, lwork);
    return mak::sbr_chase_t<cplx>(h, n, b, (const cplx*)A, lda, d, e, (cplx*)V2, ldv, (cplx*)tau2, ldt, work, lwork);
}


// ---- experimental: first stage (dense -> band) and the Q2 application, exposed for bring-up ----------
size_t makb200_sy2sb_worksize(makb200_handle_t* h, int dtype, int n, int b) {
    if (!h || !dtype_ok(dtype) || n < 0 || b < 1) return 0;
    return dtype == MAKB200_F64 ? mak::sy2sb_worksize_t<double>(h, n, b) : mak::sy2sb_worksize_t<cplx>(h, n, b);
}
int makb200_sy2sb(makb200_handle_t* h, int dtype, int n, int b, void* A, int lda, void* tau1, void* work, size_t lwork) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (n < 0) return -3;
    if (b < 1 || b > 128) return -4;
    if (lda < maxi(1, n)) return -6;
    if (n == 0) return 0;
    if (!A) return -5;
    if (!tau1) return -7;
    if (dtype == MAKB200_F64) return mak::sy2sb_t<double>(h, n, b, (double*)A, lda, (double*)tau1, work, lwork);
    return mak::sy2sb_t<cplx>(h, n, b, (cplx*)A, lda, (cplx*)tau1, work, lwork);
}
size_t makb200_sbr_apply_q2_worksize(makb200_handle_t* h, int dtype, int n, int b, int g, int ncols) {
    if (!h || !dtype_ok(dtype) || n < 0 || b < 1 || g < 1) return 0;
    return dtype == MAKB200_F64 ? mak::sbr_apply_q2_worksize_t<double>(n, b, g, ncols)
                                : mak::sbr_apply_q2_worksize_t<cplx>(n, b, g, ncols);
}
int makb200_sbr_apply_q2(makb200_handle_t* h, int dtype, int n, int b, int g, const void* V2, int ldv, const void* tau2,
                         int ldt, void* Z, int ldz, int ncols, void* work, size_t lwork) {
    if (!h) return -1;
    if (!dtype_ok(dtype)) return -2;
    if (n < 0) return -3;
    if (b < 1 || b > 64) return -4;
    if (g < 1 || g > 128) return -5;
    if (ldv < maxi(1, n)) return -7;
    if (ldt < (n + b - 1) / b + 1) return -9;
    if (ldz < maxi(1, n)) return -11;
    if (ncols < 0) return -12;
    if (n == 0 || ncols == 0) return 0;
    if (!V2) return -6;
    if (!tau2) return -8;
    if (!Z) return -10;
    if (dtype == MAKB200_F64)
        return mak::sbr_apply_q2_t<double>(h, n, b, g, (const double*)V2, ldv, (const double*)tau2, ldt, (double*)Z, ldz,
                                           ncols, work, lwork);
    return mak::sbr_apply_q2_t<cplx>(h, n, b, g, (const cplx*)V2, ldv, (const cplx*)tau2, ldt, (cplx*)Z, ldz, ncols, work,
                                     lwork);
}

}  // extern "C"

// ---- NCCL loader ---------------------------------------------------------------------------
namespace mak {
const NcclApi* nccl_api(const char** why) {
    static NcclApi api;
    static int state = 0;   // 0 untried, 1 ok, -1 failed
    static char msg[256];
    if (state == 0) {
        void* lib = nullptr;
        const char* from = nullptr;
        const char* env = getenv("MAKB200_NCCL_LIB");
        if (env && env[0]) { lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL); from = env; }
        if (!lib) { lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); from = "libnccl.so.2 (already mapped by the process)"; }
        if (!lib) { lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL); from = "libnccl.so.2 (library search path)"; }
        bool ok = lib != nullptr;
        if (!ok) { const char* e = dlerror(); snprintf(msg, sizeof(msg), "%s", e ? e : "dlopen failed"); }
#define MAK_SYM(field, name)                                                             \
        if (ok) {                                                                            \
            *(void**)(&api.field) = dlsym(lib, name);                                        \
            if (!api.field) { ok = false; snprintf(msg, sizeof(msg), "missing symbol %s", name); } \
        }
        MAK_SYM(GetUniqueId, "ncclGetUniqueId")
        MAK_SYM(CommInitRank, "ncclCommInitRank")
        MAK_SYM(CommDestroy, "ncclCommDestroy")
        MAK_SYM(CommCount, "ncclCommCount")
        MAK_SYM(CommUserRank, "ncclCommUserRank")
        MAK_SYM(Send, "ncclSend")
        MAK_SYM(Recv, "ncclRecv")
        MAK_SYM(Broadcast, "ncclBroadcast")
        MAK_SYM(AllGather, "ncclAllGather")
        MAK_SYM(GroupStart, "ncclGroupStart")
        MAK_SYM(GroupEnd, "ncclGroupEnd")
        MAK_SYM(GetErrorString, "ncclGetErrorString")
#undef MAK_SYM
        api.loaded_from = from;
        state = ok ? 1 : -1;
    }
    if (state < 0 && why) *why = msg;
    return state > 0 ? &api : nullptr;
}
}  // namespace mak
