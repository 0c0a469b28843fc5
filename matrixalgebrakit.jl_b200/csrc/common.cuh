// Shared device/host helpers for the makb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <stdlib.h>
#include "../../include/makb200.h"
#include "scalar.h"
#include "devutil.cuh"

namespace mak {

// ---------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------
}  // namespace mak

constexpr int MAK_NPOOL = 64;
struct makb200_handle {
    int device;
    cudaStream_t stream;
    int num_sms;
    int max_cluster;  // largest usable cluster size for the panel kernel
    cudaStream_t aux_stream;   // internal second stream (look-ahead in the blocked QR)
    cudaEvent_t ev[8];         // fork/join and look-ahead events
    cudaStream_t pool[MAK_NPOOL];     // stream pool: mid-size blocks of a batch run concurrently (MAK_NPOOL)
    cudaEvent_t pool_ev[MAK_NPOOL];
    bool no_lookahead;         // set while a pooled call is in flight (aux stream/events are shared)
    void* stage;               // pinned host staging for descriptor uploads (batched entry points)
    size_t stage_bytes;
    cudaEvent_t stage_ev;      // recorded after the last upload out of `stage`
    void* graph_cache;         // replayable CUDA graphs of the per-block paths (capi.cu: GraphCache), shared by handle copies
    double* defect_dev;        // non-null (set while a per-block path is captured into a graph): svd_tall copies its rank
                               // defect indicator here instead of reading it on the host
    char err[256];
};

namespace mak {

inline int cuda_fail(makb200_handle* h, cudaError_t e, const char* where) {
    if (h) snprintf(h->err, sizeof(h->err), "%s: %s", where, cudaGetErrorString(e));
    return MAKB200_ERR_CUDA;
}
#define MAK_CUDA(h, call)                                         \
    do {                                                          \
        cudaError_t _e = (call);                                  \
        if (_e != cudaSuccess) return mak::cuda_fail(h, _e, #call); \
    } while (0)
#define MAK_LAUNCH_CHECK(h, name)                                   \
    do {                                                            \
        cudaError_t _e = cudaGetLastError();                        \
        if (_e != cudaSuccess) return mak::cuda_fail(h, _e, name);  \
    } while (0)

// process-wide count of kernels launched by this library (bench.py reports it as gpu_launches)
extern unsigned long long g_launches;
extern double g_gemm_flops;
inline void count_launch(int n = 1) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }  // pooled host threads

// per-kernel-class device-time accumulation for the roofline line of bench.py
// (enabled through makb200_kernel_timing(1); CUDA events on the launching stream)
struct KernelClock {
    bool on = false;
    static constexpr int MAXEV = 1 << 16;
    cudaEvent_t* ev = nullptr;  // pairs
    int n = 0;
    void begin(cudaStream_t s) {
        if (!on || n + 2 > MAXEV) return;
        if (!ev) {
            ev = (cudaEvent_t*)malloc(sizeof(cudaEvent_t) * MAXEV);
            for (int i = 0; i < MAXEV; ++i) cudaEventCreate(&ev[i]);
        }
        cudaEventRecord(ev[n], s);
    }
    void end(cudaStream_t s) {
        if (!on || n + 2 > MAXEV || !ev) return;
        cudaEventRecord(ev[n + 1], s);
        n += 2;
    }
    // optional per-launch tags (GEMM: m, n, k, flags) for the shape log (env MAKB200_GEMM_LOG=<file>)
    struct Tag { int m, n, k, flags; double flops; };
    Tag* tags = nullptr;
    void tag(int m, int n, int k, int flags, double flops) {
        if (!on || n_tags_ok() == false) return;
        if (!tags) tags = (Tag*)malloc(sizeof(Tag) * (MAXEV / 2));
        tags[this->n / 2] = Tag{m, n, k, flags, flops};
    }
    bool n_tags_ok() const { return n + 2 <= MAXEV; }
    // total milliseconds and launch count since the last reset (synchronises)
    void collect(double* ms, int* launches) {
        double t = 0;
        const char* logp = tags ? getenv("MAKB200_GEMM_LOG") : nullptr;
        FILE* lf = (logp && logp[0]) ? fopen(logp, "w") : nullptr;
        for (int i = 0; i + 1 < n; i += 2) {
            cudaEventSynchronize(ev[i + 1]);
            float f = 0;
            cudaEventElapsedTime(&f, ev[i], ev[i + 1]);
            t += f;
            if (lf) fprintf(lf, "%d %d %d %d %.6e %.6f\n", tags[i / 2].m, tags[i / 2].n, tags[i / 2].k, tags[i / 2].flags, tags[i / 2].flops, (double)f);
        }
        if (lf) fclose(lf);
        *ms = t;
        *launches = n / 2;
        n = 0;
    }
};
extern KernelClock g_clock_dots;  // trd_dots_kernel (dominant kernel of eigh_full!)
extern KernelClock g_clock_gemm;  // DMMA GEMM launches
extern KernelClock g_clock_w;     // trd_w_kernel

// optional phase timing (env MAKB200_PROFILE=1): prints device time per phase to stderr.
struct PhaseTimer {
    bool on;
    cudaStream_t s;
    cudaEvent_t ev[48];
    const char* names[48];
    int n;
    explicit PhaseTimer(cudaStream_t st) : s(st), n(0) {
        const char* e = getenv("MAKB200_PROFILE");
        on = e && e[0] == '1';
    }
    void mark(const char* name) {
        if (!on || n >= 48) return;
        cudaEventCreate(&ev[n]);
        cudaEventRecord(ev[n], s);
        names[n++] = name;
    }
    void report(const char* what) {
        if (!on || n < 2) return;
        cudaEventSynchronize(ev[n - 1]);
        fprintf(stderr, "[makb200 profile] %s:", what);
        for (int i = 1; i < n; ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
            fprintf(stderr, " %s=%.3fms", names[i], ms);
        }
        fprintf(stderr, "\n");
        for (int i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
        n = 0;
    }
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Descriptor upload through the handle's pinned staging buffer: no pageable-memory staging inside
// the driver, no stream synchronisation; the only wait is on the previous call's last upload.
struct Stager {
    makb200_handle* h;
    size_t off;
    bool ok;
    Stager(makb200_handle* hh, size_t total) : h(hh), off(0), ok(true) {
        cudaEventSynchronize(h->stage_ev);
        if (total > h->stage_bytes) {
            if (h->stage) cudaFreeHost(h->stage);
            h->stage = nullptr;
            h->stage_bytes = 0;
            size_t want = align_up(total + total / 2 + 4096, 4096);
            if (cudaMallocHost(&h->stage, want) != cudaSuccess) { ok = false; cudaGetLastError(); return; }
            h->stage_bytes = want;
        }
    }
    cudaError_t put(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t s) {
        if (bytes == 0) return cudaSuccess;
        if (!ok || off + bytes > h->stage_bytes) return cudaErrorMemoryAllocation;
        char* p = (char*)h->stage + off;
        memcpy(p, src_host, bytes);
        off += align_up(bytes, 256);
        return cudaMemcpyAsync(dst_dev, p, bytes, cudaMemcpyHostToDevice, s);
    }
    ~Stager() { cudaEventRecord(h->stage_ev, h->stream); }
};

// bump allocator over the caller-provided workspace
struct Arena {
    char* base;
    size_t cap, off;
    bool ok;
    Arena(void* p, size_t n) : base((char*)p), cap(n), off(0), ok(true) {}
    template <typename T> T* get(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        if (off + bytes > cap || (base == nullptr && bytes > 0)) {
            ok = false;
            off += bytes;
            return nullptr;
        }
        T* r = (T*)(base + off);
        off += bytes;
        return r;
    }
};
// sizing twin of Arena (no memory)
struct ArenaSize {
    size_t off = 0;
    bool ok = true;
    template <typename T> T* get(size_t count) {
        off += align_up(count * sizeof(T), 256);
        return nullptr;
    }
};

}  // namespace mak
