// CUDA-on-CPU emulation for kernel LOGIC tests (test infrastructure, never shipped, never timed).
//
// The GPU box is a metered resource; this header lets the product's kernel headers (csrc/*_kernels.cuh,
// written against the small subset below) compile with plain g++ and run one thread block at a time:
// every CUDA thread is a ucontext fiber on ONE OS thread, a barrier (__syncthreads, __syncwarp, any
// *_sync warp collective) parks the fiber until all live participants arrived.  The scheduler visits
// fibers in forward, reverse or seeded-random order (emu::set_order), so a missing barrier shows up
// as an order-dependent result, and a barrier some threads never reach is reported as a deadlock
// instead of hanging.  Supported: threadIdx/blockIdx/blockDim/gridDim, static and dynamic shared
// memory (one block resident at a time), __syncthreads, __syncwarp, __shfl{,_xor,_down,_up}_sync,
// __ballot/__any/__all_sync, atomicAdd/Exch/Max on global or shared words, and the warp-level DMMA
// m8n8k4 fragment layout (emu::dmma).  Not supported: clusters, cp.async/TMA, partial-mask collectives
// inside divergent code (they would deadlock here, as they are undefined on the device).
#pragma once
#include <ucontext.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <functional>
#include <random>
#include <vector>

#define MAK_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(x) alignas(x)
#define __shared__ static

struct emu_dim3 {
    unsigned x, y, z;
    emu_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef emu_dim3 dim3;

namespace emu {

struct Bar {
    int expected = 0, arrived = 0;
    uint64_t gen = 0;
};
struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
    dim3 tid;
    int lane = 0, warp = 0;
};
struct Block {
    std::vector<Fiber> fibers;
    Bar block_bar;
    std::vector<Bar> warp_bars;
    std::vector<unsigned> warp_alive;        // bit per live lane
    std::vector<unsigned char> warp_slots;   // 32 x 16 bytes per warp: exchange buffer of the collectives
    ucontext_t sched;
    int cur = -1;
    std::function<void()> body;
    uint64_t progress = 0;
    emu_dim3 bidx;                           // co-resident launches: this block's blockIdx
    std::vector<unsigned char> dyn;          // co-resident launches: this block's dynamic shared memory
};

inline Block* g_blk = nullptr;
inline dim3 g_blockIdx, g_blockDim, g_gridDim;
inline std::vector<unsigned char> g_dyn_smem;
inline int g_order = 0;            // 0 forward, 1 reverse, 2 random
inline std::mt19937_64 g_rng(1);
constexpr size_t STACK_BYTES = 256 * 1024;
inline std::vector<char*> g_stacks;

inline void set_order(int order, uint64_t seed = 1) {
    g_order = order;
    g_rng.seed(seed);
}
inline Fiber& cur() { return g_blk->fibers[g_blk->cur]; }
inline void yield_() { swapcontext(&cur().ctx, &g_blk->sched); }
inline uint64_t g_progress = 0;    // co-resident launches: barrier arrivals/releases and finished fibers of ALL blocks
inline void bar_release(Bar& b) {
    b.arrived = 0;
    b.gen++;
    g_blk->progress++;
    g_progress++;
}
inline void bar_wait(Bar& b) {
    b.arrived++;
    g_progress++;
    if (b.arrived >= b.expected) {
        bar_release(b);
        return;
    }
    const uint64_t g = b.gen;
    while (b.gen == g) yield_();
}
inline void bar_drop(Bar& b) {
    b.expected--;
    if (b.expected > 0 && b.arrived >= b.expected) bar_release(b);
}
inline void trampoline() {
    Block& B = *g_blk;
    B.body();
    Fiber& f = cur();
    f.done = true;
    B.progress++;
    g_progress++;
    B.warp_alive[f.warp] &= ~(1u << f.lane);
    bar_drop(B.block_bar);
    bar_drop(B.warp_bars[f.warp]);
    swapcontext(&f.ctx, &B.sched);
}

inline unsigned char* dyn_smem() { return (g_blk && !g_blk->dyn.empty()) ? g_blk->dyn.data() : g_dyn_smem.data(); }

// run ONE block of `nthreads` threads
inline void run_block(dim3 bdim, const std::function<void()>& body) {
    Block B;
    g_blk = &B;
    const int nt = (int)(bdim.x * bdim.y * bdim.z), nw = (nt + 31) / 32;
    B.fibers.resize(nt);
    B.warp_bars.resize(nw);
    B.warp_alive.assign(nw, 0u);
    B.warp_slots.assign((size_t)nw * 32 * 16, 0);
    B.block_bar.expected = nt;
    B.body = body;
    while ((int)g_stacks.size() < nt) g_stacks.push_back((char*)malloc(STACK_BYTES));
    for (int t = 0; t < nt; ++t) {
        Fiber& f = B.fibers[t];
        f.tid = dim3(t % bdim.x, (t / bdim.x) % bdim.y, t / (bdim.x * bdim.y));
        f.lane = t & 31;
        f.warp = t >> 5;
        B.warp_bars[f.warp].expected++;
        B.warp_alive[f.warp] |= 1u << f.lane;
        f.stack = g_stacks[t];
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = STACK_BYTES;
        f.ctx.uc_link = &B.sched;
        makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    std::vector<int> order(nt);
    for (int t = 0; t < nt; ++t) order[t] = t;
    int remaining = nt;
    while (remaining > 0) {
        if (g_order == 1) {
            for (int t = 0; t < nt; ++t) order[t] = nt - 1 - t;
        } else if (g_order == 2) {
            std::shuffle(order.begin(), order.end(), g_rng);
        }
        const uint64_t before = B.progress;
        for (int i = 0; i < nt; ++i) {
            Fiber& f = B.fibers[order[i]];
            if (f.done) continue;
            B.cur = order[i];
            swapcontext(&B.sched, &f.ctx);
            if (f.done) --remaining;
        }
        if (remaining > 0 && B.progress == before) {
            fprintf(stderr, "cuda_emu: DEADLOCK in block (%u,%u,%u): %d threads wait on a barrier that the others never reach\n",
                    g_blockIdx.x, g_blockIdx.y, g_blockIdx.z, remaining);
            abort();
        }
    }
    g_blk = nullptr;
}

// launch<<<grid, block, smem>>>: blocks run one after the other
template <typename K, typename... Args>
inline void launch(K kernel, dim3 grid, dim3 block, size_t smem_bytes, Args... args) {
    g_gridDim = grid;
    g_blockDim = block;
    if (g_dyn_smem.size() < smem_bytes + 64) g_dyn_smem.resize(smem_bytes + 64);
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                g_blockIdx = dim3(x, y, z);
                run_block(block, [&]() { kernel(args...); });
            }
}

// launch_coresident<<<grid, block, smem>>>: ALL blocks are live at once (what a persistent kernel with
// inter-CTA flags needs: one block at a time would spin forever).  Every pass visits every live fiber
// of every block in forward / reverse / random order; a thread that polls global memory must call
// MAK_SPIN_PAUSE() (= emu::spin_pause, a yield) inside its loop.  Kernels launched this way may use
// dynamic shared memory only (`__shared__` statics are one per process here).  A full pass without a
// barrier arrival, a barrier release or a finished fiber is a deadlock (pollers change nothing).
inline void spin_pause() { yield_(); }
template <typename K, typename... Args>
inline void launch_coresident(K kernel, dim3 grid, dim3 block, size_t smem_bytes, Args... args) {
    g_gridDim = grid;
    g_blockDim = block;
    const int nb = (int)(grid.x * grid.y * grid.z);
    const int nt = (int)(block.x * block.y * block.z), nw = (nt + 31) / 32;
    constexpr size_t CO_STACK = 64 * 1024;
    std::vector<Block> blocks(nb);
    std::vector<char*> stacks((size_t)nb * nt);
    for (int bi = 0; bi < nb; ++bi) {
        Block& B = blocks[bi];
        B.bidx = dim3(bi % grid.x, (bi / grid.x) % grid.y, bi / (grid.x * grid.y));
        B.dyn.assign(smem_bytes + 64, 0);
        B.fibers.resize(nt);
        B.warp_bars.resize(nw);
        B.warp_alive.assign(nw, 0u);
        B.warp_slots.assign((size_t)nw * 32 * 16, 0);
        B.block_bar.expected = nt;
        B.body = [&]() { kernel(args...); };
        for (int t = 0; t < nt; ++t) {
            Fiber& f = B.fibers[t];
            f.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            f.lane = t & 31;
            f.warp = t >> 5;
            B.warp_bars[f.warp].expected++;
            B.warp_alive[f.warp] |= 1u << f.lane;
            f.stack = stacks[(size_t)bi * nt + t] = (char*)malloc(CO_STACK);
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = CO_STACK;
            f.ctx.uc_link = &B.sched;
            makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
    }
    const int total = nb * nt;
    std::vector<int> order(total);
    for (int i = 0; i < total; ++i) order[i] = i;
    int remaining = total;
    while (remaining > 0) {
        if (g_order == 1) {
            for (int i = 0; i < total; ++i) order[i] = total - 1 - i;
        } else if (g_order == 2) {
            std::shuffle(order.begin(), order.end(), g_rng);
        }
        const uint64_t before = g_progress;
        for (int i = 0; i < total; ++i) {
            Block& B = blocks[order[i] / nt];
            Fiber& f = B.fibers[order[i] % nt];
            if (f.done) continue;
            g_blk = &B;
            g_blockIdx = B.bidx;
            B.cur = order[i] % nt;
            swapcontext(&B.sched, &f.ctx);
            if (f.done) --remaining;
        }
        if (remaining > 0 && g_progress == before) {
            fprintf(stderr, "cuda_emu: DEADLOCK in a co-resident launch: %d threads poll or wait and nothing changes\n", remaining);
            abort();
        }
    }
    for (char* st : stacks) free(st);
    g_blk = nullptr;
}

// ---- warp collectives ------------------------------------------------------------------------
template <typename T>
inline T exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 16, "exchange payload");
    Block& B = *g_blk;
    Fiber& f = cur();
    unsigned char* slot = B.warp_slots.data() + (size_t)f.warp * 32 * 16;
    memcpy(slot + f.lane * 16, &v, sizeof(T));
    bar_wait(B.warp_bars[f.warp]);
    T r;
    memcpy(&r, slot + (src_lane & 31) * 16, sizeof(T));
    bar_wait(B.warp_bars[f.warp]);
    return r;
}
// D(8x8) += A(8x4, row) * B(4x8, col); lane holds a = A[lane/4][lane%4], b = B[lane%4][lane/4],
// d0/d1 = D[lane/4][2*(lane%4) + {0,1}]  (PTX mma.sync.aligned.m8n8k4.row.col.f64)
inline void dmma(double& d0, double& d1, double a, double b) {
    Block& B = *g_blk;
    Fiber& f = cur();
    unsigned char* slot = B.warp_slots.data() + (size_t)f.warp * 32 * 16;
    double ab[2] = {a, b};
    memcpy(slot + f.lane * 16, ab, 16);
    bar_wait(B.warp_bars[f.warp]);
    const int r = f.lane >> 2, c0 = (f.lane & 3) * 2;
    for (int k = 0; k < 4; ++k) {
        double ark, b0, b1;
        memcpy(&ark, slot + (r * 4 + k) * 16, 8);
        memcpy(&b0, slot + (c0 * 4 + k) * 16 + 8, 8);
        memcpy(&b1, slot + ((c0 + 1) * 4 + k) * 16 + 8, 8);
        d0 = fma(ark, b0, d0);
        d1 = fma(ark, b1, d1);
    }
    bar_wait(B.warp_bars[f.warp]);
}

}  // namespace emu

#define threadIdx (emu::cur().tid)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
// dynamic shared memory: `MAK_DYN_SMEM(name);` declares `unsigned char* name` / the extern array
#define MAK_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()

inline void __syncthreads() { emu::bar_wait(emu::g_blk->block_bar); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::bar_wait(emu::g_blk->warp_bars[emu::cur().warp]); }
#define MAK_SPIN_PAUSE() emu::spin_pause()
inline void __threadfence() {}
inline void __threadfence_block() {}

template <typename T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    const int lane = emu::cur().lane;
    return emu::exchange(v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lanemask, int width = 32) {
    const int lane = emu::cur().lane;
    int src = lane ^ lanemask;
    if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
    return emu::exchange(v, src);
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
    const int lane = emu::cur().lane;
    int src = lane + (int)delta;
    if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
    return emu::exchange(v, src);
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
    const int lane = emu::cur().lane;
    int src = lane - (int)delta;
    if (src < 0 || (src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
    return emu::exchange(v, src);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    emu::Block& B = *emu::g_blk;
    emu::Fiber& f = emu::cur();
    unsigned char* slot = B.warp_slots.data() + (size_t)f.warp * 32 * 16;
    int p = pred ? 1 : 0;
    memcpy(slot + f.lane * 16, &p, sizeof(int));
    emu::bar_wait(B.warp_bars[f.warp]);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) {
        if (!(B.warp_alive[f.warp] >> l & 1u)) continue;
        int q;
        memcpy(&q, slot + l * 16, sizeof(int));
        if (q) r |= 1u << l;
    }
    emu::bar_wait(B.warp_bars[f.warp]);
    return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) {
    const unsigned alive = emu::g_blk->warp_alive[emu::cur().warp];
    return (__ballot_sync(m, pred) & alive) == alive;
}
inline unsigned __activemask() { return emu::g_blk->warp_alive[emu::cur().warp]; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }

template <typename T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <typename T> inline T __ldg(const T* p) { return *p; }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
