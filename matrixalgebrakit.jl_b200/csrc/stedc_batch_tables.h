// Host-side tables of the batched tridiagonal D&C (stedc.cu: stedc_batched): the blocks laid end to end as one
// block-diagonal problem, every block with its own binary tree.  Plain C++ (shared with tests/cpu_harness/
// stedc_batched_host.cpp, which replays the batched solver on the CPU with the same tables and the same strip storage).
//
// Storage of the "Ntot x Ntot" eigenvector matrices: strips with leading dimension ld = nmax.  The kernels address
// Z[(lo + c) * ld + lo + r] with GLOBAL positions and only touch diagonal sub-blocks, so block i (offset o_i, order n_i)
// owns the addresses o_i (ld + 1) + [0, n_i ld): disjoint from every other block because ld >= n_i.
#pragma once
#include "stedc_core.h"
#include <algorithm>
#include <cstddef>
#include <vector>

namespace mak {
namespace dc {

inline int dc_tree_levels(int n) {
    int L = 0;
    while ((n + (1 << L) - 1) / (1 << L) > DC_LEAF) ++L;
    return L;
}

struct BatchLevel { size_t first; int nm, maxN, maxH; };   // merges [first, first + nm) of `merges`
struct BatchTables {
    size_t ntot = 0, nleaves = 0;
    int nmax = 1, Lmax = 0;
    std::vector<int> off, lev;       // per block: global offset, tree depth (the block ends in ping-pong buffer lev & 1)
    std::vector<int> bnd, cuts;      // global leaf boundaries (nleaves + 1) and tear positions
    std::vector<Merge> merges;       // all levels, level by level
    std::vector<BatchLevel> levels;  // levels 1 .. Lmax
    size_t strip_elems() const { return (ntot > 0 ? ntot : 1) * ((size_t)nmax + 1); }
    size_t block_base(int i) const { return (size_t)off[i] * ((size_t)nmax + 1); }
};

inline BatchTables dc_batch_tables(int nblk, const int* n) {
    BatchTables t;
    t.off.resize(nblk); t.lev.resize(nblk);
    std::vector<std::vector<int>> lb(nblk);
    for (int i = 0; i < nblk; ++i) {
        t.off[i] = (int)t.ntot;
        t.ntot += (size_t)n[i];
        t.lev[i] = dc_tree_levels(n[i]);
        t.nleaves += (size_t)1 << t.lev[i];
        t.nmax = std::max(t.nmax, n[i]);
        t.Lmax = std::max(t.Lmax, t.lev[i]);
    }
    for (int i = 0; i < nblk; ++i) {
        const int nl = 1 << t.lev[i];
        lb[i].resize(nl + 1);
        for (int k = 0; k <= nl; ++k) lb[i][k] = t.off[i] + (int)((long long)k * n[i] / nl);
        for (int k = 0; k < nl; ++k) t.bnd.push_back(lb[i][k]);
        for (int k = 1; k < nl; ++k) t.cuts.push_back(lb[i][k]);
    }
    t.bnd.push_back((int)t.ntot);
    for (int l = 1; l <= t.Lmax; ++l) {
        BatchLevel li{t.merges.size(), 0, 0, 0};
        const int step = 1 << l;
        for (int i = 0; i < nblk; ++i) {
            if (t.lev[i] < l) continue;
            const int nm = (1 << t.lev[i]) / step;
            for (int k = 0; k < nm; ++k) {
                Merge m{};
                m.lo = lb[i][k * step]; m.mid = lb[i][k * step + step / 2]; m.hi = lb[i][(k + 1) * step];
                t.merges.push_back(m);
                li.maxN = std::max(li.maxN, m.hi - m.lo);
                li.maxH = std::max(li.maxH, std::max(m.mid - m.lo, m.hi - m.mid));
                ++li.nm;
            }
        }
        t.levels.push_back(li);
    }
    return t;
}

}  // namespace dc
}  // namespace mak
