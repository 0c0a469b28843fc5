// CPU replay of the batched tridiagonal D&C (csrc/stedc.cu: stedc_batched) with the PRODUCT tables
// (csrc/stedc_batch_tables.h) and the product work-item bodies (csrc/stedc_core.h): the tridiagonals of all blocks laid end
// to end, every block with its own tree, the eigenvector "matrices" stored as strips with leading dimension nmax that are
// NaN-filled first (a read outside a block's own region poisons the result), blocks finishing in ping-pong buffer
// (levels & 1).  Test infrastructure only (tests/test_stedc_batched_cpu.py).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#include "../../matrixalgebrakit.jl_b200/csrc/stedc_batch_tables.h"

using namespace mak::dc;

// d, e: concatenated inputs (block i at offset off_i, e of length n_i per block with the last entry unused);
// w: concatenated eigenvalues out; V: concatenated n_i x n_i eigenvector matrices out (column-major, ld n_i)
extern "C" int stedc_batched_host(int nblk, const int* n, const double* d_in, const double* e_in, double* w, double* V,
                                  int* stats) {
    const BatchTables t = dc_batch_tables(nblk, n);
    const size_t N = t.ntot;
    const int ld = t.nmax;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    std::vector<double> D(N), E(N, 0.0), Dn(N, nan), z(N), dl(N), zl(N), rc(N), rs(N), tau(N), zhat(N), scale(nblk);
    std::vector<int> perm(N), src(N), ctype(N), rowpos(N), rp(N), rq(N), rtp(N), rtq(N), orig(N), pos(N);
    std::vector<double> Za(t.strip_elems(), nan), Zb(t.strip_elems(), nan), Pack(t.strip_elems(), nan), S(t.strip_elems(), nan);
    // structural checks of the tables
    if (t.bnd.size() != t.nleaves + 1 || t.merges.size() + (size_t)nblk != t.nleaves) return -1;
    for (size_t i = 0; i + 1 < t.bnd.size(); ++i)
        if (t.bnd[i + 1] <= t.bnd[i] || t.bnd[i + 1] - t.bnd[i] > DC_LEAF) return -2;
    for (int i = 0; i < nblk; ++i) {
        // the strip regions of different blocks are disjoint and inside the strip
        const size_t b0 = t.block_base(i), b1 = b0 + (size_t)(n[i] - 1) * ld + n[i];
        if (b1 > t.strip_elems()) return -3;
        if (i + 1 < nblk && b1 > t.block_base(i + 1)) return -4;
    }
    // per-block scaling
    for (int i = 0; i < nblk; ++i) {
        const int o = t.off[i];
        double nrm = 0.0;
        for (int k = 0; k < n[i]; ++k) nrm = std::max(nrm, std::fabs(d_in[o + k]));
        for (int k = 0; k + 1 < n[i]; ++k) nrm = std::max(nrm, std::fabs(e_in[o + k]));
        scale[i] = nrm > 0 ? nrm : 1.0;
        for (int k = 0; k < n[i]; ++k) { D[o + k] = d_in[o + k] / scale[i]; E[o + k] = (k + 1 < n[i]) ? e_in[o + k] / scale[i] : 0.0; }
    }
    std::vector<double> rho_cut(N + 1, 0.0), sgn_cut(N + 1, 1.0);
    for (int c : t.cuts) {
        const double e = E[c - 1];
        rho_cut[c] = std::fabs(e); sgn_cut[c] = e < 0 ? -1.0 : 1.0;
        D[c - 1] -= std::fabs(e); D[c] -= std::fabs(e);
    }
    double* Zin = Za.data(); double* Zo = Zb.data();
    for (size_t i = 0; i < t.nleaves; ++i) {
        const int lo = t.bnd[i], sz = t.bnd[i + 1] - lo;
        double dd[DC_LEAF + 1], ee[DC_LEAF + 1];
        for (int k = 0; k < sz; ++k) { dd[k] = D[lo + k]; ee[k] = (k + 1 < sz) ? E[lo + k] : 0.0; }
        for (int c = 0; c < sz; ++c) for (int r = 0; r < sz; ++r) Zin[(size_t)(lo + c) * ld + lo + r] = (r == c);
        const int rcq = leaf_ql(sz, dd, ee, Zin + (size_t)lo * ld + lo, ld);
        if (rcq) return 1000 + rcq;
        for (int k = 0; k < sz; ++k) D[lo + k] = dd[k];
    }
    Ctx c;
    c.n = (int)N; c.D = D.data(); c.Dn = Dn.data(); c.z = z.data(); c.perm = perm.data(); c.dl = dl.data(); c.zl = zl.data();
    c.src = src.data(); c.ctype = ctype.data(); c.rowpos = rowpos.data(); c.rot_p = rp.data(); c.rot_q = rq.data();
    c.rot_c = rc.data(); c.rot_s = rs.data(); c.rot_tp = rtp.data(); c.rot_tq = rtq.data(); c.tau = tau.data();
    c.orig = orig.data(); c.zhat = zhat.data(); c.pos = pos.data();
    double* D0 = c.D; double* D1 = c.Dn;
    int merges_done = 0;
    for (const BatchLevel& li : t.levels) {
        for (int q = 0; q < li.nm; ++q) {
            Merge mg = t.merges[li.first + q];
            mg.rho = 2.0 * rho_cut[mg.mid]; mg.sgn = sgn_cut[mg.mid];
            const int Nn = mg.hi - mg.lo, N1 = mg.mid - mg.lo, N2 = mg.hi - mg.mid, lo = mg.lo;
            if (Nn > li.maxN || std::max(N1, N2) > li.maxH) return -5;
            for (int i = 0; i < Nn; ++i) merge_z_item(c, mg, Zin, ld, i);
            for (int i = 0; i < Nn; ++i) merge_rank_item(c, mg, i);
            deflate_scan(c, mg);
            for (int r = 0; r < Nn; ++r) rotate_row_item(c, mg, Zin, ld, r);
            const int K = mg.K;
            for (int j = 0; j < K; ++j) secular_root(K, j, c.dl + lo, c.zl + lo, mg.rho, c.tau + lo + j, c.orig + lo + j);
            for (int i = 0; i < K; ++i) zhat_item(K, c.dl + lo, c.zl + lo, mg.rho, c.tau + lo, c.orig + lo, c.zhat + lo, i);
            for (int j = 0; j < Nn; ++j) final_pos_item(c, mg, j);
            for (int j = 0; j < K; ++j) {
                double nn = 0.0;
                for (int i = 0; i < K; ++i) {
                    const double v = c.zhat[lo + i] / sec_delta(c.dl + lo, c.tau + lo, c.orig + lo, i, j);
                    nn += v * v;
                }
                nn = std::sqrt(nn);
                for (int i = 0; i < K; ++i) {
                    const double v = c.zhat[lo + i] / sec_delta(c.dl + lo, c.tau + lo, c.orig + lo, i, j);
                    S[(size_t)(lo + j) * ld + lo + c.rowpos[lo + i]] = v / nn;
                }
            }
            for (int j = 0; j < Nn; ++j) {
                const int sc = c.src[lo + j], ty = c.ctype[lo + j];
                const double* col = Zin + (size_t)(lo + sc) * ld + lo;
                if (j < K) {
                    const int p = c.rowpos[lo + j];
                    if (ty != 3) for (int r = 0; r < N1; ++r) Pack[(size_t)(lo + p) * ld + lo + r] = col[r];
                    if (ty != 1) for (int r = 0; r < N2; ++r) Pack[(size_t)(lo + p - mg.k1) * ld + mg.mid + r] = col[N1 + r];
                } else {
                    double* dst = Zo + (size_t)(lo + c.pos[lo + j]) * ld + lo;
                    for (int r = 0; r < Nn; ++r) {
                        const bool top = r < N1;
                        dst[r] = ((ty == 1 && !top) || (ty == 3 && top)) ? 0.0 : col[r];
                    }
                }
            }
            const int k12 = mg.k1 + mg.k2, k23 = mg.k2 + mg.k3;
            std::vector<double> Tmp((size_t)Nn * std::max(K, 1), 0.0);
            for (int j = 0; j < K; ++j) {
                for (int r = 0; r < N1; ++r) {
                    double s = 0.0;
                    for (int p = 0; p < k12; ++p) s += Pack[(size_t)(lo + p) * ld + lo + r] * S[(size_t)(lo + j) * ld + lo + p];
                    Tmp[(size_t)j * Nn + r] = s;
                }
                for (int r = 0; r < N2; ++r) {
                    double s = 0.0;
                    for (int p = 0; p < k23; ++p) s += Pack[(size_t)(lo + p) * ld + mg.mid + r] * S[(size_t)(lo + j) * ld + lo + mg.k1 + p];
                    Tmp[(size_t)j * Nn + N1 + r] = s;
                }
            }
            for (int j = 0; j < K; ++j) {
                double* dst = Zo + (size_t)(lo + c.pos[lo + j]) * ld + lo;
                for (int r = 0; r < Nn; ++r) dst[r] = Tmp[(size_t)j * Nn + r];
            }
            ++merges_done;
        }
        std::swap(Zin, Zo);
        std::swap(c.D, c.Dn);
    }
    // finishing pass: each block from ITS ping-pong buffer
    size_t vo = 0;
    for (int i = 0; i < nblk; ++i) {
        const int par = t.lev[i] & 1, o = t.off[i];
        const double* Dd = par ? D1 : D0;
        const double* Z = (par ? Zb.data() : Za.data()) + t.block_base(i);
        for (int k = 0; k < n[i]; ++k) w[o + k] = Dd[o + k] * scale[i];
        for (int cc = 0; cc < n[i]; ++cc)
            for (int r = 0; r < n[i]; ++r) V[vo + (size_t)cc * n[i] + r] = Z[(size_t)cc * ld + r];
        vo += (size_t)n[i] * n[i];
    }
    if (stats) { stats[0] = merges_done; stats[1] = t.Lmax; stats[2] = (int)t.nleaves; }
    return 0;
}
