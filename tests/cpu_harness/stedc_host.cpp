// CPU harness for the product's D&C work-item bodies (matrixalgebrakit.jl_b200/csrc/stedc_core.h).
// Mirrors the kernel sequence of stedc.cu with plain loops so the deflation / secular-equation /
// Loewner logic can be unit-tested on the CPU-only build box.  Test infrastructure only.
#include <vector>
#include <algorithm>
#include <cstring>
#include <cstdio>
#include "../../matrixalgebrakit.jl_b200/csrc/stedc_core.h"

using namespace mak::dc;

extern "C" int stedc_host(int n, const double* d_in, const double* e_in, double* w, double* Zout, int* stats) {
    if (n == 0) return 0;
    std::vector<double> D(n), E(n, 0.0), Dn(n), z(n), dl(n), zl(n), rc(n), rs(n), tau(n), zhat(n);
    std::vector<int> perm(n), src(n), ctype(n), rowpos(n), rp(n), rq(n), rtp(n), rtq(n), orig(n), pos(n);
    double nrm = 0.0;
    for (int i = 0; i < n; ++i) nrm = std::max(nrm, std::fabs(d_in[i]));
    for (int i = 0; i + 1 < n; ++i) nrm = std::max(nrm, std::fabs(e_in[i]));
    double scale = nrm > 0 ? nrm : 1.0;
    for (int i = 0; i < n; ++i) D[i] = d_in[i] / scale;
    for (int i = 0; i + 1 < n; ++i) E[i] = e_in[i] / scale;
    int L = 0;
    while ((n + (1 << L) - 1) / (1 << L) > DC_LEAF) ++L;
    int nleaf = 1 << L;
    std::vector<int> bnd(nleaf + 1);
    for (int i = 0; i <= nleaf; ++i) bnd[i] = (int)((long long)i * n / nleaf);
    std::vector<double> Za((size_t)n * n, 0.0), Zb((size_t)n * n, 0.0), Pack((size_t)n * n, 0.0), S((size_t)n * n, 0.0);
    // tears
    std::vector<double> rho_cut(n, 0.0), sgn_cut(n, 1.0);
    for (int i = 1; i < nleaf; ++i) {
        int c = bnd[i];
        double e = E[c - 1];
        rho_cut[c] = std::fabs(e); sgn_cut[c] = e < 0 ? -1.0 : 1.0;
        D[c - 1] -= std::fabs(e); D[c] -= std::fabs(e);
    }
    double* Zin = Za.data(); double* Zo = Zb.data();
    int ld = n;
    for (int i = 0; i < nleaf; ++i) {
        int lo = bnd[i], sz = bnd[i + 1] - lo;
        double dd[DC_LEAF + 1], ee[DC_LEAF + 1];
        for (int k = 0; k < sz; ++k) { dd[k] = D[lo + k]; ee[k] = (k + 1 < sz) ? E[lo + k] : 0.0; }
        for (int c = 0; c < sz; ++c) for (int r = 0; r < sz; ++r) Zin[(size_t)(lo + c) * ld + lo + r] = (r == c);
        int rcq = leaf_ql(sz, dd, ee, Zin + (size_t)lo * ld + lo, ld);
        if (rcq) return 1000 + rcq;
        for (int k = 0; k < sz; ++k) D[lo + k] = dd[k];
    }
    Ctx c;
    c.n = n; c.D = D.data(); c.Dn = Dn.data(); c.z = z.data(); c.perm = perm.data(); c.dl = dl.data(); c.zl = zl.data();
    c.src = src.data(); c.ctype = ctype.data(); c.rowpos = rowpos.data(); c.rot_p = rp.data(); c.rot_q = rq.data();
    c.rot_c = rc.data(); c.rot_s = rs.data(); c.rot_tp = rtp.data(); c.rot_tq = rtq.data(); c.tau = tau.data();
    c.orig = orig.data(); c.zhat = zhat.data(); c.pos = pos.data();
    int total_defl = 0, total_rot = 0;
    for (int lev = 1; lev <= L; ++lev) {
        int step = 1 << lev;
        for (int b = 0; b < nleaf; b += step) {
            Merge mg;
            mg.lo = bnd[b]; mg.mid = bnd[b + step / 2]; mg.hi = bnd[b + step];
            mg.rho = 2.0 * rho_cut[mg.mid]; mg.sgn = sgn_cut[mg.mid];
            int N = mg.hi - mg.lo, N1 = mg.mid - mg.lo, N2 = mg.hi - mg.mid, lo = mg.lo;
            for (int i = 0; i < N; ++i) merge_z_item(c, mg, Zin, ld, i);
            for (int i = 0; i < N; ++i) merge_rank_item(c, mg, i);
            deflate_scan(c, mg);
            total_defl += N - mg.K; total_rot += mg.nrot;
            for (int r = 0; r < N; ++r) rotate_row_item(c, mg, Zin, ld, r);
            int K = mg.K;
            for (int j = 0; j < K; ++j) secular_root(K, j, c.dl + lo, c.zl + lo, mg.rho, c.tau + lo + j, c.orig + lo + j);
            for (int i = 0; i < K; ++i) zhat_item(K, c.dl + lo, c.zl + lo, mg.rho, c.tau + lo, c.orig + lo, c.zhat + lo, i);
            for (int j = 0; j < N; ++j) final_pos_item(c, mg, j);
            // S
            for (int j = 0; j < K; ++j) {
                double nn = 0.0;
                for (int i = 0; i < K; ++i) {
                    double v = c.zhat[lo + i] / sec_delta(c.dl + lo, c.tau + lo, c.orig + lo, i, j);
                    nn += v * v;
                }
                nn = std::sqrt(nn);
                for (int i = 0; i < K; ++i) {
                    double v = c.zhat[lo + i] / sec_delta(c.dl + lo, c.tau + lo, c.orig + lo, i, j);
                    S[(size_t)(lo + j) * ld + lo + c.rowpos[lo + i]] = v / nn;
                }
            }
            // pack + deflated copy
            for (int j = 0; j < N; ++j) {
                int sc = c.src[lo + j], t = c.ctype[lo + j];
                const double* col = Zin + (size_t)(lo + sc) * ld + lo;
                if (j < K) {
                    int p = c.rowpos[lo + j];
                    if (t != 3) for (int r = 0; r < N1; ++r) Pack[(size_t)(lo + p) * ld + lo + r] = col[r];
                    if (t != 1) for (int r = 0; r < N2; ++r) Pack[(size_t)(lo + p - mg.k1) * ld + mg.mid + r] = col[N1 + r];
                } else {
                    double* dst = Zo + (size_t)(lo + c.pos[lo + j]) * ld + lo;
                    for (int r = 0; r < N; ++r) {
                        bool top = r < N1;
                        dst[r] = ((t == 1 && !top) || (t == 3 && top)) ? 0.0 : col[r];
                    }
                }
            }
            // GEMMs into Tmp = Zin[lo:hi, lo:lo+K]
            int k12 = mg.k1 + mg.k2, k23 = mg.k2 + mg.k3;
            std::vector<double> Tmp((size_t)N * std::max(K, 1), 0.0);
            for (int j = 0; j < K; ++j) {
                for (int r = 0; r < N1; ++r) {
                    double s = 0.0;
                    for (int p = 0; p < k12; ++p) s += Pack[(size_t)(lo + p) * ld + lo + r] * S[(size_t)(lo + j) * ld + lo + p];
                    Tmp[(size_t)j * N + r] = s;
                }
                for (int r = 0; r < N2; ++r) {
                    double s = 0.0;
                    for (int p = 0; p < k23; ++p) s += Pack[(size_t)(lo + p) * ld + mg.mid + r] * S[(size_t)(lo + j) * ld + lo + mg.k1 + p];
                    Tmp[(size_t)j * N + N1 + r] = s;
                }
            }
            for (int j = 0; j < K; ++j) {
                double* dst = Zo + (size_t)(lo + c.pos[lo + j]) * ld + lo;
                for (int r = 0; r < N; ++r) dst[r] = Tmp[(size_t)j * N + r];
            }
        }
        std::swap(Zin, Zo);
        std::swap(c.D, c.Dn);
    }
    for (int i = 0; i < n; ++i) w[i] = c.D[i] * scale;
    std::memcpy(Zout, Zin, sizeof(double) * (size_t)n * n);
    if (stats) { stats[0] = total_defl; stats[1] = total_rot; stats[2] = L; }
    return 0;
}
