// CPU replay of the persistent tridiagonalisation column step's tile geometry (csrc/trd_tiles.h, csrc/trd2.cuh):
// the same chunking, band/strip walk, partial-slot addressing and partial sums as the kernels, with plain loops in
// place of the TMA ring and the warp reductions.  y = A22 v over the LOWER triangle only; compared with numpy in
// tests/test_trd_tiles_cpu.py.  Test infrastructure only.
#include <vector>
#include <cstring>
#include "../../matrixalgebrakit.jl_b200/csrc/trd_tiles.h"

using namespace mak;

extern "C" {

// A: n x n column-major (only entries with row >= col >= row0 may be read), v: n entries (global index, zero outside
// [row0, n)), y: n entries out (rows >= row0).  Returns the number of geometry violations found (0 = ok):
// a tile visited twice, an element of the trailing lower triangle never visited, a slot written twice, a slot read
// that was never written.
int trd_tiles_replay(int n, int row0, int BH, int CW, int G, const double* A, const double* v, double* y, double* vAv) {
    const TrdTiling tl = trd_tiling(n, row0, BH, CW, G);
    int bad = 0;
    std::vector<double> yrow((size_t)G * n, 0.0), ycol((size_t)tl.JB * n, 0.0);
    std::vector<char> wrow((size_t)G * n, 0), wcol((size_t)tl.JB * n, 0);
    std::vector<int> seen((size_t)n * n, 0);
    double q = 0.0;
    int tiles_done = 0;
    for (int g = 0; g < G; ++g) {
        int t_beg = g * tl.q, t_end = t_beg + tl.q;
        if (t_beg > tl.NT) t_beg = tl.NT;
        if (t_end > tl.NT) t_end = tl.NT;
        if (t_beg >= t_end) continue;
        int J = trd_tile_band(tl, t_beg);
        int S = tl.S0 + (t_beg - trd_band_first_tile(tl, J));
        int Slast = trd_band_last_strip(tl, J);
        std::vector<double> rowacc(BH, 0.0), adiag(BH, 0.0);
        bool open = false;
        for (int t = t_beg; t < t_end; ++t) {
            if (J >= tl.JB || S < tl.S0 || S > Slast) { ++bad; break; }
            if (!open) { std::fill(rowacc.begin(), rowacc.end(), 0.0); std::fill(adiag.begin(), adiag.end(), 0.0); open = true; }
            ++tiles_done;
            for (int k = 0; k < CW; ++k) {
                const int gc = CW * S + k;
                double colacc = 0.0;
                for (int rr = 0; rr < BH; ++rr) {
                    const int gr = BH * J + rr;
                    if (gr >= n || gc >= n) continue;
                    const double vr = (gr >= row0) ? v[gr] : 0.0, vc = (gc >= row0) ? v[gc] : 0.0;
                    if (gr > gc) {
                        const double a = A[(size_t)gc * n + gr];
                        if (gc >= row0) { if (seen[(size_t)gc * n + gr]++) ++bad; }
                        rowacc[rr] += a * vc;
                        colacc += a * vr;
                    } else if (gr == gc) {
                        adiag[rr] = A[(size_t)gc * n + gr];
                        if (gc >= row0) { if (seen[(size_t)gc * n + gr]++) ++bad; }
                    }
                }
                if (gc < n) {
                    if (wcol[(size_t)J * n + gc]++) ++bad;
                    ycol[(size_t)J * n + gc] = colacc;
                }
            }
            const bool band_end = (S == Slast) || (t + 1 == t_end);
            if (band_end) {
                for (int rr = 0; rr < BH; ++rr) {
                    const int gr = BH * J + rr;
                    if (gr >= n) continue;
                    const double vr = (gr >= row0) ? v[gr] : 0.0;
                    if (wrow[(size_t)g * n + gr]++) ++bad;
                    yrow[(size_t)g * n + gr] = rowacc[rr] + vr * adiag[rr];
                    q += 2.0 * vr * rowacc[rr] + adiag[rr] * vr * vr;
                }
                open = false;
            }
            if (S == Slast) { ++J; S = tl.S0; Slast = (J < tl.JB) ? trd_band_last_strip(tl, J) : 0; }
            else ++S;
        }
    }
    if (tiles_done != tl.NT) ++bad;
    for (int c = row0; c < n; ++c)
        for (int r = c; r < n; ++r)
            if (seen[(size_t)c * n + r] != 1) ++bad;
    // sums as in trd2_ysum
    for (int r = row0; r < n; ++r) {
        const int J = r / tl.BH;
        int g_lo, g_hi;
        trd_band_chunks(tl, J, g_lo, g_hi);
        double s = 0.0;
        for (int gg = g_lo; gg <= g_hi; ++gg) {
            if (!wrow[(size_t)gg * n + r]) ++bad;
            s += yrow[(size_t)gg * n + r];
        }
        for (int JJ = J; JJ < tl.JB; ++JJ) {
            if (!wcol[(size_t)JJ * n + r]) ++bad;
            s += ycol[(size_t)JJ * n + r];
        }
        y[r] = s;
    }
    // the panel-dot slices cover [0, mt) exactly once, each at most 64 rows when G >= mt/64
    {
        const int mt = n - row0;
        int next = 0;
        for (int g = 0; g < G; ++g) {
            int lo, hi;
            trd_slice(tl, g, lo, hi);
            if (lo != next && lo != hi) ++bad;
            if (hi > lo) next = hi;
            if (64 * G >= mt && hi - lo > 64) ++bad;
        }
        if (next != mt) ++bad;
    }
    *vAv = q;
    return bad;
}

}  // extern "C"
