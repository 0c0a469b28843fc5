// Tile geometry of the persistent tridiagonalisation column kernel (trd2.cuh).  Host/device, no CUDA
// dependency: compiled with g++ by tests/cpu_harness/trd_tiles_host.cpp, which replays the whole column
// step (tiles -> partial buffers -> sums) on the CPU and compares it with a dense y = A22 v.
//
// Column step c of the reduction works on the trailing block A22 = A[row0:n, row0:n], row0 = c + 1, of
// which only the LOWER triangle is read.  The triangle is cut into tiles on a grid that is fixed in GLOBAL
// coordinates (so every TMA box is 16-byte aligned for every c and the tile <-> address map does not move
// between columns):
//     band  J : global rows [BH J, BH J + BH)          BH = 256
//     strip S : global cols [CW S, CW S + CW)          CW = 16 (Float64) / 8 (ComplexF64)
// Tile (J, S) is needed iff it holds an element with row >= col inside the trailing block:
//     J0 = row0 / BH <= J < JB = ceil(n / BH),   S0 = row0 / CW <= S <= min(SPB J + SPB - 1, SN - 1),
// SPB = BH / CW strips per band, SN = ceil(n / CW).  Tiles are numbered band-major (J ascending, S ascending
// inside the band) and dealt to the G persistent CTAs as contiguous chunks of q = ceil(NT / G) tiles, so a CTA
// streams consecutive strips of one band (the row part of y stays in a register per thread) and every tile is
// one stage of its shared-memory ring.  Partial results are written to fixed slots (no atomics, bit-reproducible):
//     yrow[g][r]  row part of y for global row r from chunk g   (chunks g_lo(J) .. g_hi(J) touch band J)
//     ycol[J][k]  column part of y for global column k from band J   (bands J >= k / BH)
#pragma once
#include "scalar.h"

namespace mak {

struct TrdTiling {
    int n, row0, BH, CW, SPB;
    int J0, JB, S0, SN;
    int G, NT, q;
};

__host__ __device__ __forceinline__ int trd_band_last_strip(const TrdTiling& t, int J) {
    const int s = t.SPB * J + t.SPB - 1;
    return s < t.SN - 1 ? s : t.SN - 1;
}
__host__ __device__ __forceinline__ int trd_band_ntiles(const TrdTiling& t, int J) {
    return trd_band_last_strip(t, J) - t.S0 + 1;
}
// number of tiles in bands J0 .. J-1   (J0 <= J <= JB)
__host__ __device__ __forceinline__ int trd_band_first_tile(const TrdTiling& t, int J) {
    // bands before the last one are never clipped by SN:  nt(J') = SPB (J' + 1) - S0
    const int a = J - t.J0;
    int s = t.SPB * (J * (J + 1) / 2 - t.J0 * (t.J0 + 1) / 2) - a * t.S0;
    if (J == t.JB && J > t.J0) {
        // the sum above used the unclipped count for band JB-1: correct it
        s += trd_band_ntiles(t, t.JB - 1) - (t.SPB * t.JB - t.S0);
    }
    return s;
}

__host__ __device__ __forceinline__ TrdTiling trd_tiling(int n, int row0, int BH, int CW, int G) {
    TrdTiling t;
    t.n = n; t.row0 = row0; t.BH = BH; t.CW = CW; t.SPB = BH / CW;
    t.J0 = row0 / BH; t.JB = (n + BH - 1) / BH;
    t.S0 = row0 / CW; t.SN = (n + CW - 1) / CW;
    t.G = G;
    t.NT = (row0 < n) ? trd_band_first_tile(t, t.JB) : 0;
    t.q = (t.NT + G - 1) / G;
    if (t.q < 1) t.q = 1;
    return t;
}

// band of linear tile index `tile` (0 <= tile < NT)
__host__ __device__ __forceinline__ int trd_tile_band(const TrdTiling& t, int tile) {
    int J = t.J0;
    while (J + 1 < t.JB && trd_band_first_tile(t, J + 1) <= tile) ++J;
    return J;
}
// chunks that hold tiles of band J: [g_lo, g_hi]
__host__ __device__ __forceinline__ void trd_band_chunks(const TrdTiling& t, int J, int& g_lo, int& g_hi) {
    const int f = trd_band_first_tile(t, J), l = f + trd_band_ntiles(t, J) - 1;
    g_lo = f / t.q;
    g_hi = l / t.q;
}
// rows of the panel-dot slice of chunk g (local to the trailing block): [r_lo, r_hi)
__host__ __device__ __forceinline__ void trd_slice(const TrdTiling& t, int g, int& r_lo, int& r_hi) {
    const int mt = t.n - t.row0, sl = (mt + t.G - 1) / t.G;
    r_lo = g * sl;
    r_hi = r_lo + sl;
    if (r_lo > mt) r_lo = mt;
    if (r_hi > mt) r_hi = mt;
}

}  // namespace mak
