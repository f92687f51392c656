// Slab decomposition layout math, shared by host and device code (single source of truth).
//
// G ranks.  Stage 1 (generation, x and z transforms) is sharded over y: rank g owns the
// h = N/(2G) primary rows y = g, g+G, g+2G, ... < N/2 and their Hermitian partners N-y, 2h rows in
// all — the reference keeps +ky and -ky planes together for the same reason
// (reference src/zeldovich.cpp:558-587, yblock and numblock-1-yblock).  The assignment is CYCLIC
// because the work of a row is not uniform: modes outside the k_cutoff sphere are masked, so a
// rank holding only high-|ky| rows would draw a fraction of the modes (and transform a fraction of
// the non-zero rows) that a rank holding low-|ky| rows does — with contiguous blocks of rows the
// slowest rank of 8 had 8.5x the generation work of the fastest and everybody waited for it.
// Row y = 0 has no partner; rank 0 uses that free slot for the all-zero Nyquist row y = N/2
// (reference src/zeldovich.cpp:640-650).
//
//   slot s of rank g:  s <  h : y = s*G + g
//                      s >= h : y = N - ((s - h)*G + g)      (rank 0, s == h : y = N/2)
//
// Stage-1 buffer (per rank):  B1[z][a][slot][x]            N * na * 2h * N complex
//   = G contiguous blocks along z, block r = planes z in [r*N/G, (r+1)*N/G): what rank r needs.
// One all-to-all (block r of every rank goes to rank r) — the y<->z transpose that the
// reference's BlockArray StoreBlock/LoadBlock perform through RAM or disk
// (reference src/block_array.cpp:387-414, 466-504) — gives
// Stage-2 buffer (per rank):  B2[src][zl][a][slot][x]      G * (N/G) * na * 2h * N complex
// in which rank r finds every y of its planes zl in [0, N/G) (global z = r*N/G + zl).
#pragma once

#ifdef __CUDACC__
#define ZPLT_HD __host__ __device__ __forceinline__
#else
#define ZPLT_HD inline
#endif

namespace zplt {

struct SlabGeom {
    int N;     // ppd
    int G;     // ranks
    int rank;  // this rank
    int h;     // primary rows per rank = N / (2G)  (a power of two, like N and G)
    int na;    // packed arrays
    int log2h;
    int log2G;
    // stage 1 can run in groups of primary rows [ly0, ly0+nly) (and their partners' slots
    // h+ly0 ...), so that generating one group overlaps with sending the previous one
    int ly0, nly;
    // receive layout of the fused exchange (the sender computes every address, so it is a choice): rows at their true y,
    // [zl][a][y][x] with b2_zstride elements between planes (>= na*N*N: padding breaks the power-of-two plane stride), or
    // (b2_persrc) the per-source blocks [src][zl][a][slot][x] of the caller-run all-to-all
    long long b2_zstride;
    int b2_persrc;
};

// which rank owns row y, and in which of its 2h slots
ZPLT_HD void slab_owner(int N, int G, int y, int &rank, int &slot) {
    const int h = N / (2 * G), half = N / 2;
    if (y < half) {
        rank = y % G;
        slot = y / G;
    } else if (y == half) {
        rank = 0;
        slot = h;
    } else {
        const int yp = N - y;  // 1 .. N/2-1
        rank = yp % G;
        slot = h + yp / G;
    }
}

// the row a slot of a rank holds
ZPLT_HD int slab_row(int N, int G, int rank, int slot) {
    const int h = N / (2 * G);
    if (slot < h) return slot * G + rank;
    if (rank == 0 && slot == h) return N / 2;
    return N - ((slot - h) * G + rank);
}

// element offsets (in complex numbers) of the start of an x-row
ZPLT_HD long long slab_b1_row(const SlabGeom &s, int a, int z, int slot) {
    return (((long long) z * s.na + a) * (2 * s.h) + slot) * (long long) s.N;
}
ZPLT_HD long long slab_b2_row(const SlabGeom &s, int a, int zl, int y) {
    // slab_owner() with the divisions by G done as shifts
    const int half = s.N / 2, gm = s.G - 1;
    int src, slot;
    if (y < half) {
        src = y & gm, slot = y >> s.log2G;
    } else if (y == half) {
        src = 0, slot = s.h;
    } else {
        const int yp = s.N - y;
        src = yp & gm, slot = s.h + (yp >> s.log2G);
    }
    return ((((long long) src * (s.N / s.G) + zl) * s.na + a) * (2 * s.h) + slot) * (long long) s.N;
}
// complex elements every rank sends to every other rank
ZPLT_HD long long slab_block_elems(const SlabGeom &s) { return (long long) (s.N / s.G) * s.na * (2 * s.h) * s.N; }

}  // namespace zplt
