// FFT tile kernels: in-place strided/row backward FFTs (z and y axes) and the fused
// x-axis FFT + particle-record emission epilogue.
//
//   fft_tile_kernel  replaces InverseFFT_Yonly / the column half of Inverse2dFFT
//                    (reference src/zeldovich.cpp:93-114, :88-92 as used at :508-511, :653-658)
//   fft_emit_kernel  replaces the row half of Inverse2dFFT plus WriteParticlesSlab
//                    (reference src/output.cpp:41-234): unpack Re/Im of the packed arrays
//                    into displacement/velocity, cast to the ICFormat record, accumulate
//                    density_variance and max_disp.
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

// measured in the generation kernel: table-loaded twiddles 35.4 ms vs multiplication tree 34.5 ms
#define ZPLT_GENX_TWLOAD false
#include "zplt_fft.cuh"
#include "zplt_internal.h"
#include "zplt_kernel_util.cuh"

namespace zplt {

// One CTA per tile (x tile fastest, so that CTAs running at the same time cover neighbouring 128-byte runs).  This is the
// general form — every length, row tiles, slab geometries; the unit-stride passes of the large sizes use the
// ring-prefetched kernel below.  Measured on this kernel at PPD=1024 before the ring existed (z pass 34.6 ms): persistent
// CTAs walking tiles 36.0 ms (the hardware CTA scheduler balances better); prefetch.global.L2 of the next tile during the
// transform 44.9 ms (1024 distinct 2 MB pages per tile: the extra translations cost more than the prefetch saves); the bare
// load/store pattern without the transform 23.8 ms (5.8 TB/s, 89 % of the measured copy peak) with 128-byte runs and
// 46.5 ms with 64-byte runs; a cp.async landing buffer for the next tile (12/8/4 of 16 elements: 36.7/34.9/32.4 ms).
template <int N, int T>
__global__ void __launch_bounds__(T *(N / 16), min_ctas(T *(N / 16))) fft_tile_kernel(cplx *__restrict__ data, TileGeom g, const cplx *__restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *S         = reinterpret_cast<cplx *>(smem_raw);
    constexpr int M = N / 16;
    const int tid = threadIdx.x, p = tid % T, b = tid / T;
    const long long t  = blockIdx.x;
    const long long tx = t % g.grid_x, ty = (t / g.grid_x) % g.grid_y, tz = t / ((long long) g.grid_x * g.grid_y);
    const long long base = tz * g.astride + ty * g.ostride + tx * g.tstride + (long long) (p % g.pa) * g.plo_stride
                           + (long long) (p / g.pa) * g.phi_stride;
    cplx v[16];
#pragma unroll
    for (int e = 0; e < 16; e++) v[e] = ld_stream(&data[base + (long long) (b + M * e) * g.nstride]);
    const int bo = fft_pencil<N, T>(v, S + p * FftSmem<N, T>::PSTRIDE, b, tw);
#pragma unroll
    for (int e = 0; e < 16; e++) st_stream(&data[base + (long long) (bo + M * e) * g.nstride], v[e]);
}

// ------------------------------------------------------------------ ring-prefetched strided pass
// Same transform as fft_tile_kernel, organised so that the SM always has loads in flight: the register
// file holds exactly one tile (512 threads x 16 complex) and shared memory holds its exchange image, so a
// second resident tile is impossible — but the 96 KB of shared memory next to the exchange image can
// receive KP of the next tile's 16 slices (slice e = rows b + M*e, one 128-byte run per row) while the
// current tile is transformed and stored.  The copies are cp.async.bulk (the TMA unit's 1-D form, one
// per row, completion counted in bytes on an mbarrier), so they cost no registers and no LSU issue
// slots; the remaining 16 - KP slices are ordinary loads at the top of the iteration.  CTAs are
// persistent and take tiles from a device counter (the hardware scheduler's balance, kept).
template <int N, int T>
struct RingSmem {
    static constexpr size_t EXCHANGE = ((size_t) T * FftSmem<N, T>::PSTRIDE * sizeof(cplx) + 127) / 128 * 128;
};

template <int N, int T, int KP>
__global__ void __launch_bounds__(T *(N / 16), 1)
   fft_tile_ring_kernel(cplx *__restrict__ data, TileGeom g, const cplx *__restrict__ tw, unsigned int *__restrict__ counter,
                        const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char smem_ring[];
    constexpr int M  = N / 16;
    static_assert(KP >= 1 && KP <= 16, "slices");
    cplx *S            = reinterpret_cast<cplx *>(smem_ring);
    cplx *L            = reinterpret_cast<cplx *>(smem_ring + RingSmem<N, T>::EXCHANGE);  // [KP][M][T]
    uint64_t *mbar     = reinterpret_cast<uint64_t *>(smem_ring + RingSmem<N, T>::EXCHANGE + (size_t) KP * M * T * sizeof(cplx));
    unsigned int *s_nn = reinterpret_cast<unsigned int *>(mbar + 1);
    const int tid = threadIdx.x, p = tid % T, b = tid / T;
    const unsigned int ntiles = (unsigned int) (g.grid_x * g.grid_y * g.grid_z);
    auto tile_base = [&](unsigned int t) {
        const long long tx = t % g.grid_x, ty = (t / g.grid_x) % g.grid_y, tz = t / ((unsigned int) g.grid_x * g.grid_y);
        return tz * g.astride + ty * g.ostride + tx * g.tstride;
    };
    auto issue_ring = [&](unsigned int t) {
        if (tid == 0) {
            mbar_expect_tx(mbar, (unsigned) (KP * M * T * sizeof(cplx)));
            const int tx = t % g.grid_x, ty = (t / g.grid_x) % g.grid_y, tz = t / ((unsigned int) g.grid_x * g.grid_y);
#pragma unroll
            for (int e = 0; e < KP; e++) tma_load_4d(L + (size_t) e * M * T, &tmap, tx * 2 * T, ty, M * e, tz, mbar);
        }
    };
    if (tid == 0) {
        mbar_init(mbar, 1);
        s_nn[0] = atomicAdd(counter, 1u);
        s_nn[1] = atomicAdd(counter, 1u);
    }
    __syncthreads();
    unsigned int cur = s_nn[0], nxt = s_nn[1];
    __syncthreads();
    if (cur < ntiles) issue_ring(cur);
    unsigned int parity = 0;
    while (cur < ntiles) {
        if (tid == 0) s_nn[0] = atomicAdd(counter, 1u);  // the tile after next
        const long long base = tile_base(cur) + p;
        cplx v[16];
#pragma unroll
        for (int e = KP; e < 16; e++) v[e] = ld_stream(&data[base + (long long) (b + M * e) * g.nstride]);
        mbar_wait(mbar, parity);
        parity ^= 1u;
#pragma unroll
        for (int e = 0; e < KP; e++) v[e] = L[(size_t) (e * M + b) * T + p];
        __syncthreads();  // the ring has been read and the exchange image of the previous tile is no longer in use
        const unsigned int nn = s_nn[0];
        if (nxt < ntiles) issue_ring(nxt);
        const int bo = fft_pencil<N, T>(v, S + p * FftSmem<N, T>::PSTRIDE, b, tw);
#pragma unroll
        for (int e = 0; e < 16; e++) st_stream(&data[base + (long long) (bo + M * e) * g.nstride], v[e]);
        cur = nxt;
        nxt = nn;
    }
}

// ------------------------------------------------------------------ generation + x FFT
// Fused mode generation and x-axis FFT (the first of the three axes; the transform is
// separable so the axis order is free).  One CTA owns R row pairs: row (y, z) of primary
// modes and its conjugate-structured twin row (N-y, N-z) (reference
// src/zeldovich.cpp:447-466).  Phase 1: the CTA draws the R*N primary modes of its rows once
// and keeps their state (D, s0, s1, s2, f) in shared memory.  Phase 2: the na*2*R pencils
// (na packed arrays x {primary, twin} x R rows) are formed from that state, transformed
// in registers, and written as whole contiguous rows.  Nothing is read from HBM except
// the RNG/P(k)/eigenmode tables (L2-resident).
//   y == 0   : the plane is its own twin (reference :485-503): rows z <= N/2 are canonical;
//              row (0,0) mixes primary (x <= N/2) and twin (x > N/2) entries, origin = 0
//   y == N/2 : the Nyquist row is zero (reference :640-650 with src/block_array.cpp:487-491)
template <int N, int NP>
__global__ void __launch_bounds__(NP *(N / 16), min_ctas(NP *(N / 16)))
   gen_xfft_kernel(GenParams g, SlabGeom sg, cplx *__restrict__ cube, const cplx *__restrict__ tw, int skip_fft) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int M  = N / 16;
    constexpr int NT = NP * M;
    cplx *S          = reinterpret_cast<cplx *>(smem_raw);
    double *state    = reinterpret_cast<double *>(smem_raw + (size_t) NP * FftSmem<N, NP>::PSTRIDE * sizeof(cplx));
    const int na = g.na;
    const int tid = threadIdx.x, p = tid % NP, b = tid / NP;
    constexpr int half = N / 2;
    // single GPU: y = 0 .. N/2.  Slab rank: its h primary rows y = slot*G + rank, plus the Nyquist row on rank 0.
    // (in groups: blockIdx.y counts the group's nly rows; one extra block on rank 0's first group is the Nyquist row)
    const int y = (sg.G == 1) ? (int) blockIdx.y
                              : ((int) blockIdx.y < sg.nly ? (sg.ly0 + (int) blockIdx.y) * sg.G + sg.rank : half);
    const int z = blockIdx.x;
    if (y == 0 && z > half) return;  // produced as the twin row of (0, N-z)

    const bool origin_row = (y == 0 && z == 0);
    const bool has_twin   = (y > 0 && y < half) || (y == 0 && z > 0 && z < half);
    const int zh = (N - z) % N, yh = (N - y) % N;
    auto row_of = [&](int a, int side) -> long long {
        if (sg.G == 1)
            return (side == 0) ? ((long long) a * N + z) * N * (long long) N + (long long) y * N
                               : ((long long) a * N + zh) * N * (long long) N + (long long) yh * N;
        const int ly = (y == half) ? sg.h : (y >> sg.log2G);  // slot of the primary row (cyclic ownership, zplt_slab.h)
        return (side == 0) ? slab_b1_row(sg, a, z, ly) : slab_b1_row(sg, a, zh, (y == 0) ? 0 : sg.h + ly);
    };
    constexpr int RUN = N / NT;  // = 16 / NP
    static_assert(RUN * NT == N && (RUN == 2 || RUN == 4 || RUN == 8), "whole runs");

    // the row's generator state and this thread's jump along x: dependent table loads (L2 latency) and two 128-bit
    // multiply-adds, issued before the mask test so that they overlap with it
    RowConst rc;
    Affine xj0;
    if (y < half) {
        xj0 = g.xjump[tid * RUN];
        rc  = row_const(g, y, z);
    }

    // ---- rows without a single unmasked mode (outside the k_cutoff sphere: 1 - pi/4 of all rows; the Nyquist
    //      row) are zero in every packed array, and so are their transforms: store zeros, skip everything else ----
    {
        bool any = (g.phi != nullptr) && y < half;  // ZD_f_NL: the density comes from the potential, nothing is masked
        if (y < half && !any) {
            const int kz = wrap_k(z, N, half);
#pragma unroll
            for (int j = 0; j < RUN; j++) {
                const int kx = wrap_k(tid * RUN + j, N, half);
                any |= !mode_masked(g, kx, y, kz, kx * kx + y * y + kz * kz);
            }
        }
        // (skipping the all-masked 64-element slices of the other rows in the pencil builder was tried: the 16 extra
        // predicates cost more than the skipped loads save, generation + x pass 31.5 -> 33.7 ms)
        if (!__syncthreads_or(any)) {
            for (int P = 0; P < 2 * na; P++) {
                const int a = P % na, side = P / na;
                if (side == 1 && !has_twin) continue;
                const long long row = row_of(a, side);
                for (int i = tid; i < N; i += NT) st_stream(&cube[row + i], make_double2(0.0, 0.0));
            }
            return;
        }
    }

    // ---- phase 1: draw the primary modes of row (y, z) once, RUN consecutive x per thread ----
    {
        double q[6][RUN];
#pragma unroll
        for (int c = 0; c < 6; c++)
#pragma unroll
            for (int j = 0; j < RUN; j++) q[c][j] = 0.0;
        if (y < half) primary_run<RUN>(g, rc, tid * RUN, xj0, q[0], q[1], q[2], q[3], q[4], q[5]);
#pragma unroll
        for (int c = 0; c < 6; c++)
#pragma unroll
            for (int j = 0; j < RUN; j += 2)
                *reinterpret_cast<double2 *>(&state[c * N + tid * RUN + j]) = make_double2(q[c][j], q[c][j + 1]);
    }
    __syncthreads();

    // ---- phase 2: the 2*na pencils (na arrays x {row, twin row}), NP at a time ----
    const int rounds = 2 * na / NP;
    for (int rnd = 0; rnd < rounds; rnd++) {
        const int P = rnd * NP + p, a = P % na, side = P / na;
        // Every packed entry is a fixed real-linear form of (Dr, Di) (reference
        // src/zeldovich.cpp:432-434, :447-466 with F,G,H = i s_c D):
        //   re =      base*Dr - sgn*(c1*Dr)*F - (c2*Di)*F
        //   im = sgn* base*Di -     (c1*Di)*F + sgn*(c2*Dr)*F
        // A0: base=1, c1=s0, c2=0, F=1 | A1: base=0, c1=s2, c2=s1, F=1 | A2 / A3: the same with
        // base=0 and F=f; sgn = -1 for the conjugate-structured twin.  The selectors are
        // per-thread constants, so the element loop is branch-free.
        const bool odd     = a & 1;
        const double base  = (a == 0) ? 1.0 : 0.0;
        const double *pc1  = state + (odd ? 4 : 2) * N;
        const double *pc2  = state + 3 * N;
        const double *pF   = state + 5 * N;
        const bool useF    = a >= 2;
        cplx v[16];
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int x = b + M * e;
            int mx;
            bool twin;
            if (side == 0) {
                twin = origin_row && x > half;  // second half of the (0,0) row comes from the twin
                mx   = twin ? N - x : x;
            } else {
                twin = true;
                mx   = (N - x) % N;
            }
            const double sgn = twin ? -1.0 : 1.0;
            const double Dr = state[mx], Di = state[N + mx];
            const double c1 = pc1[mx];
            const double c2 = odd ? pc2[mx] : 0.0;
            const double F  = useF ? pF[mx] : 1.0;
            cplx val;
            val.x = base * Dr - sgn * ((c1 * Dr) * F) - (c2 * Di) * F;
            val.y = sgn * (base * Di) - (c1 * Di) * F + sgn * ((c2 * Dr) * F);
            if (origin_row && side == 0 && x == 0) val = make_double2(0.0, 0.0);
            v[e] = val;
        }
        // natural order: whole rows are stored.  skip_fft (CTA-uniform): the packed arrays before the transform (introspection)
        const int bo = skip_fft ? b : fft_pencil<N, NP, false, ZPLT_GENX_TWLOAD>(v, S + p * FftSmem<N, NP>::PSTRIDE, b, tw);
        const bool live     = (side == 0) || has_twin;
        const long long row = row_of(a, side);
        if (live) {
#pragma unroll
            for (int e = 0; e < 16; e++) st_stream(&cube[row + bo + M * e], v[e]);
        }
        if (rnd + 1 < rounds) __syncthreads();  // the exchange buffer is reused by the next round
    }
}

// z-axis pass of a slab rank with the exchange fused in: the transformed pencil is not written
// back to the local stage-1 buffer but straight into the stage-2 buffers of the ranks that own
// its planes — peer stores over NVLink (peer[r] is rank r's stage-2 base, opened through CUDA
// IPC; peer[rank] is local; a NULL entry discards that rank's share: single-GPU tests of one rank).
// This is BlockArray::StoreBlock + LoadBlock (reference src/block_array.cpp:387-414, 466-504) done by
// the FFT epilogue, 128-byte runs.  The receiver's layout is free because the sender computes every
// address (SlabGeom::b2_persrc / b2_zstride): rows at their true y, B2[zl][a][y][x] — stage 2 of a slab rank then
// reads exactly what a single GPU reads (the y shift of LoadBlock, :487-491, is the slot -> y map here) — or
// per-source blocks B2[src][zl][a][slot][x], which sustain ~25 % more NVLink bandwidth on 4 and 8 ranks at N <= 1024
// (DESIGN.md section 6).
struct PeerTable {
    cplx *recv[16];
};
// tile t of the stage-1 buffer -> x tile, packed array, slot, row group.  Tiles are numbered group by group (sg.nly primary
// rows per group, starting at row sg.ly0), so that a launch covering several groups meets them in the order they are generated.
__device__ __forceinline__ void p2p_tile(const SlabGeom &sg, int XT, long long t, int &xt, int &a, int &slot, int &grp) {
    const int nsl       = 2 * sg.nly;
    const long long tpg = (long long) XT * nsl * sg.na;
    grp                 = (int) (t / tpg);
    const long long tt  = t - grp * tpg;
    xt                  = (int) (tt % XT);
    const int rr = (int) (tt / XT), sidx = rr % nsl;
    a             = rr / nsl;
    const int ly0 = sg.ly0 + grp * sg.nly;
    slot          = sidx < sg.nly ? ly0 + sidx : sg.h + ly0 + (sidx - sg.nly);
}
template <int N, int T>
__global__ void __launch_bounds__(T *(N / 16), min_ctas(T *(N / 16)))
   fft_tile_p2p_kernel(const cplx *__restrict__ b1, SlabGeom sg, const __grid_constant__ PeerTable peers, const cplx *__restrict__ tw,
                       GroupSync gs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_ok;
    cplx *S         = reinterpret_cast<cplx *>(smem_raw);
    constexpr int M = N / 16;
    const int tid = threadIdx.x, p = tid % T, b = tid / T;
    const int rows = sg.na * 2 * sg.h;  // x-rows per z plane of the stage-1 buffer
    const long long nstride = (long long) rows * N;
    const int np = N / sg.G, lognp = FftLog2<N>::value - sg.log2G;
    // persistent CTAs over the tiles (x tile, array, slot): the pass is NVLink-bound, so a limited number of CTAs
    // saturates the links and leaves the other SMs to the generation kernels, which run concurrently on another stream
    constexpr int XT = N / T;
    const long long ntiles = (long long) XT * 2 * sg.nly * sg.na * gs.J;
    int ready = gs.flags == nullptr ? gs.J : -1;  // groups known to be generated
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int xt, a, slot, grp;
        p2p_tile(sg, XT, t, xt, a, slot, grp);
        if (grp > ready) {  // CTA-uniform
            if (tid == 0) s_ok = wait_group(gs.flags, grp) ? 1 : 0;
            __syncthreads();
            if (!s_ok) {
                if (tid == 0) atomicExch(gs.err, 1u);
                return;
            }
            ready = grp;
        }
        const int row  = a * 2 * sg.h + slot;
        const int x    = xt * T + p;
        const int y    = slab_row(N, sg.G, sg.rank, slot);
        const long long base = (long long) row * N + x;
        // where the rows of this tile land in an owner's buffer: plane zl at zl*zstride + rowoff
        const long long zstride = sg.b2_persrc ? (long long) rows * N : sg.b2_zstride;
        const long long rowoff  = sg.b2_persrc ? ((long long) sg.rank * np * rows + row) * N + x : ((long long) a * N + y) * N + x;
        cplx v[16];
#pragma unroll
        for (int e = 0; e < 16; e++) v[e] = ld_stream(&b1[base + (long long) (b + M * e) * nstride]);
        const int bo = fft_pencil<N, T>(v, S + p * FftSmem<N, T>::PSTRIDE, b, tw);
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int z = bo + M * e;
            cplx *dst = peers.recv[z >> lognp];
            if (dst != nullptr) st_stream(&dst[(long long) (z & (np - 1)) * zstride + rowoff], v[e]);
        }
        __syncthreads();  // the exchange image is reused by the next tile
    }
}

// The same pass with the TMA ring of fft_tile_ring_kernel on the load side: while one tile is transformed and sent,
// KP of the 16 slices of the CTA's next tile land in shared memory, so the links never wait for the local HBM reads
// (the plain kernel alternates: load, transform, store).  Tiles come from a device counter.  With GroupSync::flags the
// launch covers every row group of stage 1: the CTAs hold their SMs from the start (a static split of the SMs between this
// pass and the generation kernels — CTAs launched per group had to win whole SMs back from two-per-SM generation CTAs and
// mostly ran after them) and wait, one tile ahead, for the group of the next tile to be generated.
template <int N, int T, int KP>
__global__ void __launch_bounds__(T *(N / 16), 1)
   fft_tile_p2p_ring_kernel(const cplx *__restrict__ b1, SlabGeom sg, const __grid_constant__ PeerTable peers, const cplx *__restrict__ tw,
                            unsigned int *__restrict__ counter, const __grid_constant__ CUtensorMap tmap, GroupSync gs) {
    extern __shared__ __align__(128) unsigned char smem_ring[];
    constexpr int M  = N / 16;
    constexpr int XT = N / T;
    cplx *S            = reinterpret_cast<cplx *>(smem_ring);
    cplx *L            = reinterpret_cast<cplx *>(smem_ring + RingSmem<N, T>::EXCHANGE);  // [KP][M][T]
    uint64_t *mbar     = reinterpret_cast<uint64_t *>(smem_ring + RingSmem<N, T>::EXCHANGE + (size_t) KP * M * T * sizeof(cplx));
    unsigned int *s_nn = reinterpret_cast<unsigned int *>(mbar + 1);  // [0], [1]: tiles handed out; [2]: a wait timed out
    const int tid = threadIdx.x, p = tid % T, b = tid / T;
    const int rows = sg.na * 2 * sg.h;
    const long long nstride = (long long) rows * N;
    const int np = N / sg.G, lognp = FftLog2<N>::value - sg.log2G;
    const unsigned int tpg = (unsigned int) (XT * 2 * sg.nly * sg.na), ntiles = tpg * (unsigned int) gs.J;
    int ready = gs.flags == nullptr ? gs.J : -1;  // thread 0: groups known to be generated
    auto ensure = [&](unsigned int t) {           // thread 0: the rows of tile t exist
        if (t < ntiles) {
            const int grp = (int) (t / tpg);
            if (grp > ready) {
                if (!wait_group(gs.flags, grp)) s_nn[2] = 1u;
                ready = grp;
            }
        }
    };
    auto issue_ring = [&](unsigned int t) {
        if (tid == 0) {
            int xt, a, slot, grp;
            p2p_tile(sg, XT, t, xt, a, slot, grp);
            mbar_expect_tx(mbar, (unsigned) (KP * M * T * sizeof(cplx)));
#pragma unroll
            for (int e = 0; e < KP; e++) tma_load_4d(L + (size_t) e * M * T, &tmap, xt * 2 * T, a * 2 * sg.h + slot, M * e, 0, mbar);
        }
    };
    if (tid == 0) {
        mbar_init(mbar, 1);
        s_nn[0] = atomicAdd(counter, 1u);
        s_nn[1] = atomicAdd(counter, 1u);
        s_nn[2] = 0u;
        ensure(s_nn[0]);
    }
    __syncthreads();
    unsigned int cur = s_nn[0], nxt = s_nn[1];
    if (s_nn[2]) {
        if (tid == 0) atomicExch(gs.err, 1u);
        return;
    }
    __syncthreads();
    if (cur < ntiles) issue_ring(cur);
    unsigned int parity = 0;
    while (cur < ntiles) {
        if (tid == 0) {
            s_nn[0] = atomicAdd(counter, 1u);  // the tile after next
            ensure(nxt);                       // before anybody passes this iteration's barrier: its ring and plain loads come after it
        }
        int xt, a, slot, grp;
        p2p_tile(sg, XT, cur, xt, a, slot, grp);
        const int row = a * 2 * sg.h + slot;
        const int x   = xt * T + p;
        const int y   = slab_row(N, sg.G, sg.rank, slot);
        const long long base = (long long) row * N + x;
        // where the rows of this tile land in an owner's buffer: plane zl at zl*zstride + rowoff
        const long long zstride = sg.b2_persrc ? (long long) rows * N : sg.b2_zstride;
        const long long rowoff  = sg.b2_persrc ? ((long long) sg.rank * np * rows + row) * N + x : ((long long) a * N + y) * N + x;
        cplx v[16];
#pragma unroll
        for (int e = KP; e < 16; e++) v[e] = ld_stream(&b1[base + (long long) (b + M * e) * nstride]);
        mbar_wait(mbar, parity);
        parity ^= 1u;
#pragma unroll
        for (int e = 0; e < KP; e++) v[e] = L[(size_t) (e * M + b) * T + p];
        __syncthreads();  // the ring has been read and the exchange image of the previous tile is no longer in use
        const unsigned int nn = s_nn[0];
        if (s_nn[2]) {  // CTA-uniform: nothing is in flight here (the ring of `cur` is consumed, that of `nxt` not yet requested)
            if (tid == 0) atomicExch(gs.err, 1u);
            return;
        }
        if (nxt < ntiles) issue_ring(nxt);
        const int bo = fft_pencil<N, T>(v, S + p * FftSmem<N, T>::PSTRIDE, b, tw);
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int z = bo + M * e;
            cplx *dst = peers.recv[z >> lognp];
            if (dst != nullptr) st_stream(&dst[(long long) (z & (np - 1)) * zstride + rowoff], v[e]);
        }
        cur = nxt;
        nxt = nn;
    }
}

// byte offsets inside one record, per ICFormat (reference include/output.h:19-42)
struct RecLayout {
    int off_ijk;   // -1: no ids
    int off_d[3];  // displ[0..2]
    int off_v[3];  // vel[0..2], -1: none
    int dbl;       // fields are double (else float)
};
__device__ __forceinline__ RecLayout rec_layout(int fmt) {
    RecLayout L;
    switch (fmt) {
        case 0: L = {0, {8, 16, 24}, {-1, -1, -1}, 1}; break;  // Zeldovich
        case 1: L = {0, {8, 12, 16}, {20, 24, 28}, 0}; break;  // RVZel
        case 2: L = {0, {8, 16, 24}, {32, 40, 48}, 1}; break;  // RVdoubleZel
        default: L = {-1, {0, 4, 8}, {-1, -1, -1}, 0}; break;  // ZelSimple
    }
    return L;
}
__device__ __forceinline__ void put(unsigned char *rec, int off, double val, int dbl) {
    if (off < 0) return;
    if (dbl)
        *reinterpret_cast<double *>(rec + off) = val;
    else
        *reinterpret_cast<float *>(rec + off) = (float) val;
}

// One packed array A of the tile: load, transform along y, then either park its values or
// complete record fields.  A is a compile-time constant, the format test is CTA-uniform and
// sits outside the element loops.
template <int N, int T, int A, bool RVZEL, bool ACC>
__device__ __forceinline__ void emit_finish(cplx (&v)[16], int zl, cplx *S, float *keep, const cplx *__restrict__ tw,
                                            const EmitParams &ep, const RecLayout &L, unsigned char *rec0, long long z, int x,
                                            int tid, int p, int b, double (*s_red)[8], unsigned pslot);

template <int N, int T, int A, bool SLAB, bool RVZEL>
__device__ __forceinline__ void emit_array(const cplx *__restrict__ src, const cplx *__restrict__ next, const SlabGeom &sg, int zl,
                                           cplx *S, float *keep,
                                           const cplx *__restrict__ tw, const EmitParams &ep, const RecLayout &L,
                                           unsigned char *rec0, long long z, int x,
                                           int tid, int p, int b, bool first, double (*s_red)[8], unsigned pslot) {
    constexpr int M = N / 16;
    cplx v[16];
#pragma unroll
    for (int e = 0; e < 16; e++) {
        // single GPU: rows of the [a][z][y][x] cube; slab rank: rows of the exchanged buffer B2
        const long long off = SLAB ? slab_b2_row(sg, A, zl, b + M * e) : (long long) (b + M * e) * N;
        v[e] = ld_stream(&src[off]);
    }
    if (!SLAB && next != nullptr && (p & 7) == 0) {
        // pull the next array's rows of this tile into L2 while this one is transformed (the y stride
        // stays inside a few 2 MB pages, unlike the z pass where this prefetch costs more than it gains)
#pragma unroll
        for (int e = 0; e < 16; e++) {
            prefetch_l2(&next[(long long) (b + M * e) * N]);
        }
    }
    if (!first) __syncthreads();  // the previous array's last exchange read is complete
    emit_finish<N, T, A, RVZEL, false>(v, zl, S, keep, tw, ep, L, rec0, z, x, tid, p, b, s_red, pslot);
}

// Transform one packed array of the tile along y (v[e] = row b + M*e on entry), then either park its values or
// complete record fields.  `keep` is 64 KB of parking space private to the CTA: shared memory in the
// one-tile-per-CTA kernel, an L2-resident global area in the persistent ring kernel.
template <int N, int T, int A, bool RVZEL, bool ACC>
__device__ __forceinline__ void emit_finish(cplx (&v)[16], int zl, cplx *S, float *keep, const cplx *__restrict__ tw,
                                            const EmitParams &ep, const RecLayout &L, unsigned char *rec0, long long z, int x,
                                            int tid, int p, int b, double (*s_red)[8], unsigned pslot) {
    constexpr int M  = N / 16;
    constexpr int NT = T * M;
    const int rb = ep.record_bytes, dbl = L.dbl;
    constexpr bool rvzel = RVZEL;
    const bool qplt = ep.qPLT;
    const double vn = ep.vnorm;
    b = fft_pencil<N, T>(v, S + p * FftSmem<N, T>::PSTRIDE, b, tw);  // from here on b is the OUTPUT slot
    const uint64_t pol = l2_evict_last();
    // default (wide_records < 0): on (measured in the persistent ring kernel: y pass 28.2 -> 24.9 ms at PPD=1024)
    const bool wide    = ep.wide_records != 0 && (reinterpret_cast<size_t>(ep.out) & 31) == 0;
    double *keepd = reinterpret_cast<double *>(keep);  // the same 64 KB seen as [16][NT] doubles (non-RVZel formats)
    if (A == 0) {  // Re = density, Im = pos[0] -> displ[2]: parked until A1 completes the displacement
        {
            double var = 0.0, mp = 0.0, mn = 0.0;
#pragma unroll
            for (int e = 0; e < 16; e++) {
                var += v[e].x * v[e].x;
                mp = fmax(mp, v[e].y), mn = fmax(mn, -v[e].y);
            }
            fold_stats<NT, ACC>(s_red, tid, var, 0, mp, 1, mn, 4);
        }
        if (ep.dens != nullptr) {  // density planes (reference src/output.cpp:196,217-224): float(Re A0)
            float *dp = ep.dens + ((size_t) (zl - ep.z0) * N) * N + x;
#pragma unroll
            for (int e = 0; e < 16; e++) dp[(size_t) (b + M * e) * N] = (float) v[e].x;
        }
        if (rvzel) {
#pragma unroll
            for (int e = 0; e < 16; e++) {
                park_st<ACC>(&keep[(0 * 16 + e) * NT + tid], (float) v[e].y, pol);
                if (!qplt) park_st<ACC>(&keep[(1 * 16 + e) * NT + tid], (float) (v[e].y * vn), pol);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 16; e++) keepd[e * NT + tid] = v[e].y;
        }
    } else if (A == 2) {  // Im = vel[0] -> vel[2]: parked until A3 completes the velocity
        if (rvzel) {
#pragma unroll
            for (int e = 0; e < 16; e++) park_st<ACC>(&keep[(1 * 16 + e) * NT + tid], (float) v[e].y, pol);
        } else {
#pragma unroll
            for (int e = 0; e < 16; e++) keepd[e * NT + tid] = v[e].y;
        }
    } else if (A == 1) {  // Re = pos[1] -> displ[1], Im = pos[2] -> displ[0]
        {
            double mp1 = 0.0, mn1 = 0.0, mp2 = 0.0, mn2 = 0.0;
#pragma unroll
            for (int e = 0; e < 16; e++) {
                mp1 = fmax(mp1, v[e].x), mn1 = fmax(mn1, -v[e].x);
                mp2 = fmax(mp2, v[e].y), mn2 = fmax(mn2, -v[e].y);
            }
            fold_stats<NT, ACC>(s_red, tid, mp1, 2, mn1, 5, mp2, 3);
            fold_stats<NT, ACC>(s_red, tid, mn2, 6, 0.0, 7, 0.0, 7);
        }
        if (ep.out == nullptr) {
            // density-only run: nothing to store
        } else if (rvzel) {
            const unsigned int w1 = (unsigned int) (unsigned short) x;
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int y = b + M * e;
                unsigned char *rec = rec0 + (size_t) y * N * rb;
                const unsigned int w0 = (unsigned int) (unsigned short) z | ((unsigned int) (unsigned short) y << 16);
                if (qplt && ep.scratch != nullptr) {
                    // park (displ0, displ1) in this SM's L2-resident scratch: the record is then written whole, as
                    // two back-to-back 16-byte stores, when A3 is done — no partially written 32-byte sectors in L2
                    park_st2(&reinterpret_cast<float2 *>(ep.scratch)[((size_t) pslot * 16 + e) * NT + tid],
                             make_float2((float) v[e].y, (float) v[e].x), pol);
                    continue;
                }
                *reinterpret_cast<float4 *>(rec) = make_float4(__uint_as_float(w0), __uint_as_float(w1), (float) v[e].y, (float) v[e].x);
                if (!qplt)
                    *reinterpret_cast<float4 *>(rec + 16) =
                       make_float4(park_ld<ACC>(&keep[(0 * 16 + e) * NT + tid], pol), (float) (v[e].y * vn), (float) (v[e].x * vn),
                                   park_ld<ACC>(&keep[(1 * 16 + e) * NT + tid], pol));
            }
        } else {
            // ids and the whole displacement (and, without qPLT, the whole velocity) in one burst of
            // back-to-back stores to consecutive bytes, so that L2 sees complete sectors
            double *sd = reinterpret_cast<double *>(ep.scratch);
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int y = b + M * e;
                unsigned char *rec = rec0 + (size_t) y * N * rb;
                const double pos0 = keepd[e * NT + tid];
                if (qplt && sd != nullptr) {
                    // qPLT (RVdoubleZel): park the displacement in this SM's L2-resident scratch, the whole
                    // 56-byte record is written in one burst when A3 is done
                    double *q = sd + (((size_t) pslot * 16 + e) * NT + tid) * 3;
                    q[0] = v[e].y, q[1] = v[e].x, q[2] = pos0;
                    continue;
                }
                if (L.off_ijk >= 0)
                    *reinterpret_cast<ushort4 *>(rec + L.off_ijk) =
                       make_ushort4((unsigned short) z, (unsigned short) y, (unsigned short) x, 0);
                put(rec, L.off_d[0], v[e].y, dbl);
                put(rec, L.off_d[1], v[e].x, dbl);
                put(rec, L.off_d[2], pos0, dbl);
                if (!qplt) {
                    put(rec, L.off_v[0], v[e].y * vn, dbl);
                    put(rec, L.off_v[1], v[e].x * vn, dbl);
                    put(rec, L.off_v[2], pos0 * vn, dbl);
                }
            }
        }
    } else {  // A == 3: Re = vel[1] -> vel[1], Im = vel[2] -> vel[0]
        if (ep.out == nullptr) {
        } else if (rvzel) {
            const unsigned int w1 = (unsigned int) (unsigned short) x;
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int y = b + M * e;
                unsigned char *rec = rec0 + (size_t) y * N * rb;
                if (wide && ep.scratch != nullptr) {
                    const float2 d01 = park_ld2(&reinterpret_cast<const float2 *>(ep.scratch)[((size_t) pslot * 16 + e) * NT + tid], pol);
                    const unsigned int w0 = (unsigned int) (unsigned short) z | ((unsigned int) (unsigned short) y << 16);
                    st_record32(rec, make_float4(__uint_as_float(w0), __uint_as_float(w1), d01.x, d01.y),
                                make_float4(park_ld<ACC>(&keep[(0 * 16 + e) * NT + tid], pol), (float) v[e].y, (float) v[e].x,
                                            park_ld<ACC>(&keep[(1 * 16 + e) * NT + tid], pol)));
                    continue;
                }
                if (ep.scratch != nullptr) {
                    const float2 d01 = park_ld2(&reinterpret_cast<const float2 *>(ep.scratch)[((size_t) pslot * 16 + e) * NT + tid], pol);
                    const unsigned int w0 = (unsigned int) (unsigned short) z | ((unsigned int) (unsigned short) y << 16);
                    *reinterpret_cast<float4 *>(rec) = make_float4(__uint_as_float(w0), __uint_as_float(w1), d01.x, d01.y);
                }
                *reinterpret_cast<float4 *>(rec + 16) =
                   make_float4(park_ld<ACC>(&keep[(0 * 16 + e) * NT + tid], pol), (float) v[e].y, (float) v[e].x,
                               park_ld<ACC>(&keep[(1 * 16 + e) * NT + tid], pol));
            }
        } else {
            const double *sd = reinterpret_cast<const double *>(ep.scratch);
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int y = b + M * e;
                unsigned char *rec = rec0 + (size_t) y * N * rb;
                if (sd != nullptr) {
                    const double *q = sd + (((size_t) pslot * 16 + e) * NT + tid) * 3;
                    if (L.off_ijk >= 0)
                        *reinterpret_cast<ushort4 *>(rec + L.off_ijk) =
                           make_ushort4((unsigned short) z, (unsigned short) y, (unsigned short) x, 0);
                    put(rec, L.off_d[0], q[0], dbl);
                    put(rec, L.off_d[1], q[1], dbl);
                    put(rec, L.off_d[2], q[2], dbl);
                }
                put(rec, L.off_v[0], v[e].y, dbl);
                put(rec, L.off_v[1], v[e].x, dbl);
                put(rec, L.off_v[2], keepd[e * NT + tid], dbl);
            }
        }
    }
}

// RVdoubleZel (BASELINE configs[1]'s format: u16 i,j,k; pad; double displ[3]; double vel[3] = 56 bytes) in the persistent ring
// kernel, with the layout known at compile time — the generic path above carries the record layout in registers and spilled
// (up to 1.2 KB) when it was inlined four times into the ring kernel.  Four parked doubles per site, all L2-resident and
// coalesced ([16][NT] planes): pos[0] in `keepd`, then displ[0], displ[1] and vel[2] in `park3`; A3 writes the record whole.
// One 56-byte record per lane, 8 lanes = 8 consecutive x = 448 contiguous bytes.  Written field by field that is seven 8-byte
// stores at a 56-byte stride — four times the sector requests the bytes need, which is what bounded this format (y pass 51.6 ms
// against 24.9 ms for RVZel at PPD=1024).  With STAGE the 8 lanes assemble their records in a piece of the exchange image that
// belongs to their warp alone after the transform (3-pass lengths: passes 2 and 3 are warp-local, zplt_fft.cuh) and write the
// 448 bytes as 28 coalesced 16-byte chunks.
template <bool STAGE>
__device__ __forceinline__ void store_record56(unsigned char *rec, double *stage, int p, unsigned long long ids, double f0, double f1,
                                               double f2, double f3, double f4, double f5) {
    if constexpr (STAGE) {
        double *st = stage + p * 7;
        st[0] = __longlong_as_double((long long) ids);
        st[1] = f0, st[2] = f1, st[3] = f2, st[4] = f3, st[5] = f4, st[6] = f5;
        __syncwarp();
        const double2 *sv = reinterpret_cast<const double2 *>(stage);
        double2 *g        = reinterpret_cast<double2 *>(rec - p * 56);
        __stcs(&g[p], sv[p]);
        __stcs(&g[p + 8], sv[p + 8]);
        __stcs(&g[p + 16], sv[p + 16]);
        if (p < 4) __stcs(&g[p + 24], sv[p + 24]);
        __syncwarp();
    } else {
        *reinterpret_cast<unsigned long long *>(rec) = ids;
        double *f = reinterpret_cast<double *>(rec + 8);
        f[0] = f0, f[1] = f1, f[2] = f2, f[3] = f3, f[4] = f4, f[5] = f5;
    }
}

template <int N, int T, int A>
__device__ __forceinline__ void emit_finish_rvdouble(cplx (&v)[16], int zl, cplx *S, double *keepd, double *park3, const cplx *__restrict__ tw,
                                                     const EmitParams &ep, unsigned char *rec0, long long z, int x, int tid, int p, int b,
                                                     double (*s_red)[8]) {
    constexpr int M  = N / 16;
    constexpr int NT = T * M;
    // staging needs a warp-private piece of the exchange image (3-pass lengths), 8-lane groups of consecutive x and aligned records
    constexpr bool CAN_STAGE = FftPlan<N>::PASSES == 3 && T == 8;
    constexpr int R3         = CAN_STAGE ? FftPlan<N>::R3 : 1;
    double *stage = reinterpret_cast<double *>(S + (b % R3) * FftSmem<N, T>::PSTRIDE + (b / R3) * (16 * R3));  // 256*R3 bytes >= 448
    const bool aligned = (reinterpret_cast<size_t>(ep.out) & 15) == 0;
    b = fft_pencil<N, T>(v, S + p * FftSmem<N, T>::PSTRIDE, b, tw);  // from here on b is the OUTPUT slot
    if (CAN_STAGE) __syncwarp();  // every lane of the warp has done its last read of the exchange image
    const uint64_t pol = l2_evict_last();
    const bool qplt = ep.qPLT;
    if (A == 0) {  // Re = density, Im = pos[0] -> displ[2]
        double var = 0.0, mp = 0.0, mn = 0.0;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            var += v[e].x * v[e].x;
            mp = fmax(mp, v[e].y), mn = fmax(mn, -v[e].y);
        }
        fold_stats<NT, true>(s_red, tid, var, 0, mp, 1, mn, 4);
        if (ep.dens != nullptr) {
            float *dp = ep.dens + ((size_t) (zl - ep.z0) * N) * N + x;
#pragma unroll
            for (int e = 0; e < 16; e++) dp[(size_t) (b + M * e) * N] = (float) v[e].x;
        }
        if (ep.out != nullptr) {
#pragma unroll
            for (int e = 0; e < 16; e++) park_std(&keepd[e * NT + tid], v[e].y, pol);
        }
    } else if (A == 2) {  // Im = vel[0] -> vel[2]
        if (ep.out != nullptr) {
#pragma unroll
            for (int e = 0; e < 16; e++) park_std(&park3[(2 * 16 + e) * NT + tid], v[e].y, pol);
        }
    } else if (A == 1) {  // Re = pos[1] -> displ[1], Im = pos[2] -> displ[0]
        double mp1 = 0.0, mn1 = 0.0, mp2 = 0.0, mn2 = 0.0;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            mp1 = fmax(mp1, v[e].x), mn1 = fmax(mn1, -v[e].x);
            mp2 = fmax(mp2, v[e].y), mn2 = fmax(mn2, -v[e].y);
        }
        fold_stats<NT, true>(s_red, tid, mp1, 2, mn1, 5, mp2, 3);
        fold_stats<NT, true>(s_red, tid, mn2, 6, 0.0, 7, 0.0, 7);
        if (ep.out == nullptr) {
        } else if (qplt) {
#pragma unroll
            for (int e = 0; e < 16; e++) {
                park_std(&park3[(0 * 16 + e) * NT + tid], v[e].y, pol);
                park_std(&park3[(1 * 16 + e) * NT + tid], v[e].x, pol);
            }
        } else {
            const double vn = ep.vnorm;
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int y = b + M * e;
                unsigned char *rec = rec0 + (size_t) y * N * 56;
                const double pos0 = park_ldd(&keepd[e * NT + tid], pol);
                const unsigned long long ids = (unsigned long long) (unsigned short) z | ((unsigned long long) (unsigned short) y << 16)
                                               | ((unsigned long long) (unsigned short) x << 32);
                if (CAN_STAGE && aligned)
                    store_record56<CAN_STAGE>(rec, stage, p, ids, v[e].y, v[e].x, pos0, v[e].y * vn, v[e].x * vn, pos0 * vn);
                else
                    store_record56<false>(rec, stage, p, ids, v[e].y, v[e].x, pos0, v[e].y * vn, v[e].x * vn, pos0 * vn);
            }
        }
    } else if (ep.out != nullptr) {  // A == 3: Re = vel[1], Im = vel[2] -> vel[0]; the record is complete
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int y = b + M * e;
            unsigned char *rec = rec0 + (size_t) y * N * 56;
            const unsigned long long ids = (unsigned long long) (unsigned short) z | ((unsigned long long) (unsigned short) y << 16)
                                           | ((unsigned long long) (unsigned short) x << 32);
            const double d0 = park_ldd(&park3[(0 * 16 + e) * NT + tid], pol), d1 = park_ldd(&park3[(1 * 16 + e) * NT + tid], pol);
            const double d2 = park_ldd(&keepd[e * NT + tid], pol), w2 = park_ldd(&park3[(2 * 16 + e) * NT + tid], pol);
            if (CAN_STAGE && aligned)
                store_record56<CAN_STAGE>(rec, stage, p, ids, d0, d1, d2, v[e].y, v[e].x, w2);
            else
                store_record56<false>(rec, stage, p, ids, d0, d1, d2, v[e].y, v[e].x, w2);
        }
    }
}

template <int N, int T, bool SLAB, bool RVZEL>
__global__ void __launch_bounds__(T *(N / 16), min_ctas(T *(N / 16)))
   fft_emit_strided_kernel(const cplx *__restrict__ cube, SlabGeom sg, long long z_first, EmitParams ep,
                           const cplx *__restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_red[32][8];
    constexpr int M  = N / 16;
    constexpr int NT = T * M;
    cplx *S      = reinterpret_cast<cplx *>(smem_raw);
    float *keep  = reinterpret_cast<float *>(smem_raw + (size_t) T * FftSmem<N, T>::PSTRIDE * sizeof(cplx));  // [2][16][NT]
    const int tid = threadIdx.x, p = tid % T, b = tid / T;
    // z_first counts this rank's planes; the particle id carries the global plane index
    const long long zl = z_first + blockIdx.y;
    const long long z  = SLAB ? (long long) sg.rank * (N / sg.G) + zl : zl + ep.zglobal0;
    const int x        = blockIdx.x * T + p;
    const long long N3 = SLAB ? 0 : ep.astride;
    const cplx *src    = SLAB ? cube + x : cube + zl * ep.zstride + x;  // + B2 row offset  |  + a*astride + y*N
    // parking slot of this CTA in the L2-resident scratch: the SM it runs on, read once (the launcher only hands the
    // scratch out when at most one CTA of the kernel fits an SM)
    const unsigned pslot = smid();
    const RecLayout L  = rec_layout(ep.icformat);
    unsigned char *rec0 = ep.out + ((size_t) ((zl - ep.z0) * N) * N + x) * ep.record_bytes;  // + y*N*rb
    const bool pf = ep.prefetch;
    const cplx *s0 = src, *s1 = src + N3, *s2 = src + 2 * N3, *s3 = src + 3 * N3;
#define ZPLT_EA(A, SRC, NEXT, FIRST) \
    emit_array<N, T, A, SLAB, RVZEL>(SRC, pf ? (NEXT) : nullptr, sg, (int) zl, S, keep, tw, ep, L, rec0, z, x, tid, p, b, FIRST, s_red, pslot)
    if constexpr (RVZEL) {
        // RVZel: A0 and A2 first (two parked floats), then A1 and A3 complete the two 16-byte halves
        if (ep.qPLT) {
            ZPLT_EA(0, s0, s2, true);
            ZPLT_EA(2, s2, s1, false);
            ZPLT_EA(1, s1, s3, false);
            ZPLT_EA(3, s3, nullptr, false);
        } else {
            ZPLT_EA(0, s0, s1, true);
            ZPLT_EA(1, s1, nullptr, false);
        }
    } else {
        // other formats: one parked double at a time (A0 -> A1 writes the displacement, A2 -> A3 the velocity)
        ZPLT_EA(0, s0, s1, true);
        ZPLT_EA(1, s1, ep.qPLT ? s2 : nullptr, false);
        if (ep.qPLT) {
            ZPLT_EA(2, s2, s3, false);
            ZPLT_EA(3, s3, nullptr, false);
        }
    }
#undef ZPLT_EA
    __syncthreads();
    if (tid < 7) {
        constexpr int NW = (NT >= 32) ? NT / 32 : NT;
        double a7 = s_red[0][tid];
        for (int w = 1; w < NW; w++) a7 = (tid == 0) ? a7 + s_red[w][tid] : fmax(a7, s_red[w][tid]);
        double *slot = ep.stats + 8 * ((blockIdx.x + blockIdx.y * gridDim.x) % ZPLT_STAT_SLOTS);
        if (tid == 0)
            atomicAdd(&slot[0], a7);
        else
            atomicMax(reinterpret_cast<unsigned long long *>(&slot[tid]), (unsigned long long) __double_as_longlong(a7));
    }
}

// Persistent, ring-prefetched form of the y pass + emission (single GPU): the CTA walks (tile, array)
// units; while one array of a tile is transformed and its record fields are formed, KP of the 16 slices of
// the next unit — the next array of the tile, or the first array of the CTA's next tile — land in shared
// memory through the TMA unit (see fft_tile_ring_kernel).  The 64 KB of parked values move from shared
// memory to an L2-resident global area to make room for the ring.
#define ZPLT_KEEP_BYTES 65536
template <int N, int T, bool RVZEL, int KP>
__global__ void __launch_bounds__(T *(N / 16), 1)
   fft_emit_ring_kernel(const cplx *__restrict__ cube, long long z_first, long long nz, EmitParams ep, const cplx *__restrict__ tw,
                        unsigned int *__restrict__ counter, float *__restrict__ keep_all, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char smem_ring[];
    __shared__ double s_red[32][8];
    constexpr int M  = N / 16;
    constexpr int NT = T * M;
    constexpr int XT = N / T;
    cplx *S            = reinterpret_cast<cplx *>(smem_ring);
    cplx *L            = reinterpret_cast<cplx *>(smem_ring + RingSmem<N, T>::EXCHANGE);  // [KP][M][T]
    uint64_t *mbar     = reinterpret_cast<uint64_t *>(smem_ring + RingSmem<N, T>::EXCHANGE + (size_t) KP * M * T * sizeof(cplx));
    unsigned int *s_nn = reinterpret_cast<unsigned int *>(mbar + 1);
    // parking areas are indexed by CTA (grid <= SM count <= 256 slots), not by %smid: nothing guarantees one CTA per SM
    float *keep        = keep_all + (size_t) blockIdx.x * (ZPLT_KEEP_BYTES / sizeof(float));
    const unsigned pslot = blockIdx.x;
    double *park3 = reinterpret_cast<double *>(ep.scratch) + (size_t) blockIdx.x * 3 * 16 * NT;  // RVdoubleZel: [3][16][NT] per CTA
    const int tid = threadIdx.x, p = tid % T, b = tid / T;
    const unsigned int ntiles = (unsigned int) (XT * nz);
    const RecLayout Lr = rec_layout(ep.icformat);
    const long long N3 = ep.astride;
    const bool qplt    = ep.qPLT;
    const bool za_order = ep.zstride > ep.astride;  // slab layout [zl][a][y][x]: the tensor map's dimensions are (x, y, a, zl)
    auto issue_ring = [&](unsigned int t, int a) {
        if (tid == 0) {
            mbar_expect_tx(mbar, (unsigned) (KP * M * T * sizeof(cplx)));
            const int tx = t % XT, zl = (int) z_first + (int) (t / XT);
#pragma unroll
            for (int e = 0; e < KP; e++) tma_load_4d(L + (size_t) e * M * T, &tmap, tx * 2 * T, M * e, za_order ? a : zl, za_order ? zl : a, mbar);
        }
    };
    for (int i = tid; i < 32 * 8; i += NT) (&s_red[0][0])[i] = 0.0;
    if (tid == 0) {
        mbar_init(mbar, 1);
        s_nn[0] = atomicAdd(counter, 1u);
        s_nn[1] = atomicAdd(counter, 1u);
    }
    __syncthreads();
    unsigned int cur = s_nn[0], nxt = s_nn[1], nn = 0;
    __syncthreads();
    if (cur < ntiles) issue_ring(cur, 0);
    unsigned int parity = 0;
    while (cur < ntiles) {
        if (tid == 0) s_nn[0] = atomicAdd(counter, 1u);  // the tile after next
        const int x        = (int) (cur % XT) * T + p;
        const long long zl = z_first + cur / XT;
        const cplx *src    = cube + zl * ep.zstride + x;
        unsigned char *rec0 = ep.out + ((size_t) ((zl - ep.z0) * N) * N + x) * ep.record_bytes;
        // one unit: array A of this tile; NEXT_T / NEXT_A name the unit whose slices are requested meanwhile
#define ZPLT_UNIT(A, NEXT_OK, NEXT_T, NEXT_A)                                                                          \
    {                                                                                                                  \
        cplx v[16];                                                                                                    \
        _Pragma("unroll") for (int e = KP; e < 16; e++) v[e] = ld_stream(&src[(long long) (A) * N3 + (long long) (b + M * e) * N]); \
        mbar_wait(mbar, parity);                                                                                       \
        parity ^= 1u;                                                                                                  \
        _Pragma("unroll") for (int e = 0; e < KP; e++) v[e] = L[(size_t) (e * M + b) * T + p];                          \
        if (ep.prefetch == 5) { _Pragma("unroll") for (int e = 0; e < KP; e++) v[e] = ld_stream(&src[(long long) (A) * N3 + (long long) (b + M * e) * N]); } \
        __syncthreads(); /* ring consumed; the previous unit's last exchange read is complete */                      \
        if (ep.prefetch == 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                            \
        if ((A) == 0) nn = s_nn[0];                                                                                    \
        if (NEXT_OK) issue_ring(NEXT_T, NEXT_A);                                                                       \
        if constexpr (RVZEL)                                                                                           \
            emit_finish<N, T, A, true, true>(v, (int) zl, S, keep, tw, ep, Lr, rec0, zl + ep.zglobal0, x, tid, p, b, s_red, pslot); \
        else                                                                                                           \
            emit_finish_rvdouble<N, T, A>(v, (int) zl, S, reinterpret_cast<double *>(keep), park3, tw, ep, rec0, zl + ep.zglobal0, x, tid, p, \
                                          b, s_red);                                                                   \
    }
        if constexpr (RVZEL) {
            // RVZel: A0 and A2 first (two parked floats), then A1 and A3 complete the two 16-byte halves
            if (qplt) {
                ZPLT_UNIT(0, true, cur, 2)
                ZPLT_UNIT(2, true, cur, 1)
                ZPLT_UNIT(1, true, cur, 3)
                ZPLT_UNIT(3, nxt < ntiles, nxt, 0)
            } else {
                ZPLT_UNIT(0, true, cur, 1)
                ZPLT_UNIT(1, nxt < ntiles, nxt, 0)
            }
        } else {
            if (qplt) {
                ZPLT_UNIT(0, true, cur, 1)
                ZPLT_UNIT(1, true, cur, 2)
                ZPLT_UNIT(2, true, cur, 3)
                ZPLT_UNIT(3, nxt < ntiles, nxt, 0)
            } else {
                ZPLT_UNIT(0, true, cur, 1)
                ZPLT_UNIT(1, nxt < ntiles, nxt, 0)
            }
        }
#undef ZPLT_UNIT
        cur = nxt;
        nxt = nn;
    }
    __syncthreads();
    if (tid < 7) {
        constexpr int NW = NT / 32;
        double a7 = s_red[0][tid];
        for (int w = 1; w < NW; w++) a7 = (tid == 0) ? a7 + s_red[w][tid] : fmax(a7, s_red[w][tid]);
        double *slot = ep.stats + 8 * (blockIdx.x % ZPLT_STAT_SLOTS);
        if (tid == 0)
            atomicAdd(&slot[0], a7);
        else
            atomicMax(reinterpret_cast<unsigned long long *>(&slot[tid]), (unsigned long long) __double_as_longlong(a7));
    }
}

// ------------------------------------------------------------------ dispatch -------
// pencils per CTA for the in-place strided/row passes
int fft_tile_T(int N) {
    switch (N) {
        case 16: return 16;
        case 32: return 32;
        case 64: return 32;
        case 128: return 16;
        case 256: return 16;
        case 512: return 8;
        case 1024: return 8;
        case 2048: return 4;
    }
    return 0;
}
size_t fft_tile_smem(int N, int T) { return (size_t) T * (N + (T >= 8 ? 1 : (T == 4 ? 2 : 4))) * sizeof(cplx); }

// CTAs that fit on the device at once (persistent kernels launch exactly that many)
static int persistent_ctas(const void *func, int threads, size_t smem, int sms) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, func, threads, smem);
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

// one 4-byte work counter per launch, rotating through the context's small device array so that launches in flight on
// different streams never share one
static unsigned int *tile_counter(LaunchRes &lr, cudaStream_t st) {
    if (!lr.counters) return nullptr;
    unsigned int *c = lr.counters + (lr.next_counter++ & 63);
    if (cudaMemsetAsync(c, 0, sizeof(unsigned int), st) != cudaSuccess) return nullptr;
    return c;
}

// 4-D FP64 tensor map (no swizzle, no interleave) through the driver entry point; 0 on success
static int encode_tmap4(CUtensorMap *tmap, const void *base, const cuuint64_t dims[4], const cuuint64_t strides[3],
                        const cuuint32_t box[4]) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
            return (int) cudaErrorNotSupported;
        encode = (encode_fn) fn;
    }
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return (int) cudaErrorInvalidValue;
    return 0;
}

template <int N, int T, int KP>
static int launch_tiles_ring_t(cplx *data, const TileGeom &g, const cplx *tw, LaunchRes &lr, cudaStream_t st) {
    constexpr int M = N / 16;
    const size_t smem = RingSmem<N, T>::EXCHANGE + (size_t) KP * M * T * sizeof(cplx) + 16;
    cudaError_t e = cudaFuncSetAttribute(fft_tile_ring_kernel<N, T, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return (int) e;
    unsigned int *ctr = tile_counter(lr, st);
    if (!ctr) return (int) cudaErrorMemoryAllocation;
    const long long ntiles = (long long) g.grid_x * g.grid_y * g.grid_z;
    if (ntiles >= (1ll << 31)) return (int) cudaErrorInvalidValue;
    const long long nctas = ntiles < lr.sms ? ntiles : lr.sms;
    // the data as a 4-D tensor of doubles: (x, rows of the outer index, transform axis, array)
    CUtensorMap tmap;
    const cuuint64_t dims[4]    = {(cuuint64_t) 2 * T * g.grid_x, (cuuint64_t) g.grid_y, (cuuint64_t) N, (cuuint64_t) g.grid_z};
    const cuuint64_t astr       = g.grid_z > 1 ? (cuuint64_t) g.astride : (cuuint64_t) g.nstride * N;
    const cuuint64_t ostr       = g.grid_y > 1 ? (cuuint64_t) g.ostride : (cuuint64_t) g.nstride * N;  // extent-1 axes: any valid stride
    const cuuint64_t strides[3] = {ostr * sizeof(cplx), (cuuint64_t) g.nstride * sizeof(cplx), astr * sizeof(cplx)};
    const cuuint32_t box[4]     = {2 * T, 1, (cuuint32_t) M, 1};
    if (int rc = encode_tmap4(&tmap, data, dims, strides, box)) return rc;
    fft_tile_ring_kernel<N, T, KP><<<(unsigned) nctas, T *(N / 16), smem, st>>>(data, g, tw, ctr, tmap);
    return (int) cudaGetLastError();
}

// sizes that have the ring-prefetched kernels instantiated (register file = one tile, shared memory = exchange image + ring)
template <int N, int T>
constexpr bool has_ring() {
    // (N = 2048 with 4-pencil tiles was tried: y pass of a rank of 8 30.0 -> 28.7 ms only, and its records failed parity; the
    // 2048 passes use the decimation kernels of zplt_fft2048_kernels.cu / the one-tile-per-CTA kernels)
    return (N == 1024 && T == 8) || (N == 512 && T == 8) || (N == 256 && T == 16) || (N == 64 && T == 32);
}

template <int N, int T>
static int launch_tiles_t(cplx *data, const TileGeom &g, const cplx *tw, const Tuning &tn, LaunchRes &lr, cudaStream_t st) {
    // ring-prefetched variant for unit-stride tiles (the z pass and the plain y pass); Tuning::zring = slices to prefetch
    if constexpr (has_ring<N, T>()) {
        const int kp = tn.zring;  // measured at PPD=1024: z pass 33.9 ms without, 29.9 / 28.1 / 27.5 ms with 4 / 8 / 12 slices
        if (kp > 0 && lr.counters && g.plo_stride == 1 && g.pa == T && g.tstride == T && g.phi_stride == 0) {
            if constexpr (N == 1024) {
                if (kp < 8) return launch_tiles_ring_t<N, T, 4>(data, g, tw, lr, st);
                if (kp < 12) return launch_tiles_ring_t<N, T, 8>(data, g, tw, lr, st);
            }
            return launch_tiles_ring_t<N, T, 12>(data, g, tw, lr, st);
        }
    }
    size_t smem = fft_tile_smem(N, T);
    cudaError_t e = cudaFuncSetAttribute(fft_tile_kernel<N, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return (int) e;
    const long long ntiles = (long long) g.grid_x * g.grid_y * g.grid_z;
    fft_tile_kernel<N, T><<<(unsigned) ntiles, T *(N / 16), smem, st>>>(data, g, tw);
    return (int) cudaGetLastError();
}

template <int N, int NP>
static int launch_genx_t(const GenParams &g, const SlabGeom &sg, cplx *cube, const cplx *tw, const Tuning &tn, LaunchRes &lr,
                         bool skip_fft, cudaStream_t st) {
    if ((2 * g.na) % NP) return (int) cudaErrorInvalidValue;
    size_t smem = fft_tile_smem(N, NP) + (size_t) 6 * N * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(gen_xfft_kernel<N, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return (int) e;
    if (cube == nullptr) {
        // prepare only: make sure the kernel's code is on the device.  With lazy module loading the first launch of a kernel loads
        // it, which synchronises with running work — fatal when the running work is the resident z pass waiting for this kernel.
        cudaFuncAttributes fa;
        return (int) cudaFuncGetAttributes(&fa, gen_xfft_kernel<N, NP>);
    }
    dim3 grid(N, sg.G == 1 ? N / 2 + 1 : sg.nly + ((sg.rank == 0 && sg.ly0 == 0) ? 1 : 0), 1);
    gen_xfft_kernel<N, NP><<<grid, NP *(N / 16), smem, st>>>(g, sg, cube, tw, skip_fft ? 1 : 0);
    return (int) cudaGetLastError();
}

#define ZPLT_CASE(FN, NN, TT, ...) \
    if (N == NN && T == TT) return FN<NN, TT>(__VA_ARGS__);

// pencils transformed concurrently by one CTA of the generation + x-FFT kernel (divides 2*narray)
int gen_xfft_T(int N, int na) {
    switch (N) {
        case 512: return na == 4 ? 8 : 4;
        default: return 4;  // N = 2048: 4 pencils (131 KB) + mode state (96 KB) = 229.5 KB, just inside the 227 KiB limit
    }
}

int launch_gen_xfft(int N, int T, const GenParams &g, const SlabGeom &sg, cplx *cube, const cplx *tw, const Tuning &tn, LaunchRes &lr,
                    bool skip_fft, cudaStream_t st) {
    ZPLT_CASE(launch_genx_t, 16, 4, g, sg, cube, tw, tn, lr, skip_fft, st)
    ZPLT_CASE(launch_genx_t, 32, 4, g, sg, cube, tw, tn, lr, skip_fft, st)
    ZPLT_CASE(launch_genx_t, 64, 4, g, sg, cube, tw, tn, lr, skip_fft, st)
    ZPLT_CASE(launch_genx_t, 128, 4, g, sg, cube, tw, tn, lr, skip_fft, st)
    ZPLT_CASE(launch_genx_t, 256, 4, g, sg, cube, tw, tn, lr, skip_fft, st)
    ZPLT_CASE(launch_genx_t, 512, 4, g, sg, cube, tw, tn, lr, skip_fft, st)
    ZPLT_CASE(launch_genx_t, 512, 8, g, sg, cube, tw, tn, lr, skip_fft, st)
    ZPLT_CASE(launch_genx_t, 1024, 4, g, sg, cube, tw, tn, lr, skip_fft, st)
    ZPLT_CASE(launch_genx_t, 2048, 4, g, sg, cube, tw, tn, lr, skip_fft, st)
    return (int) cudaErrorInvalidValue;
}

// gs.flags != NULL: one launch for all gs.J row groups of stage 1 (sg.nly rows each), gated by the flags; its CTAs stay resident
// while the generation kernels run, so they must leave SMs to them: at most half of the device
template <int N, int T>
static int launch_tiles_p2p_t(const cplx *b1, const SlabGeom &sg, cplx *const *peer_recv, const cplx *tw, const Tuning &tn,
                              LaunchRes &lr, const GroupSync &gs, cudaStream_t st) {
    PeerTable pt;
    for (int i = 0; i < 16; i++) pt.recv[i] = i < sg.G ? peer_recv[i] : nullptr;
    const long long ntiles = (long long) (N / T) * 2 * sg.nly * sg.na * gs.J;
    // the CTAs of this pass and those of the generation kernels share the SMs (see run_generate)
    int lim = tn.p2p_ctas;  // 0: as many as fit
    if (gs.flags != nullptr && (lim <= 0 || lim > lr.sms / 2)) lim = lr.sms / 2;
    if constexpr (has_ring<N, T>()) {
        if (tn.slab_ring > 0 && lr.counters && ntiles < (1ll << 31)) {
            constexpr int M = N / 16, KP = 12;
            const size_t smem = RingSmem<N, T>::EXCHANGE + (size_t) KP * M * T * sizeof(cplx) + 32;
            cudaError_t e = cudaFuncSetAttribute(fft_tile_p2p_ring_kernel<N, T, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
            if (e != cudaSuccess) return (int) e;
            unsigned int *ctr = gs.counter ? gs.counter : tile_counter(lr, st);
            if (!ctr) return (int) cudaErrorMemoryAllocation;
            long long nctas = lr.sms;
            if (lim > 0 && lim < nctas) nctas = lim;
            if (nctas > ntiles) nctas = ntiles;
            // the stage-1 buffer B1[z][row][x] as a tensor of doubles (x, row, z, 1)
            const int rows = sg.na * 2 * sg.h;
            CUtensorMap tmap;
            const cuuint64_t dims[4]    = {(cuuint64_t) 2 * N, (cuuint64_t) rows, (cuuint64_t) N, 1};
            const cuuint64_t strides[3] = {(cuuint64_t) N * sizeof(cplx), (cuuint64_t) rows * N * sizeof(cplx),
                                           (cuuint64_t) rows * N * N * sizeof(cplx)};
            const cuuint32_t box[4]     = {2 * T, 1, (cuuint32_t) M, 1};
            if (int rc = encode_tmap4(&tmap, b1, dims, strides, box)) return rc;
            fft_tile_p2p_ring_kernel<N, T, KP><<<(unsigned) nctas, T *(N / 16), smem, st>>>(b1, sg, pt, tw, ctr, tmap, gs);
            return (int) cudaGetLastError();
        }
    }
    size_t smem = fft_tile_smem(N, T);
    cudaError_t e = cudaFuncSetAttribute(fft_tile_p2p_kernel<N, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return (int) e;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft_tile_p2p_kernel<N, T>, T * (N / 16), smem);
    if (per_sm < 1) per_sm = 1;
    long long nctas = (long long) (lim > 0 && lim < lr.sms ? lim : lr.sms) * per_sm;  // lim counts SMs
    if (nctas > ntiles) nctas = ntiles;
    fft_tile_p2p_kernel<N, T><<<(unsigned) nctas, T *(N / 16), smem, st>>>(b1, sg, pt, tw, gs);
    return (int) cudaGetLastError();
}

bool fft_tiles_p2p_shares_tiles(int N, const Tuning &tn) {
    return tn.slab_ring > 0 && (N == 1024 || N == 512 || N == 256 || N == 64);  // has_ring sizes: fft_tile_p2p_ring_kernel
}

int launch_fft_tiles_p2p(int N, int T, const cplx *b1, const SlabGeom &sg, cplx *const *peer_recv, const cplx *tw, const Tuning &tn,
                         LaunchRes &lr, const GroupSync &gs, cudaStream_t st) {
    if (sg.G > 16) return (int) cudaErrorInvalidValue;
    ZPLT_CASE(launch_tiles_p2p_t, 32, 32, b1, sg, peer_recv, tw, tn, lr, gs, st)
    ZPLT_CASE(launch_tiles_p2p_t, 64, 32, b1, sg, peer_recv, tw, tn, lr, gs, st)
    ZPLT_CASE(launch_tiles_p2p_t, 128, 16, b1, sg, peer_recv, tw, tn, lr, gs, st)
    ZPLT_CASE(launch_tiles_p2p_t, 256, 16, b1, sg, peer_recv, tw, tn, lr, gs, st)
    ZPLT_CASE(launch_tiles_p2p_t, 512, 8, b1, sg, peer_recv, tw, tn, lr, gs, st)
    ZPLT_CASE(launch_tiles_p2p_t, 1024, 8, b1, sg, peer_recv, tw, tn, lr, gs, st)
    ZPLT_CASE(launch_tiles_p2p_t, 2048, 4, b1, sg, peer_recv, tw, tn, lr, gs, st)
    return (int) cudaErrorInvalidValue;
}

int launch_fft_tiles(int N, int T, cplx *data, const TileGeom &g, const cplx *tw, const Tuning &tn, LaunchRes &lr, cudaStream_t st) {
    ZPLT_CASE(launch_tiles_t, 16, 16, data, g, tw, tn, lr, st)
    ZPLT_CASE(launch_tiles_t, 32, 32, data, g, tw, tn, lr, st)
    ZPLT_CASE(launch_tiles_t, 64, 32, data, g, tw, tn, lr, st)
    ZPLT_CASE(launch_tiles_t, 128, 16, data, g, tw, tn, lr, st)
    ZPLT_CASE(launch_tiles_t, 256, 16, data, g, tw, tn, lr, st)
    ZPLT_CASE(launch_tiles_t, 512, 8, data, g, tw, tn, lr, st)
    ZPLT_CASE(launch_tiles_t, 1024, 8, data, g, tw, tn, lr, st)
    ZPLT_CASE(launch_tiles_t, 2048, 4, data, g, tw, tn, lr, st)
    return (int) cudaErrorInvalidValue;
}

template <int N, int T, bool RVZEL, int KP>
static int launch_emit_ring_t(const cplx *cube, long long z_first, long long nz, const EmitParams &ep, const cplx *tw, LaunchRes &lr,
                              cudaStream_t st) {
    constexpr int M = N / 16;
    const size_t smem = RingSmem<N, T>::EXCHANGE + (size_t) KP * M * T * sizeof(cplx) + 16;
    cudaError_t e = cudaFuncSetAttribute(fft_emit_ring_kernel<N, T, RVZEL, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return (int) e;
    unsigned int *ctr = tile_counter(lr, st);
    if (!ctr) return (int) cudaErrorMemoryAllocation;
    const long long ntiles = (long long) (N / T) * nz;
    long long nctas        = ntiles < lr.sms ? ntiles : lr.sms;
    if (nctas > 256) nctas = 256;  // parking slots (ZPLT_SCRATCH_BYTES)
    // the planes as a tensor of doubles: (x, y = transform axis, then z and a in the order of their strides)
    CUtensorMap tmap;
    const bool za = ep.zstride > ep.astride;
    const cuuint64_t d2 = za ? (cuuint64_t) ep.na : (cuuint64_t) ep.nzl, d3 = za ? (cuuint64_t) ep.nzl : (cuuint64_t) ep.na;
    const cuuint64_t s2 = za ? (cuuint64_t) ep.astride : (cuuint64_t) ep.zstride, s3 = za ? (cuuint64_t) ep.zstride : (cuuint64_t) ep.astride;
    const cuuint64_t dims[4]    = {(cuuint64_t) 2 * N, (cuuint64_t) N, d2, d3};
    const cuuint64_t strides[3] = {(cuuint64_t) N * sizeof(cplx), s2 * sizeof(cplx), s3 * sizeof(cplx)};
    const cuuint32_t box[4]     = {2 * T, (cuuint32_t) M, 1, 1};
    if (int rc = encode_tmap4(&tmap, cube, dims, strides, box)) return rc;
    float *keep_all = reinterpret_cast<float *>(static_cast<unsigned char *>(ep.scratch) + ZPLT_SCRATCH_PARK_BYTES);
    fft_emit_ring_kernel<N, T, RVZEL, KP><<<(unsigned) nctas, T *(N / 16), smem, st>>>(cube, z_first, nz, ep, tw, ctr, keep_all, tmap);
    return (int) cudaGetLastError();
}

// slab: the planes are in the per-source layout B2[src][zl][a][slot][x] of the caller-run all-to-all (zplt_slab.h); otherwise
// EmitParams describes them (single-GPU cube, or a slab rank's [zl][a][y][x] after the fused exchange)
template <int N, int T>
static int launch_emit_strided_t(const cplx *cube, const SlabGeom &sg, bool slab, long long z_first, long long nz, const EmitParams &ep,
                                 const cplx *tw, const Tuning &tn, LaunchRes &lr, cudaStream_t st, int *launches) {
    if constexpr (has_ring<N, T>()) {
        // persistent ring-prefetched form (records wanted, parking space present); Tuning::yring = 0 disables
        // The planes of a slab rank after the fused exchange ([zl][a][y][x]) go through the ring kernel only with slab_ring >= 2:
        // on that layout it produced wrong records in a few tiles per run (2-GPU run and one-GPU emulation alike, tools/diag_slab.py;
        // never on the single-GPU cube), so those ranks use the one-tile-per-CTA kernel (y pass of a rank of 8: 3.6 against 3.1 ms).
        const bool natural = ep.zstride > ep.astride;
        if (!slab && (!natural || tn.slab_ring >= 2) && ep.out != nullptr && ep.scratch != nullptr && lr.counters && tn.yring > 0) {
            // measured at PPD=1024: RVZel qPLT 28.9 -> 28.4 ms, ZA 20.5 -> 19.3 ms.  RVdoubleZel has its own record code in the ring
            // kernel (emit_finish_rvdouble); Zeldovich and ZelSimple keep the one-tile-per-CTA kernel
            if (ep.icformat == 1 || ep.icformat == 2) {  // RVZel, RVdoubleZel (parking areas are sized for up to 512 threads)
                int rc = (ep.icformat == 1) ? launch_emit_ring_t<N, T, true, 12>(cube, z_first, nz, ep, tw, lr, st)
                                            : launch_emit_ring_t<N, T, false, 12>(cube, z_first, nz, ep, tw, lr, st);
                if (launches) *launches += 1;
                return rc;
            }
        }
    }
    size_t smem = fft_tile_smem(N, T) + (size_t) 2 * 16 * T * (N / 16) * sizeof(float);
    dim3 grid(N / T, (unsigned) nz, 1);
    const bool rvzel = ep.icformat == 1;
    EmitParams ep2 = ep;
    if (ep2.scratch != nullptr) {
        // the per-SM parking space is only safe when CTAs of this kernel never share an SM, and large enough only up to 512 threads
        int per_sm = 0;
        if (rvzel && slab)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft_emit_strided_kernel<N, T, true, true>, T * (N / 16), smem);
        else if (rvzel)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft_emit_strided_kernel<N, T, false, true>, T * (N / 16), smem);
        else if (slab)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft_emit_strided_kernel<N, T, true, false>, T * (N / 16), smem);
        else
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft_emit_strided_kernel<N, T, false, false>, T * (N / 16), smem);
        if (per_sm != 1 || T * (N / 16) > 512 || !ep.qPLT || lr.sms > 256) ep2.scratch = nullptr;
    }
#define ZPLT_EMIT_LAUNCH(SL, RV)                                                                                              \
    {                                                                                                                         \
        cudaError_t e = cudaFuncSetAttribute(fft_emit_strided_kernel<N, T, SL, RV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int) smem);                                                                     \
        if (e != cudaSuccess) return (int) e;                                                                                 \
        fft_emit_strided_kernel<N, T, SL, RV><<<grid, T *(N / 16), smem, st>>>(cube, sg, z_first, ep2, tw);                    \
    }
    if (slab && rvzel) ZPLT_EMIT_LAUNCH(true, true)
    else if (slab) ZPLT_EMIT_LAUNCH(true, false)
    else if (rvzel) ZPLT_EMIT_LAUNCH(false, true)
    else ZPLT_EMIT_LAUNCH(false, false)
#undef ZPLT_EMIT_LAUNCH
    if (launches) *launches += 1;
    return (int) cudaGetLastError();
}

// y-axis FFT + record emission for planes [z_first, z_first+nz); the planes hold the x- and z-transformed arrays.
// ep.astride == 0 selects the per-source slab layout (SlabGeom), otherwise EmitParams::astride/zstride describe the planes.
int launch_fft_emit_strided(int N, int T, const cplx *cube, const SlabGeom &sg, long long z_first, long long nz,
                            const EmitParams &ep, const cplx *tw, const Tuning &tn, LaunchRes &lr, cudaStream_t st, int *launches) {
    const bool slab = ep.astride == 0;
    if (N == 2048) {
        const int rc = launch_fft2048_emit(cube, z_first, nz, ep, tw, tn, lr, st);
        if (rc >= 0) {
            if (launches) *launches += 1;
            return rc;
        }
    }
    ZPLT_CASE(launch_emit_strided_t, 16, 16, cube, sg, slab, z_first, nz, ep, tw, tn, lr, st, launches)
    ZPLT_CASE(launch_emit_strided_t, 32, 32, cube, sg, slab, z_first, nz, ep, tw, tn, lr, st, launches)
    ZPLT_CASE(launch_emit_strided_t, 64, 32, cube, sg, slab, z_first, nz, ep, tw, tn, lr, st, launches)
    ZPLT_CASE(launch_emit_strided_t, 128, 16, cube, sg, slab, z_first, nz, ep, tw, tn, lr, st, launches)
    ZPLT_CASE(launch_emit_strided_t, 256, 16, cube, sg, slab, z_first, nz, ep, tw, tn, lr, st, launches)
    ZPLT_CASE(launch_emit_strided_t, 512, 8, cube, sg, slab, z_first, nz, ep, tw, tn, lr, st, launches)
    ZPLT_CASE(launch_emit_strided_t, 1024, 8, cube, sg, slab, z_first, nz, ep, tw, tn, lr, st, launches)
    ZPLT_CASE(launch_emit_strided_t, 2048, 4, cube, sg, slab, z_first, nz, ep, tw, tn, lr, st, launches)
    return (int) cudaErrorInvalidValue;
}

}  // namespace zplt
