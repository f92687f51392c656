// 8-pencil decimation kernels for the N = 2048 strided passes (Tuning::dit2048).
// The index arithmetic (exchange patterns, swizzle, slot permutation, combine) is checked thread by thread in numpy by
// tools/proto_dit2048.py.
//
// 2048-point strided transforms with 8-pencil tiles.  8 pencils x 2048 complex are 256 KB, the whole
// register file, so the regular kernels fall back to 4-pencil tiles at N = 2048: 64-byte runs, which halve the DRAM
// efficiency of the strided passes and the NVLink efficiency of the peer stores (measured: 360 GB/s per GPU against
// 685 GB/s with 128-byte stores).  Here one CTA decimates in time once:
//   (a) the 1024 even rows of an 8-pencil tile (whole 128-byte rows) -> registers -> 1024-point transform whose
//       exchanges go through shared memory in a real and an imaginary round (8-byte image: 66 KB instead of 131 KB;
//       pencil stride N+2 words and index swizzle a ^ ((a >> 2) & 1) make all three passes conflict-free,
//       `python tools/bank_sim.py 1024 8 split`);
//   (b) the result E is parked in 128 KB of shared memory — the same thread owns the same output later, no barrier;
//   (c) the same for the odd rows, O stays in registers;
//   (d) X[k] = E[k] + W_2048^k O[k],  X[k+1024] = E[k] - W_2048^k O[k], stored as rows k and k+1024 (128-byte runs),
//       in place or straight into the peers' stage-2 buffers.
// Replaces, for N = 2048, fft_tile_kernel<2048,4> / fft_tile_p2p_kernel<2048,4> (reference InverseFFT_Yonly,
// src/zeldovich.cpp:93-114, and BlockArray::StoreBlock/LoadBlock, src/block_array.cpp:387-414, 466-504).
#include "zplt_fft.cuh"
#include "zplt_internal.h"
#include "zplt_kernel_util.cuh"

namespace zplt {
namespace {

struct Split {
    static constexpr int N = 1024, T = 8, M = 64, R3 = 4;
    static constexpr int PSTRIDE = N + 2;  // 8-byte words per pencil image
    static constexpr int NT = T * M;       // 512 threads
    static constexpr size_t IMAGE_BYTES = (size_t) T * PSTRIDE * sizeof(double);            // 65,664
    static constexpr size_t PARK_BYTES  = (size_t) 16 * NT * sizeof(cplx);                   // 131,072
    __device__ static __forceinline__ int at(int a) { return a ^ ((a >> 2) & 1); }
};

// 1024-point backward transform of one pencil, radices 16 x 16 x 4 as fft_pencil<1024, 8> (zplt_fft.cuh), with every
// exchange split into a real and an imaginary round through the 8-byte image S.  tw is the W_2048 table, tws = 2 its
// stride for W_1024.  v[e] = x[b + 64 e] on entry, X[bo + 64 e] on exit; every thread of the CTA must call it, and a CTA
// barrier must separate two calls (the first exchange of the next call overwrites what pass 3 of this one reads).
__device__ __forceinline__ int fft1024_split(cplx (&v)[16], double *S, int b, const cplx *__restrict__ tw, int tws) {
    constexpr int M = Split::M, R3 = Split::R3;
    // pass 1: radix 16 over stride M
    dft16(v);
    {
        cplx pw[16];
        twiddle_powers<16>(__ldg(&tw[b * tws]), pw);
#pragma unroll
        for (int k = 1; k < 16; k++) v[k] = cmul(v[k], pw[k]);
    }
    const int k1 = b / R3, i = b % R3;
    const int base = k1 * (16 * R3);
    // exchange 1: a transpose across the whole pencil (CTA barriers)
#pragma unroll
    for (int k = 0; k < 16; k++) S[Split::at(k * M + b)] = v[k].x;
    __syncthreads();
#pragma unroll
    for (int n = 0; n < 16; n++) v[n].x = S[Split::at(base + n * R3 + i)];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) S[Split::at(k * M + b)] = v[k].y;
    __syncthreads();
#pragma unroll
    for (int n = 0; n < 16; n++) v[n].y = S[Split::at(base + n * R3 + i)];
    // pass 2: radix 16 over stride R3, in place on this thread's own 16 locations
    dft16(v);
    {
        cplx pw[16];
        twiddle_powers<16>(__ldg(&tw[16 * i * tws]), pw);
#pragma unroll
        for (int k = 1; k < 16; k++) v[k] = cmul(v[k], pw[k]);
    }
    // exchange 2: pass 3 only needs what the R3 neighbouring slots of the same warp wrote (warp barriers)
#pragma unroll
    for (int k = 0; k < 16; k++) S[Split::at(base + k * R3 + i)] = v[k].x;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16 / R3; j++)
#pragma unroll
        for (int n = 0; n < R3; n++) v[j * R3 + n].x = S[Split::at(base + (i + R3 * j) * R3 + n)];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++) S[Split::at(base + k * R3 + i)] = v[k].y;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16 / R3; j++)
#pragma unroll
        for (int n = 0; n < R3; n++) v[j * R3 + n].y = S[Split::at(base + (i + R3 * j) * R3 + n)];
    // pass 3: radix R3 butterflies
    dft_groups<R3>(v);
    // butterfly j: v[j*R3 + k] = X[bo + M*(j + (16/R3)*k)]  -> slot order
    constexpr int NB = 16 / R3;
    cplx t[16];
#pragma unroll
    for (int j = 0; j < NB; j++)
#pragma unroll
        for (int k = 0; k < R3; k++) t[j + NB * k] = v[j * R3 + k];
#pragma unroll
    for (int e = 0; e < 16; e++) v[e] = t[e];
    return k1 + 16 * i;
}

// One 8-pencil tile of 2048 points: src + base is element (z = 0) of this thread's pencil, rows are nstride apart;
// store(k, lo, hi) receives the transformed elements k (lo) and k + 1024 (hi) of the pencil, k < 1024.
template <class Store>
__device__ __forceinline__ void dit2048_tile(const cplx *__restrict__ src, long long base, long long nstride, double *S_pencil, cplx *park,
                                             const cplx *__restrict__ tw, int tid, int b, Store store) {
    constexpr int M = Split::M, NT = Split::NT;
    // one inlined copy of the transform, run twice: even rows (h = 0, result parked), then odd rows (h = 1, combined).
    // (Two inlined copies keep v[] in local memory: ~1.2 KB of spills; this form has none.)
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
        cplx v[16];
#pragma unroll
        for (int e = 0; e < 16; e++) v[e] = ld_stream(&src[base + (long long) (2 * (b + M * e) + h) * nstride]);
        __syncthreads();  // the image of the previous transform is no longer read
        const int bo = fft1024_split(v, S_pencil, b, tw, 2);  // v[e] = E[bo + 64 e] or O[bo + 64 e]: the same slot permutation
        if (h == 0) {
#pragma unroll
            for (int e = 0; e < 16; e++) park[e * NT + tid] = v[e];
        } else {
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int k  = bo + M * e;
                const cplx t = cmul(v[e], __ldg(&tw[k]));  // W_2048^k O[k]
                const cplx E = park[e * NT + tid];
                store(k, cadd(E, t), csub(E, t));  // elements k and k + 1024
            }
        }
    }
}

__global__ void __launch_bounds__(512, 1) fft2048_dit_kernel(cplx *__restrict__ data, TileGeom g, const cplx *__restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem_dit[];
    double *S  = reinterpret_cast<double *>(smem_dit);
    cplx *park = reinterpret_cast<cplx *>(smem_dit + Split::IMAGE_BYTES);
    const int tid = threadIdx.x, p = tid % Split::T, b = tid / Split::T;
    const long long t  = blockIdx.x;
    const long long tx = t % g.grid_x, ty = (t / g.grid_x) % g.grid_y, tz = t / ((long long) g.grid_x * g.grid_y);
    const long long base = tz * g.astride + ty * g.ostride + tx * g.tstride + p;
    cplx *dst            = data;
    const long long ns   = g.nstride;
    dit2048_tile(data, base, ns, S + p * Split::PSTRIDE, park, tw, tid, b, [=](int k, cplx lo, cplx hi) {
        __stcs(&dst[base + (long long) k * ns], lo);
        __stcs(&dst[base + (long long) (k + 1024) * ns], hi);
    });
}

struct PeerTable2 {
    cplx *recv[16];
};

// the z pass of a slab rank at N = 2048 with the exchange fused in (see fft_tile_p2p_kernel): 128-byte peer stores into
// the owners' stage-2 buffers B2[zl][a][y][x]; a NULL peer discards that rank's share
__global__ void __launch_bounds__(512, 1)
   fft2048_dit_p2p_kernel(const cplx *__restrict__ b1, SlabGeom sg, const __grid_constant__ PeerTable2 peers, const cplx *__restrict__ tw,
                          GroupSync gs) {
    extern __shared__ __align__(16) unsigned char smem_dit[];
    __shared__ int s_ok;
    double *S  = reinterpret_cast<double *>(smem_dit);
    cplx *park = reinterpret_cast<cplx *>(smem_dit + Split::IMAGE_BYTES);
    constexpr int N = 2048, T = Split::T, XT = N / T;
    const int tid = threadIdx.x, p = tid % T, b = tid / T;
    const int rows = sg.na * 2 * sg.h;  // x-rows per z plane of the stage-1 buffer
    const long long nstride = (long long) rows * N;
    const int np  = N / sg.G, lognp = 11 - sg.log2G;
    const int nsl = 2 * sg.nly;
    const long long tpg = (long long) XT * nsl * sg.na, ntiles = tpg * gs.J;  // tiles are numbered group by group (see p2p_tile)
    int ready = gs.flags == nullptr ? gs.J : -1;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int grp      = (int) (t / tpg);
        const long long tt = t - grp * tpg;
        if (grp > ready) {  // CTA-uniform: wait until the generation kernel of this row group has completed
            if (tid == 0) s_ok = wait_group(gs.flags, grp) ? 1 : 0;
            __syncthreads();
            if (!s_ok) {
                if (tid == 0) atomicExch(gs.err, 1u);
                return;
            }
            ready = grp;
        }
        const int xt = (int) (tt % XT);
        const int rr = (int) (tt / XT), sidx = rr % nsl, a = rr / nsl;
        const int ly0  = sg.ly0 + grp * sg.nly;
        const int slot = sidx < sg.nly ? ly0 + sidx : sg.h + ly0 + (sidx - sg.nly);
        const int row  = a * 2 * sg.h + slot;
        const int x    = xt * T + p;
        const int y    = slab_row(N, sg.G, sg.rank, slot);
        const long long base = (long long) row * N + x;
        const long long zstride = sg.b2_persrc ? (long long) rows * N : sg.b2_zstride;
        const long long rowoff  = sg.b2_persrc ? ((long long) sg.rank * np * rows + row) * N + x : ((long long) a * N + y) * N + x;
        // planes k and k + 1024 have the same local index on ranks G/2 apart (N/G divides 1024 for G >= 2)
        const int rhalf = sg.G >> 1;
        dit2048_tile(b1, base, nstride, S + p * Split::PSTRIDE, park, tw, tid, b, [&](int k, cplx lo, cplx hi) {
            const int r = k >> lognp;
            const long long off = (long long) (k & (np - 1)) * zstride + rowoff;
            cplx *d0 = peers.recv[r], *d1 = peers.recv[r + rhalf];
            if (d0 != nullptr) __stcs(&d0[off], lo);
            if (d1 != nullptr) __stcs(&d1[off], hi);
        });
    }
}

constexpr size_t DIT_SMEM = Split::IMAGE_BYTES + Split::PARK_BYTES;

// ------------------------------------------------------------------ y pass + emission at N = 2048
// The y pass of the north-star size with 8-pencil tiles (128-byte runs on the loads, whole 32-byte sectors on the record
// stores) instead of the 4-pencil tiles of fft_emit_strided_kernel<2048, 4>: the decimation transform above, whose two
// outputs per butterfly (rows y and y + 1024) go straight into the record logic of fft_emit_*_kernel
// (zplt_fft_kernels.cu; reference WriteParticlesSlab, src/output.cpp:41-234).  RVZel records only (BASELINE configs[3]
// and [4]); the other formats keep the 4-pencil kernel.  One inlined copy of the transform inside a non-unrolled loop
// over the packed arrays (A0, A2, A1, A3: the two parked floats first, then the two record halves), the array index is
// CTA-uniform.  Persistent CTAs over (plane, x tile); parked fields live in an L2-resident area indexed by CTA.
//   per CTA: keep0 [32][512] float (displ[2]) | keep1 [32][512] float (vel[2]) | d01 [32][512] float2 (displ[0], displ[1])
#define ZPLT_EMIT2048_SLOT_BYTES (262144)
__global__ void __launch_bounds__(512, 1)
   fft2048_emit_kernel(const cplx *__restrict__ planes, long long z_first, long long nz, const __grid_constant__ EmitParams ep,
                       const cplx *__restrict__ tw, unsigned int *__restrict__ counter) {
    extern __shared__ __align__(16) unsigned char smem_dit[];
    __shared__ double s_red[32][8];
    __shared__ unsigned int s_next;
    double *S  = reinterpret_cast<double *>(smem_dit);
    cplx *park = reinterpret_cast<cplx *>(smem_dit + Split::IMAGE_BYTES);
    constexpr int N = 2048, T = Split::T, XT = N / T, NT = Split::NT;
    const int tid = threadIdx.x, p = tid % T, b = tid / T;
    // this thread's parking column: keep0 = kp, keep1 = kp + 32*NT floats, d01 = the float2 area behind them
    float *kp = reinterpret_cast<float *>(static_cast<unsigned char *>(ep.scratch) + (size_t) blockIdx.x * ZPLT_EMIT2048_SLOT_BYTES) + tid;
#define keep0 kp
#define keep1 (kp + 32 * NT)
#define d01 (reinterpret_cast<float2 *>(kp + 64 * NT - tid) + tid)
    const uint64_t pol = l2_evict_last();
    const bool qplt = ep.qPLT;
#define vn ep.vnorm
    const int nsteps = qplt ? 4 : 2;
    const unsigned int ntiles = (unsigned int) (XT * nz);
    for (int i = tid; i < 32 * 8; i += NT) (&s_red[0][0])[i] = 0.0;
    if (tid == 0) s_next = atomicAdd(counter, 1u);
    __syncthreads();
    unsigned int cur = s_next;
    while (cur < ntiles) {
        __syncthreads();  // everybody has read s_next
        if (tid == 0) s_next = atomicAdd(counter, 1u);
        const int x        = (int) (cur % XT) * T + p;
        const long long zl = z_first + cur / XT;
        const long long z  = zl + ep.zglobal0;
        const cplx *src    = planes + zl * ep.zstride + x;
        unsigned char *rec0 = ep.out + ((size_t) ((zl - ep.z0) * N) * N + x) * 32;
        const unsigned int w1 = (unsigned int) (unsigned short) x;
#pragma unroll 1
        for (int s = 0; s < nsteps; s++) {
            const int A = qplt ? ((s == 0) ? 0 : (s == 1) ? 2 : (s == 2) ? 1 : 3) : s;
            double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;  // A0: sum dens^2, +max, -max of pos[0]; A1: +-max of pos[1], pos[2]
            int j = 0;                                      // parking index of this butterfly's two values
            dit2048_tile(src + (long long) A * ep.astride, 0, N, S + p * Split::PSTRIDE, park, tw, tid, b, [&](int k, cplx lo, cplx hi) {
                const int jl = j * NT, jh = (j + 1) * NT;
                j += 2;
                if (A == 0) {  // Re = density, Im = pos[0] -> displ[2] (and vel[2] without qPLT): parked
                    q0 += lo.x * lo.x + hi.x * hi.x;
                    q1 = fmax(q1, fmax(lo.y, hi.y)), q2 = fmax(q2, fmax(-lo.y, -hi.y));
                    if (ep.dens != nullptr) {
                        float *dens0 = ep.dens + ((size_t) (zl - ep.z0) * N) * N + x;
                        dens0[(size_t) k * N] = (float) lo.x, dens0[(size_t) (k + 1024) * N] = (float) hi.x;
                    }
                    if (ep.out != nullptr) {
                        park_st<true>(keep0 + jl, (float) lo.y, pol), park_st<true>(keep0 + jh, (float) hi.y, pol);
                        if (!qplt) park_st<true>(keep1 + jl, (float) (lo.y * vn), pol), park_st<true>(keep1 + jh, (float) (hi.y * vn), pol);
                    }
                } else if (A == 2) {  // Im = vel[0] -> vel[2]: parked
                    if (ep.out != nullptr) park_st<true>(keep1 + jl, (float) lo.y, pol), park_st<true>(keep1 + jh, (float) hi.y, pol);
                } else if (A == 1) {  // Re = pos[1] -> displ[1], Im = pos[2] -> displ[0]
                    q0 = fmax(q0, fmax(lo.x, hi.x)), q1 = fmax(q1, fmax(-lo.x, -hi.x));
                    q2 = fmax(q2, fmax(lo.y, hi.y)), q3 = fmax(q3, fmax(-lo.y, -hi.y));
                    if (ep.out == nullptr) {
                    } else if (qplt) {
                        park_st2(d01 + jl, make_float2((float) lo.y, (float) lo.x), pol);
                        park_st2(d01 + jh, make_float2((float) hi.y, (float) hi.x), pol);
                    } else {
                        const unsigned int wl = (unsigned int) (unsigned short) z | ((unsigned int) (unsigned short) k << 16);
                        const unsigned int wh = (unsigned int) (unsigned short) z | ((unsigned int) (unsigned short) (k + 1024) << 16);
                        st_record32(rec0 + (size_t) k * N * 32, make_float4(__uint_as_float(wl), __uint_as_float(w1), (float) lo.y, (float) lo.x),
                                    make_float4(park_ld<true>(keep0 + jl, pol), (float) (lo.y * vn), (float) (lo.x * vn), park_ld<true>(keep1 + jl, pol)));
                        st_record32(rec0 + (size_t) (k + 1024) * N * 32, make_float4(__uint_as_float(wh), __uint_as_float(w1), (float) hi.y, (float) hi.x),
                                    make_float4(park_ld<true>(keep0 + jh, pol), (float) (hi.y * vn), (float) (hi.x * vn), park_ld<true>(keep1 + jh, pol)));
                    }
                } else if (ep.out != nullptr) {  // A == 3: Re = vel[1], Im = vel[2] -> vel[0]: the record is complete
                    const float2 dl = park_ld2(d01 + jl, pol), dh = park_ld2(d01 + jh, pol);
                    const unsigned int wl = (unsigned int) (unsigned short) z | ((unsigned int) (unsigned short) k << 16);
                    const unsigned int wh = (unsigned int) (unsigned short) z | ((unsigned int) (unsigned short) (k + 1024) << 16);
                    st_record32(rec0 + (size_t) k * N * 32, make_float4(__uint_as_float(wl), __uint_as_float(w1), dl.x, dl.y),
                                make_float4(park_ld<true>(keep0 + jl, pol), (float) lo.y, (float) lo.x, park_ld<true>(keep1 + jl, pol)));
                    st_record32(rec0 + (size_t) (k + 1024) * N * 32, make_float4(__uint_as_float(wh), __uint_as_float(w1), dh.x, dh.y),
                                make_float4(park_ld<true>(keep0 + jh, pol), (float) hi.y, (float) hi.x, park_ld<true>(keep1 + jh, pol)));
                }
            });
            if (A == 0) {
                fold_stats<NT, true>(s_red, tid, q0, 0, q1, 1, q2, 4);
            } else if (A == 1) {
                fold_stats<NT, true>(s_red, tid, q0, 2, q1, 5, q2, 3);
                fold_stats<NT, true>(s_red, tid, q3, 6, 0.0, 7, 0.0, 7);
            }
        }
        cur = s_next;
    }
    __syncthreads();
    if (tid < 7) {
        constexpr int NW = NT / 32;
        double a7 = s_red[0][tid];
        for (int w = 1; w < NW; w++) a7 = (tid == 0) ? a7 + s_red[w][tid] : fmax(a7, s_red[w][tid]);
        double *sl = ep.stats + 8 * (blockIdx.x % ZPLT_STAT_SLOTS);
        if (tid == 0)
            atomicAdd(&sl[0], a7);
        else
            atomicMax(reinterpret_cast<unsigned long long *>(&sl[tid]), (unsigned long long) __double_as_longlong(a7));
    }
#undef keep0
#undef keep1
#undef d01
#undef vn
}

}  // namespace

// In-place strided pass with the N = 2048 decimation kernel when it is enabled and the geometry is the 4-pencil
// unit-stride tiling the regular launcher would use; otherwise the regular launcher.
int launch_fft_tiles_any(int N, int T, cplx *data, const TileGeom &g, const cplx *tw, const Tuning &tn, LaunchRes &lr, cudaStream_t st) {
    if (N == 2048 && tn.dit2048 > 0 && g.plo_stride == 1 && g.phi_stride == 0 && g.pa == 4 && g.tstride == 4 && (g.grid_x % 2) == 0) {
        TileGeom g8 = g;
        g8.pa = 8, g8.tstride = 8, g8.grid_x = g.grid_x / 2;
        cudaError_t e = cudaFuncSetAttribute(fft2048_dit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) DIT_SMEM);
        if (e != cudaSuccess) return (int) e;
        const long long ntiles = (long long) g8.grid_x * g8.grid_y * g8.grid_z;
        fft2048_dit_kernel<<<(unsigned) ntiles, 512, DIT_SMEM, st>>>(data, g8, tw);
        return (int) cudaGetLastError();
    }
    return launch_fft_tiles(N, T, data, g, tw, tn, lr, st);
}

int launch_fft_tiles_p2p_any(int N, int T, const cplx *b1, const SlabGeom &sg, cplx *const *peer_recv, const cplx *tw,
                             const Tuning &tn, LaunchRes &lr, const GroupSync &gs, cudaStream_t st) {
    if (N == 2048 && tn.dit2048 > 0 && sg.G <= 16) {
        cudaError_t e = cudaFuncSetAttribute(fft2048_dit_p2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) DIT_SMEM);
        if (e != cudaSuccess) return (int) e;
        PeerTable2 pt;
        for (int i = 0; i < 16; i++) pt.recv[i] = i < sg.G ? peer_recv[i] : nullptr;
        const long long ntiles = (long long) (N / 8) * 2 * sg.nly * sg.na * gs.J;
        long long nctas = lr.sms;
        int lim         = tn.p2p_ctas;  // as fft_tile_p2p_kernel: leave SMs to the overlapped generation kernels
        if (gs.flags != nullptr && (lim <= 0 || lim > (lr.sms * 2) / 3)) lim = (lr.sms * 2) / 3;
        if (lim > 0 && lim < nctas) nctas = lim;
        if (nctas > ntiles) nctas = ntiles;
        fft2048_dit_p2p_kernel<<<(unsigned) nctas, 512, DIT_SMEM, st>>>(b1, sg, pt, tw, gs);
        return (int) cudaGetLastError();
    }
    return launch_fft_tiles_p2p(N, T, b1, sg, peer_recv, tw, tn, lr, gs, st);
}

// y pass + emission of planes [z_first, z_first + nz) with the 8-pencil decimation kernel; returns -1 when this launch
// is not its case (then the caller uses the regular kernels)
int launch_fft2048_emit(const cplx *planes, long long z_first, long long nz, const EmitParams &ep, const cplx *tw, const Tuning &tn,
                        LaunchRes &lr, cudaStream_t st) {
    if (tn.dit2048_emit <= 0 || ep.icformat != 1 || ep.astride == 0 || ep.scratch == nullptr || !lr.counters) return -1;
    if (ep.out != nullptr && (reinterpret_cast<size_t>(ep.out) & 31)) return -1;  // 256-bit record stores
    const long long ntiles = (long long) (2048 / 8) * nz;
    if (ntiles >= (1ll << 31)) return -1;
    long long nctas = ntiles < lr.sms ? ntiles : lr.sms;
    if ((size_t) nctas * ZPLT_EMIT2048_SLOT_BYTES > ZPLT_SCRATCH_BYTES) nctas = ZPLT_SCRATCH_BYTES / ZPLT_EMIT2048_SLOT_BYTES;
    cudaError_t e = cudaFuncSetAttribute(fft2048_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) DIT_SMEM);
    if (e != cudaSuccess) return (int) e;
    unsigned int *ctr = lr.counters + (lr.next_counter++ & 63);
    if (cudaMemsetAsync(ctr, 0, sizeof(unsigned int), st) != cudaSuccess) return (int) cudaErrorMemoryAllocation;
    fft2048_emit_kernel<<<(unsigned) nctas, 512, DIT_SMEM, st>>>(planes, z_first, nz, ep, tw, ctr);
    return (int) cudaGetLastError();
}

}  // namespace zplt
