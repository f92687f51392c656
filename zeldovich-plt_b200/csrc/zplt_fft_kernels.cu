// FFT tile kernels: in-place strided/row backward FFTs (z and y axes) and the fused
// x-axis FFT + particle-record emission epilogue.
//
//   fft_tile_kernel  replaces InverseFFT_Yonly / the column half of Inverse2dFFT
//                    (reference src/zeldovich.cpp:93-114, :88-92 as used at :508-511, :653-658)
//   fft_emit_kernel  replaces the row half of Inverse2dFFT plus WriteParticlesSlab
//                    (reference src/output.cpp:41-234): unpack Re/Im of the packed arrays
//                    into displacement/velocity, cast to the ICFormat record, accumulate
//                    density_variance and max_disp.
#include "zplt_fft.cuh"
#include "zplt_internal.h"

namespace zplt {

template <int N, int T>
__global__ void __launch_bounds__(T *(N / 16)) fft_tile_kernel(cplx *__restrict__ data, TileGeom g, const cplx *__restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx *S         = reinterpret_cast<cplx *>(smem_raw);
    constexpr int M = N / 16;
    const int tid = threadIdx.x, p = tid % T, b = tid / T;
    const long long base = (long long) blockIdx.z * g.astride + (long long) blockIdx.y * g.ostride +
                           (long long) blockIdx.x * g.tstride + (long long) (p % g.pa) * g.plo_stride +
                           (long long) (p / g.pa) * g.phi_stride;
    cplx v[16];
#pragma unroll
    for (int e = 0; e < 16; e++) v[e] = data[base + (long long) (b + M * e) * g.nstride];
    fft_pencil<N>(v, S + p * FftPlan<N>::PSTRIDE, b, tw);
#pragma unroll
    for (int e = 0; e < 16; e++) data[base + (long long) (b + M * e) * g.nstride] = v[e];
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

__device__ __forceinline__ void track(double v, double &mp, double &mn) {
    if (v > mp) mp = v;
    if (-v > mn) mn = -v;
}

template <int N, int T>
__global__ void __launch_bounds__(T *(N / 16))
   fft_emit_kernel(const cplx *__restrict__ cube, long long z_first, EmitParams ep, const cplx *__restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_var;
    __shared__ unsigned long long s_max[6];
    cplx *S         = reinterpret_cast<cplx *>(smem_raw);
    constexpr int M = N / 16;
    const int tid = threadIdx.x, p = tid % T, b = tid / T;
    const int na = ep.na, RT = T / na;
    const int a = p % na, r = p / na;
    const long long z = z_first + blockIdx.y;
    const int y       = blockIdx.x * RT + r;
    const long long base = (long long) a * N * N * N + (z * N + y) * (long long) N;
    if (tid == 0) s_var = 0.0;
    if (tid < 6) s_max[tid] = 0ull;
    cplx v[16];
#pragma unroll
    for (int e = 0; e < 16; e++) v[e] = cube[base + b + M * e];
    fft_pencil<N>(v, S + p * FftPlan<N>::PSTRIDE, b, tw);
    __syncthreads();  // the pencil images are dead; reuse shared memory for the record image

    const RecLayout L  = rec_layout(ep.icformat);
    const int rb       = ep.record_bytes;
    unsigned char *img = smem_raw + (size_t) r * N * rb;
    double var = 0.0, mp0 = 0.0, mn0 = 0.0, mp1 = 0.0, mn1 = 0.0;
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const int x        = b + M * e;
        unsigned char *rec = img + (size_t) x * rb;
        const double re = v[e].x, im = v[e].y;
        if (a == 0) {
            // A0: Re = density, Im = pos[0] -> displ[2]
            if (L.off_ijk >= 0) {
                ushort4 id = make_ushort4((unsigned short) z, (unsigned short) y, (unsigned short) x, 0);
                *reinterpret_cast<ushort4 *>(rec + L.off_ijk) = id;
            }
            put(rec, L.off_d[2], im, L.dbl);
            if (!ep.qPLT) put(rec, L.off_v[2], im * ep.vnorm, L.dbl);
            var += re * re;
            track(im, mp0, mn0);
        } else if (a == 1) {
            // A1: Re = pos[1] -> displ[1], Im = pos[2] -> displ[0]
            put(rec, L.off_d[1], re, L.dbl);
            put(rec, L.off_d[0], im, L.dbl);
            if (!ep.qPLT) {
                put(rec, L.off_v[1], re * ep.vnorm, L.dbl);
                put(rec, L.off_v[0], im * ep.vnorm, L.dbl);
            }
            track(re, mp0, mn0);
            track(im, mp1, mn1);
        } else if (a == 2) {
            // A2: Im = vel[0] -> vel[2]
            put(rec, L.off_v[2], im, L.dbl);
        } else {
            // A3: Re = vel[1] -> vel[1], Im = vel[2] -> vel[0]
            put(rec, L.off_v[1], re, L.dbl);
            put(rec, L.off_v[0], im, L.dbl);
        }
    }
    // statistics: shared-memory atomics, then one global atomic per CTA and quantity
    if (a == 0) {
        atomicAdd(&s_var, var);
        atomicMax(&s_max[0], (unsigned long long) __double_as_longlong(mp0));
        atomicMax(&s_max[3], (unsigned long long) __double_as_longlong(mn0));
    } else if (a == 1) {
        atomicMax(&s_max[1], (unsigned long long) __double_as_longlong(mp0));
        atomicMax(&s_max[4], (unsigned long long) __double_as_longlong(mn0));
        atomicMax(&s_max[2], (unsigned long long) __double_as_longlong(mp1));
        atomicMax(&s_max[5], (unsigned long long) __double_as_longlong(mn1));
    }
    __syncthreads();
    // coalesced copy-out of RT consecutive rows of records
    {
        const size_t bytes = (size_t) RT * N * rb;
        unsigned char *dst = ep.out + ((size_t) ((z - ep.z0) * N + (long long) blockIdx.x * RT) * N) * rb;
        const int4 *s4     = reinterpret_cast<const int4 *>(smem_raw);
        int4 *d4           = reinterpret_cast<int4 *>(dst);
        for (size_t i = tid; i < bytes / 16; i += blockDim.x) d4[i] = s4[i];
    }
    if (tid < 7) {
        double *slot = ep.stats + 8 * ((blockIdx.x + blockIdx.y * gridDim.x) % ZPLT_STAT_SLOTS);
        if (tid == 0)
            atomicAdd(&slot[0], s_var);
        else
            atomicMax(reinterpret_cast<unsigned long long *>(&slot[tid]), s_max[tid - 1]);
    }
}

// ------------------------------------------------------------------ dispatch -------
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
size_t fft_tile_smem(int N, int T) { return (size_t) T * (N + 1) * sizeof(cplx); }

template <int N, int T>
static int launch_tiles_t(cplx *data, const TileGeom &g, const cplx *tw, cudaStream_t st) {
    size_t smem = fft_tile_smem(N, T);
    cudaError_t e = cudaFuncSetAttribute(fft_tile_kernel<N, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return (int) e;
    dim3 grid(g.grid_x, g.grid_y, g.grid_z);
    fft_tile_kernel<N, T><<<grid, T *(N / 16), smem, st>>>(data, g, tw);
    return (int) cudaGetLastError();
}

int launch_fft_tiles(int N, cplx *data, const TileGeom &g, const cplx *tw, cudaStream_t st) {
    switch (N) {
        case 16: return launch_tiles_t<16, 16>(data, g, tw, st);
        case 32: return launch_tiles_t<32, 32>(data, g, tw, st);
        case 64: return launch_tiles_t<64, 32>(data, g, tw, st);
        case 128: return launch_tiles_t<128, 16>(data, g, tw, st);
        case 256: return launch_tiles_t<256, 16>(data, g, tw, st);
        case 512: return launch_tiles_t<512, 8>(data, g, tw, st);
        case 1024: return launch_tiles_t<1024, 8>(data, g, tw, st);
        case 2048: return launch_tiles_t<2048, 4>(data, g, tw, st);
    }
    return (int) cudaErrorInvalidValue;
}

template <int N, int T>
static int launch_emit_t(const cplx *cube, long long z_first, long long nz, const EmitParams &ep, const cplx *tw,
                         cudaStream_t st, int *launches) {
    const int RT   = T / ep.na;
    size_t smem    = fft_tile_smem(N, T);
    size_t recs    = (size_t) RT * N * ep.record_bytes;
    if (recs > smem) smem = recs;
    cudaError_t e = cudaFuncSetAttribute(fft_emit_kernel<N, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return (int) e;
    // grid.y is limited to 65535: fine for nz <= 2048
    dim3 grid(N / RT, (unsigned) nz, 1);
    fft_emit_kernel<N, T><<<grid, T *(N / 16), smem, st>>>(cube, z_first, ep, tw);
    if (launches) *launches += 1;
    return (int) cudaGetLastError();
}

int launch_fft_emit(int N, const cplx *cube, long long z_first, long long nz, const EmitParams &ep, const cplx *tw,
                    cudaStream_t st, int *launches) {
    switch (N) {
        case 16: return launch_emit_t<16, 16>(cube, z_first, nz, ep, tw, st, launches);
        case 32: return launch_emit_t<32, 32>(cube, z_first, nz, ep, tw, st, launches);
        case 64: return launch_emit_t<64, 32>(cube, z_first, nz, ep, tw, st, launches);
        case 128: return launch_emit_t<128, 16>(cube, z_first, nz, ep, tw, st, launches);
        case 256: return launch_emit_t<256, 16>(cube, z_first, nz, ep, tw, st, launches);
        case 512: return launch_emit_t<512, 8>(cube, z_first, nz, ep, tw, st, launches);
        case 1024: return launch_emit_t<1024, 8>(cube, z_first, nz, ep, tw, st, launches);
        case 2048: return launch_emit_t<2048, 4>(cube, z_first, nz, ep, tw, st, launches);
    }
    return (int) cudaErrorInvalidValue;
}

}  // namespace zplt
