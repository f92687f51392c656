// Particle grids whose ppd is not a power of two.  The reference takes any even ppd (reference src/block_array.cpp:38-40,
// src/parameters.cpp:123-126; Abacus production grids are 2^a 3^b); the fused kernels of this library are instantiated for
// powers of two.  This file is the general path for the other sizes (single GPU, ppd <= 1024): the plain generation kernel,
// each axis transformed by Bluestein's chirp-z algorithm on top of the power-of-two FFT kernels, and an unfused emission
// kernel.  It is a correctness path (same oracle, same tolerances), several times slower per particle than the fused one.
//
// Backward DFT of length N (reference sign +1, unnormalised, src/zeldovich.cpp:61-62) through circular convolution of length
// M = 2^m >= 2N-1:  n k = (n^2 + k^2 - (k-n)^2)/2, so with w[n] = exp(+i pi n^2/N)
//     X[k] = w[k] * sum_n (x[n] w[n]) conj(w)[k-n]  =  w[k] * (a (*) b)[k],   a = x w (zero padded),  b[m] = conj(w[|m|]).
// With only backward transforms F:  a (*) b = conj(F(conj(F(a) F(b)))) / M.
#include "zplt_internal.h"
#include "zplt_kernel_util.cuh"

namespace zplt {

__device__ __forceinline__ cplx cmul_g(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// start of pencil q of an axis pass, in elements: (q / div) * hstride + (q % div)
struct PencilGeom {
    long long div, hstride, estride;
};
__device__ __forceinline__ long long pencil_base(const PencilGeom &pg, long long q) { return (q / pg.div) * pg.hstride + (q % pg.div); }

// W[n][qb] = x[q0 + qb][n] w[n] for n < N, 0 for N <= n < M
__global__ void __launch_bounds__(256) blu_gather_kernel(const cplx *__restrict__ cube, PencilGeom pg, cplx *__restrict__ W,
                                                         const cplx *__restrict__ w, int N, long long Qb, long long q0, long long nq) {
    const long long qb = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const int n        = blockIdx.y;
    if (qb >= Qb) return;
    cplx v = make_double2(0.0, 0.0);
    if (n < N && qb < nq) v = cmul_g(cube[pencil_base(pg, q0 + qb) + (long long) n * pg.estride], w[n]);
    W[(long long) n * Qb + qb] = v;
}
// W <- conj(W * Bhat[n])
__global__ void __launch_bounds__(256) blu_mul_kernel(cplx *__restrict__ W, const cplx *__restrict__ Bhat, long long Qb) {
    const long long qb = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const int n        = blockIdx.y;
    if (qb >= Qb) return;
    const cplx t = cmul_g(W[(long long) n * Qb + qb], Bhat[n]);
    W[(long long) n * Qb + qb] = make_double2(t.x, -t.y);
}
// X[k] = w[k] conj(W[k]) / M
__global__ void __launch_bounds__(256) blu_scatter_kernel(cplx *__restrict__ cube, PencilGeom pg, const cplx *__restrict__ W,
                                                          const cplx *__restrict__ w, int N, double invM, long long Qb, long long q0, long long nq) {
    const long long qb = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const int k        = blockIdx.y;
    if (qb >= nq) return;
    const cplx t = W[(long long) k * Qb + qb];
    const cplx r = cmul_g(w[k], make_double2(t.x * invM, -t.y * invM));
    cube[pencil_base(pg, q0 + qb) + (long long) k * pg.estride] = r;
}

// One axis (0 = x, 1 = y, 2 = z) of the cube [a][z][y][x], in batches of Qb pencils through the work array W[M][Qb].
int launch_bluestein_axis(cplx *cube, int N, int M, int na, int axis, cplx *W, long long Qb, const cplx *w, const cplx *Bhat,
                          const cplx *twM, const Tuning &tn, LaunchRes &lr, cudaStream_t st) {
    const long long N2 = (long long) N * N, total = (long long) na * N2;
    PencilGeom pg;
    if (axis == 0)
        pg = {1, N, 1};
    else if (axis == 1)
        pg = {N, N2, N};
    else
        pg = {N2, N2 * N, N2};
    const int T = fft_tile_T(M);
    if (T == 0 || Qb % T) return (int) cudaErrorInvalidValue;
    TileGeom g;  // pencils along n of W[n][qb]
    g.nstride = Qb, g.plo_stride = 1, g.phi_stride = 0, g.pa = T, g.tstride = T, g.grid_x = (int) (Qb / T);
    g.ostride = 0, g.grid_y = 1, g.astride = 0, g.grid_z = 1;
    const dim3 blk(256), grdM((unsigned) ((Qb + 255) / 256), (unsigned) M), grdN((unsigned) ((Qb + 255) / 256), (unsigned) N);
    for (long long q0 = 0; q0 < total; q0 += Qb) {
        const long long nq = total - q0 < Qb ? total - q0 : Qb;
        blu_gather_kernel<<<grdM, blk, 0, st>>>(cube, pg, W, w, N, Qb, q0, nq);
        if (int rc = launch_fft_tiles(M, T, W, g, twM, tn, lr, st)) return rc;
        blu_mul_kernel<<<grdM, blk, 0, st>>>(W, Bhat, Qb);
        if (int rc = launch_fft_tiles(M, T, W, g, twM, tn, lr, st)) return rc;
        blu_scatter_kernel<<<grdN, blk, 0, st>>>(cube, pg, W, w, N, 1.0 / M, Qb, q0, nq);
    }
    return (int) cudaGetLastError();
}

// ------------------------------------------------------------------ unfused emission
// WriteParticlesSlab (reference src/output.cpp:41-234) from the fully transformed cube, one thread per particle.
__device__ __forceinline__ void put_field(unsigned char *rec, int off, double val, int dbl) {
    if (off < 0) return;
    if (dbl)
        *reinterpret_cast<double *>(rec + off) = val;
    else
        *reinterpret_cast<float *>(rec + off) = (float) val;
}
__global__ void __launch_bounds__(256) emit_plain_kernel(const cplx *__restrict__ cube, int N, long long z_first, long long nz, EmitParams ep) {
    __shared__ double s_red[8][8];
    const long long N2 = (long long) N * N, N3 = N2 * N;
    const long long i  = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    double q[7] = {0, 0, 0, 0, 0, 0, 0};  // sum dens^2, +max pos[0..2], -max pos[0..2]
    if (i < nz * N2) {
        const long long z = z_first + i / N2, yx = i % N2;
        const int y = (int) (yx / N), x = (int) (yx % N);
        const cplx a0 = cube[z * N2 + yx], a1 = cube[N3 + z * N2 + yx];
        const double dens = a0.x, pos[3] = {a0.y, a1.x, a1.y};
        double vel[3];
        if (ep.qPLT) {
            const cplx a2 = cube[2 * N3 + z * N2 + yx], a3 = cube[3 * N3 + z * N2 + yx];
            vel[0] = a2.y, vel[1] = a3.x, vel[2] = a3.y;
        } else {
            vel[0] = pos[0] * ep.vnorm, vel[1] = pos[1] * ep.vnorm, vel[2] = pos[2] * ep.vnorm;
        }
        q[0] = dens * dens;
        for (int j = 0; j < 3; j++) q[1 + j] = fmax(pos[j], 0.0), q[4 + j] = fmax(-pos[j], 0.0);
        if (ep.dens != nullptr) ep.dens[(size_t) (z - ep.z0) * N2 + yx] = (float) dens;
        if (ep.out != nullptr) {
            unsigned char *rec = ep.out + ((size_t) (z - ep.z0) * N2 + yx) * ep.record_bytes;
            // byte offsets per ICFormat (reference include/output.h:19-42): displ = (pos[2], pos[1], pos[0]), vel likewise
            int off_ijk = 0, od[3] = {8, 16, 24}, ov[3] = {-1, -1, -1}, dbl = 1;
            if (ep.icformat == 1) od[0] = 8, od[1] = 12, od[2] = 16, ov[0] = 20, ov[1] = 24, ov[2] = 28, dbl = 0;
            if (ep.icformat == 2) ov[0] = 32, ov[1] = 40, ov[2] = 48;
            if (ep.icformat == 3) off_ijk = -1, od[0] = 0, od[1] = 4, od[2] = 8, dbl = 0;
            if (off_ijk >= 0) *reinterpret_cast<ushort4 *>(rec) = make_ushort4((unsigned short) (z + ep.zglobal0), (unsigned short) y, (unsigned short) x, 0);
            for (int j = 0; j < 3; j++) {
                put_field(rec, od[j], pos[2 - j], dbl);
                put_field(rec, ov[j], vel[2 - j], dbl);
            }
        }
    }
    q[0] = warp_sum(q[0]);
    for (int j = 1; j < 7; j++) q[j] = warp_max(q[j]);
    if ((threadIdx.x & 31) == 0)
        for (int j = 0; j < 7; j++) s_red[threadIdx.x >> 5][j] = q[j];
    __syncthreads();
    if (threadIdx.x < 7) {
        double a = s_red[0][threadIdx.x];
        for (int wq = 1; wq < 8; wq++) a = threadIdx.x == 0 ? a + s_red[wq][threadIdx.x] : fmax(a, s_red[wq][threadIdx.x]);
        double *slot = ep.stats + 8 * (blockIdx.x % ZPLT_STAT_SLOTS);
        if (threadIdx.x == 0)
            atomicAdd(&slot[0], a);
        else
            atomicMax(reinterpret_cast<unsigned long long *>(&slot[threadIdx.x]), (unsigned long long) __double_as_longlong(a));
    }
}

int launch_emit_plain(const cplx *cube, int N, long long z_first, long long nz, const EmitParams &ep, cudaStream_t st) {
    const long long n = nz * N * (long long) N;
    emit_plain_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(cube, N, z_first, nz, ep);
    return (int) cudaGetLastError();
}

}  // namespace zplt
