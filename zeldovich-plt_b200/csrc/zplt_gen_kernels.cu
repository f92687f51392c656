// Mode-generation kernels: P(k) table from the spline, the packed spectral arrays
// (reference LoadPlane, src/zeldovich.cpp:278-503, without its FFT), and the RNG
// introspection kernels the parity tests use.
#include "zplt_internal.h"

namespace zplt {

// SplineFunction::val (reference include/spline_function.h:141-163) followed by
// PowerSpectrum::power (reference src/power_spectrum.cpp:225-261), tabulated at the only
// wavenumbers the lattice can ask for: k = sqrt(m)*fundamental, m = kx^2+ky^2+kz^2.
__global__ void power_table_kernel(double *__restrict__ ptab, long long count, double fundamental2, int is_powerlaw,
                                   double index, int n, const double *__restrict__ xs, const double *__restrict__ ys,
                                   const double *__restrict__ y2s, double normalization, double smooth2) {
    long long m = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= count) return;
    double k2 = (double) m * fundamental2;
    double k  = sqrt(k2);
    double P;
    if (k <= 0.0) {
        P = 0.0;
    } else if (is_powerlaw) {
        P = pow(k, index) * exp(-k * k * smooth2) * normalization;
    } else {
        double v = log(k);
        int klo = 0, khi = n - 1;
        while (khi - klo > 1) {
            int mid = (khi + klo) >> 1;
            if (xs[mid] > v)
                khi = mid;
            else
                klo = mid;
        }
        double h = xs[khi] - xs[klo];
        double a = (xs[khi] - v) / h;
        double b = (v - xs[klo]) / h;
        double s = a * ys[klo] + b * ys[khi] + ((a * a * a - a) * y2s[klo] + (b * b * b - b) * y2s[khi]) * (h * h) / 6.0;
        P        = exp(s - k * k * smooth2) * normalization;
    }
    ptab[m] = P;
}

int launch_power_table(double *ptab, long long count, double fundamental2, int is_powerlaw, double index, int n,
                       const double *x, const double *y, const double *y2, double normalization, double smooth2,
                       cudaStream_t st) {
    int threads = 256;
    long long blocks = (count + threads - 1) / threads;
    power_table_kernel<<<(unsigned) blocks, threads, 0, st>>>(ptab, count, fundamental2, is_powerlaw, index, n, x, y, y2,
                                                              normalization, smooth2);
    return (int) cudaGetLastError();
}

// Packed spectral arrays, layout [a][z][y][x].  One thread per primary mode for
// 0 < y < N/2: it writes the primary entry and its conjugate-structured twin at
// (N-x, N-y, N-z) (reference src/zeldovich.cpp:447-466).  The y = 0 plane is resolved per
// site (reference :485-503), the y = N/2 row is zero (reference :640-650 after the y
// shift of src/block_array.cpp:487-491).
__global__ void __launch_bounds__(128) generate_kernel(GenParams g, cplx *__restrict__ cube) {
    const int N = g.N;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = blockIdx.y;
    const int y = blockIdx.z;  // 0 .. N/2
    if (x >= N) return;
    const size_t N3 = (size_t) N * N * N;
    cplx a[4];
    if (y == g.half) {
        const size_t idx = ((size_t) z * N + y) * N + x;
        for (int i = 0; i < g.na; i++) cube[i * N3 + idx] = make_double2(0.0, 0.0);
        return;
    }
    if (y == 0) {
        const size_t idx = ((size_t) z * N) * N + x;
        const bool twin  = (z > g.half) || (z == 0 && x > g.half);
        Mode m;
        if (x == 0 && z == 0) {
            for (int i = 0; i < 4; i++) a[i] = make_double2(0.0, 0.0);
        } else if (twin) {
            primary_mode(g, (N - x) % N, 0, (N - z) % N, m);
            pack_twin(m, a);
        } else {
            primary_mode(g, x, 0, z, m);
            pack_primary(m, a);
        }
        for (int i = 0; i < g.na; i++) cube[i * N3 + idx] = a[i];
        return;
    }
    Mode m;
    primary_mode(g, x, y, z, m);
    pack_primary(m, a);
    {
        const size_t idx = ((size_t) z * N + y) * N + x;
        for (int i = 0; i < g.na; i++) cube[i * N3 + idx] = a[i];
    }
    pack_twin(m, a);
    {
        const int xh = (N - x) % N, zh = (N - z) % N;
        const size_t idx = ((size_t) zh * N + (N - y)) * N + xh;
        for (int i = 0; i < g.na; i++) cube[i * N3 + idx] = a[i];
    }
}

int launch_generate(const GenParams &g, cplx *cube, cudaStream_t st) {
    int threads = g.N < 128 ? g.N : 128;
    dim3 grid((g.N + threads - 1) / threads, g.N, g.N / 2 + 1);
    generate_kernel<<<grid, threads, 0, st>>>(g, cube);
    return (int) cudaGetLastError();
}

// ------------------------------------------------------------------ ZD_f_NL --------
// Local primordial non-Gaussianity (reference main, src/zeldovich.cpp:945-960; README "Primordial
// Non-Gaussianity").  The reference draws the Gaussian density modes, turns them into the Bardeen potential
// phi_g(k) = D / M(k) (ZeldovichZ with gen_phi = 1, :377-394), transforms to configuration space, applies
// phi = phi_g + f_NL phi_g^2 (ZeldovichXY_Phi, :699-790), transforms forward and feeds phi(k) M(k) back as
// the density of every mode (:396-400).  Here: mfactor_table_kernel (M at every |k|^2 the lattice has),
// generate_phi_kernel, the ordinary in-place FFT passes, fnl_local_kernel, the same passes again (a real
// field's forward transform is the conjugate of its backward transform), and the generation kernels read
// GenParams::phi.

// M(k, a) = 2 D(a) c^2 T(k) k^2 / (3 Omega_M H0^2) with T(k) = sqrt(P(k) / (primordial_norm k^n_s)), T(0) = 1
// (reference src/zeldovich.cpp:377-386, src/power_spectrum.cpp:263-274), same order of operations.
__global__ void mfactor_table_kernel(double *__restrict__ mtab, const double *__restrict__ ptab, long long count,
                                     double fundamental2, double primordial_norm, double n_s, double z_initial, double Omega_M) {
    long long m = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= count) return;
    double k2   = (double) m * fundamental2;
    double kmag = sqrt(k2);
    if (k2 == 0.0) k2 = 1.0;
    double Tk = 1.0;
    if (kmag > 0.0) Tk = sqrt(ptab[m] / (primordial_norm * exp(log(kmag) * n_s)));
    const double H0 = 100., c = 299792.458;
    const double growth = 1. / (1 + z_initial);
    mtab[m] = 2. * growth * c * c * Tk * k2 / (3. * Omega_M * H0 * H0);
}

// phi_g(k): D/M at primary sites, its conjugate at their twins, with the Hermitian bookkeeping of every other array (y = 0
// plane, origin, Nyquist row: see generate_kernel).  Single GPU: the full lattice [z][y][x], blockIdx.z = y = 0 .. N/2.
// Slab rank: its own rows only, [z][slot][x] with the slots of zplt_slab.h — blockIdx.z counts the rank's h primary rows
// (y = slot*G + rank, twin row in slot h + slot), plus one block for the all-zero Nyquist row on rank 0.
__global__ void __launch_bounds__(128) generate_phi_kernel(GenParams g, SlabGeom sg, cplx *__restrict__ phi) {
    const int N = g.N;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = blockIdx.y;
    const bool slab = sg.G > 1;
    const int y = !slab ? (int) blockIdx.z : ((int) blockIdx.z < sg.h ? (int) blockIdx.z * sg.G + sg.rank : g.half);
    if (x >= N) return;
    // start of row (z, y) / of the twin's row (zh, N - y)
    auto row = [&](int zz, int yy) -> size_t {
        if (!slab) return ((size_t) zz * N + yy) * N;
        int r, s;
        slab_owner(N, sg.G, yy, r, s);
        return ((size_t) zz * (2 * sg.h) + s) * N;
    };
    if (y == g.half) {
        phi[row(z, y) + x] = make_double2(0.0, 0.0);
        return;
    }
    double Dr, Di;
    int n2;
    if (y == 0) {
        const bool twin = (z > g.half) || (z == 0 && x > g.half);
        cplx v          = make_double2(0.0, 0.0);
        if (!(x == 0 && z == 0)) {
            if (twin)
                primary_density(g, (N - x) % N, 0, (N - z) % N, Dr, Di, n2);
            else
                primary_density(g, x, 0, z, Dr, Di, n2);
            const double M = __ldg(&g.mtab[n2]);
            v = make_double2(Dr / M, (twin ? -Di : Di) / M);
        }
        phi[row(z, 0) + x] = v;
        return;
    }
    primary_density(g, x, y, z, Dr, Di, n2);
    const double M = __ldg(&g.mtab[n2]);
    phi[row(z, y) + x] = make_double2(Dr / M, Di / M);
    const int xh = (N - x) % N, zh = (N - z) % N;
    phi[row(zh, N - y) + xh] = make_double2(Dr / M, -Di / M);
}

// Slab ranks, second exchange of the potential pass: after the local transformation and the y and x transforms of its
// planes, rank r holds phi[zl][y][x]; the generation kernels of the rank that owns row y need phi[z][y][x] for all z.
// Every row y < N/2 (only primary rows are read back, reference src/zeldovich.cpp:396-400) goes to its owner's
// [z][slot][x] buffer by peer stores — BlockArray::StoreBlockForward/LoadBlockForward (reference
// src/block_array.cpp:305-382, 416-464) as one copy kernel.  blockIdx = (y, zl); a NULL peer discards its rows.
struct PhiPeers {
    cplx *p1[16];
};
__global__ void __launch_bounds__(256) phi_return_kernel(const cplx *__restrict__ p2, SlabGeom sg, const __grid_constant__ PhiPeers peers) {
    const int N = sg.N, y = blockIdx.x, zl = blockIdx.y;
    const int np = N / sg.G, z = sg.rank * np + zl;
    cplx *dst = peers.p1[y & (sg.G - 1)];
    if (dst == nullptr) return;
    const cplx *src = p2 + ((size_t) zl * N + y) * N;
    cplx *out       = dst + ((size_t) z * (2 * sg.h) + (y >> sg.log2G)) * N;
    for (int x = threadIdx.x; x < N; x += blockDim.x) out[x] = src[x];
}

// phi <- (Re phi + f_NL (Re phi)^2) / ppd^3, imaginary part dropped (reference src/zeldovich.cpp:744-755)
__global__ void fnl_local_kernel(cplx *__restrict__ phi, long long n, double f_NL, double inv_ppd3) {
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const double p = phi[i].x;
        phi[i]         = make_double2((p + f_NL * p * p) * inv_ppd3, 0.0);
    }
}

int launch_mfactor_table(double *mtab, const double *ptab, long long count, double fundamental2, double primordial_norm, double n_s,
                         double z_initial, double Omega_M, cudaStream_t st) {
    int threads = 256;
    mfactor_table_kernel<<<(unsigned) ((count + threads - 1) / threads), threads, 0, st>>>(mtab, ptab, count, fundamental2,
                                                                                           primordial_norm, n_s, z_initial, Omega_M);
    return (int) cudaGetLastError();
}
int launch_generate_phi(const GenParams &g, const SlabGeom &sg, cplx *phi, cudaStream_t st) {
    int threads = g.N < 128 ? g.N : 128;
    const int ny = sg.G == 1 ? g.N / 2 + 1 : sg.h + (sg.rank == 0 ? 1 : 0);
    dim3 grid((g.N + threads - 1) / threads, g.N, ny);
    generate_phi_kernel<<<grid, threads, 0, st>>>(g, sg, phi);
    return (int) cudaGetLastError();
}
int launch_phi_return(const cplx *p2, const SlabGeom &sg, cplx *const *peer_p1, cudaStream_t st) {
    PhiPeers pp;
    for (int i = 0; i < 16; i++) pp.p1[i] = i < sg.G ? peer_p1[i] : nullptr;
    dim3 grid(sg.N / 2, sg.N / sg.G);
    phi_return_kernel<<<grid, 256, 0, st>>>(p2, sg, pp);
    return (int) cudaGetLastError();
}
int launch_fnl_local(cplx *phi, int N, long long count, double f_NL, cudaStream_t st) {
    const double inv = 1. / N / N / N;  // as the reference forms it (src/zeldovich.cpp:706)
    fnl_local_kernel<<<148 * 8, 256, 0, st>>>(phi, count, f_NL, inv);
    return (int) cudaGetLastError();
}

// ------------------------------------------------------------------ introspection --
__global__ void pcg_draws_kernel(const u128 *state0, const Affine *jump, long long n, uint64_t *out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    u128 s = apply(*jump, *state0);
    for (long long i = 0; i < n; i++) out[i] = pcg_next(s);
}
int launch_pcg_draws(const u128 *state0, const Affine *jump, long long n, uint64_t *out, cudaStream_t st) {
    pcg_draws_kernel<<<1, 32, 0, st>>>(state0, jump, n, out);
    return (int) cudaGetLastError();
}

__global__ void mode_draws_kernel(GenParams g, long long n, const int *__restrict__ k, uint64_t *__restrict__ raw,
                                  double *__restrict__ u) {
    long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int kx = k[3 * i], ky = k[3 * i + 1], kz = k[3 * i + 2];
    int x = kx < 0 ? kx + g.N : kx, z = kz < 0 ? kz + g.N : kz;
    u128 s         = mode_rng_state(g, x, ky, z);
    uint64_t r1    = pcg_next(s);
    uint64_t r2    = pcg_next(s);
    raw[2 * i]     = r1;
    raw[2 * i + 1] = r2;
    u[2 * i]       = u64_to_unit(r1);
    u[2 * i + 1]   = u64_to_unit(r2);
}
int launch_mode_draws(const GenParams &g, long long n, const int *k, uint64_t *raw, double *u, cudaStream_t st) {
    int threads = 128;
    mode_draws_kernel<<<(unsigned) ((n + threads - 1) / threads), threads, 0, st>>>(g, n, k, raw, u);
    return (int) cudaGetLastError();
}

}  // namespace zplt
