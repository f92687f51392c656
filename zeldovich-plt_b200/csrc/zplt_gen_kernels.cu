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
