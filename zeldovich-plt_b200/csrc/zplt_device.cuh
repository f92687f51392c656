// Device-side building blocks of the mode generator: 128-bit PCG64 arithmetic, the
// (0,1] mapping, Box-Muller, PLT eigenmode interpolation and the per-mode field
// amplitudes.  sm_100a only.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace zplt {

typedef double2 cplx;  // .x = re, .y = im

// ------------------------------------------------------------------ 128-bit LCG ----
struct u128 {
    uint64_t lo, hi;
};
// affine map s -> mult*s + plus (mod 2^128): a jump of the LCG by some number of draws
struct Affine {
    u128 mult, plus;
};

__host__ __device__ __forceinline__ u128 mul128(u128 a, u128 b) {
    u128 r;
    r.lo = a.lo * b.lo;
#ifdef __CUDA_ARCH__
    r.hi = __umul64hi(a.lo, b.lo) + a.hi * b.lo + a.lo * b.hi;
#else
    unsigned __int128 t = (unsigned __int128) a.lo * b.lo;
    r.hi                = (uint64_t) (t >> 64) + a.hi * b.lo + a.lo * b.hi;
#endif
    return r;
}
__host__ __device__ __forceinline__ u128 add128(u128 a, u128 b) {
    u128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
    return r;
}
__host__ __device__ __forceinline__ u128 apply(const Affine &j, u128 s) { return add128(mul128(j.mult, s), j.plus); }

// pcg64 = setseq_xsl_rr_128_64 with the default stream
// (reference include/pcg-rng/pcg_random.hpp:159-170 constants, :370 bump, :381-386
// operator() with output_previous=false, :1144-1170 xsl_rr output).
#define ZPLT_PCG_MULT_HI 2549297995355413924ull
#define ZPLT_PCG_MULT_LO 4865540595714422341ull
#define ZPLT_PCG_INC_HI 6364136223846793005ull
#define ZPLT_PCG_INC_LO 1442695040888963407ull

__host__ __device__ __forceinline__ uint64_t pcg_next(u128 &s) {
    const u128 M = {ZPLT_PCG_MULT_LO, ZPLT_PCG_MULT_HI};
    const u128 C = {ZPLT_PCG_INC_LO, ZPLT_PCG_INC_HI};
    s            = add128(mul128(s, M), C);
    uint64_t x   = s.hi ^ s.lo;
    unsigned rot = (unsigned) (s.hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}

// one_rand<2> (reference src/power_spectrum.cpp:284-308): uint64 -> (0,1];
// (double)(r+1) rounds to nearest even exactly as the host conversion does, and the
// scale by 2^-64 is exact.
__device__ __forceinline__ double u64_to_unit(uint64_t r) {
    if (r == 0xffffffffffffffffull) return 1.0;
    return __ull2double_rn(r + 1ull) * 0x1.0p-64;
}

// ------------------------------------------------------------------ parameters -----
struct GenParams {
    int N;        // ppd
    int half;     // ppd/2
    int kmax;     // (int)(half/k_cutoff + .5)       reference src/zeldovich.cpp:350
    int na;       // packed arrays: 2 (ZA) or 4 (qPLT)
    int corner_modes, qonemode, one_mode[3];
    int qPLT, qPLTrescale, fixed_power;
    double fundamental, fundamental2;  // 2*pi/L and its square (reference src/parameters.cpp:176, src/zeldovich.cpp:301)
    double k2_cutoff;                  // nyquist^2/k_cutoff^2   (reference src/zeldovich.cpp:318-319)
    double f_cluster, target_f, growth_ratio;  // target_f :305 ; growth_ratio = a_NL/a0 :307-312,428
    double log_growth_ratio;                   // log(a_NL/a0)
    // tables
    const double *ptab;    // P at k = sqrt(m)*fundamental, m = kx^2+ky^2+kz^2
    const u128 *ystate;    // [N/2] generator state at the start of plane ky (2*ky*65536^2 draws after seeding)
    const Affine *zjump;   // [N]   jump by 2*65536*(kz mod 65536) draws, indexed by lattice z
    const Affine *xjump;   // [N]   jump by 2*(kx mod 65536) draws, indexed by lattice x
    const double *eig;     // [pe][pe][pe/2+1][4]
    int pe;                // eigenmode table ppd
    int eig_direct;        // pe % N == 0 -> direct lookup (reference src/zeldovich.cpp:161-170)
    double eig_scale;      // (double)pe / N
    // local primordial non-Gaussianity (ZD_f_NL != 0; reference src/zeldovich.cpp:377-400): when phi != NULL the
    // density of every mode but the origin, masked or not, is conj(phi[z][y][x]) * mtab[m] instead of a draw.  phi
    // holds the BACKWARD transform of the real field phi_g + f_NL phi_g^2; its conjugate is the forward transform
    // the reference takes.  mtab[m] = M(k = sqrt(m) * fundamental), the potential -> density factor.
    const double2 *phi;
    const double *mtab;
    // where row (z, y < N/2) of phi starts, in elements: z*phi_zstride + (y >> phi_yshift)*N.  Single GPU: the cube
    // [z][y][x] (zstride N^2, shift 0).  Slab rank: its own rows [z][slot][x] (zstride 2h*N, slot = y / G).
    long long phi_zstride;
    int phi_yshift;
    // introspection of the hot kernel (parity tests): when non-NULL, primary_run stores the two raw 64-bit draws of every
    // site it walks at dbg_raw[((z*N/2 + y)*N + x)*2 ..] (rows it skips as all-masked stay untouched)
    unsigned long long *dbg_raw;
};

struct Mode {
    double Dr, Di;  // density mode D
    double s0, s1, s2;  // F,G,H = i * s_c * D          (reference src/zeldovich.cpp:432-434)
    double f;       // velocity factor                  (reference src/zeldovich.cpp:415)
};

__device__ __forceinline__ int wrap_k(int i, int N, int half) { return i > half ? i - N : i; }

// interp_eigmode (reference src/zeldovich.cpp:154-227); ik* are table-convention indices in [0,N)
__device__ __forceinline__ void interp_eig(const GenParams &g, int ikx, int iky, int ikz, double e[4]) {
    const int pe   = g.pe;
    const int hp1  = pe / 2 + 1;
    const int half = pe / 2;
    const double2 *tab2 = reinterpret_cast<const double2 *>(g.eig);
    auto ld4 = [tab2](size_t i) {
        double2 lo = __ldg(&tab2[2 * i]), hi = __ldg(&tab2[2 * i + 1]);
        return make_double4(lo.x, lo.y, hi.x, hi.y);
    };
    if (g.eig_direct) {
        const int q = pe / g.N;
        double4 v   = ld4(((size_t) (ikx * q) * pe + (size_t) (iky * q)) * hp1 + (size_t) (ikz * q));
        e[0] = v.x, e[1] = v.y, e[2] = v.z, e[3] = v.w;
        return;
    }
    int lo[3], hi[3];
    double fr[3];
    const int ik[3] = {ikx, iky, ikz};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        double f = g.eig_scale * ik[d];
        if (f > half && f < hp1) f = floor(f + 1);  // never interpolate across the Nyquist gap
        lo[d] = (int) f;
        hi[d] = lo[d] + 1;
        if (hi[d] == pe) hi[d] = 0;
        fr[d] = f - lo[d];
    }
    // the z corner one past the stored half axis is only reached with weight exactly 0
    // (reference src/zeldovich.cpp:187-198); clamp so that no out-of-range load is issued
    if (hi[2] > half) hi[2] = half;
    double w[8];
    w[0] = (1 - fr[0]) * (1 - fr[1]) * (1 - fr[2]);
    w[1] = (1 - fr[0]) * (1 - fr[1]) * (fr[2]);
    w[2] = (1 - fr[0]) * (fr[1]) * (1 - fr[2]);
    w[3] = (1 - fr[0]) * (fr[1]) * (fr[2]);
    w[4] = (fr[0]) * (1 - fr[1]) * (1 - fr[2]);
    w[5] = (fr[0]) * (1 - fr[1]) * (fr[2]);
    w[6] = (fr[0]) * (fr[1]) * (1 - fr[2]);
    w[7] = (fr[0]) * (fr[1]) * (fr[2]);
    double4 acc = make_double4(0, 0, 0, 0);
#pragma unroll
    for (int c = 0; c < 8; c++) {
        int cx    = (c & 4) ? hi[0] : lo[0];
        int cy    = (c & 2) ? hi[1] : lo[1];
        int cz    = (c & 1) ? hi[2] : lo[2];
        double4 v = ld4(((size_t) cx * pe + (size_t) cy) * hp1 + (size_t) cz);
        // the reference's left-to-right corner sum, with fused multiply-adds
        acc.x = fma(w[c], v.x, acc.x);
        acc.y = fma(w[c], v.y, acc.y);
        acc.z = fma(w[c], v.z, acc.z);
        acc.w = fma(w[c], v.w, acc.w);
    }
    e[0] = acc.x, e[1] = acc.y, e[2] = acc.z, e[3] = acc.w;
}

// get_eigenmode (reference src/zeldovich.cpp:229-276) folded with the amplitude factor of
// reference :432-434.  The reference forms e.vec = ehat * k2/(k.ehat) and then
// F,G,H = rescale * i * e.vec * fundamental / (k2*fundamental^2) * D; the integer k2 cancels, so
//   s_c = rescale * ehat_c / ((k.ehat) * fundamental)          (one division, one reciprocal sqrt)
// with s = 0 when k = 0 or k.ehat = 0 (the reference's isfinite guard).  ZA: ehat = k, so
// s_c = rescale * k_c / (k2 * fundamental).  Agreement with the reference's order of roundings ~1e-16.
__device__ __forceinline__ void eig_factors(const GenParams &g, int kx, int ky, int kz, int n2, double &s0, double &s1,
                                            double &s2, double &val) {
    if (!g.qPLT) {
        const double q = (n2 == 0) ? 0.0 : 1.0 / ((double) n2 * g.fundamental);
        s0 = kx * q, s1 = ky * q, s2 = kz * q, val = 1.0;
        return;
    }
    int ikx = kx < 0 ? g.N + kx : kx;
    int iky = ky < 0 ? g.N + ky : ky;
    int ikz = kz < 0 ? g.N + kz : kz;
    ikz     = ikz > g.half ? g.N - ikz : ikz;  // stored half-space is +kz
    double eh[4];
    interp_eig(g, ikx, iky, ikz, eh);
    if (kz < 0) eh[2] = -eh[2];
    // |ehat| cancels between the normalisation and k2/(k.ehat): s_c = eh_c / ((k.eh) * fundamental)
    const double dot = kx * eh[0] + ky * eh[1] + kz * eh[2];
    double q = 1.0 / (dot * g.fundamental);
    if (n2 == 0 || !isfinite(q)) q = 0.0;
    s0 = eh[0] * q, s1 = eh[1] * q, s2 = eh[2] * q, val = eh[3];
}

// Mask of reference src/zeldovich.cpp:350-358.  n2 = kx^2+ky^2+kz^2.
__device__ __forceinline__ bool mode_masked(const GenParams &g, int kx, int ky, int kz, int n2) {
    if (abs(kx) == g.kmax || abs(kz) == g.kmax || abs(ky) == g.kmax) return true;
    if (!g.corner_modes && (double) n2 * g.fundamental2 >= g.k2_cutoff) return true;
    if (g.qonemode && !(kx == g.one_mode[0] && ky == g.one_mode[1] && kz == g.one_mode[2])) return true;
    return false;
}

// Generator state positioned at the first of the two draws of the primary mode whose
// lattice indices are (x, y, z), 0 <= y < N/2 (reference nskip
// bookkeeping src/zeldovich.cpp:335,341,358-363): 2*(ky*M^2 + (kz mod M)*M + (kx mod M)).
__device__ __forceinline__ u128 mode_rng_state(const GenParams &g, int x, int y, int z) {
    u128 s = g.ystate[y];
    s      = apply(g.zjump[z], s);
    s      = apply(g.xjump[x], s);
    return s;
}

// cgauss<2> (reference src/power_spectrum.cpp:338-359) given the two uniform deviates
__device__ __forceinline__ void box_muller(const GenParams &g, double P, double u1, double u2, double &Dr, double &Di) {
    double R     = g.fixed_power ? sqrt(P) : sqrt(-P * log(u1));
    // cos/sin(2*pi*u2): sincospi reduces the argument exactly, so this is at least as accurate as
    // the reference's cos(2*M_PI*theta) and cheaper than a radian-argument sincos
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    Dr = R * cs;
    Di = R * sn;
}

// The primary mode at lattice indices (x,y,z), 0 <= y < N/2
// (reference LoadPlane body, src/zeldovich.cpp:331-438).
__device__ __forceinline__ void primary_mode(const GenParams &g, int x, int y, int z, Mode &m) {
    const int kx = wrap_k(x, g.N, g.half), ky = y, kz = wrap_k(z, g.N, g.half);
    const int n2 = kx * kx + ky * ky + kz * kz;
    m.Dr = m.Di = m.s0 = m.s1 = m.s2 = m.f = 0.0;
    if (g.phi != nullptr) {
        if (n2 == 0) return;
        const double2 ph = g.phi[(size_t) z * g.phi_zstride + (size_t) (y >> g.phi_yshift) * g.N + x];
        const double M   = __ldg(&g.mtab[n2]);
        m.Dr = ph.x * M, m.Di = -ph.y * M;
    } else {
        if (mode_masked(g, kx, ky, kz, n2)) return;
        u128 s    = mode_rng_state(g, x, y, z);
        double u1 = u64_to_unit(pcg_next(s));
        double u2 = u64_to_unit(pcg_next(s));
        double P  = __ldg(&g.ptab[n2]);
        box_muller(g, P, u1, u2, m.Dr, m.Di);
    }
    if (m.Dr == 0.0 && m.Di == 0.0) return;  // "D != 0." guard (reference src/zeldovich.cpp:403)
    double s0, s1, s2, val;
    eig_factors(g, kx, ky, kz, n2, s0, s1, s2, val);
    double rescale = 1., f = 1.0;
    if (g.qPLT) {
        f = (sqrt(1. + 24 * val * g.f_cluster) - 1) * .25;
        // pow(a_NL/a0, target_f - f) (reference src/zeldovich.cpp:428) as exp(log(ratio)*(target_f - f)):
        // |exponent| <~ 0.1, so the two agree to a few 1e-16 relative
        if (g.qPLTrescale) rescale = exp(g.log_growth_ratio * (g.target_f - f));
    }
    m.s0 = rescale * s0;
    m.s1 = rescale * s1;
    m.s2 = rescale * s2;
    m.f  = f;
}

// ------------------------------------------------------------------ runs of modes ----
// The fused generation kernel draws RUN consecutive x of one row (y, z) per thread.  Everything
// that depends on (y, z) only is formed once per thread (RowConst); consecutive kx are consecutive
// positions of the generator, so one jump serves the whole run; and when the RUN
// sites fall into one cell of the eigenmode table (always, for ppd >= RUN * ppd_e, except next to the
// table's Nyquist gap) the eight corners are loaded once and reduced along y and z before the
// per-site x interpolation.  The RUN sites are independent straight-line code: the instruction-level
// parallelism hides the FP64 latency that one mode per thread exposes.
__device__ __forceinline__ void eig_axis(const GenParams &g, int ik, int &lo, int &hi, double &fr) {
    const int half = g.pe / 2;
    double f = g.eig_scale * ik;
    if (f > half && f < half + 1) f = floor(f + 1);  // never interpolate across the Nyquist gap
    lo = (int) f;
    hi = lo + 1;
    if (hi == g.pe) hi = 0;
    fr = f - lo;
}

struct RowConst {
    int ky, kz, n2yz;
    u128 s_yz;                    // generator state of the row before the jump along x
    int ylo, yhi, zlo, zhi;       // eigenmode cell along y and z (interpolated tables only)
    double wyz[4];                // bilinear weights (ylo,zlo) (ylo,zhi) (yhi,zlo) (yhi,zhi)
};

__device__ __forceinline__ RowConst row_const(const GenParams &g, int y, int z) {
    RowConst rc;
    rc.ky   = y;
    rc.kz   = wrap_k(z, g.N, g.half);
    rc.n2yz = rc.ky * rc.ky + rc.kz * rc.kz;
    rc.s_yz = apply(g.zjump[z], g.ystate[y]);
    rc.ylo = rc.yhi = rc.zlo = rc.zhi = 0;
    rc.wyz[0] = rc.wyz[1] = rc.wyz[2] = rc.wyz[3] = 0.0;
    if (g.qPLT && !g.eig_direct) {
        double fy, fz;
        eig_axis(g, y, rc.ylo, rc.yhi, fy);
        eig_axis(g, rc.kz < 0 ? -rc.kz : rc.kz, rc.zlo, rc.zhi, fz);  // stored half-space is +kz
        if (rc.zhi > g.pe / 2) rc.zhi = g.pe / 2;                       // weight-0 corner, see interp_eig
        rc.wyz[0] = (1 - fy) * (1 - fz);
        rc.wyz[1] = (1 - fy) * fz;
        rc.wyz[2] = fy * (1 - fz);
        rc.wyz[3] = fy * fz;
    }
    return rc;
}

// xj0 = g.xjump[x0], loaded by the caller ahead of time (its latency hides behind the row's mask test)
template <int RUN>
__device__ __forceinline__ void primary_run(const GenParams &g, const RowConst &rc, int x0, const Affine &xj0, double (&Dr)[RUN],
                                            double (&Di)[RUN], double (&s0)[RUN], double (&s1)[RUN], double (&s2)[RUN], double (&ff)[RUN]) {
    const int N = g.N, half = g.half, ky = rc.ky, kz = rc.kz;
    int kx[RUN], n2[RUN];
    bool act[RUN], any = false;
    const bool from_phi = g.phi != nullptr;  // ZD_f_NL: no mode is masked, the density comes from the potential
#pragma unroll
    for (int j = 0; j < RUN; j++) {
        kx[j]  = wrap_k(x0 + j, N, half);
        n2[j]  = kx[j] * kx[j] + rc.n2yz;
        act[j] = from_phi || !mode_masked(g, kx[j], ky, kz, n2[j]);
        any |= act[j];
        Dr[j] = Di[j] = s0[j] = s1[j] = s2[j] = ff[j] = 0.0;
    }
    if (!any) return;
    if (from_phi) {
        const double2 *ph = g.phi + (size_t) (kz < 0 ? kz + N : kz) * g.phi_zstride + (size_t) (ky >> g.phi_yshift) * N + x0;
#pragma unroll
        for (int j = 0; j < RUN; j++) {
            const double2 v = ph[j];
            const double M  = __ldg(&g.mtab[n2[j]]);
            Dr[j] = (n2[j] == 0) ? 0.0 : v.x * M;
            Di[j] = (n2[j] == 0) ? 0.0 : -v.y * M;
        }
    } else {
    // the draws: masked sites consume theirs too, so the run is one walk of the generator.  The only
    // break is between x = N/2 (kx = +N/2) and x = N/2+1 (kx = -N/2+1); runs are aligned, so it can
    // only sit between the first and the second site of a run.
    double u1[RUN], u2[RUN], Pk[RUN];
#pragma unroll
    for (int j = 0; j < RUN; j++) Pk[j] = __ldg(&g.ptab[n2[j]]);  // issued before the generator arithmetic that hides their latency
    {
        u128 s = apply(xj0, rc.s_yz);
#pragma unroll
        for (int j = 0; j < RUN; j++) {
            if (j == 1 && x0 == half) s = apply(g.xjump[x0 + 1], rc.s_yz);
            const uint64_t r1 = pcg_next(s), r2 = pcg_next(s);
            u1[j] = u64_to_unit(r1);
            u2[j] = u64_to_unit(r2);
            if (g.dbg_raw != nullptr) {
                unsigned long long *o = g.dbg_raw + (((size_t) (kz < 0 ? kz + N : kz) * half + ky) * N + (x0 + j)) * 2;
                o[0] = r1, o[1] = r2;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < RUN; j++) box_muller(g, Pk[j], u1[j], u2[j], Dr[j], Di[j]);
    }  // draws

    if (!g.qPLT) {
#pragma unroll
        for (int j = 0; j < RUN; j++) {
            const double q = (n2[j] == 0) ? 0.0 : 1.0 / ((double) n2[j] * g.fundamental);
            s0[j] = kx[j] * q, s1[j] = ky * q, s2[j] = kz * q, ff[j] = 1.0;
        }
    } else {
        double eh[RUN][4];
        const int pe = g.pe, hp1 = pe / 2 + 1;
        const double2 *tab2 = reinterpret_cast<const double2 *>(g.eig);
        auto ld4 = [tab2](size_t i) {
            double2 lo = __ldg(&tab2[2 * i]), hi = __ldg(&tab2[2 * i + 1]);
            return make_double4(lo.x, lo.y, hi.x, hi.y);
        };
        const int ikz = kz < 0 ? -kz : kz;
        if (g.eig_direct) {
            const int q = pe / N;
#pragma unroll
            for (int j = 0; j < RUN; j++) {
                double4 v = ld4(((size_t) ((x0 + j) * q) * pe + (size_t) (ky * q)) * hp1 + (size_t) (ikz * q));
                eh[j][0] = v.x, eh[j][1] = v.y, eh[j][2] = v.z, eh[j][3] = v.w;
            }
        } else {
            int xlo[RUN], xhi[RUN];
            double fx[RUN];
            bool same = true;
#pragma unroll
            for (int j = 0; j < RUN; j++) {
                eig_axis(g, x0 + j, xlo[j], xhi[j], fx[j]);
                same = same && (xlo[j] == xlo[0]);
            }
            if (same) {
                // one cell: reduce the corners along y and z once, interpolate along x per site
                double4 elo = make_double4(0, 0, 0, 0), ehi = make_double4(0, 0, 0, 0);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int cy = (c & 2) ? rc.yhi : rc.ylo, cz = (c & 1) ? rc.zhi : rc.zlo;
                    const double4 a = ld4(((size_t) xlo[0] * pe + (size_t) cy) * hp1 + (size_t) cz);
                    const double4 b = ld4(((size_t) xhi[0] * pe + (size_t) cy) * hp1 + (size_t) cz);
                    const double w  = rc.wyz[c];
                    elo.x = fma(w, a.x, elo.x), elo.y = fma(w, a.y, elo.y), elo.z = fma(w, a.z, elo.z), elo.w = fma(w, a.w, elo.w);
                    ehi.x = fma(w, b.x, ehi.x), ehi.y = fma(w, b.y, ehi.y), ehi.z = fma(w, b.z, ehi.z), ehi.w = fma(w, b.w, ehi.w);
                }
#pragma unroll
                for (int j = 0; j < RUN; j++) {
                    const double wl = 1 - fx[j], wh = fx[j];
                    eh[j][0] = fma(wh, ehi.x, wl * elo.x);
                    eh[j][1] = fma(wh, ehi.y, wl * elo.y);
                    eh[j][2] = fma(wh, ehi.z, wl * elo.z);
                    eh[j][3] = fma(wh, ehi.w, wl * elo.w);
                }
            } else {
#pragma unroll
                for (int j = 0; j < RUN; j++) interp_eig(g, x0 + j, ky, ikz, eh[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < RUN; j++) {
            if (kz < 0) eh[j][2] = -eh[j][2];
            const double dot = kx[j] * eh[j][0] + ky * eh[j][1] + kz * eh[j][2];
            double q = 1.0 / (dot * g.fundamental);
            if (n2[j] == 0 || !isfinite(q)) q = 0.0;
            const double f = (sqrt(1. + 24 * eh[j][3] * g.f_cluster) - 1) * .25;
            const double rescale = g.qPLTrescale ? exp(g.log_growth_ratio * (g.target_f - f)) : 1.0;
            q *= rescale;
            s0[j] = eh[j][0] * q, s1[j] = eh[j][1] * q, s2[j] = eh[j][2] * q, ff[j] = f;
        }
    }
    // masked sites, and the reference's "D != 0." guard (src/zeldovich.cpp:403): everything is zero
#pragma unroll
    for (int j = 0; j < RUN; j++) {
        if (!act[j] || (Dr[j] == 0.0 && Di[j] == 0.0)) Dr[j] = Di[j] = s0[j] = s1[j] = s2[j] = ff[j] = 0.0;
    }
}

// The Gaussian density mode alone (mask + draw), for the phi-generation pass of ZD_f_NL
// (reference src/zeldovich.cpp:350-375 followed by :388-394).
__device__ __forceinline__ void primary_density(const GenParams &g, int x, int y, int z, double &Dr, double &Di, int &n2) {
    const int kx = wrap_k(x, g.N, g.half), ky = y, kz = wrap_k(z, g.N, g.half);
    n2 = kx * kx + ky * ky + kz * kz;
    Dr = Di = 0.0;
    if (mode_masked(g, kx, ky, kz, n2)) return;
    u128 s    = mode_rng_state(g, x, y, z);
    double u1 = u64_to_unit(pcg_next(s));
    double u2 = u64_to_unit(pcg_next(s));
    box_muller(g, __ldg(&g.ptab[n2]), u1, u2, Dr, Di);
}

// Packed entries A0..A3 for the primary site and for its conjugate-structured twin
// (reference src/zeldovich.cpp:447-466).
__device__ __forceinline__ void pack_primary(const Mode &m, cplx a[4]) {
    const double Fr = -m.s0 * m.Di, Fi = m.s0 * m.Dr;
    const double Gr = -m.s1 * m.Di, Gi = m.s1 * m.Dr;
    const double Hr = -m.s2 * m.Di, Hi = m.s2 * m.Dr;
    const double f = m.f;
    a[0] = make_double2(m.Dr - Fi, m.Di + Fr);
    a[1] = make_double2(Gr - Hi, Gi + Hr);
    a[2] = make_double2(0. - Fi * f, 0. + Fr * f);
    a[3] = make_double2(Gr * f - Hi * f, Gi * f + Hr * f);
}
__device__ __forceinline__ void pack_twin(const Mode &m, cplx a[4]) {
    const double Fr = -m.s0 * m.Di, Fi = m.s0 * m.Dr;
    const double Gr = -m.s1 * m.Di, Gi = m.s1 * m.Dr;
    const double Hr = -m.s2 * m.Di, Hi = m.s2 * m.Dr;
    const double f = m.f;
    a[0] = make_double2(m.Dr + Fi, -m.Di + Fr);
    a[1] = make_double2(Gr + Hi, -Gi + Hr);
    a[2] = make_double2(0. + Fi * f, 0. + Fr * f);
    a[3] = make_double2(Gr * f + Hi * f, -(Gi * f) + Hr * f);
}

}  // namespace zplt
