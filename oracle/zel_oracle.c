/* TEST INFRASTRUCTURE — CPU oracle for the zeldovich-PLT IC hot path.
 *
 * This file is NOT product code: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it, and only as the
 * checker.  The product path is the CUDA library and never calls into this.
 *
 * It restates, in plain C and in closed form (one independent function of
 * (x,y,z) per lattice mode instead of the reference's sequential plane walk),
 * the algorithm of the reference hot path.  Parity pinning: this restatement is
 * checked in tests/test_oracle.py against (a) the PCG64 known answers of
 * SURVEY.md §4, (b) golden ic_* records under tests/golden/ that were produced
 * by the UNMODIFIED reference sources compiled here (oracle/_ref, see
 * oracle/Makefile and tests/golden/make_golden.py) and (c) oracle/_ref itself,
 * live, when that binary is present.
 *
 * Reference citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

/* ---------------------------------------------------------------- PCG64 ---- */
/* pcg64 = setseq_xsl_rr_128_64 (include/pcg-rng/pcg_random.hpp:1965,1868);
 * constants :159-170; bump :370; seeding ctor :427-432; output of the
 * post-advance state (output_previous=false, :855) via xsl_rr_mixin :1144-1170;
 * advance :664-686. */
#define ZO_MAKE128(hi, lo) ((((u128) (hi)) << 64) | (u128) (lo))
static const u128 ZO_MULT = ZO_MAKE128(2549297995355413924ULL, 4865540595714422341ULL);
static const u128 ZO_INC  = ZO_MAKE128(6364136223846793005ULL, 1442695040888963407ULL);

static u128 zo_seed_state(uint64_t seed) { return ((u128) seed + ZO_INC) * ZO_MULT + ZO_INC; }

static u128 zo_jump(u128 state, u128 delta) {
    u128 cm = ZO_MULT, cp = ZO_INC, am = 1, ap = 0;
    while (delta > 0) {
        if (delta & 1) {
            am *= cm;
            ap = ap * cm + cp;
        }
        cp = (cm + 1) * cp;
        cm *= cm;
        delta >>= 1;
    }
    return am * state + ap;
}

static inline uint64_t zo_next(u128 *state) {
    *state          = *state * ZO_MULT + ZO_INC;
    uint64_t hi     = (uint64_t) (*state >> 64);
    uint64_t lo     = (uint64_t) *state;
    uint64_t x      = hi ^ lo;
    unsigned int rot = (unsigned int) (hi >> 58);
    return (x >> rot) | (x << ((64 - rot) & 63));
}

/* worker threads of the OpenMP loops below (a launcher such as torchrun sets OMP_NUM_THREADS=1 for its ranks) */
void zo_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void) n;
#endif
}

/* n raw 64-bit outputs starting `offset` (= off_hi*2^64 + off_lo) draws after seeding */
void zo_pcg_draws(uint64_t seed, uint64_t off_hi, uint64_t off_lo, int64_t n, uint64_t *out) {
    u128 s = zo_jump(zo_seed_state(seed), ZO_MAKE128(off_hi, off_lo));
    for (int64_t i = 0; i < n; i++) out[i] = zo_next(&s);
}

/* one_rand<2> (src/power_spectrum.cpp:284-308): uint64 -> (0,1] */
double zo_one_rand(uint64_t r) {
    if (r == UINT64_MAX) return 1.0;
    r += 1;
    return ldexp((double) r, -64);
}

/* ---------------------------------------------------------------- spline --- */
/* SplineFunction (include/spline_function.h:54-163): nodes kept in increasing
 * abscissa, natural cubic spline second derivatives, bisection + cubic evaluation
 * (which extrapolates with the end interval). */
typedef struct {
    int n;
    double *x, *y, *y2;
} zo_spline;

static int zo_cmp_node(const void *a, const void *b) {
    double u = ((const double *) a)[0], v = ((const double *) b)[0];
    return (u > v) - (u < v);
}

static void zo_spline_build(zo_spline *s, int n, const double *xs, const double *ys) {
    s->n  = n;
    s->x  = (double *) malloc(sizeof(double) * n);
    s->y  = (double *) malloc(sizeof(double) * n);
    s->y2 = (double *) malloc(sizeof(double) * n);
    double *nodes = (double *) malloc(sizeof(double) * 2 * n);
    for (int i = 0; i < n; i++) {
        nodes[2 * i]     = xs[i];
        nodes[2 * i + 1] = ys[i];
    }
    qsort(nodes, n, 2 * sizeof(double), zo_cmp_node); /* abscissae are distinct */
    for (int i = 0; i < n; i++) {
        s->x[i] = nodes[2 * i];
        s->y[i] = nodes[2 * i + 1];
    }
    free(nodes);
    /* tridiagonal sweep for a natural spline (spline_function.h:105-139) */
    double *u = (double *) malloc(sizeof(double) * n);
    double *x = s->x, *y = s->y, *y2 = s->y2;
    y2[0] = u[0] = 0.0;
    for (int i = 1; i <= n - 2; i++) {
        double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
        double p   = sig * y2[i - 1] + 2.0;
        y2[i]      = (sig - 1.0) / p;
        u[i]       = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
        u[i]       = (6.0 * u[i] / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p;
    }
    double qn = 0.0, un = 0.0;
    y2[n - 1] = (un - qn * u[n - 2]) / (qn * y2[n - 2] + 1.0);
    for (int k = n - 2; k >= 0; k--) y2[k] = y2[k] * y2[k + 1] + u[k];
    free(u);
}

static void zo_spline_free(zo_spline *s) {
    free(s->x);
    free(s->y);
    free(s->y2);
}

static double zo_spline_val(const zo_spline *s, double v) {
    int klo = 0, khi = s->n - 1;
    while (khi - klo > 1) {
        int k = (khi + klo) >> 1;
        if (s->x[k] > v)
            khi = k;
        else
            klo = k;
    }
    double h = s->x[khi] - s->x[klo];
    double a = (s->x[khi] - v) / h;
    double b = (v - s->x[klo]) / h;
    return a * s->y[klo] + b * s->y[khi] + ((a * a * a - a) * s->y2[klo] + (b * b * b - b) * s->y2[khi]) * (h * h) / 6.0;
}

/* ------------------------------------------------------------ configuration */
typedef struct {
    int64_t ppd;
    double boxsize;
    int64_t seed; /* the reference holds an int and widens it to unsigned long (power_spectrum.cpp:14) */
    double k_cutoff;
    int corner_modes;
    int qonemode;
    int one_mode[3];
    int qPLT;
    int qPLTrescale;
    double PLT_target_z;
    double z_initial;
    double f_cluster;
    int fixed_power;
    /* power spectrum */
    int is_powerlaw;
    double powerlaw_index;
    double Pk_norm, Pk_sigma, Pk_sigma_ratio, Pk_smooth, Pk_scale;
    /* output */
    int icformat; /* 0 Zeldovich, 1 RVZel, 2 RVdoubleZel, 3 ZelSimple (enum order of include/output.h:44-49) */
    /* local primordial non-Gaussianity (include/parameters.h:59-61) */
    double f_NL, n_s, Omega_M;
} zo_config;

typedef struct {
    zo_config c;
    zo_spline sp;
    int have_spline;
    double normalization, Pk_smooth2;
    double Rnorm;
    int64_t eig_ppd;
    const double *eig;
    /* f_NL: smallest positive k of the input table and the primordial normalisation
     * (src/power_spectrum.cpp:160,180,221-222); phi = the transformed potential, [z][y][x] complex, or NULL */
    double kmin, primordial_norm;
    const double *phi;
} zo_state;

/* ------------------------------------------------------------ P(k) --------- */
/* PowerSpectrum::power (src/power_spectrum.cpp:225-261) */
static double zo_power(const zo_state *st, double k) {
    if (k <= 0.0) return 0.0;
    if (st->c.is_powerlaw) return pow(k, st->c.powerlaw_index) * exp(-k * k * st->Pk_smooth2) * st->normalization;
    return exp(zo_spline_val(&st->sp, log(k)) - k * k * st->Pk_smooth2) * st->normalization;
}

/* sigmaR_integrand (src/power_spectrum.cpp:50-58) */
static double zo_sig_integrand(const zo_state *st, double k) {
    double x = k * st->Rnorm;
    double w;
    if (x <= 1e-3)
        w = 1 - x * x / 10.0;
    else
        w = 3.0 * (sin(x) - x * cos(x)) / x / x / x;
    return 0.5 / M_PI / M_PI * k * k * w * w * zo_power(st, k);
}

/* Romberg (src/power_spectrum.cpp:94-128): trapezoid refinements with Richardson
 * extrapolation, at most 32 halvings, stops on relative change < prec. */
#define ZO_MAXIT 32
static double zo_romberg(const zo_state *st, double a, double b, double prec, double *obt) {
    static double T[ZO_MAXIT + 1][ZO_MAXIT + 1];
    double h = 0.5 * (b - a);
    T[0][1]  = h * (zo_sig_integrand(st, a) + zo_sig_integrand(st, b));
    int j    = 0;
    do {
        j++;
        double s      = 0;
        uint64_t npts = 1ULL << (j - 1);
        for (uint64_t k = 1; k <= npts; k++) s += zo_sig_integrand(st, a + (2 * k - 1) * h);
        T[j][1]    = 0.5 * T[j - 1][1] + h * s;
        double f4 = 1;
        for (int k = 2; k <= j; k++) {
            f4 *= 4;
            T[j][k] = T[j][k - 1] + (T[j][k - 1] - T[j - 1][k - 1]) / (f4 - 1);
        }
        h *= 0.5;
        if (j > 1 && fabs(T[j][j] - T[j - 1][j - 1]) < prec * fabs(T[j][j])) break;
    } while (j < ZO_MAXIT);
    *obt = (T[j][j] - T[j - 1][j - 1]) / T[j][j];
    return T[j][j];
}

/* sigmaR (src/power_spectrum.cpp:60-89) */
static double zo_sigmaR(zo_state *st, double R) {
    if (!st->c.is_powerlaw) {
        double got;
        st->Rnorm = R;
        return sqrt(zo_romberg(st, 0, 10.0, 1e-6, &got));
    }
    double n = st->c.powerlaw_index;
    double r = 9 * pow(R, -n - 3) / (2 * M_PI * sqrt(M_PI)) * tgamma((3 + n) / 2.) / (tgamma((2 - n) / 2.) * (n - 3) * (n - 1));
    return sqrt(r * st->normalization);
}

/* InitFromFile + Normalize (src/power_spectrum.cpp:130-171, :186-223).
 * ks/ps: the raw table rows as sscanf("%lf %lf") delivers them. */
static void zo_power_setup(zo_state *st, int nrows, const double *ks, const double *ps) {
    st->have_spline = 0;
    st->kmin        = 1e-4; /* power law: "arbitrary; used by f_NL" (src/power_spectrum.cpp:180) */
    if (!st->c.is_powerlaw) {
        double *xs = (double *) malloc(sizeof(double) * nrows), *ys = (double *) malloc(sizeof(double) * nrows);
        int n = 0;
        st->kmin = 1.7976931348623157e308;
        for (int i = 0; i < nrows; i++) {
            double k = ks[i], P = ps[i];
            if (k < 0.0 || P < 0.0) continue;
            k *= st->c.Pk_scale;
            if (k > 0.0 && k < st->kmin) st->kmin = k;
            xs[n] = (k > 0.0) ? log(k) : -1e3;
            ys[n] = log(P);
            n++;
        }
        zo_spline_build(&st->sp, n, xs, ys);
        st->have_spline = 1;
        free(xs);
        free(ys);
    }
    st->Pk_smooth2    = 0.0;
    st->normalization = 1.0;
    if (st->c.Pk_norm > 0.0) {
        if (st->c.Pk_sigma > 0) {
            st->normalization = st->c.Pk_sigma / zo_sigmaR(st, st->c.Pk_norm);
            st->normalization *= st->normalization;
        } else if (st->c.Pk_sigma_ratio > 0) {
            st->normalization = st->c.Pk_sigma_ratio * st->c.Pk_sigma_ratio;
        }
    }
    st->normalization /= st->c.boxsize * st->c.boxsize * st->c.boxsize;
    st->Pk_smooth2 = st->c.Pk_smooth * st->c.Pk_smooth;
    /* src/power_spectrum.cpp:221-222: T(k) = 1 at the smallest k of the table */
    st->primordial_norm = zo_power(st, st->kmin) / exp(log(st->kmin) * st->c.n_s);
}

/* PowerSpectrum::infer_Tk (src/power_spectrum.cpp:268-274) for boundary tests of the f_NL scalars */
double zo_infer_Tk(const zo_config *cfg, int nrows, const double *ks, const double *ps, double k) {
    zo_state st;
    memset(&st, 0, sizeof(st));
    st.c = *cfg;
    zo_power_setup(&st, nrows, ks, ps);
    double Tk = (k <= 0.0) ? 1.0 : sqrt(zo_power(&st, k) / (st.primordial_norm * exp(log(k) * cfg->n_s)));
    if (st.have_spline) zo_spline_free(&st.sp);
    return Tk;
}

/* host scalars for boundary tests: out = {normalization, Pk_smooth2, sigmaR(Pk_norm) after normalisation * L^1.5} */
void zo_power_scalars(const zo_config *cfg, int nrows, const double *ks, const double *ps, double *out) {
    zo_state st;
    memset(&st, 0, sizeof(st));
    st.c = *cfg;
    zo_power_setup(&st, nrows, ks, ps);
    out[0] = st.normalization;
    out[1] = st.Pk_smooth2;
    out[2] = (cfg->Pk_norm > 0) ? zo_sigmaR(&st, cfg->Pk_norm) * pow(cfg->boxsize, 1.5) : 0.0;
    if (st.have_spline) zo_spline_free(&st.sp);
}

/* P(k) evaluated at k = sqrt(m * fundamental^2), m = 0..count-1 (what the mode loop feeds to power) */
void zo_power_table(const zo_config *cfg, int nrows, const double *ks, const double *ps, int64_t count, double *out) {
    zo_state st;
    memset(&st, 0, sizeof(st));
    st.c = *cfg;
    zo_power_setup(&st, nrows, ks, ps);
    double fund  = 2.0 * M_PI / cfg->boxsize;
    double fund2 = fund * fund;
    for (int64_t m = 0; m < count; m++) out[m] = zo_power(&st, sqrt((double) m * fund2));
    if (st.have_spline) zo_spline_free(&st.sp);
}

/* ------------------------------------------------------------ eigenmodes --- */
#define ZO_EIG(st, kx, ky, kz, i) \
    ((st)->eig[(int64_t) (kx) * (st)->eig_ppd * ((st)->eig_ppd / 2 + 1) * 4 + (int64_t) (ky) * ((st)->eig_ppd / 2 + 1) * 4 + (int64_t) (kz) * 4 + (i)])

/* interp_eigmode (src/zeldovich.cpp:154-227) */
static void zo_interp_eig(const zo_state *st, int ikx, int iky, int ikz, int64_t ppd, double e[4]) {
    int64_t pe   = st->eig_ppd;
    int64_t hp1  = pe / 2 + 1;
    int64_t half = pe / 2;
    if (pe % ppd == 0) {
        for (int i = 0; i < 4; i++) e[i] = ZO_EIG(st, ikx * pe / ppd, iky * pe / ppd, ikz * pe / ppd, i);
        return;
    }
    double fr[3];
    int lo[3], hi[3];
    int ik[3] = {ikx, iky, ikz};
    for (int d = 0; d < 3; d++) {
        double f = ((double) pe) / ppd * ik[d];
        if (f > half && f < hp1) f = floor(f + 1); /* never interpolate across the Nyquist gap */
        lo[d] = (int) f;
        hi[d] = lo[d] + 1;
        if (hi[d] == pe) hi[d] = 0;
        fr[d] = f - lo[d];
    }
    double w[8];
    w[0] = (1 - fr[0]) * (1 - fr[1]) * (1 - fr[2]);
    w[1] = (1 - fr[0]) * (1 - fr[1]) * (fr[2]);
    w[2] = (1 - fr[0]) * (fr[1]) * (1 - fr[2]);
    w[3] = (1 - fr[0]) * (fr[1]) * (fr[2]);
    w[4] = (fr[0]) * (1 - fr[1]) * (1 - fr[2]);
    w[5] = (fr[0]) * (1 - fr[1]) * (fr[2]);
    w[6] = (fr[0]) * (fr[1]) * (1 - fr[2]);
    w[7] = (fr[0]) * (fr[1]) * (fr[2]);
    for (int i = 0; i < 4; i++) {
        double acc = 0.0;
        for (int c = 0; c < 8; c++) {
            int cx = (c & 4) ? hi[0] : lo[0];
            int cy = (c & 2) ? hi[1] : lo[1];
            int cz = (c & 1) ? hi[2] : lo[2];
            /* a corner one past the stored half axis is only ever read with weight 0
             * (src/zeldovich.cpp:187-198); do not touch memory for it */
            double t = (w[c] == 0.0) ? 0.0 : w[c] * ZO_EIG(st, cx, cy, cz, i);
            acc      = (c == 0) ? t : acc + t;
        }
        e[i] = acc;
    }
}

/* get_eigenmode (src/zeldovich.cpp:229-276): out = {vx, vy, vz, val} */
static void zo_get_eig(const zo_state *st, int kx, int ky, int kz, int64_t ppd, double out[4]) {
    if (!st->c.qPLT) {
        out[0] = kx;
        out[1] = ky;
        out[2] = kz;
        out[3] = 1;
        return;
    }
    int ikx = kx < 0 ? (int) (ppd + kx) : kx;
    int iky = ky < 0 ? (int) (ppd + ky) : ky;
    int ikz = kz < 0 ? (int) (ppd + kz) : kz;
    ikz     = ikz > ppd / 2 ? (int) (ppd - ikz) : ikz;
    double k2 = (double) (kx * kx + ky * ky + kz * kz);
    double eh[4];
    zo_interp_eig(st, ikx, iky, ikz, ppd, eh);
    eh[2] *= copysign(1.0, (double) kz);
    double mag = sqrt(eh[0] * eh[0] + eh[1] * eh[1] + eh[2] * eh[2]);
    eh[0] /= mag;
    eh[1] /= mag;
    eh[2] /= mag;
    double norm = k2 / (kx * eh[0] + ky * eh[1] + kz * eh[2]);
    if (k2 == 0.0 || !isfinite(norm)) norm = 0.0;
    out[0] = norm * eh[0];
    out[1] = norm * eh[1];
    out[2] = norm * eh[2];
    out[3] = eh[3];
}

/* ------------------------------------------------------------ one mode ----- */
typedef struct {
    double Dr, Di;   /* density mode */
    double s[3];     /* F,G,H = i * s[c] * D   (src/zeldovich.cpp:432-434) */
    double f;        /* velocity growth factor f (src/zeldovich.cpp:415) */
    double M;        /* f_NL only: potential -> density factor of this mode */
} zo_mode;

static inline int zo_wrap(int64_t i, int64_t n) { return (int) (i > n / 2 ? i - n : i); }

/* primordial_power / infer_Tk (src/power_spectrum.cpp:263-274) and the M(k, a) factor between the Bardeen
 * potential and the density (src/zeldovich.cpp:377-386; 1108.5512 eq. 50).  k2 has the origin's 0 replaced by 1. */
static double zo_Mfactor(const zo_state *st, double kmag, double k2) {
    double Tk = 1.0;
    if (kmag > 0.0) Tk = sqrt(zo_power(st, kmag) / (st->primordial_norm * exp(log(kmag) * st->c.n_s)));
    double H0 = 100., c = 299792.458;
    double growth = 1. / (1 + st->c.z_initial);
    return 2. * growth * c * c * Tk * k2 / (3. * st->c.Omega_M * H0 * H0);
}

/* mask of src/zeldovich.cpp:350-358 for the signed integer wavevector (kx,ky,kz); k2 = |k|^2 in physical units */
static int zo_masked(const zo_config *c, int kx, int ky, int kz, double k2) {
    const int64_t N = c->ppd;
    double sep    = c->boxsize / N;
    double nyq    = M_PI / sep;
    double k2cut  = nyq * nyq / (c->k_cutoff * c->k_cutoff);
    double ikcut  = 1.0 / c->k_cutoff;
    int kmax      = (int) ((double) (N / 2) * ikcut + .5);
    return (abs(kx) == kmax || abs(kz) == kmax || abs(ky) == kmax) || (!c->corner_modes && k2 >= k2cut)
           || (c->qonemode && !(kx == c->one_mode[0] && ky == c->one_mode[1] && kz == c->one_mode[2]));
}

/* cgauss<2> (src/power_spectrum.cpp:338-359) from the mode's two raw draws and P(k) */
static void zo_cgauss(const zo_config *c, double P, uint64_t r1, uint64_t r2, zo_mode *m) {
    double R = zo_one_rand(r1);
    double t = zo_one_rand(r2);
    if (!c->fixed_power)
        R = sqrt(-P * log(R));
    else
        R = sqrt(P);
    t     = 2 * M_PI * t;
    m->Dr = R * cos(t);
    m->Di = R * sin(t);
}

/* eigenmode + growth + rescale (src/zeldovich.cpp:404-434) of a mode whose density D is already in m */
static void zo_finish_mode(const zo_state *st, int kx, int ky, int kz, double k2, zo_mode *m) {
    const zo_config *c = &st->c;
    double fund = 2.0 * M_PI / c->boxsize;
    double ik2  = 1. / k2;
    double e[4];
    zo_get_eig(st, kx, ky, kz, c->ppd, e);
    double rescale = 1., f = 1.0;
    if (c->qPLT) {
        f = (sqrt(1. + 24 * e[3] * c->f_cluster) - 1) * .25;
        if (c->qPLTrescale) {
            double target_f = (sqrt(1. + 24 * c->f_cluster) - 1) / 4.;
            double a_NL     = 1. / (1 + c->PLT_target_z);
            double a0       = 1. / (1 + c->z_initial);
            rescale         = pow(a_NL / a0, target_f - f);
        }
    }
    for (int d = 0; d < 3; d++) m->s[d] = rescale * e[d] * fund * ik2;
    m->f = f;
}

/* The primary mode at signed integer wavevector (kx,ky,kz), ky in [0, ppd/2):
 * mask (src/zeldovich.cpp:350-358), RNG position (SURVEY.md A.2 = the nskip
 * bookkeeping of :335,341,358-363), cgauss<2> (src/power_spectrum.cpp:338-359),
 * eigenmode + growth + rescale (:404-434). */
static void zo_primary(const zo_state *st, u128 seed_state, int kx, int ky, int kz, zo_mode *m) {
    const zo_config *c = &st->c;
    const int64_t N = c->ppd, M = 65536;
    memset(m, 0, sizeof(*m));
    double fund   = 2.0 * M_PI / c->boxsize;
    double fund2  = fund * fund;
    double k2     = (kx * kx + ky * ky + kz * kz) * fund2;
    double kmag   = sqrt(k2);
    if (!zo_masked(c, kx, ky, kz, k2)) {
        u128 off = 2 * ((u128) ky * M * M + (u128) (kz < 0 ? kz + M : kz) * M + (u128) (kx < 0 ? kx + M : kx));
        u128 s   = zo_jump(seed_state, off);
        double P = zo_power(st, kmag);
        uint64_t r1 = zo_next(&s);
        uint64_t r2 = zo_next(&s);
        zo_cgauss(c, P, r1, r2, m);
    }
    if (k2 == 0.0) k2 = 1.0;
    if (c->f_NL != 0.) {
        /* src/zeldovich.cpp:377-400: with an input potential the density of EVERY mode but the origin, masked or
         * not, is phi(k) M(k); in the phi-generation pass (st->phi == NULL) the caller divides D by M */
        m->M = zo_Mfactor(st, kmag, k2);
        if (st->phi) {
            if (kx == 0 && ky == 0 && kz == 0) {
                m->Dr = m->Di = 0.0;
            } else {
                int64_t x = kx < 0 ? kx + N : kx, y = ky, z = kz < 0 ? kz + N : kz;
                const double *ph = st->phi + 2 * ((z * N + y) * N + x);
                m->Dr = ph[0] * m->M;
                m->Di = ph[1] * m->M;
            }
        }
    }
    if (m->Dr == 0.0 && m->Di == 0.0) return; /* "D != 0." guard, src/zeldovich.cpp:403 */
    if (c->f_NL != 0. && !st->phi) return;     /* phi-generation pass: only D and M are used (:388-394) */
    zo_finish_mode(st, kx, ky, kz, k2, m);
}

/* packed entries A0..A3 of a mode, primary or conjugate-structured twin form (src/zeldovich.cpp:447-466) */
static void zo_pack(const zo_mode *mp, int conj, double out[8]) {
    const zo_mode m = *mp;
    /* F = i s0 D etc. */
    double Fr = -m.s[0] * m.Di, Fi = m.s[0] * m.Dr;
    double Gr = -m.s[1] * m.Di, Gi = m.s[1] * m.Dr;
    double Hr = -m.s[2] * m.Di, Hi = m.s[2] * m.Dr;
    double f = m.f;
    if (!conj) {
        out[0] = m.Dr - Fi;
        out[1] = m.Di + Fr;
        out[2] = Gr - Hi;
        out[3] = Gi + Hr;
        out[4] = 0. - Fi * f;
        out[5] = 0. + Fr * f;
        out[6] = Gr * f - Hi * f;
        out[7] = Gi * f + Hr * f;
    } else {
        out[0] = m.Dr + Fi;
        out[1] = -m.Di + Fr;
        out[2] = Gr + Hi;
        out[3] = -Gi + Hr;
        out[4] = 0. + Fi * f;
        out[5] = 0. + Fr * f;
        out[6] = Gr * f + Hi * f;
        out[7] = -(Gi * f) + Hr * f;
    }
}

/* The four packed arrays' entries at lattice site (x,y,z) of the spectral cube
 * (src/zeldovich.cpp:447-466 packing, :485-503 ky=0 plane,
 * src/block_array.cpp:487-491 + src/zeldovich.cpp:644-650 y-shift and Nyquist row). */
static void zo_site(const zo_state *st, u128 seed_state, int64_t x, int64_t y, int64_t z, double out[8]) {
    const int64_t N = st->c.ppd;
    for (int i = 0; i < 8; i++) out[i] = 0.0;
    if (y == N / 2) return;
    if (x == 0 && y == 0 && z == 0) return;
    int kx = zo_wrap(x, N), ky = zo_wrap(y, N), kz = zo_wrap(z, N);
    int conj = (ky < 0) || (ky == 0 && (z > N / 2 || (z == 0 && x > N / 2)));
    zo_mode m;
    if (conj) {
        /* the entry is the conjugate-structured twin of the primary mode at -k,
         * where -k is formed on lattice INDICES (N-i, 0->0) and then wrapped, so an
         * index of N/2 stays +N/2 */
        kx = zo_wrap((N - x) % N, N);
        ky = zo_wrap((N - y) % N, N);
        kz = zo_wrap((N - z) % N, N);
    }
    zo_primary(st, seed_state, kx, ky, kz, &m);
    zo_pack(&m, conj, out);
}

/* ------------------------------------------------------------ FFT ---------- */
/* Unnormalised backward DFT (sign +1, src/zeldovich.cpp:61-62): iterative radix-2 for powers of two, the direct
 * O(n^2) sum (twiddle index reduced mod n, so every phase is a table entry) for any other length — the reference takes
 * any even ppd (src/block_array.cpp:38-40); only small non-power-of-two sizes are run through the oracle. */
static void zo_fft1(double *re_im, int64_t n, int64_t stride, const double *tw, double *tmp) {
    if (n & (n - 1)) {
        for (int64_t k = 0; k < n; k++) {
            double sr = 0.0, si = 0.0;
            for (int64_t j = 0; j < n; j++) {
                const int64_t t = (j * k) % n;
                const double wr = tw[2 * t], wi = tw[2 * t + 1], ar = re_im[2 * j * stride], ai = re_im[2 * j * stride + 1];
                sr += ar * wr - ai * wi;
                si += ar * wi + ai * wr;
            }
            tmp[2 * k] = sr, tmp[2 * k + 1] = si;
        }
        for (int64_t i = 0; i < n; i++) {
            re_im[2 * i * stride]     = tmp[2 * i];
            re_im[2 * i * stride + 1] = tmp[2 * i + 1];
        }
        return;
    }
    int lg = 0;
    while ((1LL << lg) < n) lg++;
    for (int64_t i = 0; i < n; i++) {
        int64_t r = 0;
        for (int b = 0; b < lg; b++)
            if (i & (1LL << b)) r |= 1LL << (lg - 1 - b);
        tmp[2 * r]     = re_im[2 * i * stride];
        tmp[2 * r + 1] = re_im[2 * i * stride + 1];
    }
    for (int64_t len = 2; len <= n; len <<= 1) {
        int64_t half = len >> 1, step = n / len;
        for (int64_t i = 0; i < n; i += len)
            for (int64_t j = 0; j < half; j++) {
                double wr = tw[2 * j * step], wi = tw[2 * j * step + 1];
                double ar = tmp[2 * (i + j)], ai = tmp[2 * (i + j) + 1];
                double br = tmp[2 * (i + j + half)], bi = tmp[2 * (i + j + half) + 1];
                double tr = br * wr - bi * wi, ti = br * wi + bi * wr;
                tmp[2 * (i + j)]            = ar + tr;
                tmp[2 * (i + j) + 1]        = ai + ti;
                tmp[2 * (i + j + half)]     = ar - tr;
                tmp[2 * (i + j + half) + 1] = ai - ti;
            }
    }
    for (int64_t i = 0; i < n; i++) {
        re_im[2 * i * stride]     = tmp[2 * i];
        re_im[2 * i * stride + 1] = tmp[2 * i + 1];
    }
}

/* in-place 3-D backward DFT of a [n][n][n] complex cube (interleaved re,im); n a power of two */
void zo_fft3_backward(double *a, int64_t n) {
    double *tw = (double *) malloc(sizeof(double) * 2 * n);
    for (int64_t j = 0; j < n; j++) {
        long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double) j / (long double) n;
        tw[2 * j]       = (double) cosl(ang);
        tw[2 * j + 1]   = (double) sinl(ang);
    }
#pragma omp parallel
    {
        double *tmp = (double *) malloc(sizeof(double) * 2 * n);
#pragma omp for collapse(2)
        for (int64_t i = 0; i < n; i++)
            for (int64_t j = 0; j < n; j++) zo_fft1(a + 2 * ((i * n + j) * n), n, 1, tw, tmp);
#pragma omp for collapse(2)
        for (int64_t i = 0; i < n; i++)
            for (int64_t k = 0; k < n; k++) zo_fft1(a + 2 * (i * n * n + k), n, n, tw, tmp);
#pragma omp for collapse(2)
        for (int64_t j = 0; j < n; j++)
            for (int64_t k = 0; k < n; k++) zo_fft1(a + 2 * (j * n + k), n, n * n, tw, tmp);
        free(tmp);
    }
    free(tw);
}

/* ------------------------------------------------------------ drivers ------ */
static void zo_init_state(zo_state *st, const zo_config *cfg, int nrows, const double *ks, const double *ps, int64_t eig_ppd,
                          const double *eig) {
    memset(st, 0, sizeof(*st));
    st->c = *cfg;
    zo_power_setup(st, nrows, ks, ps);
    st->eig_ppd = eig_ppd;
    st->eig     = eig;
}

int zo_narray(const zo_config *cfg) { return cfg->qPLT ? 4 : 2; }

/* Spectral cube before any FFT: out is [narray][z][y][x] complex (interleaved). */
void zo_spectral_cube(const zo_config *cfg, int nrows, const double *ks, const double *ps, int64_t eig_ppd, const double *eig,
                      double *out) {
    zo_state st;
    zo_init_state(&st, cfg, nrows, ks, ps, eig_ppd, eig);
    const int64_t N = cfg->ppd;
    const int na    = zo_narray(cfg);
    u128 s0         = zo_seed_state((uint64_t) (int64_t) cfg->seed);
    double *phi     = NULL;
    if (cfg->f_NL != 0.) {
        /* main(), src/zeldovich.cpp:945-960: Gaussian potential phi_g(k) = D/M with the Hermitian structure of every
         * other array (ZeldovichZ with gen_phi = 1, :388-394, :485-503), backward transform, local transformation
         * phi_g + f_NL phi_g^2 on the REAL part, normalised by ppd^3 (ZeldovichXY_Phi, :744-755), forward transform
         * (Forward2dFFT :765 + ForwardFFT_Yonly :325) */
        phi = (double *) malloc(sizeof(double) * 2 * N * N * N);
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
        for (int64_t z = 0; z < N; z++)
            for (int64_t y = 0; y < N; y++)
                for (int64_t x = 0; x < N; x++) {
                    double *o = phi + 2 * ((z * N + y) * N + x);
                    o[0] = o[1] = 0.0;
                    if (y == N / 2 || (x == 0 && y == 0 && z == 0)) continue;
                    int kx = zo_wrap(x, N), ky = zo_wrap(y, N), kz = zo_wrap(z, N);
                    int conj = (ky < 0) || (ky == 0 && (z > N / 2 || (z == 0 && x > N / 2)));
                    if (conj) kx = zo_wrap((N - x) % N, N), ky = zo_wrap((N - y) % N, N), kz = zo_wrap((N - z) % N, N);
                    zo_mode m;
                    zo_primary(&st, s0, kx, ky, kz, &m);
                    o[0] = m.Dr / m.M;
                    o[1] = (conj ? -m.Di : m.Di) / m.M;
                }
        zo_fft3_backward(phi, N);
        const double inv = 1. / N / N / N;
        for (int64_t i = 0; i < N * N * N; i++) {
            double p       = phi[2 * i];
            phi[2 * i]     = (p + cfg->f_NL * p * p) * inv;
            phi[2 * i + 1] = 0.0;
        }
        /* forward transform of a real field = conjugate of its backward transform */
        zo_fft3_backward(phi, N);
        for (int64_t i = 0; i < N * N * N; i++) phi[2 * i + 1] = -phi[2 * i + 1];
        st.phi = phi;
    }
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int64_t z = 0; z < N; z++)
        for (int64_t y = 0; y < N; y++)
            for (int64_t x = 0; x < N; x++) {
                double v[8];
                zo_site(&st, s0, x, y, z, v);
                for (int a = 0; a < na; a++) {
                    int64_t idx      = ((a * N + z) * N + y) * N + x;
                    out[2 * idx]     = v[2 * a];
                    out[2 * idx + 1] = v[2 * a + 1];
                }
            }
    if (st.have_spline) zo_spline_free(&st.sp);
    free(phi);
}

size_t zo_record_bytes(int icformat) {
    switch (icformat) {
        case 0: return 32; /* ZelParticle: 3 x u16, pad, double[3] */
        case 1: return 32; /* RVZelParticle: 3 x u16, pad, float[3], float[3] */
        case 2: return 56; /* RVdoubleZelParticle */
        case 3: return 12; /* ZelSimpleParticle */
    }
    return 0;
}

/* one particle record (src/output.cpp:128-156; layouts include/output.h:19-42), padding bytes zeroed */
static void zo_write_record(int icformat, unsigned char *r, size_t rb, int64_t z, int64_t y, int64_t x, const double pos[3],
                            const double vel[3]) {
    memset(r, 0, rb);
    uint16_t ijk[3] = {(uint16_t) z, (uint16_t) y, (uint16_t) x};
    if (icformat == 0) {
        memcpy(r, ijk, 6);
        double d[3] = {pos[2], pos[1], pos[0]};
        memcpy(r + 8, d, 24);
    } else if (icformat == 1) {
        memcpy(r, ijk, 6);
        float d[6] = {(float) pos[2], (float) pos[1], (float) pos[0], (float) vel[2], (float) vel[1], (float) vel[0]};
        memcpy(r + 8, d, 24);
    } else if (icformat == 2) {
        memcpy(r, ijk, 6);
        double d[6] = {pos[2], pos[1], pos[0], vel[2], vel[1], vel[0]};
        memcpy(r + 8, d, 48);
    } else {
        float d[3] = {(float) pos[2], (float) pos[1], (float) pos[0]};
        memcpy(r, d, 12);
    }
}

/* WriteParticlesSlab (src/output.cpp:41-234) for all z: records in [z][y][x] order
 * with padding bytes zeroed; stats = {sum dens^2, max_disp[0..2]} where max_disp
 * keeps the SIGNED value of the largest |pos[j]| seen in (z,y,x) scan order. */
void zo_emit(const zo_config *cfg, const double *cube, unsigned char *records, double *stats) {
    const int64_t N = cfg->ppd;
    const int na    = zo_narray(cfg);
    const size_t rb = zo_record_bytes(cfg->icformat);
    double vnorm    = cfg->qPLT ? 1.0 : (sqrt(1. + 24 * cfg->f_cluster) - 1) * .25;
    double var = 0.0, md[3] = {0, 0, 0};
    for (int64_t z = 0; z < N; z++)
        for (int64_t y = 0; y < N; y++)
            for (int64_t x = 0; x < N; x++) {
                int64_t i0 = ((0 * N + z) * N + y) * N + x, i1 = ((1 * N + z) * N + y) * N + x;
                double dens = cube[2 * i0];
                double pos[3] = {cube[2 * i0 + 1], cube[2 * i1], cube[2 * i1 + 1]};
                double vel[3];
                if (na == 4) {
                    int64_t i2 = ((2 * N + z) * N + y) * N + x, i3 = ((3 * N + z) * N + y) * N + x;
                    vel[0] = cube[2 * i2 + 1] * vnorm;
                    vel[1] = cube[2 * i3] * vnorm;
                    vel[2] = cube[2 * i3 + 1] * vnorm;
                } else {
                    vel[0] = pos[0] * vnorm;
                    vel[1] = pos[1] * vnorm;
                    vel[2] = pos[2] * vnorm;
                }
                zo_write_record(cfg->icformat, records + rb * (size_t) ((z * N + y) * N + x), rb, z, y, x, pos, vel);
                for (int j = 0; j < 3; j++)
                    if (fabs(pos[j]) > fabs(md[j])) md[j] = pos[j];
                var += dens * dens;
            }
    stats[0] = var;
    stats[1] = md[0];
    stats[2] = md[1];
    stats[3] = md[2];
}

/* Whole hot path: records [z][y][x] + stats.  work = scratch for narray*N^3 complex (or NULL). */
int zo_run(const zo_config *cfg, int nrows, const double *ks, const double *ps, int64_t eig_ppd, const double *eig,
           unsigned char *records, double *stats) {
    const int64_t N = cfg->ppd;
    const int na    = zo_narray(cfg);
    if (N & 1) return 1;
    double *cube = (double *) malloc(sizeof(double) * 2 * na * N * N * N);
    if (!cube) return 2;
    zo_spectral_cube(cfg, nrows, ks, ps, eig_ppd, eig, cube);
    for (int a = 0; a < na; a++) zo_fft3_backward(cube + 2 * a * N * N * N, N);
    zo_emit(cfg, cube, records, stats);
    free(cube);
    return 0;
}

/* ------------------------------------------------------------ selected planes at any size ----
 * Records of nplanes chosen z planes without ever holding the cube: the z transform of those planes is summed
 * directly, A(x,y,z0) = sum_z' S(x,y,z') exp(2 pi i z' z0 / N) with S the spectral entries of zo_site, then the
 * plane goes through the 2-D transform and the record writer of zo_emit.  O(N^3) mode evaluations (each primary
 * mode is evaluated once and feeds its own site and its twin's), O(nplanes narray N^2) memory — PPD=1024 and
 * 2048 fit any host, which is what lets the benchmark configurations face the oracle.
 *
 * Same arithmetic as zo_run per mode (zo_masked / zo_cgauss / zo_finish_mode / zo_pack); what differs is
 * book-keeping only: P(k) is tabulated per integer |k|^2 with the very expression zo_primary evaluates, and the
 * generator walks a row of consecutive kx sequentially (consecutive kx are consecutive draw pairs, masked sites
 * consume theirs: the reference's nskip walk, src/zeldovich.cpp:335,341,358-363) instead of jumping per mode.
 * tests/test_oracle.py pins it to zo_run.  No f_NL (the potential pass needs the whole cube).
 * records: [nplanes][N][N] records; stats: [nplanes][4] = sum dens^2, max_disp[3] of each plane. */
int zo_planes(const zo_config *cfg, int nrows, const double *ks, const double *ps, int64_t eig_ppd, const double *eig, int nplanes,
              const int64_t *zs, unsigned char *records, double *stats) {
    const int64_t N = cfg->ppd, M = 65536, half = N / 2;
    const int na = zo_narray(cfg);
    if ((N & 1) || cfg->f_NL != 0. || nplanes < 1 || nplanes > 16) return 1;
    for (int p = 0; p < nplanes; p++)
        if (zs[p] < 0 || zs[p] >= N) return 1;
    zo_state st;
    zo_init_state(&st, cfg, nrows, ks, ps, eig_ppd, eig);
    const u128 s0 = zo_seed_state((uint64_t) (int64_t) cfg->seed);
    const double fund = 2.0 * M_PI / cfg->boxsize, fund2 = fund * fund;
    const int64_t nm = 3 * half * half + 1;
    double *ptab = (double *) malloc(sizeof(double) * nm);
    double *tw   = (double *) malloc(sizeof(double) * 2 * N);
    /* accumulators [y][x][p][a] while summing (one mode touches one contiguous run), [p][a][y][x] for the 2-D transform */
    double *acc0 = (double *) calloc((size_t) nplanes * na * N * N * 2, sizeof(double));
    double *acc  = (double *) malloc((size_t) nplanes * na * N * N * 2 * sizeof(double));
    if (!ptab || !tw || !acc || !acc0) return 2;
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < nm; m++) ptab[m] = zo_power(&st, sqrt((double) m * fund2));
    for (int64_t j = 0; j < N; j++) {
        long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double) j / (long double) N;
        tw[2 * j]       = (double) cosl(ang);
        tw[2 * j + 1]   = (double) sinl(ang);
    }
#define ZO_ACC(p, a, y, x) (acc0 + 2 * ((((size_t) (y) * N + (x)) * nplanes + (p)) * na + (a)))
    /* the ky = 0 plane (its own twin, src/zeldovich.cpp:485-503): plain zo_site per entry */
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t x = 0; x < N; x++)
        for (int64_t z = 0; z < N; z++) {
            double v[8];
            zo_site(&st, s0, x, 0, z, v);
            for (int p = 0; p < nplanes; p++) {
                const int64_t j = (z * zs[p]) % N;
                const double wr = tw[2 * j], wi = tw[2 * j + 1];
                for (int a = 0; a < na; a++) {
                    double *o = ZO_ACC(p, a, 0, x);
                    o[0] += v[2 * a] * wr - v[2 * a + 1] * wi;
                    o[1] += v[2 * a] * wi + v[2 * a + 1] * wr;
                }
            }
        }
    /* primary rows ky = 1 .. N/2-1 and their twins at y = N - ky; the row y = N/2 stays zero */
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t ky = 1; ky < half; ky++)
        for (int64_t z = 0; z < N; z++) {
            const int kz = zo_wrap(z, N);
            /* whole row masked?  The plane tests do not involve kx and the sphere cut grows with kx^2, so the row is
             * masked as soon as its kx = 0 site is (not with ZD_qonemode, whose test involves kx) */
            if (!cfg->qonemode && zo_masked(cfg, 0, (int) ky, kz, (double) (ky * ky + (int64_t) kz * kz) * fund2)) continue;
            const u128 rowoff = 2 * ((u128) ky * M * M + (u128) (kz < 0 ? kz + M : kz) * M);
            u128 s = zo_jump(s0, rowoff);
            const int64_t zt = (N - z) % N; /* the twin's plane */
            double pw[16][2], pwt[16][2];
            for (int p = 0; p < nplanes; p++) {
                int64_t j = (z * zs[p]) % N, jt = (zt * zs[p]) % N;
                pw[p][0] = tw[2 * j], pw[p][1] = tw[2 * j + 1];
                pwt[p][0] = tw[2 * jt], pwt[p][1] = tw[2 * jt + 1];
            }
            for (int64_t x = 0; x < N; x++) {
                const int kx = zo_wrap(x, N);
                if (x == half + 1) s = zo_jump(s0, rowoff + 2 * (u128) (kx + M)); /* kx = -N/2+1: position (kx mod 65536) */
                const uint64_t r1 = zo_next(&s), r2 = zo_next(&s);
                const int64_t n2 = (int64_t) kx * kx + ky * ky + (int64_t) kz * kz;
                double k2 = (kx * kx + (int) ky * (int) ky + kz * kz) * fund2;
                if (zo_masked(cfg, kx, (int) ky, kz, k2)) continue;
                zo_mode m;
                memset(&m, 0, sizeof(m));
                zo_cgauss(cfg, ptab[n2], r1, r2, &m);
                if (m.Dr == 0.0 && m.Di == 0.0) continue;
                zo_finish_mode(&st, kx, (int) ky, kz, k2, &m);
                double v[8], vt[8];
                zo_pack(&m, 0, v);
                zo_pack(&m, 1, vt);
                const int64_t xt = (N - x) % N, yt = N - ky;
                for (int p = 0; p < nplanes; p++)
                    for (int a = 0; a < na; a++) {
                        double *o = ZO_ACC(p, a, ky, x);
                        o[0] += v[2 * a] * pw[p][0] - v[2 * a + 1] * pw[p][1];
                        o[1] += v[2 * a] * pw[p][1] + v[2 * a + 1] * pw[p][0];
                        double *t = ZO_ACC(p, a, yt, xt);
                        t[0] += vt[2 * a] * pwt[p][0] - vt[2 * a + 1] * pwt[p][1];
                        t[1] += vt[2 * a] * pwt[p][1] + vt[2 * a + 1] * pwt[p][0];
                    }
            }
        }
#pragma omp parallel for schedule(static)
    for (int64_t y = 0; y < N; y++)
        for (int64_t x = 0; x < N; x++)
            for (int p = 0; p < nplanes; p++)
                for (int a = 0; a < na; a++) {
                    const double *src = ZO_ACC(p, a, y, x);
                    double *dst       = acc + 2 * ((((size_t) p * na + a) * N + y) * N + x);
                    dst[0] = src[0], dst[1] = src[1];
                }
    free(acc0);
#undef ZO_ACC
#define ZO_ACC(p, a, y, x) (acc + 2 * ((((size_t) (p) * na + (a)) * N + (y)) * N + (x)))
    /* 2-D backward transform of every accumulated plane */
#pragma omp parallel
    {
        double *tmp = (double *) malloc(sizeof(double) * 2 * N);
#pragma omp for collapse(2)
        for (int64_t q = 0; q < (int64_t) nplanes * na; q++)
            for (int64_t y = 0; y < N; y++) zo_fft1(acc + 2 * ((size_t) q * N * N + y * N), N, 1, tw, tmp);
#pragma omp for collapse(2)
        for (int64_t q = 0; q < (int64_t) nplanes * na; q++)
            for (int64_t x = 0; x < N; x++) zo_fft1(acc + 2 * ((size_t) q * N * N + x), N, N, tw, tmp);
        free(tmp);
    }
    /* WriteParticlesSlab for the chosen planes (src/output.cpp:41-234), as zo_emit */
    const size_t rb = zo_record_bytes(cfg->icformat);
    const double vnorm = cfg->qPLT ? 1.0 : (sqrt(1. + 24 * cfg->f_cluster) - 1) * .25;
    for (int p = 0; p < nplanes; p++) {
        double var = 0.0, md[3] = {0, 0, 0};
        for (int64_t y = 0; y < N; y++)
            for (int64_t x = 0; x < N; x++) {
                const double *a0 = ZO_ACC(p, 0, y, x), *a1 = ZO_ACC(p, 1, y, x);
                double dens = a0[0];
                double pos[3] = {a0[1], a1[0], a1[1]}, vel[3];
                if (na == 4) {
                    const double *a2 = ZO_ACC(p, 2, y, x), *a3 = ZO_ACC(p, 3, y, x);
                    vel[0] = a2[1] * vnorm, vel[1] = a3[0] * vnorm, vel[2] = a3[1] * vnorm;
                } else {
                    vel[0] = pos[0] * vnorm, vel[1] = pos[1] * vnorm, vel[2] = pos[2] * vnorm;
                }
                zo_write_record(cfg->icformat, records + rb * ((size_t) p * N * N + (size_t) (y * N + x)), rb, zs[p], y, x, pos, vel);
                for (int j = 0; j < 3; j++)
                    if (fabs(pos[j]) > fabs(md[j])) md[j] = pos[j];
                var += dens * dens;
            }
        stats[4 * p] = var, stats[4 * p + 1] = md[0], stats[4 * p + 2] = md[1], stats[4 * p + 3] = md[2];
    }
#undef ZO_ACC
    free(ptab), free(tw), free(acc);
    if (st.have_spline) zo_spline_free(&st.sp);
    return 0;
}
