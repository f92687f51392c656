// TEST INFRASTRUCTURE — oracle build shim, not product code.
//
// FP64 complex-to-complex FFT backing the fftw3.h shim.  Stockham autosort,
// radix-4 passes with one radix-2 pass when log2(n) is odd; twiddles come from a
// table computed once per plan with sincos in long double.  Power-of-two lengths
// only take the fast path; other lengths use a direct O(n^2) DFT (tiny test
// sizes only).  Plans are immutable after creation, so fftw_execute_dft may be
// called concurrently from OpenMP threads as the reference does
// (reference src/zeldovich.cpp:572-577, :654-657).
#include "fftw3.h"

#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef std::complex<double> cplx;

struct zshim_plan_s {
    int rank;
    int n0, n1;
    int sign;
    std::vector<cplx> tw0, tw1;  // tw[j] = exp(sign*2*pi*i*j/n)
};

static std::vector<cplx> make_twiddles(int n, int sign) {
    std::vector<cplx> tw(n);
    for (int j = 0; j < n; j++) {
        long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double) j / (long double) n;
        tw[j]         = cplx((double) cosl(a), (double) (sign * sinl(a)));
    }
    return tw;
}

static inline bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

// x: n elements with element stride 1; y: scratch of n elements.  Result ends in x.
static void fft_pow2(cplx *x, cplx *y, int n, const cplx *tw, int sign) {
    // Stockham: at each pass, l = number of already-combined sub-transforms' length,
    // m = n / (l * radix).
    cplx *src = x, *dst = y;
    int l = 1;
    int rem = n;
    const cplx isgn(0.0, (double) sign);
    while (rem > 1) {
        if (rem % 4 == 0) {
            int m = rem / 4;  // stride between the 4 inputs is m*l? (see indexing below)
            // input index:  j + l*(k + m*q)   q=0..3 (decimation in frequency style over k)
            // We use the DIT Stockham form:
            //   for k in [0,m): for j in [0,l):
            //     a_q = src[j + l*(k + m*q)]
            //     w^q with w = exp(sign*2pi*i * j*? ...
            // To keep this obviously correct we use the DIF Stockham form instead:
            //   n_cur = rem, s = l
            //   for p in [0,m): w1 = W_{rem}^p
            //     for q in [0,s):
            //        a,b,c,d = src[q + s*(p + m*{0,1,2,3})]
            //        dst[q + s*(4p+0)] =  a+b+c+d
            //        dst[q + s*(4p+1)] = (a + i*sgn*b - c - i*sgn*d) * w1
            //        dst[q + s*(4p+2)] = (a - b + c - d) * w2
            //        dst[q + s*(4p+3)] = (a - i*sgn*b - c + i*sgn*d) * w3
            int s        = l;
            int twstride = n / rem;
            for (int p = 0; p < m; p++) {
                cplx w1 = tw[(size_t) p * twstride];
                cplx w2 = tw[(size_t) 2 * p * twstride];
                cplx w3 = tw[(size_t) 3 * p * twstride];
                for (int q = 0; q < s; q++) {
                    cplx a = src[q + s * (p + m * 0)];
                    cplx b = src[q + s * (p + m * 1)];
                    cplx c = src[q + s * (p + m * 2)];
                    cplx d = src[q + s * (p + m * 3)];
                    cplx apc = a + c, amc = a - c, bpd = b + d;
                    cplx bmd = b - d;
                    cplx jbmd(-sign * bmd.imag(), sign * bmd.real());  // i*sgn*(b-d)
                    dst[q + s * (4 * p + 0)] = apc + bpd;
                    dst[q + s * (4 * p + 1)] = (amc + jbmd) * w1;
                    dst[q + s * (4 * p + 2)] = (apc - bpd) * w2;
                    dst[q + s * (4 * p + 3)] = (amc - jbmd) * w3;
                }
            }
            l *= 4;
            rem /= 4;
        } else {
            int m        = rem / 2;
            int s        = l;
            int twstride = n / rem;
            for (int p = 0; p < m; p++) {
                cplx w1 = tw[(size_t) p * twstride];
                for (int q = 0; q < s; q++) {
                    cplx a                   = src[q + s * (p + m * 0)];
                    cplx b                   = src[q + s * (p + m * 1)];
                    dst[q + s * (2 * p + 0)] = a + b;
                    dst[q + s * (2 * p + 1)] = (a - b) * w1;
                }
            }
            l *= 2;
            rem /= 2;
        }
        cplx *t = src;
        src     = dst;
        dst     = t;
    }
    if (src != x) memcpy(x, src, sizeof(cplx) * n);
}

static void dft_naive(cplx *x, cplx *y, int n, const cplx *tw) {
    for (int k = 0; k < n; k++) {
        cplx acc(0, 0);
        for (int j = 0; j < n; j++) acc += x[j] * tw[(int) (((long long) j * k) % n)];
        y[k] = acc;
    }
    memcpy(x, y, sizeof(cplx) * n);
}

static void fft_any(cplx *x, cplx *scratch, int n, const cplx *tw, int sign) {
    if (is_pow2(n))
        fft_pow2(x, scratch, n, tw, sign);
    else
        dft_naive(x, scratch, n, tw);
}

extern "C" {

fftw_plan fftw_plan_dft_1d(int n, fftw_complex *, fftw_complex *, int sign, unsigned) {
    zshim_plan_s *p = new zshim_plan_s;
    p->rank         = 1;
    p->n0           = n;
    p->n1           = 1;
    p->sign         = sign;
    p->tw0          = make_twiddles(n, sign);
    return p;
}

fftw_plan fftw_plan_dft_2d(int n0, int n1, fftw_complex *, fftw_complex *, int sign, unsigned) {
    zshim_plan_s *p = new zshim_plan_s;
    p->rank         = 2;
    p->n0           = n0;
    p->n1           = n1;
    p->sign         = sign;
    p->tw0          = make_twiddles(n0, sign);
    p->tw1          = make_twiddles(n1, sign);
    return p;
}

void fftw_execute_dft(const fftw_plan p, fftw_complex *in, fftw_complex *out) {
    cplx *x = reinterpret_cast<cplx *>(out);
    if (in != out) {
        size_t tot = (size_t) p->n0 * p->n1;
        memcpy(out, in, tot * sizeof(cplx));
    }
    if (p->rank == 1) {
        std::vector<cplx> scratch(p->n0);
        fft_any(x, scratch.data(), p->n0, p->tw0.data(), p->sign);
        return;
    }
    // rank 2: array is x[n0][n1]; transform rows (length n1) then columns (length n0)
    const int n0 = p->n0, n1 = p->n1;
    {
        std::vector<cplx> scratch(n1);
        for (int r = 0; r < n0; r++) fft_any(x + (size_t) r * n1, scratch.data(), n1, p->tw1.data(), p->sign);
    }
    {
        const int CB = 8;  // columns per block, for cache reuse
        std::vector<cplx> col((size_t) CB * n0), scratch(n0);
        for (int c0 = 0; c0 < n1; c0 += CB) {
            int cb = (n1 - c0 < CB) ? (n1 - c0) : CB;
            for (int r = 0; r < n0; r++)
                for (int c = 0; c < cb; c++) col[(size_t) c * n0 + r] = x[(size_t) r * n1 + c0 + c];
            for (int c = 0; c < cb; c++) fft_any(col.data() + (size_t) c * n0, scratch.data(), n0, p->tw0.data(), p->sign);
            for (int r = 0; r < n0; r++)
                for (int c = 0; c < cb; c++) x[(size_t) r * n1 + c0 + c] = col[(size_t) c * n0 + r];
        }
    }
}

void fftw_destroy_plan(fftw_plan p) { delete p; }

// No wisdom in the shim: report "failure" on import (the reference prints and
// carries on, src/zeldovich.cpp:49-56) and do not leave a file behind on export.
int fftw_import_wisdom_from_filename(const char *) { return 0; }
int fftw_export_wisdom_to_filename(const char *) { return 0; }
}
