#include "host_power.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>
#include <numeric>

void SplineFunction::spline() {
    const int n = size();
    // order the nodes by abscissa (the reference shell-sorts in place; abscissae are distinct
    // so any correct sort gives the same arrays)
    std::vector<int> idx(n);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return x[a] < x[b]; });
    std::vector<double> xs(n), ys(n);
    for (int i = 0; i < n; i++) xs[i] = x[idx[i]], ys[i] = y[idx[i]];
    x.swap(xs);
    y.swap(ys);
    y2.assign(n, 0.0);
    if (n < 3) return;
    // natural spline: forward elimination then back-substitution of the tridiagonal system
    std::vector<double> u(n, 0.0);
    for (int i = 1; i <= n - 2; i++) {
        const double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
        const double p   = sig * y2[i - 1] + 2.0;
        y2[i]            = (sig - 1.0) / p;
        u[i]             = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
        u[i]             = (6.0 * u[i] / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p;
    }
    const double qn = 0.0, un = 0.0;
    y2[n - 1]       = (un - qn * u[n - 2]) / (qn * y2[n - 2] + 1.0);
    for (int k = n - 2; k >= 0; k--) y2[k] = y2[k] * y2[k + 1] + u[k];
}

double SplineFunction::val(double v) const {
    int klo = 0, khi = size() - 1;
    while (khi - klo > 1) {
        const int k = (khi + klo) >> 1;
        if (x[k] > v)
            khi = k;
        else
            klo = k;
    }
    const double h = x[khi] - x[klo];
    const double a = (x[khi] - v) / h;
    const double b = (v - x[klo]) / h;
    return a * y[klo] + b * y[khi] + ((a * a * a - a) * y2[klo] + (b * b * b - b) * y2[khi]) * (h * h) / 6.0;
}

PowerSpectrum::PowerSpectrum() {
    fixed_power = 0, is_powerlaw = 0, powerlaw_index = 1000;
    normalization = 1.0, Pk_smooth2 = 0.0, Rnorm = 0.0;
    kmin = std::numeric_limits<double>::max();
    kmax = std::numeric_limits<double>::min();
}

int PowerSpectrum::InitFromFile(const fs::path &filename, const PkParams &param) {
    fprintf(stderr, "Loading power spectrum from file \"%s\"\n", filename.c_str());
    FILE *fp = fopen(filename.c_str(), "r");
    if (!fp) {
        fprintf(stderr, "Power spectrum file \"%s\" not found; exiting.\n", filename.c_str());
        return 1;
    }
    char line[200];
    double k = 0, P = 0;
    while (fgets(line, 200, fp) != NULL) {
        if (line[0] == '#') continue;
        sscanf(line, "%lf %lf", &k, &P);  // a malformed line re-uses the previous pair, as in the reference
        if (k < 0.0) continue;
        if (P < 0.0) continue;
        k *= param.Pk_scale;
        if (k > 0.0) {
            load(log(k), log(P));
            kmin = std::min(k, kmin);
        } else {
            load(-1e3, log(P));
        }
        kmax = std::max(k, kmax);
    }
    fclose(fp);
    if (size() < 3) {
        fprintf(stderr, "Power spectrum file \"%s\" has fewer than 3 usable rows.\n", filename.c_str());
        return 1;
    }
    spline();
    Normalize(param);
    return 0;
}

int PowerSpectrum::InitFromPowerLaw(double index, const PkParams &param) {
    powerlaw_index = index;
    is_powerlaw    = 1;
    fprintf(stderr, "Initializing power spectrum with power law index %g\n", powerlaw_index);
    kmin = 1e-4;
    Normalize(param);
    return 0;
}

// reference src/power_spectrum.cpp:186-223
void PowerSpectrum::Normalize(const PkParams &param) {
    Pk_smooth2    = 0.0;
    normalization = 1.0;
    if (param.Pk_norm > 0.0) {
        fprintf(stderr, "Input sigma(%f) = %.6g\n", param.Pk_norm, sigmaR(param.Pk_norm));
        if (param.Pk_sigma > 0) {
            normalization = param.Pk_sigma / sigmaR(param.Pk_norm);
            normalization *= normalization;
        } else if (param.Pk_sigma_ratio > 0) {
            normalization = param.Pk_sigma_ratio * param.Pk_sigma_ratio;
        }
        fprintf(stderr, "Final sigma(%f) = %.6g\n", param.Pk_norm, sigmaR(param.Pk_norm));
    }
    normalization /= param.boxsize * param.boxsize * param.boxsize;
    Pk_smooth2 = param.Pk_smooth * param.Pk_smooth;
    fixed_power = param.qPk_fix_to_mean;
    if (fixed_power) fprintf(stderr, "Fixing density mode amplitudes to sqrt(P(k))\n");
    // T(k) = 1 at the smallest k of the input (reference src/power_spectrum.cpp:221-222)
    n_s             = param.n_s;
    primordial_norm = 1.;
    primordial_norm = power(kmin) / primordial_power(kmin);
}

double PowerSpectrum::primordial_power(double wavenumber) {
    if (wavenumber <= 0.0) return 0.0;
    return primordial_norm * exp(log(wavenumber) * n_s);
}

double PowerSpectrum::infer_Tk(double wavenumber) {
    if (wavenumber <= 0.0) return 1.0;
    return sqrt(power(wavenumber) / primordial_power(wavenumber));
}

// reference src/power_spectrum.cpp:225-261
double PowerSpectrum::power(double wavenumber) {
    if (wavenumber <= 0.0) return 0.0;
    if (is_powerlaw) return std::pow(wavenumber, powerlaw_index) * exp(-wavenumber * wavenumber * Pk_smooth2) * normalization;
    if (wavenumber > kmax && !warned_extrapolation_) {
        fprintf(stderr,
                "\n*** WARNING: power spectrum spline interpolation was requested\npast the maximum k (%f) that was provided in the "
                "input power\nspectrum file.  The extrapolation should be well-behaved, but\nmake sure that this was expected.\n\n",
                kmax);
        warned_extrapolation_ = true;
    }
    return exp(val(log(wavenumber)) - wavenumber * wavenumber * Pk_smooth2) * normalization;
}

// reference src/power_spectrum.cpp:50-58
double PowerSpectrum::sigmaR_integrand(double k) {
    const double xx = k * Rnorm;
    double w;
    if (xx <= 1e-3)
        w = 1 - xx * xx / 10.0;
    else
        w = 3.0 * (sin(xx) - xx * cos(xx)) / xx / xx / xx;
    return 0.5 / M_PI / M_PI * k * k * w * w * power(k);
}

// reference src/power_spectrum.cpp:60-89
double PowerSpectrum::sigmaR(double R) {
    if (!is_powerlaw) {
        const double target_prec = 1e-6;
        double precision         = 1.0;
        Rnorm                    = R;
        const double retval      = sqrt(Romberg(0, 10.0, target_prec, &precision));
        if (precision > target_prec)
            throw ParameterError("actual Romberg integration precision is greater than the target precision; halting.");
        return retval;
    }
    const double n = powerlaw_index;
    double retval  = 9 * pow(R, -n - 3) / (2 * M_PI * sqrt(M_PI)) * tgamma((3 + n) / 2.) / (tgamma((2 - n) / 2.) * (n - 3) * (n - 1));
    return sqrt(retval * normalization);
}

// Romberg quadrature with up to 32 interval halvings (reference src/power_spectrum.cpp:94-128).
// The trapezoid sums run serially in index order; the reference's OpenMP reduction makes
// its own last bits depend on the thread count, so agreement is to ~1e-16 relative.
double PowerSpectrum::Romberg(double a, double b, double prec, double *obtprec) {
    const int kMaxLevel = 32;
    // tab[level][order]: order 1 is the trapezoid rule on 2^level panels, higher orders are
    // the Richardson extrapolants
    std::vector<std::vector<double>> tab(kMaxLevel + 1, std::vector<double>(kMaxLevel + 2, 0.0));
    double width = 0.5 * (b - a);
    tab[0][1]    = width * (sigmaR_integrand(a) + sigmaR_integrand(b));
    int level    = 0;
    for (;;) {
        level++;
        double midsum       = 0;
        const uint64_t npts = 1ULL << (level - 1);
        for (uint64_t i = 1; i <= npts; i++) midsum += sigmaR_integrand(a + (2 * i - 1) * width);
        tab[level][1] = 0.5 * tab[level - 1][1] + width * midsum;
        double pow4   = 1;
        for (int order = 2; order <= level; order++) {
            pow4 *= 4;
            tab[level][order] = tab[level][order - 1] + (tab[level][order - 1] - tab[level - 1][order - 1]) / (pow4 - 1);
        }
        width *= 0.5;
        const bool converged = level > 1 && fabs(tab[level][level] - tab[level - 1][level - 1]) < prec * fabs(tab[level][level]);
        if (converged || level >= kMaxLevel) break;
    }
    *obtprec = (tab[level][level] - tab[level - 1][level - 1]) / tab[level][level];
    return tab[level][level];
}
