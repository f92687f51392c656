// Host half of the power spectrum: table load, natural cubic spline in (ln k, ln P) and
// the sigma(R) normalisation.  The NUMBERS produced here (spline nodes and second
// derivatives, `normalization`) feed the device kernels, so every floating-point
// expression is evaluated in the same order as the reference does
// (reference include/spline_function.h:54-163, src/power_spectrum.cpp:50-261);
// see SURVEY.md trap T1: an independent quadrature shifts every amplitude by ~2e-8.
#pragma once
#include <string>
#include <vector>

#include "host_parameters.h"

class SplineFunction {
public:
    std::vector<double> x, y, y2;
    int size() const { return (int) x.size(); }
    void load(double xv, double yv) {
        x.push_back(xv);
        y.push_back(yv);
    }
    void spline();            // sorts nodes by abscissa and builds y2 (natural end conditions)
    double val(double v) const;
};

// the subset of Parameters the power spectrum needs
struct PkParams {
    double boxsize, Pk_scale, Pk_norm, Pk_sigma, Pk_sigma_ratio, Pk_smooth, n_s;
    int qPk_fix_to_mean;
    PkParams() : boxsize(0), Pk_scale(1), Pk_norm(0), Pk_sigma(0), Pk_sigma_ratio(0), Pk_smooth(0), n_s(1), qPk_fix_to_mean(0) {}
    explicit PkParams(const Parameters &p)
        : boxsize(p.boxsize), Pk_scale(p.Pk_scale), Pk_norm(p.Pk_norm), Pk_sigma(p.Pk_sigma), Pk_sigma_ratio(p.Pk_sigma_ratio),
          Pk_smooth(p.Pk_smooth), n_s(p.n_s), qPk_fix_to_mean(p.qPk_fix_to_mean) {}
};

class PowerSpectrum : public SplineFunction {
public:
    PowerSpectrum();
    int fixed_power, is_powerlaw;
    double powerlaw_index, normalization, Pk_smooth2, Rnorm, kmax, kmin;
    double n_s, primordial_norm;  // ZD_f_NL: P(kmin) / kmin^n_s (reference src/power_spectrum.cpp:221-222)

    int InitFromFile(const fs::path &filename, const PkParams &param);
    int InitFromPowerLaw(double index, const PkParams &param);
    void Normalize(const PkParams &param);
    double power(double wavenumber);
    double primordial_power(double wavenumber);  // reference src/power_spectrum.cpp:263-266
    double infer_Tk(double wavenumber);          // reference src/power_spectrum.cpp:268-274
    double sigmaR(double R);

private:
    double sigmaR_integrand(double k);
    double Romberg(double a, double b, double prec, double *obtprec);
    bool warned_extrapolation_ = false;
};
