// TEST INFRASTRUCTURE — oracle build shim, not product code.
//
// Minimal stand-in for the subset of the FFTW3 API that the reference calls
// (reference src/zeldovich.cpp:41-75 Setup_FFTW, :83-92 Inverse1dFFT/Inverse2dFFT,
// :116-121 forward twins).  FFTW itself is an external dependency of the
// reference (meson.build:37, no version pin) and is not installed in this image.
// The unnormalised c2c DFT is mathematically defined, so any correct FP64 FFT
// stands in at the 1e-15 level; the implementation lives in fftw3_shim.cpp.
#pragma once
#include <cstddef>

extern "C" {
typedef double fftw_complex[2];
struct zshim_plan_s;
typedef zshim_plan_s *fftw_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)

fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
fftw_plan fftw_plan_dft_2d(int n0, int n1, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
void fftw_execute_dft(const fftw_plan p, fftw_complex *in, fftw_complex *out);
void fftw_destroy_plan(fftw_plan p);
int fftw_import_wisdom_from_filename(const char *filename);
int fftw_export_wisdom_to_filename(const char *filename);
}
