// TEST INFRASTRUCTURE — oracle build shim, not product code.
//
// Abort-stubs for the four GSL calls the reference makes on its deprecated
// ZD_Version=1 path only (reference src/power_spectrum.cpp:19-23,43,279).
// GSL is an external dependency (meson.build:36) that is absent from this image;
// ZD_Version=1 is out of scope, so reaching any of these is an error.
#pragma once
#include <cstdio>
#include <cstdlib>

struct gsl_rng { int unused; };
struct gsl_rng_type { int unused; };
static const gsl_rng_type zshim_mt19937 = {0};
static const gsl_rng_type *gsl_rng_mt19937 = &zshim_mt19937;

static inline void zshim_gsl_abort(const char *fn) {
    fprintf(stderr, "oracle shim: %s called — ZD_Version=1 (GSL mt19937) is not available in the oracle build\n", fn);
    abort();
}
static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *) { zshim_gsl_abort("gsl_rng_alloc"); return NULL; }
static inline void gsl_rng_set(gsl_rng *, unsigned long) { zshim_gsl_abort("gsl_rng_set"); }
static inline double gsl_rng_uniform(gsl_rng *) { zshim_gsl_abort("gsl_rng_uniform"); return 0; }
static inline void gsl_rng_free(gsl_rng *) {}
