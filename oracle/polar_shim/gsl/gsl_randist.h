// TEST INFRASTRUCTURE ONLY -- empty stand-in for a GSL header: the vendored polar library includes it but the
// decoder path used as oracle never calls GSL.  Any accidental call aborts.
#pragma once
#include <cstdlib>
#ifndef PK_GSL_STUB
#define PK_GSL_STUB
struct gsl_rng; struct gsl_rng_type; struct gsl_interp; struct gsl_interp_accel; struct gsl_interp_type;
static const gsl_rng_type *gsl_rng_mt19937 = nullptr;
static const gsl_interp_type *gsl_interp_linear = nullptr;
static const gsl_interp_type *gsl_interp_cspline = nullptr;
inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *) { abort(); }
inline void gsl_rng_free(gsl_rng *) {}
inline void gsl_rng_set(gsl_rng *, unsigned long) { abort(); }
inline unsigned long gsl_rng_get(gsl_rng *) { abort(); }
inline double gsl_rng_uniform(gsl_rng *) { abort(); }
inline unsigned long gsl_rng_uniform_int(gsl_rng *, unsigned long) { abort(); }
inline double gsl_ran_gaussian(gsl_rng *, double) { abort(); }
inline double gsl_ran_gaussian_ziggurat(gsl_rng *, double) { abort(); }
inline gsl_interp *gsl_interp_alloc(const gsl_interp_type *, size_t) { return (gsl_interp *)malloc(8); }   // static capacity tables only
inline int gsl_interp_init(gsl_interp *, const double *, const double *, size_t) { return 0; }
inline void gsl_interp_free(gsl_interp *p) { free(p); }
inline gsl_interp_accel *gsl_interp_accel_alloc() { return (gsl_interp_accel *)malloc(8); }
inline void gsl_interp_accel_free(gsl_interp_accel *p) { free(p); }
inline double gsl_interp_eval(const gsl_interp *, const double *, const double *, double, gsl_interp_accel *) { abort(); }
inline double gsl_cdf_ugaussian_Q(double) { abort(); }
inline double gsl_cdf_ugaussian_Qinv(double) { abort(); }
#endif
